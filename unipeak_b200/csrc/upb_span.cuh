// upb_span.cuh — the region-calling kernels (sm_100a) and the small kernels that order, stitch and filter their output.
//
//   warm_span_kernel   which 4096-position spans of a (ProfileBuffer, contig) unit can hold a region at all (per-mini-tile zone map)
//   hot_span_kernel    which 32-position blocks of those spans can (per-block zone map, Tags::blkW): a block is "hot" when
//                      sum_j blkW[B+j] * (largest tap any tag of block B+j can have for any position of block B) >= threshold — an
//                      upper envelope of every score in the block (control samples and rounding only loosen it); a block that is
//                      not hot cannot hold a position of a region (misc/peakcall.cpp:57-60) and is never looked at again.
//                      Neither kernel touches a tag entry.
//   span_kernel        ONE WARP works on one span that has a hot block, four spans per ticket; no block-level barrier anywhere: the
//                      warps of a CTA only share the read-only tap table (v2 gave a whole CTA one 16384-position span and lost 63 %
//                      of its issue slots at __syncthreads — profiles/r01_v2_*).  Maximal groups of hot blocks are cut into pieces
//                      of <= 16 blocks; per piece the tags of the mini-tiles its window touches (one contiguous chunk of the
//                      tile-major copy per span, 8 bytes per entry, four loads in flight per lane) are pooled into dense
//                      per-position cells; every position is classified against the threshold with the O(1) integer moment form of
//                      the Epanechnikov sum (exact integers, window sums carried by warp scans) and re-evaluated in the reference's
//                      own floating-point order (ascending tag position, product and sum rounded separately,
//                      misc/peakcall.cpp:203-209) only where the fast value is within its error bound of the threshold or of the
//                      run maximum.  Runs of above-threshold positions are written out with their first strict maximum
//                      (misc/data.cpp:97-101) and pooled hit sum.
//
// The genome-wide profile is never materialised.  Runs cut by a piece or span boundary are stitched by merge_runs2_kernel;
// region_pool_kernel + region_moment_kernel then compute Region::exptSums and posKurtosis for the candidates that pass the hit filter.
#pragma once

#include <type_traits>

#include "upb_device.cuh"
#include "upb_edge.cuh"

#ifndef UPB_EXP_EXACT
#define UPB_EXP_EXACT 1
#endif

namespace upb {

#ifndef UPB_SP_MIN_BLOCKS
#define UPB_SP_MIN_BLOCKS 4
#endif
constexpr int SP_THREADS = 256;
constexpr int SP_MIN_BLOCKS = UPB_SP_MIN_BLOCKS;
constexpr int SP_WARPS = SP_THREADS / 32;
constexpr int HOT_JMAX = (448 - 1) / 32 + 1;        // blocks on either side whose tags can reach a block's positions (bandwidth <= 448)
constexpr int SPAN_BLOCKS = SPAN_LEN / 32;          // 128 blocks of 32 positions
constexpr int SPAN_WORDS = SPAN_BLOCKS / 32;
constexpr int PIECE_BLOCKS = 16;                    // <= 512 positions per piece
constexpr int PIECE_LEN = PIECE_BLOCKS * 32;
constexpr uint32_t RUN_CHUNK = 32;                  // run slots a warp reserves at a time
constexpr uint32_t RUN_INVALID = 0xFFFFFFFFu;       // SpanRun::span of a reserved slot that was never filled
constexpr int MAX_BW_FAST = MAX_BW_SPAN;            // piece cells = PIECE_LEN + 2*bw <= 1408
constexpr int MO_N = SPAN_MT + 2 * HALO_MT + 1;     // mini-tile offsets a span needs: its own mini-tiles, the halo on either side, and the end
constexpr int MO_STRIDE = (MO_N + 3) & ~3;
static_assert(MO_N <= 32, "one lane per mini-tile offset");

typedef Unit SpanUnit;                              // one (ProfileBuffer, contig) pair; nSpans = ceil((size + bw) / SPAN_LEN)

struct SpanRun {                                    // a run of above-threshold positions inside one piece
  uint32_t left, len, peak, wsum;
  double score;                                     // exact-order score at `peak`
  uint32_t span;                                    // work item that produced it
  uint32_t rank;                                    // ordinal among the runs of that span (position order)
};

struct SpanArgs {
  Tags T;
  const SpanUnit* units;
  const uint32_t* nTilesIdx;                        // [C]
  const uint8_t* control;                           // [nSamples] 1 = negative control (not a density source)
  uint32_t nSamples;
  int hasControl;
  const double* coeff;                              // [nSamples] scaling coefficient of every non-control sample (-z), NULL otherwise
  uint32_t nUnits, totalSpans;
  const uint2* hotList;                             // (span, unit) of the spans that hold at least one hot block (hot_span_kernel), any order
  const uint32_t* nHot;
  const uint32_t* hotMask;                          // [totalSpans * 4] hot-block bits of those spans
  uint32_t bw;
  const double* taps;                               // 2bw+1, host-computed exactly as misc/kernel.cpp:16-35
  double thr;                                       // region threshold
  double cN;                                        // taps[bw] / bw^2: moment-form scale
  double dS;                                        // error bound of the moment form per unit of window weight
  // outputs
  uint32_t* ticket;
  SpanRun* runs;
  uint32_t runCap;
  uint32_t* runCount;                               // run slots handed out (chunks of RUN_CHUNK per warp; unused slots are marked invalid)
  uint32_t* runReal;                                // runs actually written
  uint32_t* spanRunCount;                           // [totalSpans]
  uint32_t* flags;
  unsigned long long* stat;                         // [0] hot blocks, [1] pieces, [2] exact evaluations, [3] spans with tags
};

struct SpanSmem { uint32_t taps, ctl, coef, warp0, warpStride, oD, oDR, oWf, oWr, oSt, oMo, oSc, oMw, oHm, oBar, total; };

__host__ __device__ inline SpanSmem span_smem_layout (uint32_t bw, uint32_t nSamples, bool nondir, bool coeff) {
  SpanSmem L;
  uint32_t o = 0;
  L.taps = o;      o += ((2 * bw + 2) & ~1u) * 8;
  L.ctl = o;       o += (nSamples + 3) & ~3u;
  o = (o + 15) & ~15u;
  L.coef = o;      o += coeff ? nSamples * 8 : 0;
  o = (o + 15) & ~15u;
  L.warp0 = o;
  const uint32_t W = PIECE_LEN + 2 * bw + 4;
  uint32_t w = 0;
  L.oD = w;    w += W * 4;
  L.oDR = w;   w += nondir ? W * 4 : 0;
  w = (w + 7) & ~7u;
  L.oWf = w;   w += coeff ? W * 8 : 0;              // -z: per-position weights are doubles
  L.oWr = w;   w += (coeff && nondir) ? W * 8 : 0;
  w = (w + 15) & ~15u;
  L.oMo = w;   w += 2 * MO_STRIDE * 8;                   // mini-tile offsets of the span + its halo on either side, per strand
  L.oSt = w;   w += 128 * 8;                        // products of 128 taps (exact_dense_warp)
  L.oSc = w;   w += coeff ? PIECE_LEN * 8 : PIECE_LEN * 4;   // oSt + oSc together (>= 3 KB, contiguous) double as the entry staging ring of B1
  L.oMw = w;   w += (PIECE_BLOCKS + 2) * 4;
  L.oHm = w;   w += (SPAN_WORDS + 2) * 4;
  w = (w + 15) & ~15u;
  L.oBar = w;  w += 16;                             // two mbarriers: one per slot of B1's bulk-copy ring
  L.warpStride = w;
  L.total = o + w * SP_WARPS;
  return L;
}

__device__ __forceinline__ unsigned long long warp_sum_u64 (unsigned long long v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ uint32_t warp_sum_u32 (uint32_t v) { return __reduce_add_sync(0xffffffffu, v); }   // one REDUX instead of five SHFL + five IADD
template <class T> __device__ __forceinline__ T warp_incl_scan (T v, uint32_t lane) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const T n = __shfl_up_sync(0xffffffffu, v, o); if (lane >= (uint32_t)o) v += n; }
  return v;
}
// the same scan for 32-bit values with the shuffle's own in-range predicate: one SHFL + one predicated IADD per step
// (the generic form above compiles to SHFL + SEL + IADD and recomputes the five lane tests every time it is inlined in a loop)
__device__ __forceinline__ uint32_t warp_incl_scan_u32 (uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1)
    asm volatile("{ .reg .u32 t; .reg .pred p; shfl.sync.up.b32 t|p, %0, %1, 0, 0xffffffff; @p add.u32 %0, %0, t; }" : "+r"(v) : "r"(o));
  return v;
}
template <class T> __device__ __forceinline__ T warp_sum (T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// exact-order density of the position whose window starts at cell u0 (cells u0 .. u0+2bw), dense form:
// a cell without tags contributes +0.0, which leaves the running sum unchanged, so this equals the
// reference's sum over the tags in ascending position (misc/peakcall.cpp:203-209)
__device__ __forceinline__ double exact_dense (const double* __restrict__ Ks, const uint32_t* __restrict__ cells, uint32_t u0, uint32_t bw) {
  double s = 0.0;
  const uint32_t n = 2 * bw;
  for (uint32_t k = 0; k <= n; ++k) {
    const uint32_t w = cells[u0 + k];
    if (w) s = __dadd_rn(s, __dmul_rn(Ks[n - k], (double)w));
  }
  return s;
}

// -z: the weights are doubles (misc/peakcall.cpp:191-200)
__device__ __noinline__ double exact_dense_f64 (const double* __restrict__ Ks, const double* __restrict__ cells, uint32_t u0, uint32_t bw) {
  double s = 0.0;
  const uint32_t n = 2 * bw;
  for (uint32_t k = 0; k <= n; ++k) {
    const double w = cells[u0 + k];
    if (w != 0.0) s = __dadd_rn(s, __dmul_rn(Ks[n - k], w));
  }
  return s;
}

// the same sum computed by the whole warp for ONE position (u0 uniform): the products are independent, so the lanes round
// 64 of them at a time into `stage`; the additions stay a single chain in ascending tag order (a product of an empty cell
// is +0.0 and changes nothing), run by every lane on the broadcast values
__device__ __forceinline__ double exact_dense_warp (const double* __restrict__ Ks, const uint32_t* __restrict__ cells, double* stage /* [128], 16-byte aligned */,
                                                    uint32_t u0, uint32_t bw, uint32_t lane) {
  double s = 0.0;
  const uint32_t n = 2 * bw;
#if UPB_EXP_EXACT == 0
  for (uint32_t k0 = 0; k0 <= n; k0 += 32) {
    const uint32_t k = k0 + lane;
    const uint32_t w = k <= n ? cells[u0 + k] : 0u;
    const double p = w ? __dmul_rn(Ks[n - k], (double)w) : 0.0;
    uint32_t nz = __ballot_sync(0xffffffffu, w != 0);
    while (nz) {
      const int b = __ffs(nz) - 1;
      s = __dadd_rn(s, __shfl_sync(0xffffffffu, p, b));
      nz &= nz - 1;
    }
  }
  return s;
#endif
  for (uint32_t k0 = 0; k0 <= n; k0 += 128) {
#pragma unroll
    for (int h = 0; h < 4; ++h) {
      const uint32_t k = k0 + 32 * h + lane;
      const uint32_t w = k <= n ? cells[u0 + k] : 0u;
      stage[32 * h + lane] = w ? __dmul_rn(Ks[n - k], (double)w) : 0.0;
    }
    __syncwarp();
    // the chain, eight terms per step from four 16-byte loads; the stage is zero beyond tap n and s + (+0.0) == s
    const uint32_t m8 = (min(128u, n + 1 - k0) + 7) >> 3;
    const double2* __restrict__ st2 = (const double2*)stage;
    for (uint32_t q = 0; q < m8; ++q) {
      const double2 t0 = st2[4 * q], t1 = st2[4 * q + 1], t2 = st2[4 * q + 2], t3 = st2[4 * q + 3];
      s = __dadd_rn(s, t0.x); s = __dadd_rn(s, t0.y); s = __dadd_rn(s, t1.x); s = __dadd_rn(s, t1.y);
      s = __dadd_rn(s, t2.x); s = __dadd_rn(s, t2.y); s = __dadd_rn(s, t3.x); s = __dadd_rn(s, t3.y);
    }
    __syncwarp();
  }
  return s;
}

// Entries of (strand d, contig c) whose mini-tile overlaps positions [lo, hi] (1-based, inclusive; lo may be <= 0), visited by
// the 32 lanes with four loads in flight each: f(position, entry); a slot beyond the end of the chunk arrives with count 0.
// Entries carry their offset inside a 4096-position span, so the mini-tiles are walked span by span (a piece window touches at
// most two spans).  The mini-tile offsets come from a table the warp loaded once per span: tab[j] = mtOff[gb + min(mBase + j, nMt)]
// for j = 0 .. MO_N - 1 (the span's mini-tiles, HALO_MT of halo on either side, and the end) — no dependent global look-up per piece.
template <class F>
__device__ __forceinline__ void piece_entries_tab (const uint2* __restrict__ ment, const unsigned long long* __restrict__ tab, int mBase, uint32_t nMt,
                                                   int64_t lo, int64_t hi, uint32_t lane, F f) {
  if (nMt == 0 || hi < 1) return;
  if (lo < 1) lo = 1;
  const uint32_t ma = (uint32_t)((lo - 1) >> MT_SHIFT);
  if (ma >= nMt) return;
  const uint32_t mb = min((uint32_t)((hi - 1) >> MT_SHIFT), nMt - 1);
  for (uint32_t sp = ma / SPAN_MT; sp <= mb / SPAN_MT; ++sp) {
    const uint32_t m0 = max(ma, sp * SPAN_MT), m1 = min(mb, sp * SPAN_MT + SPAN_MT - 1);
    const unsigned long long e0 = tab[(int)m0 - mBase], e1 = tab[(int)m1 + 1 - mBase];
    const uint32_t base = sp * SPAN_LEN + 1;
    // the chunk holds at most one entry per (sample, position) of one span and its halo: its length fits 32 bits
    const uint32_t n = (uint32_t)(e1 - e0);
    const uint2* __restrict__ src = ment + e0 + lane;
    for (uint32_t k = lane; k - lane < n; k += 128, src += 128) {
      uint2 en[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) en[q] = (k + 32 * q < n) ? __ldg(src + 32 * q) : make_uint2(0u, 0u);
#pragma unroll
      for (int q = 0; q < 4; ++q) f(base + (en[q].y & (SPAN_LEN - 1)), en[q]);   // a slot beyond the end has count 0: f skips it
    }
  }
}

// B1 with the entries staged by the bulk-copy engine (cp.async.bulk, SASS UBLKCP — the 1-D form of TMA: the chunk is contiguous,
// no tensor map is needed).  One elected lane posts the copy of a sub-chunk of BULK_ENTRIES entries into one of two ring slots and
// arms that slot's mbarrier with the byte count; the warp waits on the barrier's phase, pulls the slot into registers with conflict-
// free LDS.64, hands the slot back (the next sub-chunk is posted before the current one is pooled) and pools from registers.  Against
// the per-lane LDG walk above (r01: 17 instructions per entry, 20 % of the kernel's stall samples on the loads' scoreboard) the lanes
// issue no global load, no address arithmetic and no bounds predicate per load.  Copies are 16-byte granular: the chunk is widened
// to even entry indices at both ends and the odd entry is masked by index.
#ifndef UPB_PREFETCH_L2
#define UPB_PREFETCH_L2 1                         // span_kernel: the entries of the next piece are prefetched into L2 while the current one is worked on
#endif
constexpr uint32_t BULK_ENTRIES = 192;                                // per ring slot: 1536 bytes; two slots = SpanSmem oSt + oSc (3 KB, contiguous)
static_assert(128 * 8 + PIECE_LEN * 4 >= 2 * BULK_ENTRIES * 8, "SpanSmem: oSt + oSc must hold the bulk ring");

__device__ __forceinline__ uint32_t smem_u32 (const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init (uint32_t bar, uint32_t count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }
__device__ __forceinline__ void mbar_expect_tx (uint32_t bar, uint32_t bytes) { asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }
__device__ __forceinline__ void bulk_g2s (uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait (uint32_t bar, uint32_t parity) {
  asm volatile("{\n .reg .pred P1;\n LAB_WAIT:\n mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n @P1 bra DONE;\n bra LAB_WAIT;\n DONE:\n }" :: "r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async () { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// seq: sub-chunks this warp has posted so far (slot = seq & 1, phase parity = (seq >> 1) & 1); persists across pieces
template <class P, class F>
__device__ __forceinline__ void piece_entries_bulk (const uint2* __restrict__ ment, const unsigned long long* __restrict__ tab, int mBase, uint32_t nMt,
                                                    int64_t lo, int64_t hi, uint32_t lane, const uint2* ring, uint32_t ringAddr, uint32_t barAddr, uint32_t& seq,
                                                    bool& preDone, P pre, F f) {
  if (nMt == 0 || hi < 1) return;
  if (lo < 1) lo = 1;
  const uint32_t ma = (uint32_t)((lo - 1) >> MT_SHIFT);
  if (ma >= nMt) return;
  const uint32_t mb = min((uint32_t)((hi - 1) >> MT_SHIFT), nMt - 1);
  for (uint32_t sp = ma / SPAN_MT; sp <= mb / SPAN_MT; ++sp) {
    const uint32_t m0 = max(ma, sp * SPAN_MT), m1 = min(mb, sp * SPAN_MT + SPAN_MT - 1);
    const unsigned long long e0 = tab[(int)m0 - mBase], e1 = tab[(int)m1 + 1 - mBase];
    if (e1 <= e0) continue;
    const uint32_t base = sp * SPAN_LEN + 1;
    const unsigned long long eA = e0 & ~1ull;                          // 16-byte granules: even entry indices
    const uint32_t nAll = (uint32_t)(((e1 + 1) & ~1ull) - eA);          // entries copied (the chunk of a piece fits 32 bits)
    const uint32_t skip = (uint32_t)(e0 - eA), nEnd = (uint32_t)(e1 - eA);   // entries [skip, nEnd) of the copy are the chunk
    const uint32_t nSub = (nAll + BULK_ENTRIES - 1) / BULK_ENTRIES;
    auto post = [&] (uint32_t j) {                                     // sub-chunk j -> slot (seq + j) & 1
      if (lane == 0) {
        const uint32_t g = seq + j, slot = g & 1u;
        const uint32_t cnt = min(BULK_ENTRIES, nAll - j * BULK_ENTRIES);
        fence_proxy_async();                                           // the slot was read (or written as stage / scf) through the generic proxy
        mbar_expect_tx(barAddr + 8 * slot, cnt * 8);
        bulk_g2s(ringAddr + slot * (BULK_ENTRIES * 8), ment + eA + (unsigned long long)j * BULK_ENTRIES, cnt * 8, barAddr + 8 * slot);
      }
    };
    post(0);
    if (nSub > 1) post(1);
    if (!preDone) { pre(); preDone = true; }
    for (uint32_t j = 0; j < nSub; ++j) {
      const uint32_t g = seq + j, slot = g & 1u;
      mbar_wait(barAddr + 8 * slot, (g >> 1) & 1u);
      const uint2* src = ring + slot * BULK_ENTRIES + lane;
      const uint32_t i0 = j * BULK_ENTRIES + lane;
      constexpr int HALF = (int)(BULK_ENTRIES / 64);                   // two batches per slot: half the registers held across the pooling
      uint2 en[HALF];
#pragma unroll
      for (int q = 0; q < HALF; ++q) { const uint32_t i = i0 + 32 * q; en[q] = (i >= skip && i < nEnd) ? src[32 * q] : make_uint2(0u, 0u); }
#pragma unroll
      for (int q = 0; q < HALF; ++q) f(base + (en[q].y & (SPAN_LEN - 1)), en[q]);   // a masked slot has count 0: f skips it
#pragma unroll
      for (int q = 0; q < HALF; ++q) { const uint32_t i = i0 + 32 * (q + HALF); en[q] = (i >= skip && i < nEnd) ? src[32 * (q + HALF)] : make_uint2(0u, 0u); }
      __syncwarp();                                                    // every lane holds the rest of its entries: the slot can be refilled
      if (j + 2 < nSub) post(j + 2);
#pragma unroll
      for (int q = 0; q < HALF; ++q) f(base + (en[q].y & (SPAN_LEN - 1)), en[q]);
    }
    seq += nSub;
  }
}

// Zone-map prefilter: a block's window (32 + 2bw positions) lies inside `nAdj` consecutive mini-tiles, so a
// span none of whose runs of nAdj mini-tiles (its own plus HALO_MT of halo on either side) weighs enough for
// maxTap * weight >= threshold has no hot block and is skipped without touching its entries — the device-side analogue
// of the reference leaping over empty stretches (misc/peakcall.cpp:172-173).  The weights count every sample (controls
// too), which only makes the bound more conservative.
struct WarmRec {                    // everything hot_span_kernel needs about one warm span: no dependent look-ups there
  uint32_t span, unit;
  uint32_t blk0, blk1;              // index into Tags::blkW of the span's first block, strand 0 / 1 of the unit (blk1 unused unless -D)
  int32_t left;                     // blocks of the contig before the span
  int32_t right0, right1;           // blocks of the contig from the span's first block to the end of the mini-tile table (may be <= 0), per strand
  uint32_t nBlk;                    // blocks of the span that hold evaluated positions (Q20: through size + bw)
};

struct WarmArgs {
  Tags T;
  const SpanUnit* units;
  uint32_t nUnits, totalSpans;
  uint32_t spanBegin, spanEnd;      // spans of the group of units this launch covers
  int nondir;
  uint32_t nAdj;
  double flatBound, thr;
  WarmRec* warmList;                // the spans that have to be looked at, any order
  uint32_t bw;
  uint32_t* nWarm;
  uint32_t* spanRunCount;           // zeroed here for the spans that are skipped
  const uint32_t* loadFlags;        // validation flags of a load that is still being waited for group by group (NULL: the load was checked before the call):
                                    //   once it has reported invalid input nothing is evaluated any more — the entries may be anything
};

constexpr int WARM_ADJ_MAX = ((31 + 2 * MAX_BW_SPAN) >> MT_SHIFT) + 2;   // upb_api.cu: nAdj = ((31 + 2 bw) >> MT_SHIFT) + 2

__global__ void warm_span_kernel (WarmArgs a) {
  const uint32_t w = a.spanBegin + blockIdx.x * blockDim.x + threadIdx.x;
  bool warm = false;
  uint32_t unit = 0;
  const bool badInput = a.loadFlags && (*a.loadFlags & (2u | 4u | 8u | 16u)) != 0;   // FLAG_UNSORTED | FLAG_BAD_POS | FLAG_ZERO_COUNT | FLAG_BAD_PACK
  if (w < a.spanEnd && badInput) a.spanRunCount[w] = 0;
  if (w < a.spanEnd && !badInput) {
    uint32_t lo = 0, hi = a.nUnits;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.units[mid].spanStart <= w) lo = mid; else hi = mid; }
    const SpanUnit U = a.units[lo];
    unit = lo;
    const uint32_t k = w - U.spanStart;
    const int m0 = (int)(k * SPAN_MT) - HALO_MT, m1 = (int)(k * SPAN_MT + SPAN_MT) - 1 + HALO_MT;   // inclusive mini-tile range with halo
    const uint32_t d0 = a.nondir ? 0u : U.buffer, d1 = a.nondir ? 1u : U.buffer;
    unsigned long long win[WARM_ADJ_MAX];                            // sliding window of the last nAdj mini-tile weights
#pragma unroll
    for (int q = 0; q < WARM_ADJ_MAX; ++q) win[q] = 0;
    unsigned long long run = 0;
    for (int m = m0; m <= m1; ++m) {
      unsigned long long wgt = 0;
      for (uint32_t d = d0; d <= d1; ++d) {
        const uint32_t gb = a.T.mtBase[d * a.T.C + U.contig], nMt = a.T.mtBase[d * a.T.C + U.contig + 1] - gb;
        if (m >= 0 && (uint32_t)m < nMt) wgt += a.T.mtW[gb + m];
      }
      const uint32_t slot = (uint32_t)(m - m0) % a.nAdj;
      run += wgt - win[slot];
      win[slot] = wgt;
      if ((double)run * a.flatBound >= a.thr) { warm = true; break; }
    }
    if (!warm) a.spanRunCount[w] = 0;
  }
  // compact the warm spans: one atomic per block (atomics on one address serialise: ~1 ns each)
  __shared__ uint32_t s_warp[8], s_base;
  const uint32_t msk = __ballot_sync(0xffffffffu, warm);
  if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = __popc(msk);
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int q = 0; q < 8; ++q) { const uint32_t c = s_warp[q]; s_warp[q] = run; run += c; }
    s_base = run ? atomicAdd(a.nWarm, run) : 0u;
  }
  __syncthreads();
  if (msk) {
    const uint32_t base = s_base + s_warp[threadIdx.x >> 5];
    if (warm) {
      const SpanUnit U = a.units[unit];
      const uint32_t k = w - U.spanStart;
      const uint32_t d0 = a.nondir ? 0u : U.buffer, d1 = a.nondir ? 1u : U.buffer;
      const uint32_t gb0 = a.T.mtBase[d0 * a.T.C + U.contig], gb1 = a.T.mtBase[d1 * a.T.C + U.contig];
      const uint32_t nB0 = (a.T.mtBase[d0 * a.T.C + U.contig + 1] - gb0) * BLK_PER_MT, nB1 = (a.T.mtBase[d1 * a.T.C + U.contig + 1] - gb1) * BLK_PER_MT;
      const uint64_t lastPos = (uint64_t)U.size + a.bw;
      const uint32_t np = (uint32_t)min((uint64_t)SPAN_LEN, lastPos - (uint64_t)k * SPAN_LEN);
      WarmRec r;
      r.span = w; r.unit = unit;
      r.blk0 = gb0 * BLK_PER_MT + k * SPAN_BLOCKS; r.blk1 = gb1 * BLK_PER_MT + k * SPAN_BLOCKS;
      r.left = (int32_t)(k * SPAN_BLOCKS);
      r.right0 = (int32_t)nB0 - (int32_t)(k * SPAN_BLOCKS); r.right1 = (int32_t)nB1 - (int32_t)(k * SPAN_BLOCKS);
      r.nBlk = (np + 31) >> 5;
      a.warmList[base + __popc(msk & ((1u << (threadIdx.x & 31)) - 1))] = r;
    }
  }
}

// Hot blocks of the warm spans from the block-weight zone map, one warp per span, without touching the entries.  Block B of the
// contig = positions 32B+1 .. 32B+32; a tag of block B+j is at least 32(|j|-1)+1 positions away from every position of block B
// (0 for j = 0), so  bound(B) = sum_j blkW[B+j] * tap(that distance)  is an upper envelope of every score in B: a block whose
// bound stays below the threshold cannot hold a position of a region (misc/peakcall.cpp:57-60).  Control samples, float
// rounding (upwards, margins folded into the coefficients on the host) and -z coefficients only loosen the bound.
struct HotArgs {
  Tags T;
  const SpanUnit* units;
  const WarmRec* warmList;
  const uint32_t* nWarm;
  uint32_t bw;
  int nondir;
  const float* hotCoefF;            // [hotJ+1] largest tap a tag |j| blocks away can have for a position of the block, rounded up
  uint32_t hotJ;                    // blocks on either side that can reach a block: (bw-2)/32 + 1
  float thrLow;                     // threshold rounded down to float
  uint32_t* hotMask;                // [totalSpans * 4]
  uint2* hotList;
  uint32_t* nHot;
  uint32_t* spanRunCount;           // zeroed here for the spans without a hot block
  unsigned long long* stat;         // [0] hot blocks
};

constexpr int HOT_THREADS = 256;
constexpr int HOT_WB = (SPAN_BLOCKS + 2 * HOT_JMAX + 3) & ~1;

__global__ void __launch_bounds__(HOT_THREADS, 6) hot_span_kernel (const __grid_constant__ HotArgs a) {
  __shared__ float hotf[HOT_JMAX + 2];
  __shared__ float wbAll[HOT_THREADS / 32][HOT_WB];
  __shared__ uint2 outAll[HOT_THREADS / 32][32];                      // hot spans found by this warp, flushed with one atomic per 32
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* wb = wbAll[warp];
  uint2* outBuf = outAll[warp];
  uint32_t nOut = 0;
  for (uint32_t i = threadIdx.x; i <= a.hotJ; i += HOT_THREADS) hotf[i] = a.hotCoefF[i];
  __syncthreads();
  const int J = (int)a.hotJ;
  const uint32_t nWarm = *a.nWarm;
  const uint32_t nWarps = gridDim.x * (HOT_THREADS / 32);
  unsigned long long stHot = 0;
  uint32_t it = blockIdx.x * (HOT_THREADS / 32) + warp;
  uint4 ra = make_uint4(0, 0, 0, 0), rb = ra;
  if (it < nWarm) { const uint4* rp = (const uint4*)(a.warmList + it); ra = __ldg(rp); rb = __ldg(rp + 1); }
  for (; it < nWarm; it += nWarps) {
    const uint32_t span = ra.x, unitOf = ra.y, blk0 = ra.z, blk1 = ra.w, nBlk = rb.w;
    const int left = (int)rb.x, right0 = (int)rb.y, right1 = (int)rb.z;
    if (it + nWarps < nWarm) { const uint4* rp = (const uint4*)(a.warmList + it + nWarps); ra = __ldg(rp); rb = __ldg(rp + 1); }   // next record: in flight during this span
    // one round trip: 128 + 2J floats per strand, straight into shared memory
    const bool interior = left >= J && right0 >= SPAN_BLOCKS + J && (!a.nondir || right1 >= SPAN_BLOCKS + J);
    if (interior && J <= 16) {                                         // the common case: no bounds to check, 160 consecutive floats
      const float* __restrict__ p0 = a.T.blkW + (size_t)blk0 - J;
      const float* __restrict__ p1 = a.T.blkW + (size_t)blk1 - J;
      float w0 = __ldg(p0 + lane), w1 = __ldg(p0 + 32 + lane), w2 = __ldg(p0 + 64 + lane), w3 = __ldg(p0 + 96 + lane);
      float w4 = lane < (uint32_t)(2 * J) ? __ldg(p0 + 128 + lane) : 0.0f;
      if (a.nondir) {
        w0 += __ldg(p1 + lane); w1 += __ldg(p1 + 32 + lane); w2 += __ldg(p1 + 64 + lane); w3 += __ldg(p1 + 96 + lane);   // may round down by 2^-24: inside the margin
        if (lane < (uint32_t)(2 * J)) w4 += __ldg(p1 + 128 + lane);
      }
      wb[lane] = w0; wb[32 + lane] = w1; wb[64 + lane] = w2; wb[96 + lane] = w3;
      if (128 + lane < (uint32_t)HOT_WB) wb[128 + lane] = w4;
    } else {
#pragma unroll
      for (int q = 0; q < (SPAN_BLOCKS + 2 * HOT_JMAX + 31) / 32; ++q) {
        const int i = 32 * q + (int)lane, r = i - J;                 // block r of the span (r < 0: halo before it)
        float w = 0.0f;
        if (i < SPAN_BLOCKS + 2 * J && r >= -left) {
          if (r < right0) w = __ldg(a.T.blkW + (size_t)blk0 + r);
          if (a.nondir && r < right1) w += __ldg(a.T.blkW + (size_t)blk1 + r);
        }
        if (i < HOT_WB) wb[i] = w;
      }
    }
    __syncwarp();
    uint32_t any = 0, mine = 0;
    // the envelope sum with the number of neighbour blocks known at compile time for the usual bandwidths (J <= 4: bandwidth <= 129):
    // the j loop unrolls and its coefficients stay in registers (the run-time loop made this kernel issue-bound: 377 instructions per span)
    auto words = [&] (auto jt) {
      constexpr int JT = decltype(jt)::value;
      const int JJ = JT ? JT : J;
      float hf[JT ? JT + 1 : 1];
      if (JT) {
#pragma unroll
        for (int j = 0; j <= JT; ++j) hf[j] = hotf[j];
      }
#pragma unroll
      for (uint32_t m = 0; m < (uint32_t)SPAN_WORDS; ++m) {
        const uint32_t kblk = 32 * m + lane;
        float bound = wb[kblk + JJ] * (JT ? hf[0] : hotf[0]);
        if (JT) {
#pragma unroll
          for (int j = 1; j <= JT; ++j) bound = fmaf(wb[kblk + JT - j] + wb[kblk + JT + j], hf[j], bound);
        } else {
          for (int j = 1; j <= JJ; ++j) bound = fmaf(wb[kblk + JJ - j] + wb[kblk + JJ + j], hotf[j], bound);
        }
        const uint32_t msk = __ballot_sync(0xffffffffu, kblk < nBlk && !(bound < a.thrLow));   // a NaN would count as hot
        any |= msk;
        if (lane == m) mine = msk;
        stHot += __popc(msk);
      }
    };
    switch (J) {
      case 1: words(std::integral_constant<int, 1>()); break;
      case 2: words(std::integral_constant<int, 2>()); break;
      case 3: words(std::integral_constant<int, 3>()); break;
      case 4: words(std::integral_constant<int, 4>()); break;
      default: words(std::integral_constant<int, 0>()); break;
    }
    __syncwarp();
    if (any) {
      if (lane < (uint32_t)SPAN_WORDS) a.hotMask[(size_t)span * SPAN_WORDS + lane] = mine;
      if (lane == 0) outBuf[nOut] = make_uint2(span, unitOf);
      if (++nOut == 32) {
        __syncwarp();
        uint32_t base = 0;
        if (lane == 0) base = atomicAdd(a.nHot, 32u);
        base = __shfl_sync(0xffffffffu, base, 0);
        a.hotList[base + lane] = outBuf[lane];
        nOut = 0;
        __syncwarp();
      }
    } else if (lane == 0) a.spanRunCount[span] = 0;
  }
  if (nOut) {
    __syncwarp();
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.nHot, nOut);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (lane < nOut) a.hotList[base + lane] = outBuf[lane];
  }
  if (lane == 0 && stHot) atomicAdd(a.stat + 0, stHot);
}

// MINB: CTAs per SM the register budget is sized for.  Four (64 registers) when the shared memory of four CTAs fits; where it
// does not (bandwidth > ~150, or -D with its second cell array) the kernel is built for three (no spills, 78 registers): measured
// on C3 at bandwidth 200 0.69 -> 0.60 ms, C2 0.57 -> 0.47 ms, while at bandwidth <= 150 four CTAs win (profiles/r01_v26_ab.txt).
template <bool NONDIR, bool COEFF, int MINB = SP_MIN_BLOCKS, bool BULK = false>
__global__ void __launch_bounds__(SP_THREADS, COEFF ? 2 : MINB) span_kernel (const __grid_constant__ SpanArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  const uint32_t bw = a.bw;
  const SpanSmem L = span_smem_layout(bw, a.nSamples, NONDIR, COEFF);
  double* Ks = (double*)(smem + L.taps);
  double* coef = (double*)(smem + L.coef);
  uint8_t* isCtl = (uint8_t*)(smem + L.ctl);                          // control flag of every sample (CTA-shared, read-only)
  const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  unsigned char* wbase = smem + L.warp0 + (size_t)warp * L.warpStride;
  uint32_t* D = (uint32_t*)(wbase + L.oD);
  uint32_t* DR = (uint32_t*)(wbase + L.oDR);
  float* scf = (float*)(wbase + L.oSc);
  double* scd = (double*)(wbase + L.oSc);                             // -z keeps exact doubles instead
  double* stage = (double*)(wbase + L.oSt);
  unsigned long long* mo = (unsigned long long*)(wbase + L.oMo);
  double* Wf = (double*)(wbase + L.oWf);
  double* Wr = (double*)(wbase + L.oWr);
  uint32_t* mw = (uint32_t*)(wbase + L.oMw);
  uint32_t* hm = (uint32_t*)(wbase + L.oHm);
  const uint32_t barAddr = smem_u32(wbase + L.oBar), ringAddr = smem_u32(stage);
  uint32_t bulkSeq = 0;
  if (BULK) {
    if (lane == 0) { mbar_init(barAddr, 1); mbar_init(barAddr + 8, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    __syncwarp();
  }

  for (uint32_t i = tid; i < 2 * bw + 1; i += blockDim.x) Ks[i] = a.taps[i];      // (blockDim.x: a launch may bring fewer than SP_WARPS warps, launch_span)
  for (uint32_t i = tid; i < a.nSamples; i += blockDim.x) isCtl[i] = a.control[i];
  if (COEFF) for (uint32_t i = tid; i < a.nSamples; i += blockDim.x) coef[i] = a.coeff[i];
  __syncthreads();                                                    // the only block-level barrier
  const unsigned long long bw2 = (unsigned long long)bw * bw;
  unsigned long long stPieces = 0, stExact = 0, stSpans = 0;
  const uint32_t nHot = *a.nHot;
  uint32_t rBase = 0, rUsed = RUN_CHUNK, rTotal = 0;                   // current chunk of run slots: [rBase, rBase + RUN_CHUNK), rUsed taken
  bool rRoom = true;
  constexpr uint32_t BATCH = 4;                                        // hot spans per ticket: one atomic and two dependent loads per four spans ...
  const uint32_t batch = nHot > gridDim.x * (blockDim.x >> 5) * 16u ? BATCH : (nHot > gridDim.x * (blockDim.x >> 5) * 4u ? 2u : 1u);   // ... fewer where a rank's shard leaves a warp only a handful (N = 8)

  for (;;) {
    uint32_t base = 0;
    if (lane == 0) base = atomicAdd(a.ticket, batch);
    base = __shfl_sync(0xffffffffu, base, 0);
    if (base >= nHot) break;
    const uint32_t nIn = min(batch, nHot - base);
    uint2 item = make_uint2(0u, 0u);
    if (lane < nIn) item = __ldg(a.hotList + base + lane);
    uint32_t maskWord = 0;                                            // lane 4t+m holds word m of the t-th span's hot-block mask
    { const uint32_t sp = __shfl_sync(0xffffffffu, item.x, lane >> 2);
      if ((lane >> 2) < nIn) maskWord = __ldg(a.hotMask + (size_t)sp * SPAN_WORDS + (lane & 3)); }
  for (uint32_t tb = 0; tb < nIn; ++tb) {
    const uint32_t work = __shfl_sync(0xffffffffu, item.x, tb);        // span index
    const SpanUnit U = a.units[__shfl_sync(0xffffffffu, item.y, tb)];
    const uint32_t k = work - U.spanStart;                            // span index inside the contig
    const uint32_t ts = k * SPAN_LEN + 1;                             // first position of the span
    const uint64_t lastPos = (uint64_t)U.size + bw;                   // Q20: evaluated through lastTag+bw
    const uint32_t np = (uint32_t)min((uint64_t)SPAN_LEN, lastPos - (ts - 1));
    const uint32_t nBlk = (np + 31) >> 5;
    __syncwarp();
    if ((lane >> 2) == tb) hm[lane & 3] = maskWord;
    if (lane == 0) hm[SPAN_WORDS] = 0;
    const uint32_t nzWords = (__ballot_sync(0xffffffffu, maskWord != 0) >> (4 * tb)) & ((1u << SPAN_WORDS) - 1);   // which of the span's mask words hold a hot block
    const int mBase = (int)(k * SPAN_MT) - HALO_MT;
    uint32_t nMtS[2] = {0u, 0u};
    unsigned long long* const moSpan = mo;
#pragma unroll
    for (int str = 0; str < (NONDIR ? 2 : 1); ++str) {
      const uint32_t d = NONDIR ? (uint32_t)str : U.buffer;
      const uint32_t gb = __ldg(a.T.mtBase + d * a.T.C + U.contig);
      nMtS[str] = __ldg(a.T.mtBase + d * a.T.C + U.contig + 1) - gb;
      if (lane < (uint32_t)MO_N)
        moSpan[MO_STRIDE * str + lane] = __ldg(a.T.mtOff + gb + (uint32_t)min(max(mBase + (int)lane, 0), (int)nMtS[str]));
    }
    __syncwarp();
    ++stSpans;

    // ---- B: pieces, in position order -------------------------------------------------------------------------------
    uint32_t spanRuns = 0;
    // the pieces of the span come from a small scanner over the hot-block mask that runs ONE PIECE AHEAD of the work: while piece k is
    // pooled, classified and cut into runs (~10 us of a warp's time), the entries of piece k+1 are already on their way into L2
    // (cp.async.bulk.prefetch.L2, one instruction of one lane) — its loads then pay an L2 round trip instead of a DRAM one
    uint32_t kb = 0, gEnd = 0, gNext = 0;                              // scanner state: current group of hot blocks is [.., gEnd), next piece starts at gNext
    auto nextPiece = [&] (uint32_t& pbOut, uint32_t& nbOut) -> bool {
      if (gNext >= gEnd) {                                             // next maximal group of hot blocks (uniform scalar code over the mask words)
        for (;;) {
          if (kb >= nBlk) return false;
          const uint32_t m = hm[kb >> 5] >> (kb & 31);
          if (!m) {                                                    // nothing left in this word: on to the next word that has a hot block
            const uint32_t w1 = (kb >> 5) + 1, rest = w1 < (uint32_t)SPAN_WORDS ? nzWords >> w1 : 0u;
            if (!rest) { kb = nBlk; return false; }
            kb = (w1 + (uint32_t)__ffs(rest) - 1) << 5;
            continue;
          }
          kb += __ffs(m) - 1;
          break;
        }
        gNext = kb;
        for (;;) {
          const uint32_t x = ~(hm[kb >> 5] >> (kb & 31));
          kb += x ? (uint32_t)(__ffs(x) - 1) : 32u;
          if ((kb & 31) != 0 || kb >= nBlk || !(hm[kb >> 5] & 1u)) break;
        }
        if (kb > nBlk) kb = nBlk;
        gEnd = kb;
      }
      pbOut = gNext;
      nbOut = min((uint32_t)PIECE_BLOCKS, gEnd - gNext);
      gNext += nbOut;
      return true;
    };
#if UPB_PREFETCH_L2
    auto prefetchPiece = [&] (uint32_t pbN, uint32_t nbN) {
      if (lane != 0) return;
      const int64_t lo = (int64_t)ts + 32 * pbN - bw, hi = (int64_t)ts + 32 * pbN + 32 * nbN - 1 + bw;
#pragma unroll
      for (int str = 0; str < (NONDIR ? 2 : 1); ++str) {
        const uint32_t nMt = nMtS[str];
        if (nMt == 0 || hi < 1) continue;
        const uint32_t ma = lo < 1 ? 0u : (uint32_t)((lo - 1) >> MT_SHIFT);
        if (ma >= nMt) continue;
        const uint32_t mb = min((uint32_t)((hi - 1) >> MT_SHIFT), nMt - 1);
        const unsigned long long* tab = moSpan + MO_STRIDE * str;
        const unsigned long long e0 = tab[(int)ma - mBase] & ~1ull, e1 = (tab[(int)mb + 1 - mBase] + 1) & ~1ull;   // the chunk is contiguous across a span boundary too
        if (e1 > e0) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(a.T.ment + e0), "r"((uint32_t)min(e1 - e0, 8192ull) * 8u) : "memory");
      }
    };
#endif
    uint32_t pb = 0, nb = 0;
    bool havePiece = nextPiece(pb, nb);
    while (havePiece) {
      uint32_t pbNext = 0, nbNext = 0;
      const bool haveNext = nextPiece(pbNext, nbNext);
#if UPB_PREFETCH_L2
      if (haveNext) prefetchPiece(pbNext, nbNext);
#endif
      do {                                                             // `continue` below leaves the piece
        ++stPieces;
        const uint32_t P = min(32 * nb, np - 32 * pb);                // positions of this piece
        const uint32_t q0 = ts + 32 * pb;                             // first position
        const uint32_t W = 32 * nb + 2 * bw;                          // cells
        const uint32_t nbEff = (P + 31) >> 5;

        auto zeroCells = [&] () {
          // cells 0 .. W+1, eight bytes per store (D and DR are 8-byte aligned and PIECE_LEN + 2 bw + 4 cells long: W + 3 fits)
          for (uint32_t u = 2 * lane; u < W + 2; u += 64) { *(uint2*)(D + u) = make_uint2(0u, 0u); if (NONDIR) *(uint2*)(DR + u) = make_uint2(0u, 0u); }
          __syncwarp();
        };

        // B1: the tags within bandwidth of the piece -> dense cells: cell u holds position q0 - bw + u
        unsigned long long tot = 0;                                    // pooled weight of the piece's cells
        {
          const int64_t lo = (int64_t)q0 - bw, hi = (int64_t)q0 + 32 * nb - 1 + bw;
          const uint32_t cell0 = q0 - bw;                              // may wrap below 0: the unsigned difference below is still right
          if (BULK && !COEFF) {
            bool zeroed = false;                                       // the cells are zeroed behind the first bulk copies
            __syncwarp();                                              // the ring (stage + scf) was read by other lanes in B4 of the previous piece
#pragma unroll
            for (int str = 0; str < (NONDIR ? 2 : 1); ++str) {
              uint32_t* dst = str ? DR : D;
              piece_entries_bulk(a.T.ment, moSpan + MO_STRIDE * str, mBase, nMtS[str], lo, hi, lane, (const uint2*)stage, ringAddr, barAddr, bulkSeq, zeroed, zeroCells,
                                 [&] (uint32_t pos, const uint2& en) {
                const uint32_t u = pos - cell0;
                if ((en.x != 0) & (u < W) && !(a.hasControl && isCtl[en.y >> SPAN_SHIFT])) { atomicAdd(&dst[u], en.x); tot += en.x; }
              });
            }
            if (!zeroed) zeroCells();
          } else
          {
          zeroCells();
#pragma unroll
          for (int str = 0; str < (NONDIR ? 2 : 1); ++str) {
            uint32_t* dst = str ? DR : D;
            piece_entries_tab(a.T.ment, moSpan + MO_STRIDE * str, mBase, nMtS[str], lo, hi, lane, [&] (uint32_t pos, const uint2& en) {
              const uint32_t u = pos - cell0;
              if ((en.x != 0) & (u < W) && !(a.hasControl && isCtl[en.y >> SPAN_SHIFT])) { atomicAdd(&dst[u], en.x); tot += en.x; }
            });
          }
          }
        }
        tot = warp_sum(tot);                                           // also orders the atomics before the reads below
        __syncwarp();

        double dPiece = 0.0;
        if (COEFF) {
          // -z (misc/peakcall.cpp:191-200): the weight of a position is  sum_s c_s*coeff_s  followed by  + sum_s c_s,
          // both accumulated in sample order with every product and sum rounded.  One pass over the chunk per sample and
          // phase keeps that order (a sample has at most one entry per strand and position, so a pass has no conflicts).
          for (uint32_t u = lane; u < W + 2; u += 32) { Wf[u] = 0.0; if (NONDIR) Wr[u] = 0.0; }
          __syncwarp();
          for (int phase = 0; phase < 2; ++phase)
            for (uint32_t sm = 0; sm < a.nSamples; ++sm) {
              if (isCtl[sm]) continue;
              const double cf = coef[sm];
#pragma unroll
              for (int str = 0; str < (NONDIR ? 2 : 1); ++str) {
                double* dst = str ? Wr : Wf;
                const uint32_t cell0 = q0 - bw;
                piece_entries_tab(a.T.ment, moSpan + MO_STRIDE * str, mBase, nMtS[str], (int64_t)q0 - bw, (int64_t)q0 + 32 * nb - 1 + bw, lane, [&] (uint32_t pos, const uint2& en) {
                  if (en.x == 0 || (en.y >> SPAN_SHIFT) != sm) return;
                  const uint32_t u = pos - cell0;
                  if (u < W) dst[u] = __dadd_rn(dst[u], phase ? (double)en.x : __dmul_rn((double)en.x, cf));
                });
              }
              __syncwarp();
            }
          // every position of the piece in the reference's own order
          for (uint32_t b = 0; b < nbEff; ++b) {
            const uint32_t i = 32 * b + lane;
            const bool valid = i < P;
            double v = 0.0;
            if (valid) {
              const double f = exact_dense_f64(Ks, Wf, i, bw);
              v = NONDIR ? __dadd_rn(f, exact_dense_f64(Ks, Wr, i, bw)) : f;
              scd[i] = v;
              stExact += 32;
            }
            const uint32_t m = __ballot_sync(0xffffffffu, valid && v >= a.thr);
            if (lane == 0) mw[b] = m;
          }
        } else {
        dPiece = __dmul_rn((double)tot, a.dS);           // one conservative error bound for the whole piece
        // B2 + B3 with the window sums S0 = sum w, S1 = sum w*u, S2 = sum w*u^2 (u = cell index) held in type S: they are bounded by
        // tot, tot*W and tot*W^2, so 32-bit (wrapping differences, exact sums) is enough for all but the heaviest pieces
        auto classify = [&] (auto tag) {
          typedef decltype(tag) S;
          // B2: window sums of position 0 (cells 0 .. 2bw)
          S c0 = 0, c1 = 0, c2 = 0;
          for (uint32_t u = lane; u <= 2 * bw; u += 32) {
            const S w = (S)D[u] + (NONDIR ? (S)DR[u] : (S)0);
            c0 += w; c1 += w * (S)u; c2 += w * (S)u * (S)u;
          }
          if constexpr (sizeof(S) == 4) { c0 = warp_sum_u32(c0); c1 = warp_sum_u32(c1); c2 = warp_sum_u32(c2); }
          else { c0 = warp_sum(c0); c1 = warp_sum(c1); c2 = warp_sum(c2); }
          // B3: classify every position: moment form, exact order only inside the error band
          const double thrHi = a.thr + dPiece, thrLo = a.thr - dPiece;
          for (uint32_t b = 0; b < nbEff; ++b) {
            const uint32_t i = 32 * b + lane;
            S d0 = 0, d1 = 0, d2 = 0;
            if (i >= 1) {
              const S ue = (S)(i + 2 * bw), ul = (S)(i - 1);
              const S we = (S)D[i + 2 * bw] + (NONDIR ? (S)DR[i + 2 * bw] : (S)0);
              const S wl = (S)D[i - 1] + (NONDIR ? (S)DR[i - 1] : (S)0);
              d0 = we - wl; d1 = we * ue - wl * ul; d2 = we * ue * ue - wl * ul * ul;
            }
            S S0, S1, S2;
            double v;
            if constexpr (sizeof(S) == 4) {
              // tot * W^2 < 2^32: the sums and N itself (0 <= N <= bw^2 * S0) fit 32 bits, intermediate products may wrap
              S0 = c0 + warp_incl_scan_u32(d0); S1 = c1 + warp_incl_scan_u32(d1); S2 = c2 + warp_incl_scan_u32(d2);
              const uint32_t x = i + bw;
              const uint32_t N = (uint32_t)bw2 * S0 - (x * x * S0 - 2u * x * S1 + S2);
              v = __dmul_rn((double)N, a.cN);
            } else {
              S0 = c0 + warp_incl_scan(d0, lane); S1 = c1 + warp_incl_scan(d1, lane); S2 = c2 + warp_incl_scan(d2, lane);
              const unsigned long long x = i + bw;
              const unsigned long long N = bw2 * (unsigned long long)S0 - (x * x * (unsigned long long)S0 - 2ull * x * (unsigned long long)S1 + (unsigned long long)S2);
              v = __dmul_rn((double)(long long)N, a.cN);
            }
            c0 = __shfl_sync(0xffffffffu, S0, 31); c1 = __shfl_sync(0xffffffffu, S1, 31); c2 = __shfl_sync(0xffffffffu, S2, 31);
            const bool valid = i < P;
            bool above = valid && v >= thrHi;
            if (valid && !above && v >= thrLo) {                       // too close to call: the reference's own arithmetic decides
              const double f = exact_dense(Ks, D, i, bw);
              v = NONDIR ? __dadd_rn(f, exact_dense(Ks, DR, i, bw)) : f; // misc/peakcall.cpp:57
              above = v >= a.thr;
              stExact += 32;
            }
            if (valid) scf[i] = (float)v;
            const uint32_t m = __ballot_sync(0xffffffffu, above);
            if (lane == 0) mw[b] = m;
          }
        };
        if (tot * (unsigned long long)W * (unsigned long long)W < 0x100000000ull) classify((uint32_t)0); else classify((unsigned long long)0);
        }
        if (lane == 0) mw[nbEff] = 0;
        __syncwarp();

        // B4: runs of above-threshold positions
        uint32_t nRuns = 0;
        {
          uint32_t n = 0;
          if (lane < nbEff) { const uint32_t m = mw[lane]; const uint32_t prev = lane > 0 ? (mw[lane - 1] >> 31) : 0u; n = __popc(m & ~((m << 1) | prev)); }
          nRuns = warp_sum_u32(n);
        }
        if (nRuns == 0) continue;
        uint32_t base;
        bool room;
        if (nRuns > RUN_CHUNK) {                                       // rare: more runs than a chunk holds get their own reservation
          base = 0;
          if (lane == 0) base = atomicAdd(a.runCount, nRuns);
          base = __shfl_sync(0xffffffffu, base, 0);
          room = (uint64_t)base + nRuns <= a.runCap;
        } else {
          if (rUsed + nRuns > RUN_CHUNK) {                             // next chunk; what is left of the old one is marked unused
            if (rRoom && rUsed + lane < RUN_CHUNK) a.runs[rBase + rUsed + lane].span = RUN_INVALID;
            if (lane == 0) rBase = atomicAdd(a.runCount, RUN_CHUNK);
            rBase = __shfl_sync(0xffffffffu, rBase, 0);
            rUsed = 0;
            rRoom = (uint64_t)rBase + RUN_CHUNK <= a.runCap;
          }
          base = rBase + rUsed;
          rUsed += nRuns;
          room = rRoom;
        }
        if (!room && lane == 0) atomicOr(a.flags, FLAG_RUN_OVERFLOW);
        if (room) rTotal += nRuns;

        uint32_t cursor = 0;
        for (uint32_t r = 0; r < nRuns; ++r) {
          // next run [r0, r1) in position order (uniform across the warp)
          { uint32_t w = cursor >> 5; uint32_t m = mw[w] >> (cursor & 31);
            while (!m) { ++w; cursor = w << 5; m = mw[w]; }
            cursor += __ffs(m) - 1; }
          const uint32_t r0 = cursor;
          for (;;) {
            const uint32_t x = ~(mw[cursor >> 5] >> (cursor & 31));
            cursor += x ? (uint32_t)(__ffs(x) - 1) : 32u;
            if ((cursor & 31) != 0 || cursor >= 32 * nbEff || !(mw[cursor >> 5] & 1u)) break;
          }
          const uint32_t r1 = cursor;
          if (!room) continue;

          // pooled hit sum of the density sources == nonControlSum of misc/peakcall.cpp:36-40 (HitCount wraps mod 2^32)
          uint32_t wsum = 0;
          double best = -1.0; uint32_t bestI = r0;
          if (COEFF) {
            for (uint32_t i = r0 + lane; i < r1; i += 32) {
              wsum += D[bw + i] + (NONDIR ? DR[bw + i] : 0u);
              if (scd[i] > best) { best = scd[i]; bestI = i; }          // ascending i per lane: first strict maximum
            }
            wsum = warp_sum_u32(wsum);
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              const double oe = __shfl_xor_sync(0xffffffffu, best, o);
              const uint32_t oi = __shfl_xor_sync(0xffffffffu, bestI, o);
              if (oe > best || (oe == best && oi < bestI)) { best = oe; bestI = oi; }
            }
          } else {
          float fmax_ = 0.0f;                                         // the fast values are >= 0: their bit patterns order like the values
          for (uint32_t i = r0 + lane; i < r1; i += 32) { wsum += D[bw + i] + (NONDIR ? DR[bw + i] : 0u); fmax_ = fmaxf(fmax_, scf[i]); }
          wsum = warp_sum_u32(wsum);
          fmax_ = __uint_as_float(__reduce_max_sync(0xffffffffu, __float_as_uint(fmax_)));

          // peak: first strict maximum of the exact-order score (misc/data.cpp:97-101).  Candidates are the positions
          // whose fast value (kept as float) could still be the maximum; they alone are evaluated exactly:
          // exact_i <= f_i (1+2^-24) + dPiece and exact_max >= f_max (1-2^-24) - dPiece.
          const double cut = (double)fmax_ * (1.0 - 0x1p-22) - 2.5 * dPiece;
          for (uint32_t g = r0; g < r1; g += 32) {
            const uint32_t i = g + lane;
            uint32_t cm = __ballot_sync(0xffffffffu, i < r1 && (double)scf[i] >= cut);
            while (cm) {                                               // ascending position: a later tie never replaces an earlier maximum
              const uint32_t ci = g + (uint32_t)(__ffs(cm) - 1);
              const double f = exact_dense_warp(Ks, D, stage, ci, bw, lane);
              const double e = NONDIR ? __dadd_rn(f, exact_dense_warp(Ks, DR, stage, ci, bw, lane)) : f;
              if (e > best) { best = e; bestI = ci; }
              ++stExact;
              cm &= cm - 1;
            }
          }
          }
          if (lane == 0) {
            SpanRun rr;
            rr.left = q0 + r0; rr.len = r1 - r0; rr.peak = q0 + bestI; rr.wsum = wsum; rr.score = best; rr.span = work; rr.rank = spanRuns + r;
            a.runs[base + r] = rr;
          }
        }
        spanRuns += nRuns;
      } while (0);
      pb = pbNext; nb = nbNext; havePiece = haveNext;
    }
    if (lane == 0) a.spanRunCount[work] = spanRuns;
  }
  }
  if (rRoom && rUsed + lane < RUN_CHUNK) a.runs[rBase + rUsed + lane].span = RUN_INVALID;
  if (lane == 0 && rTotal) atomicAdd(a.runReal, rTotal);
  stExact = warp_sum_u64(stExact);                                    // lane-divergent (uncertain positions) + 32 x uniform (peak candidates)
  if (lane == 0) {
    if (stPieces) atomicAdd(a.stat + 1, stPieces);
    if (stExact) atomicAdd(a.stat + 2, stExact / 32);
    if (stSpans) atomicAdd(a.stat + 3, stSpans);
  }
}

// ------------------------------------------------------------------------------------------------
// counters of the region pipeline.  They live in device memory (upb_ctx::d_small) and the host reads them ONCE, when the call
// is over: every kernel below takes its element counts from here and walks them with a grid stride, so the whole pipeline is
// enqueued without a single synchronising read-back (r01 had four per call).
// ------------------------------------------------------------------------------------------------
struct GroupCtr {                   // zeroed before every group of units (a call works through its units in one or more groups)
  uint32_t flags;                   // FLAG_RUN_OVERFLOW / FLAG_CAP_OVERFLOW of this group
  uint32_t ticket;                  // span_kernel work ticket
  uint32_t runCount;                // run slots handed out
  uint32_t runReal;                 // runs written
  uint32_t nWarm, poolTicket, nHot, pad0;
  unsigned long long totSpanRuns;   // scan totals: runs by span order,
  unsigned long long nCand;         //   candidate regions (= Region objects of the reference),
  unsigned long long pv;            //   hit-filter survivors << PACK_SHIFT | their positions,
  unsigned long long n1, cells;     //   the same two unpacked (n1 clamped to the capacity of the per-survivor arrays),
  unsigned long long rows;          //   accepted rows of span_kernel's candidates
  unsigned long long pad1[4];
};
static_assert(sizeof(GroupCtr) <= 128, "GroupCtr: the scan totals of dump_and_corr start at byte 128 of d_small");

struct CallCtr {                    // zeroed once per call; accumulated over the groups
  unsigned long long stat[4];       // hot blocks, pieces, exact evaluations, spans with a hot block
  unsigned long long rows;          // table rows so far (left-edge replay rows included)
  unsigned long long cands, runs, n1, cells, warm;
  unsigned long long needSlots;     // largest runCount of a group (what runCap has to be when a group overflowed)
  uint32_t flags, pad;
};
constexpr size_t CALLCTR_OFFSET = 512;   // byte offset of CallCtr inside d_small
constexpr size_t LANE1_GROUPCTR_OFFSET = 1024;   // the second lane's GroupCtr (upb_ctx::laneStream); lane 0's is at byte 0
constexpr size_t TAIL_GROUPCTR_OFFSET = 1536;    // two more: the halves of a split statistics + emit tail (compact_pv_split_kernel), 256 bytes apart

__device__ __forceinline__ bool pipe_dead (const GroupCtr* g) { return (g->flags & FLAG_RUN_OVERFLOW) != 0; }   // runs are incomplete: nothing downstream may index through them

// ------------------------------------------------------------------------------------------------
// ordering, stitching, filtering
// ------------------------------------------------------------------------------------------------
struct Cand2 {                      // a candidate region == one Region object of the reference
  uint32_t unit, left, len, peak, wsum;
  uint32_t slot;                    // run slot of the first piece
  uint32_t nPieces, pad;
  double score;
};

// position of every run in (span, rank) order; spanOrd = exclusive scan of spanRunCount over the group's spans
__global__ void run_order_kernel (const SpanRun* __restrict__ runs, const GroupCtr* __restrict__ g, uint32_t cap, const unsigned long long* __restrict__ spanOrd,
                                  uint32_t* __restrict__ ordIdx) {
  if (pipe_dead(g)) return;
  const uint32_t n = min(g->runCount, cap);                           // slots handed out; unfilled ones are skipped
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t sp = runs[i].span;
    if (sp == RUN_INVALID) continue;
    ordIdx[spanOrd[sp] + runs[i].rank] = i;
  }
}

__device__ __forceinline__ uint32_t unit_of_span (const SpanUnit* __restrict__ units, uint32_t nUnits, uint32_t span) {
  uint32_t lo = 0, hi = nUnits;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (units[mid].spanStart <= span) lo = mid; else hi = mid; }
  return lo;
}

__global__ void run_heads2_kernel (const SpanRun* __restrict__ runs, const uint32_t* __restrict__ ordIdx, const GroupCtr* __restrict__ g,
                                   const SpanUnit* __restrict__ units, uint32_t nUnits, uint32_t* __restrict__ head) {
  if (pipe_dead(g)) return;
  const uint32_t n = g->runReal;
  for (uint32_t j = blockIdx.x * blockDim.x + threadIdx.x; j < n; j += gridDim.x * blockDim.x) {
    bool cont = false;
    if (j > 0) {
      const SpanRun a = runs[ordIdx[j]], p = runs[ordIdx[j - 1]];
      cont = a.left == p.left + p.len && unit_of_span(units, nUnits, a.span) == unit_of_span(units, nUnits, p.span);
    }
    head[j] = cont ? 0u : 1u;
  }
}

// runs -> candidates (stitching what a piece or span boundary cut), with the hit filter (misc/peakcall.cpp:41, first clause) and
// the length of the survivors (the size of their pooled position cells) written alongside: pass[i], candLen[i]
__global__ void merge_runs2_kernel (const SpanRun* __restrict__ runs, const uint32_t* __restrict__ ordIdx, const uint32_t* __restrict__ head,
                                    const unsigned long long* __restrict__ headScan, const GroupCtr* __restrict__ g, const SpanUnit* __restrict__ units, uint32_t nUnits,
                                    double hitThr, Cand2* __restrict__ cand, uint32_t* __restrict__ pass, uint32_t* __restrict__ candLen,
                                    unsigned long long* __restrict__ unitCand) {
  if (pipe_dead(g)) return;
  const uint32_t n = g->runReal;
  for (uint32_t base = blockIdx.x * blockDim.x; base < n; base += gridDim.x * blockDim.x) {   // whole warps stay together (match_any below)
    const uint32_t j = base + threadIdx.x;
    const bool isHead = j < n && head[j];
    uint32_t myUnit = 0xFFFFFFFFu;
    if (isHead) myUnit = unit_of_span(units, nUnits, runs[ordIdx[j]].span);
    {                                                                 // candidates per unit: one atomic per (warp, unit) instead of one per candidate
      const uint32_t peers = __match_any_sync(0xffffffffu, myUnit);
      if (isHead && (threadIdx.x & 31) == (uint32_t)(__ffs(peers) - 1)) atomicAdd(&unitCand[myUnit], (unsigned long long)__popc(peers));
    }
    if (!isHead) continue;
    const SpanRun f = runs[ordIdx[j]];
    Cand2 c;
    c.unit = myUnit; c.left = f.left; c.len = f.len; c.peak = f.peak; c.wsum = f.wsum; c.score = f.score;
    c.slot = ordIdx[j]; c.nPieces = 1; c.pad = 0;
    for (uint32_t q = j + 1; q < n && !head[q]; ++q) {
      const SpanRun r = runs[ordIdx[q]];
      c.len += r.len;
      c.wsum += r.wsum;                                               // HitCount wraps mod 2^32 (misc/peakcall.cpp:36-40)
      if (r.score > c.score) { c.score = r.score; c.peak = r.peak; }  // first strict maximum (misc/data.cpp:97-101)
      ++c.nPieces;
    }
    const unsigned long long i = headScan[j];
    cand[i] = c;
    const bool hit = (double)c.wsum >= hitThr;
    pass[i] = hit ? 1u : 0u;
    candLen[i] = hit ? c.len : 0u;
  }
}

// ordinal of the first candidate of every unit of the group [u0, u1) among the group's candidates
__global__ void unit_first_kernel (const unsigned long long* __restrict__ unitCand, uint32_t u0, uint32_t u1, unsigned long long* __restrict__ unitFirst) {
  if (blockIdx.x == 0 && threadIdx.x == 0) { unsigned long long run = 0; for (uint32_t u = u0; u < u1; ++u) { unitFirst[u] = run; run += unitCand[u]; } }
}

// candidates of span_kernel that lie in the stretch the left-edge replay owns (upb_edge.cuh) are not Region objects of the
// reference: they are dropped and counted per affected unit
__global__ void edge_drop_kernel (const uint32_t* __restrict__ affected, const uint32_t* __restrict__ stop, uint32_t nAffected,
                                  const Cand2* __restrict__ cand, const GroupCtr* __restrict__ g, uint32_t* __restrict__ pass, uint32_t* __restrict__ candLen,
                                  uint32_t* __restrict__ dropCount) {
  if (pipe_dead(g)) return;
  const unsigned long long n = g->nCand;
  for (unsigned long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const uint32_t unit = cand[i].unit;
    for (uint32_t k = 0; k < nAffected; ++k)
      if (affected[k] == unit) { if (cand[i].left <= stop[k]) { pass[i] = 0; candLen[i] = 0; atomicAdd(&dropCount[k], 1u); } break; }
  }
}

// hit-filter survivors -> list (candidate indices, ascending); pvScan = exclusive scan of (pass << PACK_SHIFT | candLen).
// Also unpacks the scan total and checks it against the capacities of the per-survivor arrays.
__global__ void compact_pv_kernel (const uint32_t* __restrict__ pass, const unsigned long long* __restrict__ pvScan, GroupCtr* __restrict__ g,
                                   uint32_t n1Cap, unsigned long long cellCap, uint32_t* __restrict__ list) {
  if (pipe_dead(g)) { if (blockIdx.x == 0 && threadIdx.x == 0) { g->n1 = 0; g->cells = 0; } return; }
  const unsigned long long n = g->nCand;
  for (unsigned long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (pass[i]) { const unsigned long long k = pvScan[i] >> PACK_SHIFT; if (k < n1Cap) list[k] = (uint32_t)i; }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long n1 = g->pv >> PACK_SHIFT, cells = g->pv & PACK_MASK;
    g->n1 = n1 < n1Cap ? n1 : n1Cap;
    g->cells = cells;
    if (n1 > n1Cap || cells > cellCap) atomicOr(&g->flags, FLAG_CAP_OVERFLOW);
  }
}

// The same for a call whose statistics + emit tail runs in two halves (upb_api.cu: split tail): survivors of the units before uMid go to
// listA, the others to listB, each half with its own counters — gA takes over the group's totals (so that group_end_kernel adds them to the
// call's once), gB starts from zero.
__global__ void compact_pv_split_kernel (const uint32_t* __restrict__ pass, const unsigned long long* __restrict__ pvScan, const GroupCtr* g, GroupCtr* gA, GroupCtr* gB,
                                         const unsigned long long* __restrict__ unitFirst, uint32_t uMid, uint32_t n1Cap, unsigned long long cellCap,
                                         uint32_t* __restrict__ listA, uint32_t* __restrict__ listB) {
  const bool dead = pipe_dead(g);
  const unsigned long long n = dead ? 0ull : g->nCand;
  const unsigned long long cMid = dead ? 0ull : unitFirst[uMid];       // ordinal of the first candidate of unit uMid among the group's candidates
  const unsigned long long n1 = dead ? 0ull : g->pv >> PACK_SHIFT;
  const unsigned long long nA = cMid < n ? pvScan[cMid] >> PACK_SHIFT : n1;
  for (unsigned long long i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x)
    if (pass[i]) {
      const unsigned long long k = pvScan[i] >> PACK_SHIFT;
      if (i < cMid) { if (k < n1Cap) listA[k] = (uint32_t)i; }
      else if (k - nA < n1Cap) listB[k - nA] = (uint32_t)i;
    }
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    const unsigned long long cells = dead ? 0ull : g->pv & PACK_MASK, nB = n1 - nA;
    GroupCtr a = *g;
    a.poolTicket = 0; a.n1 = nA < n1Cap ? nA : n1Cap; a.cells = cells; a.rows = 0;
    if (nA > n1Cap || nB > n1Cap || cells > cellCap) a.flags |= FLAG_CAP_OVERFLOW;
    *gA = a;
    GroupCtr b;
    memset(&b, 0, sizeof b);
    b.n1 = nB < n1Cap ? nB : n1Cap; b.cells = cells;
    *gB = b;
  }
}

// Region::exptSums (misc/data.cpp:116-131) and Region::posKurtosis (misc/data.cpp:133-182) of the candidates that passed the
// hit filter, in two kernels.
//
// region_pool_kernel: eight lanes work on one region.  The region's entries are contiguous chunks of the tile-major copy (the
// mini-tiles it touches, one chunk per 4096-position span); every lane takes entries, adds the count to its sample's slot
// (shared-memory atomics, -> the regions x samples matrix row) and to the region's pooled position cell (ALL samples, controls
// included, minus the tags the reference forgets, Q4).  The four groups of a warp run as independent state machines inside one
// lock-step loop — fetch a region / pool one batch of 64 entries / write the finished region out — so a group that finishes a
// small region takes the next one at once instead of idling until its warp-mates are through a big one (v5–v16 handed a warp
// four regions at a time: 14.7 of 32 lanes active on average, 433 M warp instructions).
// region_moment_kernel: one thread per region walks its pooled cells in position order: count and count*offset (offsets as
// UShort, both sums as HitCount), then the two central-moment sums added sequentially exactly as the reference does.
struct StatArgs {
  Tags T;
  const SpanUnit* units;
  const Cand2* cand;
  const uint32_t* list;             // hit-filter survivors (candidate indices, ascending)
  const GroupCtr* g;                // g->n1 survivors
  int nondir;
  double kurtThr;
  const unsigned long long* pvScan; // [nCand] low PACK_SHIFT bits: first pooled cell of every surviving candidate
  unsigned long long cellCap;       // capacity of `cells`
  uint32_t* cells;                  // pooled counts per position of every surviving candidate
  uint32_t* counts;                 // [n1 * nSamples]
  double* kurt;                     // [n1]
  uint32_t* pass;                   // [n1] kurtosis filter (misc/peakcall.cpp:41 second clause)
  uint32_t* ticket;
};

constexpr int ST_GROUP = 8;                         // lanes per region
constexpr int ST_PER_WARP = 32 / ST_GROUP;
constexpr int ST_CHUNK = 256;                       // regions up to this long pool their positions in shared memory; longer ones straight in global memory
constexpr int ST_WARPS = 8;
constexpr uint32_t ST_TICKET = 4;                   // regions a group takes per atomic

__global__ void __launch_bounds__(ST_WARPS * 32) region_pool_kernel (StatArgs a) {
  extern __shared__ __align__(16) unsigned char stat_smem_raw[];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t g = lane / ST_GROUP, gl = lane % ST_GROUP, leader = lane & ~(uint32_t)(ST_GROUP - 1);
  const uint32_t n = a.T.nSamples;
  const uint32_t nList = (uint32_t)a.g->n1;
  // regions a group takes per atomic: four where there are many per group (one same-address atomic per region serialises), one where a rank's
  // shard leaves a group fewer than eight (N = 8 on C3: 1.6 regions per group — with tickets of four, 60 % of the groups got none)
  const uint32_t tkSize = nList > gridDim.x * (uint32_t)(ST_WARPS * ST_PER_WARP) * 8u ? ST_TICKET : 1u;
  const size_t groupBytes = ((size_t)(ST_CHUNK + n) * 4 + 15) / 16 * 16;   // per group: cell[ST_CHUNK] u32, cnts[n] u32
  unsigned char* gbase = stat_smem_raw + ((size_t)warp * ST_PER_WARP + g) * groupBytes;
  uint32_t* cell = (uint32_t*)gbase;
  uint32_t* cnts = cell + ST_CHUNK;

  // group state (identical in the lanes of a group)
  bool active = false, exhausted = false;
  uint32_t tkNext = 0, tkEnd = 0;
  uint32_t i = 0, left = 1, right = 0, contig = 0;
  bool single = true, cellsOk = true;
  uint32_t d = 0, dEnd = 0, sp = 0, spEnd = 0, ma = 0, mb = 0, gb = 0;
  unsigned long long eb = 0, e1 = 0, cellBase = 0;

  for (;;) {
    // ---- fetch: a group without a region takes the next one -------------------------------------------------------------------
    uint32_t t = 0;
    const bool wantTicket = !active && !exhausted && tkNext == tkEnd;
    if (wantTicket && gl == 0) t = atomicAdd(a.ticket, tkSize);
    t = __shfl_sync(0xffffffffu, t, leader);
    if (wantTicket) { tkNext = t; tkEnd = min(t + tkSize, nList); if (t >= nList) { exhausted = true; tkEnd = tkNext; } }
    bool needSeg = false;                                             // (re)load the entry range of segment (d, sp)
    if (!active && tkNext < tkEnd) {
      i = tkNext++;
      const uint32_t ci = __ldg(a.list + i);
      const Cand2 c = a.cand[ci];
      const SpanUnit U = a.units[c.unit];
      left = c.left; right = c.left + c.len - 1; contig = U.contig;
      single = c.len <= (uint32_t)ST_CHUNK;
      cellBase = a.pvScan[ci] & PACK_MASK;
      cellsOk = cellBase + c.len <= a.cellCap;                        // beyond the capacity: the call is repeated with a larger array (FLAG_CAP_OVERFLOW)
      d = a.nondir ? 0u : U.buffer; dEnd = a.nondir ? 1u : U.buffer;
      for (uint32_t s2 = gl; s2 < n; s2 += ST_GROUP) cnts[s2] = 0;
      if (single) { for (uint32_t k = gl; k < (uint32_t)ST_CHUNK; k += ST_GROUP) cell[k] = 0; }
      else if (cellsOk) { for (unsigned long long k = gl; k < c.len; k += ST_GROUP) a.cells[cellBase + k] = 0; }
      active = true;
      needSeg = true;
      sp = 0xFFFFFFFFu;                                               // "strand d not opened yet"
    }
    if (!__any_sync(0xffffffffu, active)) break;
    __syncwarp();

    // ---- segment: entries of strand d, span sp that lie in the region's mini-tiles --------------------------------------------
    if (active && needSeg) {
      for (;;) {
        if (sp == 0xFFFFFFFFu) {                                      // open strand d
          gb = __ldg(a.T.mtBase + d * a.T.C + contig);
          const uint32_t nMt = __ldg(a.T.mtBase + d * a.T.C + contig + 1) - gb;
          if (nMt == 0) { if (d == dEnd) { eb = e1 = 0; break; } ++d; continue; }
          ma = min((left - 1) >> MT_SHIFT, nMt - 1); mb = min((right - 1) >> MT_SHIFT, nMt - 1);
          sp = ma / SPAN_MT; spEnd = mb / SPAN_MT;
        }
        const uint32_t m0 = max(ma, sp * SPAN_MT), m1 = min(mb, sp * SPAN_MT + SPAN_MT - 1);
        eb = __ldg(a.T.mtOff + gb + m0); e1 = __ldg(a.T.mtOff + gb + m1 + 1);
        break;
      }
    }

    // ---- pool one batch of up to 8 * ST_GROUP entries ----------------------------------------------------------------------------
    if (active) {
      uint2 en[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const unsigned long long e = eb + gl + q * ST_GROUP; en[q] = e < e1 ? __ldg(a.T.ment + e) : make_uint2(0u, 0u); }
      uint32_t eff[8];
#pragma unroll
      for (int q = 0; q < 8; ++q) { const unsigned long long e = eb + gl + q * ST_GROUP; eff[q] = (a.T.mEff && e < e1) ? __ldg(a.T.mEff + e) : en[q].x; }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const uint32_t p = sp * SPAN_LEN + (en[q].y & (SPAN_LEN - 1)) + 1;
        if (eff[q] && p >= left && p <= right) {
          atomicAdd(&cnts[en[q].y >> SPAN_SHIFT], eff[q]);
          if (single) atomicAdd(&cell[p - left], eff[q]); else if (cellsOk) atomicAdd(&a.cells[cellBase + (p - left)], eff[q]);
        }
      }
      eb += 8 * ST_GROUP;
    }
    __syncwarp();

    // ---- segment / region finished? ----------------------------------------------------------------------------------------------
    if (active && eb >= e1) {
      bool done = false;
      if (sp != 0xFFFFFFFFu && sp < spEnd) { ++sp; }
      else if (d < dEnd) { ++d; sp = 0xFFFFFFFFu; }
      else done = true;
      if (!done) {
        for (;;) {                                                    // same walk as above (kept in one place would need a lambda with many captures)
          if (sp == 0xFFFFFFFFu) {
            gb = __ldg(a.T.mtBase + d * a.T.C + contig);
            const uint32_t nMt = __ldg(a.T.mtBase + d * a.T.C + contig + 1) - gb;
            if (nMt == 0) { if (d == dEnd) { done = true; break; } ++d; continue; }
            ma = min((left - 1) >> MT_SHIFT, nMt - 1); mb = min((right - 1) >> MT_SHIFT, nMt - 1);
            sp = ma / SPAN_MT; spEnd = mb / SPAN_MT;
          }
          const uint32_t m0 = max(ma, sp * SPAN_MT), m1 = min(mb, sp * SPAN_MT + SPAN_MT - 1);
          eb = __ldg(a.T.mtOff + gb + m0); e1 = __ldg(a.T.mtOff + gb + m1 + 1);
          break;
        }
      }
      if (done) {                                                     // write the region out: matrix row and (short regions) the pooled cells
        for (uint32_t s2 = gl; s2 < n; s2 += ST_GROUP) a.counts[(uint64_t)i * n + s2] = cnts[s2];
        if (single && cellsOk) { const uint32_t len = right - left + 1; for (uint32_t k = gl; k < len; k += ST_GROUP) a.cells[cellBase + k] = cell[k]; }
        active = false;
      }
    }
    __syncwarp();
  }
}

__global__ void region_moment_kernel (StatArgs a) {
  const uint32_t nList = (uint32_t)a.g->n1;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nList; i += gridDim.x * blockDim.x) {
    const uint32_t ci = a.list[i];
    const uint32_t len = a.cand[ci].len;
    const unsigned long long cb = a.pvScan[ci] & PACK_MASK;
    if (cb + len > a.cellCap) { a.kurt[i] = 0.0; a.pass[i] = 0; continue; }   // capacity overflow: the call is repeated
    const uint32_t* __restrict__ cells = a.cells + cb;
    uint32_t count = 0, sum = 0;                                        // HitCount arithmetic (wraps), offsets as UShort (Q8)
    for (uint32_t o = 0; o < len; ++o) { const uint32_t k = cells[o]; count += k; sum += k * (o & 0xFFFFu); }
    const double xbar = (double)sum / (double)count;                   // misc/data.cpp:147
    double sum2 = 0.0, sum4 = 0.0;
    for (uint32_t o = 0; o < len; ++o) {
      const uint32_t pc = cells[o];
      if (!pc) continue;                                                // positions without a stored hit vector are not visited (misc/data.cpp:171-173)
      const double dd = __dsub_rn((double)(o & 0xFFFFu), xbar);
      const double d2 = __dmul_rn(dd, dd);
      const double d4 = __dmul_rn(d2, d2);                              // pow(x,4) == (x*x)*(x*x) in the reference build (Q9)
      sum2 = __dadd_rn(sum2, __dmul_rn((double)pc, d2));
      sum4 = __dadd_rn(sum4, __dmul_rn((double)pc, d4));
    }
    const double k = __ddiv_rn(__dmul_rn(__dsub_rn((double)count, 1.0), sum4), __dmul_rn(sum2, sum2));   // misc/data.cpp:181
    a.kurt[i] = k;
    a.pass[i] = (a.kurtThr == 0.0 || (len > 1 && k <= a.kurtThr)) ? 1u : 0u;
  }
}

// strand-correlation filter (third clause); list2 holds positions in list1 space, corr[k] belongs to list2[k]
__global__ void cand_corr_kernel (const uint32_t* __restrict__ list2, uint32_t n2, const double* __restrict__ corr, double corrThr,
                                  uint32_t* __restrict__ pass, double* __restrict__ rowCorr) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n2) return;
  const uint32_t i = list2[k];
  rowCorr[i] = corr[k];
  if (corrThr > -1.0 && !(corr[k] >= corrThr)) pass[i] = 0;         // NaN fails
}

__global__ void dump_desc2_kernel (const Cand2* __restrict__ cand, const uint32_t* __restrict__ list1, const uint32_t* __restrict__ list2, uint32_t n,
                                   uint32_t* __restrict__ dUnit, uint32_t* __restrict__ dLeft, uint32_t* __restrict__ dLen, uint32_t* __restrict__ dChunks) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Cand2 c = cand[list1[list2[i]]];
  dUnit[i] = c.unit; dLeft[i] = c.left; dLen[i] = c.len; dChunks[i] = (c.len + TILE - 1) / TILE;
}

// ------------------------------------------------------------------------------------------------
// where the rows go.  A table row has the place (buffer, contig, left) gives it: the rows of unit u follow those of all units
// before it.  row_place_kernel (one block) counts the group's accepted rows per unit and turns that into the table row of each
// unit's first row.  With several GPUs (one rank per GPU, contigs sharded, upb_call_regions_shared) the per-unit counts of ALL
// ranks are needed for that: every rank publishes its counts in one host segment all ranks map and waits — on the device, inside
// this kernel — until the others have published theirs.  No collective, no host round trip (r01: an NCCL all-reduce + .cpu()).
// ------------------------------------------------------------------------------------------------
struct XchgArgs {                   // the exchange area of a shared table (mapped host memory, one per job)
  volatile unsigned long long* counts;   // [world][2*C] accepted rows per (buffer, contig), written by the rank that owns the contig
  volatile unsigned long long* ready;    // [world] step number once counts[rank][] of that step are complete
  uint32_t rank, world, C;
  unsigned long long step;          // 1, 2, 3 ... (all ranks call in lock step)
};

struct PlaceArgs {
  const Cand2* cand;
  const uint32_t* list;             // survivors
  const unsigned long long* pass2Scan;   // [n1] exclusive scan of the final pass flags
  GroupCtr* g;
  CallCtr* call;
  const SpanUnit* units;
  uint32_t u0, u1;                  // units of this group
  const uint32_t* edgeRows;         // [nUnits] rows the left-edge replay contributes to each unit (NULL: none)
  unsigned long long* normalFirst;  // [nUnits + 1] scratch
  long long* rowShift;              // [nUnits] out: table row of a unit's candidate i = rowShift[u] + pass2Scan[i]
  unsigned long long* edgeRowBase;  // [nUnits] out: table row of the unit's first replay row
  unsigned long long* unitRows;     // [nUnits] out: rows of the unit (replay rows included)
  int shared;                       // 0: one GPU (rows follow the earlier groups' rows); 1: exchange
  int countOnly;                    // shared table, a group that is not the call's last: count this group's rows per unit and leave; the last group's
                                    // launch (ordered behind the others) publishes and places the units [allU0, allU1) of all groups
  uint32_t allU0, allU1;
  XchgArgs x;
};

__global__ void __launch_bounds__(256) row_place_kernel (PlaceArgs a) {
  const uint32_t nu = a.u1 - a.u0;
  const uint32_t n1 = (uint32_t)a.g->n1;
  const unsigned long long rows = a.g->rows;
  // first survivor of every unit (survivors are in unit order): rows of span_kernel's candidates before the unit
  for (uint32_t t = threadIdx.x; t <= nu; t += blockDim.x) {
    const uint32_t u = a.u0 + t;
    uint32_t lo = 0, hi = n1;
    if (t == nu) lo = n1;
    else while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (a.cand[a.list[mid]].unit < u) lo = mid + 1; else hi = mid; }
    a.normalFirst[u] = lo < n1 ? a.pass2Scan[lo] : rows;
  }
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < nu; t += blockDim.x) {
    const uint32_t u = a.u0 + t;
    a.unitRows[u] = a.normalFirst[u + 1] - a.normalFirst[u] + (a.edgeRows ? a.edgeRows[u] : 0u);
  }
  __syncthreads();
  if (!a.shared) {
    if (threadIdx.x == 0) {
      unsigned long long base = a.call->rows;
      for (uint32_t u = a.u0; u < a.u1; ++u) {
        const unsigned long long er = a.edgeRows ? a.edgeRows[u] : 0u;
        a.edgeRowBase[u] = base;
        a.rowShift[u] = (long long)(base + er) - (long long)a.normalFirst[u];
        base += a.unitRows[u];
      }
      a.call->rows = base;                                               // the next group's rows follow (its row_place_kernel is ordered behind this one)
    }
    return;
  }
  if (a.countOnly) return;
  // shared table: publish this rank's counts, wait for the others, place by (buffer, contig) over all ranks
  const uint32_t C2 = 2 * a.x.C;
  const uint32_t nAll = a.allU1 - a.allU0;
  volatile unsigned long long* mine = a.x.counts + (size_t)a.x.rank * C2;
  for (uint32_t k = threadIdx.x; k < C2; k += blockDim.x) mine[k] = 0;
  __syncthreads();
  for (uint32_t t = threadIdx.x; t < nAll; t += blockDim.x) { const SpanUnit U = a.units[a.allU0 + t]; mine[U.buffer * a.x.C + U.contig] = a.unitRows[a.allU0 + t]; }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) { a.x.ready[a.x.rank] = a.x.step; __threadfence_system(); }
  for (uint32_t r = threadIdx.x; r < a.x.world; r += blockDim.x) {
    uint32_t spins = 0;                                               // bounded (a few seconds): a rank that died must not hang the others' GPUs
    while (a.x.ready[r] < a.x.step) { __nanosleep(256); if (++spins > (1u << 24)) { atomicOr(&a.g->flags, FLAG_XCHG_TIMEOUT); break; } }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    // table row of the first row of every (buffer, contig), summed over the ranks in table order; this rank's units pick theirs
    uint32_t t = 0;
    unsigned long long base = 0;
    for (uint32_t k = 0; k < C2 && t < nAll; ++k) {
      const SpanUnit U = a.units[a.allU0 + t];
      if (U.buffer * a.x.C + U.contig == k) {
        const uint32_t u = a.allU0 + t;
        const unsigned long long er = a.edgeRows ? a.edgeRows[u] : 0u;
        a.edgeRowBase[u] = base;
        a.rowShift[u] = (long long)(base + er) - (long long)a.normalFirst[u];
        ++t;
      }
      for (uint32_t r = 0; r < a.x.world; ++r) base += a.x.counts[(size_t)r * C2 + k];
    }
    unsigned long long mineRows = 0;                                   // this rank's rows of the call
    for (uint32_t u = a.allU0; u < a.allU1; ++u) mineRows += a.unitRows[u];
    a.call->rows += mineRows;
  }
}

// accepted candidates -> the column arrays of the result table (canonical order: buffer, contig, left).  The columns are MAPPED
// HOST memory: the rows cross PCIe as they are written (coalesced 4- and 8-byte stores), there is no device copy of the table
// and no D2H copy whose size the host would have to know.
struct EmitArgs {
  const Cand2* cand;
  const uint32_t* list;                  // hit-filter survivors
  const uint32_t* pass;                  // [n1] after the kurtosis (and correlation) filters
  const unsigned long long* passScan;
  GroupCtr* g;
  const SpanUnit* units;
  const unsigned long long* unitFirst;   // [nUnits] ordinal of the first candidate of each unit among the group's candidates
  const long long* rowShift;             // [nUnits] row_place_kernel
  const int32_t* affOfUnit;              // [nUnits] index into the left-edge replay's arrays, -1 if the unit is not replayed (NULL: no replay)
  const uint32_t* edgeCand;              // [nAffected] candidates the replay produced
  const uint32_t* edgeDrop;              // [nAffected] candidates of span_kernel it replaced
  const uint32_t* statCounts;            // [n1 * nSamples]
  const double* statKurt;                // [n1]
  const double* rowCorr;                 // [n1] or NULL
  uint32_t nSamples;
  unsigned long long rowCap;             // rows the table has room for
  // outputs (mapped host memory)
  uint8_t* buffer; uint32_t* contig; uint32_t* left; uint32_t* len; uint32_t* peak; unsigned long long* candIndex;
  double* peakScore; double* kurt; double* corr; uint32_t* counts; uint32_t* wsum;
};

__global__ void emit_rows_kernel (EmitArgs a) {
  const uint32_t n1 = (uint32_t)a.g->n1;
  bool over = false;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n1; i += gridDim.x * blockDim.x) {
    if (!a.pass[i]) continue;
    const uint32_t ci = a.list[i];
    const Cand2 c = a.cand[ci];
    const unsigned long long o = (unsigned long long)(a.rowShift[c.unit] + (long long)a.passScan[i]);
    if (o >= a.rowCap) { over = true; continue; }
    long long candShift = 0;
    if (a.affOfUnit) { const int32_t k = a.affOfUnit[c.unit]; if (k >= 0) candShift = (long long)a.edgeCand[k] - (long long)a.edgeDrop[k]; }
    a.buffer[o] = (uint8_t)a.units[c.unit].buffer;
    a.contig[o] = a.units[c.unit].contig;
    a.left[o] = c.left; a.len[o] = c.len; a.peak[o] = c.peak; a.wsum[o] = c.wsum;
    a.candIndex[o] = (unsigned long long)((long long)((unsigned long long)ci - a.unitFirst[c.unit]) + candShift);
    a.peakScore[o] = c.score;
    a.kurt[o] = a.statKurt[i];
    a.corr[o] = a.rowCorr ? a.rowCorr[i] : __longlong_as_double(0x7ff8000000000000ll);
  }
  if (over) atomicOr(&a.g->flags, FLAG_CAP_OVERFLOW);
}

// the regions x samples matrix, one thread per ELEMENT: consecutive threads write consecutive words of the (mapped host) table, so a
// warp's stores leave as whole 128-byte lines — one thread per row wrote 32 scattered 4-byte words per instruction, each its own
// PCIe transaction (measured: 3.5 ms for 180 K rows x 16 samples instead of 0.2)
__global__ void emit_counts_kernel (EmitArgs a) {
  const unsigned long long total = (unsigned long long)a.g->n1 * a.nSamples;
  for (unsigned long long gI = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gI < total; gI += (unsigned long long)gridDim.x * blockDim.x) {
    const uint32_t i = (uint32_t)(gI / a.nSamples);
    if (!a.pass[i]) continue;
    const unsigned long long o = (unsigned long long)(a.rowShift[a.cand[a.list[i]].unit] + (long long)a.passScan[i]);
    if (o >= a.rowCap) continue;                                      // reported by emit_rows_kernel
    a.counts[o * a.nSamples + (gI - (unsigned long long)i * a.nSamples)] = a.statCounts[gI];
  }
}

// rows of the left-edge replay -> their slots in the table (they precede span_kernel's rows of the same unit)
struct EdgeEmitArgs {
  const EdgeRow* rows; const uint32_t* rowCounts;
  uint32_t rowCap, nSamples, nAffected;
  const uint32_t* affected; const uint32_t* nRows;
  const unsigned long long* edgeRowBase;   // [nUnits] first table row of the unit's replay rows (row_place_kernel)
  const SpanUnit* units;
  unsigned long long tableCap;
  GroupCtr* g;
  uint8_t* buffer; uint32_t* contig; uint32_t* left; uint32_t* len; uint32_t* peak; unsigned long long* candIndex;
  double* peakScore; double* kurt; double* corr; uint32_t* counts; uint32_t* wsum;
};

__global__ void edge_emit_kernel (EdgeEmitArgs a) {
  const uint32_t ai = blockIdx.y, k = blockIdx.x * blockDim.x + threadIdx.x;
  if (ai >= a.nAffected || k >= a.nRows[ai]) return;
  const EdgeRow r = a.rows[(size_t)ai * a.rowCap + k];
  const uint32_t u = a.affected[ai];
  const SpanUnit U = a.units[u];
  const unsigned long long o = a.edgeRowBase[u] + k;
  if (o >= a.tableCap) { atomicOr(&a.g->flags, FLAG_CAP_OVERFLOW); return; }
  a.buffer[o] = (uint8_t)U.buffer; a.contig[o] = U.contig;
  a.left[o] = r.left; a.len[o] = r.len; a.peak[o] = r.peak; a.wsum[o] = r.wsum; a.candIndex[o] = r.candIndex;
  a.peakScore[o] = r.score; a.kurt[o] = r.kurt; a.corr[o] = r.corr;
  for (uint32_t s = 0; s < a.nSamples; ++s) a.counts[o * a.nSamples + s] = a.rowCounts[((size_t)ai * a.rowCap + k) * a.nSamples + s];
}

// per-sample totals of the accepted rows (ProfileBuffer::nTagsInRegions, misc/peakcall.cpp:44) from the per-survivor matrix rows,
// plus the end-of-group bookkeeping: the call-wide counters take this group's numbers
struct GroupEndArgs {
  const uint32_t* statCounts; const uint32_t* pass; GroupCtr* g; CallCtr* call;
  uint32_t nSamples;
  const uint32_t* edgeRowCounts; const uint32_t* edgeNRows; uint32_t edgeRowCap, nAffected;   // left-edge replay rows count too
  unsigned long long edgeRowsTotal;
  unsigned long long* tir;          // [nSamples]
};

__global__ void group_end_kernel (GroupEndArgs a) {
  extern __shared__ unsigned long long sh[];
  const uint32_t n = a.nSamples;
  for (uint32_t s = threadIdx.x; s < n; s += blockDim.x) sh[s] = 0;
  __syncthreads();
  const unsigned long long total = (unsigned long long)a.g->n1 * n;
  for (unsigned long long gI = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; gI < total; gI += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = gI / n;
    if (a.pass[i]) { const uint32_t v = a.statCounts[gI]; if (v) atomicAdd(&sh[gI - i * n], (unsigned long long)v); }
  }
  if (blockIdx.x == 0)
    for (uint32_t ai = 0; ai < a.nAffected; ++ai) {
      const unsigned long long tot2 = (unsigned long long)a.edgeNRows[ai] * n;
      for (unsigned long long q = threadIdx.x; q < tot2; q += blockDim.x) atomicAdd(&sh[q % n], (unsigned long long)a.edgeRowCounts[(size_t)ai * a.edgeRowCap * n + q]);
    }
  __syncthreads();
  for (uint32_t s = threadIdx.x; s < n; s += blockDim.x) if (sh[s]) atomicAdd(&a.tir[s], sh[s]);
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    // (atomics: the groups of the two lanes of a call end independently of each other; CallCtr::rows is row_place_kernel's)
    GroupCtr* g = a.g; CallCtr* c = a.call;
    atomicAdd(&c->cands, g->nCand); atomicAdd(&c->runs, (unsigned long long)g->runReal);
    atomicAdd(&c->n1, g->pv >> PACK_SHIFT); atomicAdd(&c->cells, g->pv & PACK_MASK); atomicAdd(&c->warm, (unsigned long long)g->nWarm);
    atomicMax(&c->needSlots, (unsigned long long)g->runCount);
    atomicOr(&c->flags, g->flags);
  }
}

}  // namespace upb
