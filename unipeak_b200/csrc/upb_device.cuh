// upb_device.cuh — sm_100a kernels of the UniPeak region-calling hot path.
//
// Reference semantics (paths relative to the reference root): the density at position p of one
// ProfileBuffer is the sum, over the tag positions t within `bandwidth` of p IN ASCENDING ORDER, of
// round(K[p-t+bw] * w_t), each product and each sum rounded separately (misc/peakcall.cpp:203-209);
// a candidate region is a maximal run of positions with score >= regionThreshold
// (misc/peakcall.cpp:55-86); its peak is the first strict maximum (misc/data.cpp:92-102).
// All FP64 arithmetic that feeds a decision uses __dmul_rn/__dadd_rn (never FMA) in that order, so
// results are bit-identical to the reference's SSE2 build.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace upb {

constexpr int TILE = 2048;          // output positions per KDE work item
constexpr int TILE_BLOCKS = TILE / 32;
constexpr int KDE_THREADS = 256;
constexpr int KDE_WARPS = KDE_THREADS / 32;
constexpr int MAX_BW = 1024;        // halo must fit inside the neighbouring tile

struct Unit {                       // one (ProfileBuffer, contig) pair
  uint32_t contig;
  uint32_t buffer;                  // 0 forwardBuffer, 1 reverseBuffer
  uint32_t size;                    // contig size
  uint32_t nSpans;                  // ceil((size + bw) / SPAN_LEN): the reference evaluates up to lastTag+bw (Q20)
  uint32_t spanStart;               // first global span index of this unit
};

#ifndef UPB_MT_SHIFT
#define UPB_MT_SHIFT 9
#endif
constexpr int MT_SHIFT = UPB_MT_SHIFT;  // mini-tile: 512 positions, the granule of the tile-major copy.  256 (UPB_MT_SHIFT=8) is supported and passes the same
                                    //   tests, but buys nothing: the entries a piece reads sit in the peak itself, and the zone map doubles (profiles/r01_v27_ab.txt)
constexpr int SPAN_SHIFT = 12;      // span: 4096 positions = 8 mini-tiles; entries carry their offset inside the span
constexpr int SPAN_LEN = 1 << SPAN_SHIFT;
constexpr int SPAN_MT = 1 << (SPAN_SHIFT - MT_SHIFT);
constexpr int MAX_BW_SPAN = 448;    // largest bandwidth of the region-calling kernels (upb_span.cuh: MAX_BW_FAST)
constexpr int HALO_MT = (MAX_BW_SPAN + (1 << MT_SHIFT) - 1) >> MT_SHIFT;   // mini-tiles on either side of a span that a window of its positions can reach
static_assert(MT_SHIFT >= 7 && MT_SHIFT <= 10, "mini-tile: 128 .. 1024 positions");

struct Tags {                       // resident tag-frequency data (see upb_load_tags)
  const uint32_t* pos;              // sample-major arrays exactly as loaded (density kernel, validation)
  const uint32_t* cnt;
  // tile-major copy used by region calling: all samples' entries of one (strand, contig, mini-tile) are contiguous.
  //   ment[e].x = count, ment[e].y = (sample << 12) | ((pos - 1) & 4095);   order inside a mini-tile is arbitrary
  const uint2* ment;
  const unsigned long long* mtOff;  // [nMtTotal + 1] first entry of every mini-tile
  const uint32_t* mtBase;           // [2*C + 1] first mini-tile of (strand d, contig c) at d*C + c
  const unsigned long long* mtW;    // [nMtTotal] zone map: total count (all samples) per mini-tile — lets whole spans be skipped
  const float* blkW;                // [nMtTotal * 16] fine zone map: total count (all samples) per 32-position block, rounded UP to float
                                    //   (+inf where a 32-bit sum could have wrapped) — hot blocks are found without touching the entries
  const uint32_t* mEff;             // [E] counts with the control tags the reference forgets zeroed (Q4); NULL when no controls
  const uint64_t* off;              // [(s*2+d)*(C+1)+c]
  const uint32_t* tidx;             // per (s,d): tidxStride entries; per contig: nTiles0+3 entries
  const uint32_t* tidxBase;         // [C] offset of contig c inside one (s,d) index block
  uint64_t tidxStride;
  uint32_t C;
  uint32_t nSamples;
};

// flags word (device -> host)
enum : uint32_t { FLAG_RUN_OVERFLOW = 1u, FLAG_UNSORTED = 2u, FLAG_BAD_POS = 4u, FLAG_ZERO_COUNT = 8u, /* 16: FLAG_BAD_PACK (upb_pack.cuh) */
                  FLAG_CAP_OVERFLOW = 32u,    // more survivors / positions / rows than the arrays sized from the previous call hold: the call is repeated
                  FLAG_XCHG_TIMEOUT = 64u };  // a rank of a shared-table job never published its row counts

// ------------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t lane_id () { return threadIdx.x & 31; }
__device__ __forceinline__ uint32_t lanemask_lt () { uint32_t m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

// first index in [lo, hi) with p[idx] >= key
__device__ __forceinline__ uint64_t lower_bound_u32 (const uint32_t* __restrict__ p, uint64_t lo, uint64_t hi, uint32_t key) {
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(p + mid) < key) lo = mid + 1; else hi = mid; }
  return lo;
}
// first index in [lo, hi) with p[idx] > key
__device__ __forceinline__ uint64_t upper_bound_u32 (const uint32_t* __restrict__ p, uint64_t lo, uint64_t hi, uint32_t key) {
  while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (__ldg(p + mid) <= key) lo = mid + 1; else hi = mid; }
  return lo;
}

// Entries of segment (s, d, c) with position in [lo, hi] (inclusive, 1-based; lo may be <= 0).
// Uses the per-tile index to narrow the search to one tile's worth of entries at each end.
// Returns absolute entry indices [a, b).
__device__ __forceinline__ void seg_range (const Tags& T, uint32_t s, uint32_t d, uint32_t c, uint32_t nTilesIdx,
                                           int64_t lo, int64_t hi, uint64_t& a, uint64_t& b) {
  const uint64_t base = T.off[(uint64_t)(s * 2 + d) * (T.C + 1) + c];
  const uint32_t* ip = T.tidx + (uint64_t)(s * 2 + d) * T.tidxStride + T.tidxBase[c];
  if (hi < 1 || hi < lo) { a = b = base; return; }
  if (lo < 1) lo = 1;
  uint32_t tl = (uint32_t)((lo - 1) / TILE), th = (uint32_t)((hi - 1) / TILE);
  const uint32_t tmax = nTilesIdx + 1;            // index holds entries 0 .. nTilesIdx+2
  if (tl > tmax) tl = tmax;
  if (th > tmax) th = tmax;
  const uint64_t la = ip[tl], lb = ip[tl + 1];
  a = base + lower_bound_u32(T.pos + base, la, lb, (uint32_t)lo);
  const uint64_t ha = ip[th], hb = ip[th + 1];
  const uint32_t hik = hi > 0xFFFFFFFFll ? 0xFFFFFFFFu : (uint32_t)hi;
  b = base + upper_bound_u32(T.pos + base, ha, hb, hik);
  if (b < a) b = a;
}

// ------------------------------------------------------------------------------------------------
// per-tile search index (built once per upb_load_tags)
// idx[t] = number of entries of the segment with pos <= t*TILE, for t = 0 .. nTilesIdx+2
// ------------------------------------------------------------------------------------------------
struct IndexArgs {
  const uint32_t* pos;
  const uint32_t* cnt;
  const uint64_t* segStart;         // [nSeg+1] absolute start of segment g = (s*2+d)*C + c (monotone)
  const uint32_t* sizes;            // [C]
  const uint32_t* tidxBase;         // [C+1]
  uint32_t* tidx;
  uint64_t tidxStride;
  uint32_t C, nSeg;
  uint64_t nEntries;
  uint32_t* flags;
};

__global__ void index_fill_kernel (IndexArgs a, uint64_t total) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= total) return;
  const uint64_t sd = e / a.tidxStride;
  const uint32_t r = (uint32_t)(e % a.tidxStride);
  uint32_t lo = 0, hi = a.C;        // largest c with tidxBase[c] <= r
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.tidxBase[mid] <= r) lo = mid; else hi = mid; }
  const uint64_t g = sd * a.C + lo;
  a.tidx[e] = (uint32_t)(a.segStart[g + 1] - a.segStart[g]);
}

__global__ void index_build_kernel (IndexArgs a) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nEntries) return;
  uint32_t lo = 0, hi = a.nSeg;     // largest g with segStart[g] <= i (empty segments are skipped by taking the largest)
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.segStart[mid] <= i) lo = mid; else hi = mid; }
  const uint32_t g = lo;
  const uint32_t c = g % a.C;
  const uint64_t sd = g / a.C;
  const uint64_t li = i - a.segStart[g];
  const uint32_t p = a.pos[i];
  uint32_t f = 0;
  if (p < 1 || p > a.sizes[c]) f |= FLAG_BAD_POS;
  if (a.cnt[i] == 0) f |= FLAG_ZERO_COUNT;
  int64_t tq = -1;
  if (li > 0) { const uint32_t q = a.pos[i - 1]; if (q >= p) f |= FLAG_UNSORTED; tq = q ? (int64_t)((q - 1) / TILE) : -1; }
  if (f) { atomicOr(a.flags, f); if (f & (FLAG_BAD_POS | FLAG_UNSORTED)) return; }
  const int64_t tp = (int64_t)((p - 1) / TILE);
  uint32_t* ip = a.tidx + sd * a.tidxStride + a.tidxBase[c];
  for (int64_t t = tq + 1; t <= tp; ++t) ip[t] = (uint32_t)li;
}

// ------------------------------------------------------------------------------------------------
// tile-major copy (built once per upb_load_tags): a counting sort of all entries by (strand, contig, mini-tile)
// ------------------------------------------------------------------------------------------------
struct MtArgs {
  const uint32_t* pos;
  const uint32_t* cnt;
  const uint64_t* segStart;         // [nSeg+1]
  const uint32_t* mtBase;           // [2*C+1]
  uint32_t* mtCount;                // [nMtTotal] histogram, then per-tile cursors
  const unsigned long long* mtOff;  // [nMtTotal+1]
  uint2* ment;
  const uint32_t* sizes;            // [C]
  uint32_t* flags;                  // validation results (FLAG_*)
  unsigned long long* mtW;          // [nMtTotal] zone map per mini-tile (written by the scatter pass)
  float* blkW;                      // [nMtTotal * 16] zone map per 32-position block
  const uint32_t* winBase;          // [2*C+1] first window (WIN_MT mini-tiles) of (strand d, contig c)
  uint32_t* wEnt;                   // [nSamples * nWinTotal] first entry (relative to its segment) of sample s at or after each window
  uint32_t nWinTotal;
  uint32_t C, nSeg;
  uint64_t first;                   // entries [first, first + nEntries) are processed (pass 1 runs sample by sample behind the H2D copies)
  uint64_t nEntries;
};

__device__ __forceinline__ uint32_t seg_of_entry (const uint64_t* __restrict__ segStart, uint32_t nSeg, uint64_t i) {
  uint32_t lo = 0, hi = nSeg;       // largest g with segStart[g] <= i (empty segments are skipped by taking the largest)
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (segStart[mid] <= i) lo = mid; else hi = mid; }
  return lo;
}

constexpr int WIN_SHIFT = 16;       // scatter window: 65536 positions = 128 mini-tiles, one CTA of mt_scatter_kernel
constexpr int WIN_MT = 1 << (WIN_SHIFT - MT_SHIFT);

// Pass 1 (validation, per-mini-tile histogram, per-sample window index) is load_tile_kernel / load_seam_kernel in upb_pack.cuh, fused
// with the decoder of the compact host form; pass 2 (mt_scatter_kernel) scatters and builds the zone maps.

// Pass 2 of the counting sort, window-major: one CTA per 65536-position window of one (strand, contig) moves the entries of
// ALL samples that fall into it (one warp per sample at a time; each sample's entries of the window are one contiguous range of
// its sorted segment, found through the window index) to their mini-tiles.  The CTA's output is one contiguous stretch of the
// tile-major copy (~28 KB on C3), so the 8-byte stores merge into full sectors in L2, and the per-mini-tile cursors are
// shared-memory atomics — the sample-major version (v1–v19) did 326 M global atomics and partial-sector writes: 5.3 ms.
__global__ void __launch_bounds__(512) mt_scatter_kernel (MtArgs a, uint32_t nSamples, uint32_t w0 /* first window of the load group */) {
  __shared__ uint32_t cursor[WIN_MT];
  __shared__ unsigned long long base[WIN_MT];
  __shared__ uint32_t big[WIN_MT];                                    // a count large enough that a 32-bit block sum could wrap
  __shared__ uint32_t blk[WIN_MT * (1 << (MT_SHIFT - 5))];            // zone map per 32-position block
  if (*a.flags & (FLAG_UNSORTED | FLAG_BAD_POS)) return;             // pass 1 found invalid input: its entries are never used as indices
  const uint32_t gw = w0 + blockIdx.x;
  uint32_t lo = 0, hi = 2 * a.C;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.winBase[mid] <= gw) lo = mid; else hi = mid; }
  const uint32_t d = lo / a.C, c = lo % a.C, w = gw - a.winBase[lo];
  const uint32_t gb = a.mtBase[lo], nMt = a.mtBase[lo + 1] - gb;
  const uint32_t m0 = w * WIN_MT, nHere = min((uint32_t)WIN_MT, nMt - m0);
  const bool lastWin = gw + 1 == a.winBase[lo + 1];
  for (uint32_t k = threadIdx.x; k < (uint32_t)WIN_MT; k += blockDim.x) { cursor[k] = 0; big[k] = 0; base[k] = k < nHere ? a.mtOff[gb + m0 + k] : 0ull; }
  for (uint32_t k = threadIdx.x; k < (uint32_t)(WIN_MT << (MT_SHIFT - 5)); k += blockDim.x) blk[k] = 0;
  __syncthreads();
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t kBig = 0xFFFFFFFFu / (32u * nSamples);               // a block holds at most 32 * nSamples entries of one strand
  for (uint32_t s = warp; s < nSamples; s += blockDim.x / 32) {
    const uint64_t g = (uint64_t)(s * 2 + d) * a.C + c;
    const uint64_t seg = a.segStart[g];
    const uint32_t* we = a.wEnt + (size_t)s * a.nWinTotal + gw;
    const uint64_t e0 = seg + we[0], e1 = lastWin ? a.segStart[g + 1] : seg + we[1];
    for (uint64_t eb = e0; eb < e1; eb += 128) {                      // four loads of each array in flight per lane (one at a time: 64 % of the stall samples on them, r02)
      uint32_t pp[4], kk[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) { const uint64_t e = eb + lane + 32 * q; const bool ok = e < e1; pp[q] = ok ? __ldg(a.pos + e) : 0u; kk[q] = ok ? __ldg(a.cnt + e) : 0u; }
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        if (!pp[q]) continue;                                         // beyond the range (validated positions are >= 1)
        const uint32_t p = pp[q] - 1, k = kk[q];
        const uint32_t lm = (p >> MT_SHIFT) - m0;
        if (lm >= nHere) continue;                                    // cannot happen on validated input
        const unsigned long long slot = base[lm] + atomicAdd(&cursor[lm], 1u);
        a.ment[slot] = make_uint2(k, (s << SPAN_SHIFT) | (p & (SPAN_LEN - 1)));
        if (k >= kBig) big[lm] = 1;
        atomicAdd(&blk[(lm << (MT_SHIFT - 5)) | ((p >> 5) & ((1u << (MT_SHIFT - 5)) - 1))], k);
      }
    }
  }
  // zone maps of the window: mtW[m] = total count of the mini-tile, blkW[16m + j] = total count of its j-th 32-position block as a
  // float rounded up (+inf where a 32-bit block sum could have wrapped).  Both only ever feed upper bounds (a span / block whose
  // bound stays below the threshold is never evaluated), so rounding up is safe.
  __syncthreads();
  for (uint32_t k = threadIdx.x; k < nHere; k += blockDim.x) {
    unsigned long long t = 0;
    for (uint32_t j = 0; j < (1u << (MT_SHIFT - 5)); ++j) t += blk[(k << (MT_SHIFT - 5)) + j];
    a.mtW[gb + m0 + k] = big[k] ? (1ull << 56) : t;
  }
  for (uint32_t k = threadIdx.x; k < (nHere << (MT_SHIFT - 5)); k += blockDim.x)
    a.blkW[((size_t)(gb + m0) << (MT_SHIFT - 5)) + k] = big[k >> (MT_SHIFT - 5)] ? __int_as_float(0x7f800000) : __uint2float_ru(blk[k]);
}

constexpr int BLK_PER_MT = 1 << (MT_SHIFT - 5);

// control filter: a control sample's tag is remembered by the reference only when the same add()
// call carried a non-zero non-control sum, i.e. when a non-control sample has a tag at the same (strand,
// contig, position) (misc/peakcall.cpp:203-220).  mEff = count with the others zeroed.  One thread per mini-tile.
__global__ void mt_control_kernel (const uint2* __restrict__ ment, const unsigned long long* __restrict__ mtOff, uint32_t nMt,
                                   const uint8_t* __restrict__ control, uint32_t* __restrict__ mEff) {
  const uint32_t m = blockIdx.x * blockDim.x + threadIdx.x;
  if (m >= nMt) return;
  const unsigned long long e0 = mtOff[m], e1 = mtOff[m + 1];
  for (unsigned long long e = e0; e < e1; ++e) {
    const uint2 x = ment[e];
    uint32_t v = x.x;
    if (control[x.y >> SPAN_SHIFT]) {
      bool found = false;
      const uint32_t rel = x.y & (SPAN_LEN - 1);
      for (unsigned long long f = e0; f < e1 && !found; ++f) { const uint32_t y = ment[f].y; found = (y & (SPAN_LEN - 1)) == rel && !control[y >> SPAN_SHIFT]; }
      if (!found) v = 0;
    }
    mEff[e] = v;
  }
}

// ------------------------------------------------------------------------------------------------
// density kernel: exact F/R densities of selected position ranges, written to global memory.  Used where
// the reference itself needs every value: strandCorr (misc/data.cpp:184-193), the strand-shift sweep
// (src/strand_shift.cpp:211-218) and the -w profile (misc/peakcall.cpp:81-84).  Region calling never
// goes through here (upb_span.cuh does not materialise the profile).
// ------------------------------------------------------------------------------------------------
struct DensArgs {
  Tags T;
  const Unit* units;
  const uint32_t* nTilesIdx;        // [C] tiles covered by the search index of contig c
  const uint32_t* srcSample;        // non-control sample ids (they alone feed the density)
  uint32_t nSrc;
  uint32_t bw;
  const double* taps;               // 2bw+1 (host-computed, misc/kernel.cpp:16-35)
  const double* srcCoeff;           // [nSrc] -z scaling coefficients of the sources, NULL otherwise
  uint32_t* ticket;
  const uint32_t* dumpUnit;         // [nDump] unit of each range
  const uint32_t* dumpLeft;         // first position
  const uint32_t* dumpLen;
  const uint64_t* dumpOff;          // offset into dumpF/dumpR
  const uint64_t* dumpChunkStart;   // [nDump+1] prefix of ceil(len/TILE)
  uint32_t nDump;
  double* dumpF;
  double* dumpR;
};

struct DensSmem { uint32_t taps, wdF, wdR, wfF, wfR, ranges, total; };

__host__ __device__ inline DensSmem dens_smem_layout (uint32_t bw, uint32_t nRanges, bool nondir, bool coeff) {
  DensSmem L;
  const uint32_t W = TILE + 2 * bw + 4;
  uint32_t o = 0;
  L.taps = o;   o += ((2 * bw + 2) & ~1u) * 8;
  L.wfF = o;    o += coeff ? W * 8 : 0;
  L.wfR = o;    o += (coeff && nondir) ? W * 8 : 0;
  L.wdF = o;    o += W * 4;
  L.wdR = o;    o += nondir ? W * 4 : 0;
  o = (o + 15) & ~15u;
  L.ranges = o; o += nRanges * 16;
  L.total = o;
  return L;
}

template <bool NONDIR>
__global__ void __launch_bounds__(KDE_THREADS) density_kernel (const __grid_constant__ DensArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t s_ticket;
  const uint32_t bw = a.bw;
  const uint32_t W = TILE + 2 * bw;
  const uint32_t nStr = NONDIR ? 2u : 1u;
  const uint32_t nRanges = a.nSrc * nStr;
  const bool coeff = a.srcCoeff != nullptr;
  const DensSmem L = dens_smem_layout(bw, nRanges, NONDIR, coeff);
  double* Ks = (double*)(smem + L.taps);
  double* wfF = (double*)(smem + L.wfF);
  double* wfR = (double*)(smem + L.wfR);
  uint32_t* wdF = (uint32_t*)(smem + L.wdF);
  uint32_t* wdR = (uint32_t*)(smem + L.wdR);
  ulonglong2* ranges = (ulonglong2*)(smem + L.ranges);
  const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  for (uint32_t i = tid; i < 2 * bw + 1; i += KDE_THREADS) Ks[i] = a.taps[i];
  const uint32_t totalWork = (uint32_t)a.dumpChunkStart[a.nDump];

  for (;;) {
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t work = s_ticket;
    if (work >= totalWork) break;
    uint32_t lo = 0, hi = a.nDump;
    while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.dumpChunkStart[mid] <= work) lo = mid; else hi = mid; }
    const uint32_t r = lo;
    const uint32_t chunk = work - (uint32_t)a.dumpChunkStart[r];
    const Unit U = a.units[a.dumpUnit[r]];
    const uint32_t ts = a.dumpLeft[r] + chunk * TILE;
    const uint32_t np = min((uint32_t)TILE, a.dumpLen[r] - chunk * TILE);
    const uint64_t dumpBase = a.dumpOff[r] + (uint64_t)chunk * TILE;
    const int64_t cell0 = (int64_t)ts - bw;                           // position of cell 0

    uint32_t myCount = 0;
    for (uint32_t j = tid; j < nRanges; j += KDE_THREADS) {
      const uint32_t s = a.srcSample[j / nStr];
      const uint32_t d = NONDIR ? (j % nStr) : U.buffer;
      uint64_t x, y;
      seg_range(a.T, s, d, U.contig, a.nTilesIdx[U.contig], cell0, (int64_t)ts + np - 1 + bw, x, y);
      ranges[j] = make_ulonglong2(x, y - x);
      myCount += (uint32_t)(y - x);
    }
    for (uint32_t i = tid; i < W + 1; i += KDE_THREADS) { wdF[i] = 0; if (NONDIR) wdR[i] = 0; }
    const uint32_t any = __syncthreads_or(myCount != 0);
    if (!any) {                                                      // nothing within bw of this chunk
      for (uint32_t i = tid; i < np; i += KDE_THREADS) { a.dumpF[dumpBase + i] = 0.0; a.dumpR[dumpBase + i] = 0.0; }
      continue;
    }
    if (coeff) {
      // -z (misc/peakcall.cpp:191-200): weight = sum_s c_s*coeff_s, then + sum_s c_s, in sample order, every step rounded.
      // One source range at a time (a range has at most one entry per position, so there are no conflicts).
      for (uint32_t i = tid; i < W + 1; i += KDE_THREADS) { wfF[i] = 0.0; if (NONDIR) wfR[i] = 0.0; }
      __syncthreads();
      for (int phase = 0; phase < 2; ++phase)
        for (uint32_t j = 0; j < nRanges; ++j) {
          const ulonglong2 rg = ranges[j];
          double* wf = (NONDIR && (j & 1)) ? wfR : wfF;
          const double cf = a.srcCoeff[j / nStr];
          for (uint64_t i = tid; i < rg.y; i += KDE_THREADS) {
            const uint32_t cell = (uint32_t)((int64_t)__ldg(a.T.pos + rg.x + i) - cell0);
            const double c = (double)__ldg(a.T.cnt + rg.x + i);
            wf[cell] = __dadd_rn(wf[cell], phase ? c : __dmul_rn(c, cf));
          }
          __syncthreads();
        }
      for (uint32_t po = tid; po < np; po += KDE_THREADS) {
        double sF = 0.0, sR = 0.0;
        const uint32_t n2 = 2 * bw;
        for (uint32_t k = 0; k <= n2; ++k) {
          const double w = wfF[po + k];
          if (w != 0.0) sF = __dadd_rn(sF, __dmul_rn(Ks[n2 - k], w));
          if (NONDIR) { const double v = wfR[po + k]; if (v != 0.0) sR = __dadd_rn(sR, __dmul_rn(Ks[n2 - k], v)); }
        }
        const bool rev = (!NONDIR) && U.buffer == 1;
        a.dumpF[dumpBase + po] = rev ? 0.0 : sF;
        a.dumpR[dumpBase + po] = NONDIR ? sR : (rev ? sF : 0.0);
      }
      continue;
    }
    // pool the samples: dense per-position weight (integer sum of non-control counts)
    for (uint32_t j = warp; j < nRanges; j += KDE_WARPS) {
      const ulonglong2 rg = ranges[j];
      uint32_t* wd = (NONDIR && (j & 1)) ? wdR : wdF;
      for (uint64_t i = lane; i < rg.y; i += 32) {
        const uint32_t p = __ldg(a.T.pos + rg.x + i);
        const uint32_t c = __ldg(a.T.cnt + rg.x + i);
        atomicAdd(&wd[(uint32_t)((int64_t)p - cell0)], c);
      }
    }
    __syncthreads();
    // exact-order evaluation (misc/peakcall.cpp:203-209); cells without tags add +0.0
    for (uint32_t po = tid; po < np; po += KDE_THREADS) {
      double sF = 0.0, sR = 0.0;
      const uint32_t n2 = 2 * bw;
      for (uint32_t k = 0; k <= n2; ++k) {
        const uint32_t w = wdF[po + k];
        if (w) sF = __dadd_rn(sF, __dmul_rn(Ks[n2 - k], (double)w));
        if (NONDIR) { const uint32_t v = wdR[po + k]; if (v) sR = __dadd_rn(sR, __dmul_rn(Ks[n2 - k], (double)v)); }
      }
      // a forward-strand buffer in directional mode has only F; reverseBuffer only R (src/regions.cpp:356-358)
      const bool rev = (!NONDIR) && U.buffer == 1;
      a.dumpF[dumpBase + po] = rev ? 0.0 : sF;
      a.dumpR[dumpBase + po] = NONDIR ? sR : (rev ? sF : 0.0);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// generic exclusive scan of uint32 -> uint64 (three small kernels; n up to a few 10^8)
// ------------------------------------------------------------------------------------------------
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 8;
constexpr int SCAN_BLOCK = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ unsigned long long scan_block_excl_u64 (unsigned long long v, unsigned long long* sh /*[9]*/, unsigned long long* total) {
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  unsigned long long inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const unsigned long long n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += n; }
  if (lane == 31) sh[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) { unsigned long long run = 0; for (int w = 0; w < SCAN_THREADS / 32; ++w) { const unsigned long long t = sh[w]; sh[w] = run; run += t; } sh[SCAN_THREADS / 32] = run; }
  __syncthreads();
  const unsigned long long res = sh[warp] + inc - v;
  *total = sh[SCAN_THREADS / 32];
  __syncthreads();
  return res;
}

__global__ void scan_partial_kernel (const uint32_t* __restrict__ in, uint64_t n, unsigned long long* __restrict__ partial) {
  __shared__ unsigned long long sh[SCAN_THREADS / 32 + 1];
  const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_BLOCK;
  unsigned long long s = 0;
  for (int k = 0; k < SCAN_ITEMS; ++k) { const uint64_t i = b0 + (uint64_t)k * SCAN_THREADS + threadIdx.x; if (i < n) s += in[i]; }
  unsigned long long total;
  scan_block_excl_u64(s, sh, &total);
  if (threadIdx.x == 0) partial[blockIdx.x] = total;
}

__global__ void scan_top_kernel (unsigned long long* partial, uint32_t nBlocks, unsigned long long* grandTotal) {
  __shared__ unsigned long long sh[SCAN_THREADS / 32 + 1];
  __shared__ unsigned long long carry;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint32_t b0 = 0; b0 < nBlocks; b0 += SCAN_THREADS) {
    const uint32_t i = b0 + threadIdx.x;
    const unsigned long long v = i < nBlocks ? partial[i] : 0ull;
    unsigned long long total;
    const unsigned long long ex = scan_block_excl_u64(v, sh, &total);
    if (i < nBlocks) partial[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grandTotal = carry;
}

__global__ void scan_final_kernel (const uint32_t* __restrict__ in, uint64_t n, const unsigned long long* __restrict__ partial, unsigned long long* __restrict__ out) {
  __shared__ unsigned long long sh[SCAN_THREADS / 32 + 1];
  const uint64_t b0 = (uint64_t)blockIdx.x * SCAN_BLOCK + (uint64_t)threadIdx.x * SCAN_ITEMS;
  uint32_t v[SCAN_ITEMS];
  unsigned long long s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { const uint64_t i = b0 + k; v[k] = i < n ? in[i] : 0u; s += v[k]; }
  unsigned long long total;
  unsigned long long run = partial[blockIdx.x] + scan_block_excl_u64(s, sh, &total);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; ++k) { const uint64_t i = b0 + k; if (i < n) out[i] = run; run += v[k]; }
}

// ------------------------------------------------------------------------------------------------
// The same scan with the element count taken from DEVICE memory (or an immediate) and any grid: every kernel walks its tiles
// with a grid stride, so the host can enqueue the whole region pipeline without reading a single size back
// (four synchronising read-backs per call).  The input value comes from a functor: a plain u32 array,
// or two arrays packed into one 64-bit word (flag << 40 | length) so that one scan yields both prefixes.
// ------------------------------------------------------------------------------------------------
struct DevN {                       // a count the host may not know: *p32, *p64, or imm
  unsigned long long imm;
  const uint32_t* p32;
  const unsigned long long* p64;
  __device__ __forceinline__ unsigned long long get () const { return p32 ? (unsigned long long)*p32 : (p64 ? *p64 : imm); }
};
static inline DevN devn_imm (unsigned long long v) { DevN n; n.imm = v; n.p32 = nullptr; n.p64 = nullptr; return n; }
static inline DevN devn_u32 (const uint32_t* p) { DevN n; n.imm = 0; n.p32 = p; n.p64 = nullptr; return n; }
static inline DevN devn_u64 (const unsigned long long* p) { DevN n; n.imm = 0; n.p32 = nullptr; n.p64 = p; return n; }

constexpr int PACK_SHIFT = 40;      // packed scans: (flag << 40) | length; lengths sum to < 2^40 (a genome has < 2^33 positions per strand)
constexpr unsigned long long PACK_MASK = (1ull << PACK_SHIFT) - 1;

struct LoadU32 {
  const uint32_t* a;
  __device__ __forceinline__ unsigned long long operator() (unsigned long long i) const { return a[i]; }
};
struct LoadPacked {                 // flag[i] << 40 | len[i]
  const uint32_t* flag; const uint32_t* len;
  __device__ __forceinline__ unsigned long long operator() (unsigned long long i) const { return ((unsigned long long)flag[i] << PACK_SHIFT) | len[i]; }
};

template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) scanv_partial_kernel (Load ld, DevN nn, unsigned long long* __restrict__ partial) {
  __shared__ unsigned long long sh[SCAN_THREADS / 32 + 1];
  const unsigned long long n = nn.get();
  const unsigned long long nTiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  for (unsigned long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const unsigned long long b0 = tile * SCAN_BLOCK;
    unsigned long long s = 0;
    for (int k = 0; k < SCAN_ITEMS; ++k) { const unsigned long long i = b0 + (unsigned long long)k * SCAN_THREADS + threadIdx.x; if (i < n) s += ld(i); }
    unsigned long long total;
    scan_block_excl_u64(s, sh, &total);
    if (threadIdx.x == 0) partial[tile] = total;
  }
}

// one block; also publishes the grand total (0 when there is nothing to scan)
__global__ void __launch_bounds__(SCAN_THREADS) scanv_top_kernel (unsigned long long* partial, DevN nn, unsigned long long* grandTotal) {
  __shared__ unsigned long long sh[SCAN_THREADS / 32 + 1];
  __shared__ unsigned long long carry;
  const unsigned long long n = nn.get();
  const unsigned long long nTiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (unsigned long long b0 = 0; b0 < nTiles; b0 += SCAN_THREADS) {
    const unsigned long long i = b0 + threadIdx.x;
    const unsigned long long v = i < nTiles ? partial[i] : 0ull;
    unsigned long long total;
    const unsigned long long ex = scan_block_excl_u64(v, sh, &total);
    if (i < nTiles) partial[i] = carry + ex;
    __syncthreads();
    if (threadIdx.x == 0) carry += total;
    __syncthreads();
  }
  if (threadIdx.x == 0) *grandTotal = carry;
}

template <class Load>
__global__ void __launch_bounds__(SCAN_THREADS) scanv_final_kernel (Load ld, DevN nn, const unsigned long long* __restrict__ partial, unsigned long long* __restrict__ out,
                                                                    unsigned long long bias) {
  __shared__ unsigned long long sh[SCAN_THREADS / 32 + 1];
  const unsigned long long n = nn.get();
  const unsigned long long nTiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  for (unsigned long long tile = blockIdx.x; tile < nTiles; tile += gridDim.x) {
    const unsigned long long b0 = tile * SCAN_BLOCK + (unsigned long long)threadIdx.x * SCAN_ITEMS;
    unsigned long long v[SCAN_ITEMS];
    unsigned long long s = 0;
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { const unsigned long long i = b0 + k; v[k] = i < n ? ld(i) : 0ull; s += v[k]; }
    unsigned long long total;
    unsigned long long run = bias + partial[tile] + scan_block_excl_u64(s, sh, &total);
#pragma unroll
    for (int k = 0; k < SCAN_ITEMS; ++k) { const unsigned long long i = b0 + k; if (i < n) out[i] = run; run += v[k]; }
  }
}

// ------------------------------------------------------------------------------------------------
// small generic kernels
// ------------------------------------------------------------------------------------------------
__global__ void compact_index_kernel (const uint32_t* __restrict__ pass, const unsigned long long* __restrict__ scan, uint32_t n, uint32_t* __restrict__ list) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pass[i]) list[scan[i]] = i;
}

constexpr int KURT_CHUNK = 1024;
constexpr int KURT_WARPS = 8;

// Pearson correlation exactly as misc/data.cpp:22-58 computes it: sequential sums.  One warp per
// (region, shift); lanes fetch 32 values at a time and the additions are chained in index order.
__device__ __forceinline__ double warp_corr (const double* __restrict__ f, const double* __restrict__ r, uint32_t m) {
  const uint32_t lane = lane_id();
  double s1 = 0.0, s2 = 0.0;
  for (uint32_t g = 0; g < m; g += 32) {
    const uint32_t k = min(32u, m - g);
    const double x = (g + lane < m) ? f[g + lane] : 0.0, y = (g + lane < m) ? r[g + lane] : 0.0;
    for (uint32_t b = 0; b < k; ++b) { s1 = __dadd_rn(s1, __shfl_sync(0xffffffffu, x, b)); s2 = __dadd_rn(s2, __shfl_sync(0xffffffffu, y, b)); }
  }
  const double mean1 = __ddiv_rn(s1, (double)m), mean2 = __ddiv_rn(s2, (double)m);
  double q1 = 0.0, q2 = 0.0, q12 = 0.0;
  for (uint32_t g = 0; g < m; g += 32) {
    const uint32_t k = min(32u, m - g);
    const double x = (g + lane < m) ? __dsub_rn(f[g + lane], mean1) : 0.0, y = (g + lane < m) ? __dsub_rn(r[g + lane], mean2) : 0.0;
    const double xx = __dmul_rn(x, x), yy = __dmul_rn(y, y), xy = __dmul_rn(x, y);
    for (uint32_t b = 0; b < k; ++b) {
      q1 = __dadd_rn(q1, __shfl_sync(0xffffffffu, xx, b));
      q2 = __dadd_rn(q2, __shfl_sync(0xffffffffu, yy, b));
      q12 = __dadd_rn(q12, __shfl_sync(0xffffffffu, xy, b));
    }
  }
  const double m1 = __dsub_rn((double)m, 1.0);
  const double sd1 = __dsqrt_rn(__ddiv_rn(q1, m1)), sd2 = __dsqrt_rn(__ddiv_rn(q2, m1));
  return __ddiv_rn(q12, __dmul_rn(__dmul_rn(m1, sd1), sd2));
}

// strandCorr(shift) (misc/data.cpp:184-193) for a list of regions whose F/R densities were dumped
struct CorrArgs {
  const double* F;
  const double* R;
  const uint64_t* off;              // [n] offset of region i in F/R
  const uint32_t* len;              // [n]
  uint32_t n;
  uint32_t nShifts;                 // shifts 0 .. nShifts-1
  double* corr;                     // [n * nShifts]
};

__global__ void strand_corr_kernel (CorrArgs a) {
  const uint64_t w = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= (uint64_t)a.n * a.nShifts) return;
  const uint32_t i = (uint32_t)(w / a.nShifts), sh = (uint32_t)(w % a.nShifts);
  const uint32_t len = a.len[i];
  double c;
  if (len > 2 * sh + 3) c = warp_corr(a.F + a.off[i], a.R + a.off[i] + 2 * sh, len - 2 * sh);
  else c = __longlong_as_double(0x7ff8000000000000ll);
  if (lane_id() == 0) a.corr[w] = c;
}

// per region: first strict maximum over shifts (src/strand_shift.cpp:209-218)
__global__ void best_shift_kernel (const double* __restrict__ corr, uint32_t n, uint32_t nShifts, uint32_t* __restrict__ bestShift, double* __restrict__ bestCorr) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t bs = 0;
  double bc = -1.0;
  for (uint32_t s = 0; s < nShifts; ++s) { const double c = corr[(uint64_t)i * nShifts + s]; if (c > bc) { bc = c; bs = s; } }
  bestShift[i] = bs;
  bestCorr[i] = bc;
}

// ------------------------------------------------------------------------------------------------
// ------------------------------------------------------------------------------------------------
// bin/tags_in_regions (src/tags_in_regions.cpp:131-199).  The reference walks every extra sample's stream once,
// with a cursor that only moves forward: per region it SKIPS entries until one has the region's strand and lies at
// or after the region's left edge (a strand mismatch also skips, SURVEY Q13 i), then COUNTS entries while they are
// on the region's contig at or before its right edge — without looking at the strand (Q13 ii).
//
// Cursor positions are handled as keys in the stream's order: directional streams are strand-major
// (misc/format.cpp:503-565), key = (strand, contig, pos); NondirParseAlignStream merges the strands by position,
// forward first on ties (misc/format.cpp:832-850), key = (contig, pos, strand).  For a cursor at or before the
// region's own lower key L the outcome does not depend on the cursor, so pass 1 evaluates every (region, sample)
// independently from L and reports where the cursor ends up (E); the cursor seen by region k is then the running
// maximum of the earlier E (pass 2, one block per sample), and pass 3 re-evaluates only the rare pairs whose cursor
// had already overtaken L (overlapping or unsorted rows, or a '+' row met after the stream reached its '-' track).
// ------------------------------------------------------------------------------------------------
struct TirRegion { uint32_t contig, left, right, strand; };          // strand 0 '+', 1 '-'

__device__ __forceinline__ unsigned long long tir_key (bool directional, uint32_t d, uint32_t c, uint64_t p) {
  return directional ? (((unsigned long long)d << 62) | ((unsigned long long)c << 33) | p) : (((unsigned long long)c << 34) | (p << 1) | d);
}
constexpr unsigned long long TIR_END = ~0ull;

// first entry index (absolute) of sample s, strand d with (contig, pos) >= (c, p); strand end if none
__device__ __forceinline__ uint64_t tir_lb (const Tags& T, uint32_t s, uint32_t d, uint32_t c, uint64_t p) {
  const uint64_t* o = T.off + (uint64_t)(s * 2 + d) * (T.C + 1);
  if (c >= T.C) return o[T.C];
  if (p > 0xFFFFFFFFull) return o[c + 1];
  return lower_bound_u32(T.pos, o[c], o[c + 1], (uint32_t)p);
}
__device__ __forceinline__ uint32_t tir_contig_of (const Tags& T, uint32_t s, uint32_t d, uint64_t idx) {
  const uint64_t* o = T.off + (uint64_t)(s * 2 + d) * (T.C + 1);
  uint32_t lo = 0, hi = T.C;                                         // largest c with o[c] <= idx
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (o[mid] <= idx) lo = mid; else hi = mid; }
  return lo;
}
__device__ __forceinline__ uint32_t tir_sum (const Tags& T, uint64_t a, uint64_t b) {
  uint32_t h = 0;
  for (uint64_t e = a; e < b; ++e) h += __ldg(T.cnt + e);             // HitCount wraps mod 2^32
  return h;
}

// one region, one sample, cursor at key bound `cur` (<= L means "not past the region's own lower key")
__device__ __forceinline__ void tir_eval (const Tags& T, bool directional, uint32_t s, const TirRegion& r, unsigned long long cur,
                                          uint32_t& hits, unsigned long long& E) {
  hits = 0;
  const uint32_t C = T.C;
  uint32_t bc = r.contig; uint64_t bp = r.left;                      // bound inside the region's strand: max(cur, L)
  if (cur > tir_key(directional, r.strand, r.contig, r.left)) {
    if (cur == TIR_END) { E = TIR_END; return; }
    if (directional) {
      const uint32_t dc = (uint32_t)(cur >> 62);
      if (dc > r.strand) { E = TIR_END; return; }                    // a '+' row after the cursor reached the '-' track: nothing matches any more
      bc = (uint32_t)((cur >> 33) & 0x1FFFFFFFu); bp = cur & 0x1FFFFFFFFull;
    } else {
      bc = (uint32_t)(cur >> 34); bp = ((cur >> 1) & 0x1FFFFFFFFull) + (cur & 1ull);   // past a '-' entry at p: the next '+' entry is beyond p
    }
  }
  const uint64_t* o = T.off + (uint64_t)(s * 2 + r.strand) * (C + 1);
  const uint64_t t = tir_lb(T, s, r.strand, bc, bp);                 // SKIP ends here
  if (t >= o[C]) { E = TIR_END; return; }
  const uint32_t tc = tir_contig_of(T, s, r.strand, t);
  const uint32_t tp = T.pos[t];
  if (tc != r.contig || tp > r.right) { E = tir_key(directional, r.strand, tc, tp); return; }
  // COUNT: same contig, position <= right, strand not examined
  const uint64_t a1 = upper_bound_u32(T.pos, t, o[r.contig + 1], r.right);
  hits = tir_sum(T, t, a1);
  E = tir_key(directional, r.strand, r.contig, (uint64_t)r.right + 1);
  if (directional) {
    if (r.strand == 0 && a1 == o[C]) {                                // ran off the end of the '+' track: the '-' track follows in the stream
      const uint64_t* om = T.off + (uint64_t)(s * 2 + 1) * (C + 1);
      const uint64_t b0 = om[0];
      if (b0 >= om[C]) E = TIR_END;                                    // no '-' track: the stream has ended
      else {
        const uint32_t c0 = tir_contig_of(T, s, 1, b0);
        if (c0 == r.contig && T.pos[b0] <= r.right) {
          const uint64_t b1 = upper_bound_u32(T.pos, b0, om[r.contig + 1], r.right);
          hits += tir_sum(T, b0, b1);
          E = tir_key(true, 1, r.contig, (uint64_t)r.right + 1);
        } else E = tir_key(true, 1, c0, T.pos[b0]);                     // the cursor rests on the first '-' entry: a later '+' row skips the whole '-' track (Q13 i)
      }
    }
  } else {
    // merged stream: every '-' entry from (contig, tp) on ('-' follows '+' at equal positions) up to right
    const uint64_t* om = T.off + (uint64_t)(s * 2 + 1) * (C + 1);
    const uint64_t b0 = lower_bound_u32(T.pos, om[r.contig], om[r.contig + 1], tp);
    const uint64_t b1 = upper_bound_u32(T.pos, b0, om[r.contig + 1], r.right);
    hits += tir_sum(T, b0, b1);
  }
}

struct TirArgs {
  Tags T;
  const TirRegion* regions;
  uint64_t nRegions;
  int directional;
  uint32_t* hits;                    // [nRegions * nSamples]
  unsigned long long* E;             // [nSamples * nRegions] (sample-major for the scan)
  unsigned long long* M;             // [nSamples * nRegions] exclusive running maximum of E
  unsigned long long* firstDead;     // [nSamples] first '+' row met with the cursor already in the '-' track: it and every later row count 0
};

// a directional '+' row whose cursor already sits in the '-' track skips to the end of the stream (Q13 i)
__global__ void tir_dead_kernel (TirArgs a) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = a.T.nSamples;
  if (g >= a.nRegions * n) return;
  const uint64_t k = g / n; const uint32_t s = (uint32_t)(g % n);
  if (a.regions[k].strand == 0 && (a.M[(uint64_t)s * a.nRegions + k] >> 62) != 0) atomicMin(&a.firstDead[s], (unsigned long long)k);
}

__global__ void tir_eval_kernel (TirArgs a, int fixup) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = a.T.nSamples;
  if (g >= a.nRegions * n) return;
  const uint64_t k = g / n; const uint32_t s = (uint32_t)(g % n);
  const TirRegion r = a.regions[k];
  unsigned long long cur = 0;
  if (fixup) {
    if (a.directional && k >= a.firstDead[s]) { a.hits[g] = 0; return; }
    cur = a.M[(uint64_t)s * a.nRegions + k];
    if (cur <= tir_key(a.directional != 0, r.strand, r.contig, r.left)) return;   // pass 1 already holds the answer
  }
  uint32_t h; unsigned long long E;
  tir_eval(a.T, a.directional != 0, s, r, cur, h, E);
  a.hits[g] = h;
  if (!fixup) a.E[(uint64_t)s * a.nRegions + k] = E;
}

// exclusive running maximum per sample: one block per sample walks its row in chunks
__global__ void tir_scan_kernel (const unsigned long long* __restrict__ E, unsigned long long* __restrict__ M, uint64_t nRegions) {
  __shared__ unsigned long long sh[33];
  __shared__ unsigned long long carry;
  const uint32_t s = blockIdx.x, lane = lane_id(), warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const unsigned long long* e = E + (uint64_t)s * nRegions;
  unsigned long long* m = M + (uint64_t)s * nRegions;
  if (threadIdx.x == 0) carry = 0;
  __syncthreads();
  for (uint64_t b0 = 0; b0 < nRegions; b0 += blockDim.x) {
    const uint64_t i = b0 + threadIdx.x;
    const unsigned long long v = i < nRegions ? e[i] : 0ull;
    unsigned long long inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned long long x = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc = max(inc, x); }
    if (lane == 31) sh[warp] = inc;
    __syncthreads();
    if (threadIdx.x == 0) { unsigned long long run = carry; for (uint32_t w = 0; w < nw; ++w) { const unsigned long long t = sh[w]; sh[w] = run; run = max(run, t); } sh[32] = run; }
    __syncthreads();
    unsigned long long ex = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane == 0) ex = 0;
    if (i < nRegions) m[i] = max(sh[warp], ex);
    __syncthreads();
    if (threadIdx.x == 0) carry = sh[32];
    __syncthreads();
  }
}

}  // namespace upb

namespace upb {
// bin/regions -w: positions of a dumped range whose score F + R is not zero (misc/peakcall.cpp:57,81-84), compacted on the device so
// that only what the wiggle will print crosses PCIe
__global__ void profile_flag_kernel (const double* __restrict__ F, const double* __restrict__ R, uint32_t len, uint32_t* __restrict__ flag) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len) flag[i] = __dadd_rn(F[i], R[i]) != 0.0 ? 1u : 0u;
}
__global__ void profile_emit_kernel (const double* __restrict__ F, const double* __restrict__ R, uint32_t len, uint32_t start, const uint32_t* __restrict__ flag,
                                     const unsigned long long* __restrict__ off, uint32_t* __restrict__ outPos, double* __restrict__ outScore) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < len && flag[i]) { outPos[off[i]] = start + i; outScore[off[i]] = __dadd_rn(F[i], R[i]); }
}
}  // namespace upb
