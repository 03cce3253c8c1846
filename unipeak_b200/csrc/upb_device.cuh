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
  uint32_t nTiles;                  // ceil((size + bw) / TILE): the reference evaluates up to lastTag+bw (Q20)
  uint32_t tileStart;               // first global tile index of this unit
};

struct RunRec {                     // a run of above-threshold positions inside one tile
  uint32_t left, len, peak, wsum;
  double score;                     // score at `peak`
};

struct Tags {                       // resident tag-frequency data (see upb_load_tags)
  const uint32_t* pos;
  const uint32_t* cnt;              // counts as loaded
  const uint32_t* cntEff;           // counts with control-only positions zeroed (Q4); == cnt when no controls
  const uint64_t* off;              // [(s*2+d)*(C+1)+c]
  const uint32_t* tidx;             // per (s,d): tidxStride entries; per contig: nTiles0+3 entries
  const uint32_t* tidxBase;         // [C] offset of contig c inside one (s,d) index block
  uint64_t tidxStride;
  uint32_t C;
  uint32_t nSamples;
};

// flags word (device -> host)
enum : uint32_t { FLAG_RUN_OVERFLOW = 1u, FLAG_UNSORTED = 2u, FLAG_BAD_POS = 4u, FLAG_ZERO_COUNT = 8u };

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
// control filter: a control sample's tag is remembered by the reference only when the same
// add() call carried a non-zero non-control sum, i.e. when a non-control sample has a tag at the same
// (strand, contig, position) (misc/peakcall.cpp:203-220).  cntEff = cnt with the others zeroed.
// ------------------------------------------------------------------------------------------------
struct ControlArgs {
  Tags T;
  const uint32_t* nTilesIdx;        // [C]
  const uint64_t* segStart;         // [nSeg+1]
  const uint8_t* control;           // [nSamples]
  const uint32_t* srcSample;        // non-control samples
  uint32_t nSrc;
  uint32_t* cntEff;
  uint32_t nSeg;
  uint64_t nEntries;
};

__global__ void control_filter_kernel (ControlArgs a) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.nEntries) return;
  uint32_t lo = 0, hi = a.nSeg;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.segStart[mid] <= i) lo = mid; else hi = mid; }
  const uint32_t c = lo % a.T.C;
  const uint32_t sd = lo / a.T.C;
  const uint32_t s = sd >> 1, d = sd & 1;
  uint32_t v = a.T.cnt[i];
  if (a.control[s]) {
    const uint32_t p = a.T.pos[i];
    bool found = false;
    for (uint32_t j = 0; j < a.nSrc && !found; ++j) {
      uint64_t x, y;
      seg_range(a.T, a.srcSample[j], d, c, a.nTilesIdx[c], p, p, x, y);
      found = y > x;
    }
    if (!found) v = 0;
  }
  a.cntEff[i] = v;
}

// ------------------------------------------------------------------------------------------------
// K1: fused pool + KDE + threshold + run extraction.  Also used (DUMP mode) to write exact F/R
// densities of selected position ranges for the strand-correlation statistics.
// ------------------------------------------------------------------------------------------------
struct KdeArgs {
  Tags T;
  const Unit* units;
  const uint32_t* nTilesIdx;        // [C] tiles covered by the search index of contig c
  uint32_t nUnits;
  uint32_t totalTiles;
  const uint32_t* srcSample;        // non-control sample ids (they alone feed the density)
  uint32_t nSrc;
  uint32_t bw;
  const double* taps;               // 2bw+1 (host-computed, misc/kernel.cpp:16-35)
  double thr;
  double kmaxBound;                 // max tap * (1 + 1e-9): upper bound factor for the cold-block test
  int nondir;
  // run mode outputs
  uint32_t* ticket;
  RunRec* runs;
  uint32_t runCap;
  uint32_t* runCount;
  uint32_t* tileRunBase;            // [totalTiles]
  uint32_t* tileRunCount;           // [totalTiles]
  uint32_t* flags;
  unsigned long long* statHotTiles;
  // dump mode inputs/outputs
  const uint32_t* dumpUnit;         // [nDump] unit of each range
  const uint32_t* dumpLeft;         // first position
  const uint32_t* dumpLen;
  const uint64_t* dumpOff;          // offset into dumpF/dumpR
  const uint64_t* dumpChunkStart;   // [nDump+1] prefix of ceil(len/TILE)
  uint32_t nDump;
  double* dumpF;
  double* dumpR;
  uint32_t* dumpNonzero;            // optional [sum len] : 1 where score != 0 (profile output), else NULL
};

struct KdeSmem {                    // byte offsets inside dynamic shared memory
  uint32_t taps, score, twF, twR, wdF, wdR, cltF, cltR, tpF, tpR, mask, hot, hotList, startPref, ranges, total;
};

__host__ __device__ inline KdeSmem kde_smem_layout (uint32_t bw, uint32_t nSrcStrands, bool nondir) {
  KdeSmem L;
  const uint32_t W = TILE + 2 * bw;
  const uint32_t Wp = (W + 8) & ~7u;               // cells, padded (+1 for the prefix sentinel)
  uint32_t o = 0;
  L.taps = o;   o += ((2 * bw + 1 + 1) & ~1u) * 8;
  L.score = o;  o += TILE * 8;
  L.twF = o;    o += Wp * 8;
  L.twR = o;    o += nondir ? Wp * 8 : 0;
  L.wdF = o;    o += Wp * 4;
  L.wdR = o;    o += nondir ? Wp * 4 : 0;
  L.cltF = o;   o += Wp * 2;
  L.cltR = o;   o += nondir ? Wp * 2 : 0;
  L.tpF = o;    o += Wp * 2;
  L.tpR = o;    o += nondir ? Wp * 2 : 0;
  o = (o + 15) & ~15u;
  L.mask = o;   o += TILE_BLOCKS * 4;
  L.hot = o;    o += TILE_BLOCKS * 4;
  L.hotList = o; o += TILE_BLOCKS * 4;
  L.startPref = o; o += (TILE_BLOCKS + 1) * 4;
  o = (o + 15) & ~15u;
  L.ranges = o; o += nSrcStrands * 16;
  L.total = o;
  return L;
}

// block-wide exclusive scan of one uint32 per thread (KDE_THREADS threads); returns the exclusive prefix, total in *total
__device__ __forceinline__ uint32_t block_excl_scan (uint32_t v, uint32_t* warpTotals /*[KDE_WARPS+1] smem*/, uint32_t* total) {
  const uint32_t lane = lane_id(), warp = threadIdx.x >> 5;
  uint32_t inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) { const uint32_t n = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += n; }
  if (lane == 31) warpTotals[warp] = inc;
  __syncthreads();
  if (threadIdx.x == 0) {
    uint32_t run = 0;
    for (int w = 0; w < KDE_WARPS; ++w) { const uint32_t t = warpTotals[w]; warpTotals[w] = run; run += t; }
    warpTotals[KDE_WARPS] = run;
  }
  __syncthreads();
  const uint32_t res = warpTotals[warp] + inc - v;
  *total = warpTotals[KDE_WARPS];
  __syncthreads();
  return res;
}

// dense cells -> sorted sparse tag list (tp = cell index, tw = weight as double) and clt[cell] = number
// of tags strictly before `cell`; clt[W] = number of tags
__device__ __forceinline__ void compact_cells (const uint32_t* __restrict__ wd, uint16_t* clt, uint16_t* tp, double* tw,
                                               uint32_t W, uint32_t* warpTotals) {
  const uint32_t cpt = (W + KDE_THREADS - 1) / KDE_THREADS;
  const uint32_t c0 = threadIdx.x * cpt;
  const uint32_t c1 = min(W, c0 + cpt);
  uint32_t n = 0;
  for (uint32_t c = c0; c < c1; ++c) n += (wd[c] != 0);
  uint32_t total;
  uint32_t run = block_excl_scan(n, warpTotals, &total);
  for (uint32_t c = c0; c < c1; ++c) {
    clt[c] = (uint16_t)run;
    const uint32_t w = wd[c];
    if (w) { tp[run] = (uint16_t)c; tw[run] = (double)w; ++run; }
  }
  if (threadIdx.x == 0) clt[W] = (uint16_t)total;
}

// exact-order density of the position whose own cell index is pc (tags at cells pc-bw .. pc+bw)
__device__ __forceinline__ double eval_exact (const double* __restrict__ Ks, const uint16_t* __restrict__ clt,
                                              const uint16_t* __restrict__ tp, const double* __restrict__ tw,
                                              uint32_t pc, uint32_t bw) {
  const uint32_t lo = clt[pc - bw], hi = clt[pc + bw + 1];
  double s = 0.0;
  for (uint32_t q = lo; q < hi; ++q) s = __dadd_rn(s, __dmul_rn(Ks[pc + bw - tp[q]], tw[q]));
  return s;
}

template <bool NONDIR, bool DUMP>
__global__ void __launch_bounds__(KDE_THREADS) kde_kernel (const __grid_constant__ KdeArgs a) {
  extern __shared__ __align__(16) unsigned char smem[];
  __shared__ uint32_t s_ticket, s_nHot, s_nRuns, s_base, s_any;
  __shared__ uint32_t s_warpTotals[KDE_WARPS + 1];

  const uint32_t bw = a.bw;
  const uint32_t W = TILE + 2 * bw;
  const uint32_t nStr = NONDIR ? 2u : 1u;
  const uint32_t nRanges = a.nSrc * nStr;
  const KdeSmem L = kde_smem_layout(bw, nRanges, NONDIR);
  double* Ks = (double*)(smem + L.taps);
  double* score = (double*)(smem + L.score);
  double* twF = (double*)(smem + L.twF);
  double* twR = (double*)(smem + L.twR);
  uint32_t* wdF = (uint32_t*)(smem + L.wdF);
  uint32_t* wdR = (uint32_t*)(smem + L.wdR);
  uint16_t* cltF = (uint16_t*)(smem + L.cltF);
  uint16_t* cltR = (uint16_t*)(smem + L.cltR);
  uint16_t* tpF = (uint16_t*)(smem + L.tpF);
  uint16_t* tpR = (uint16_t*)(smem + L.tpR);
  uint32_t* mask = (uint32_t*)(smem + L.mask);
  uint32_t* hot = (uint32_t*)(smem + L.hot);
  uint32_t* hotList = (uint32_t*)(smem + L.hotList);
  uint32_t* startPref = (uint32_t*)(smem + L.startPref);
  ulonglong2* ranges = (ulonglong2*)(smem + L.ranges);

  const uint32_t tid = threadIdx.x, lane = lane_id(), warp = tid >> 5;
  for (uint32_t i = tid; i < 2 * bw + 1; i += KDE_THREADS) Ks[i] = a.taps[i];

  const uint32_t totalWork = DUMP ? (uint32_t)a.dumpChunkStart[a.nDump] : a.totalTiles;

  for (;;) {
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(a.ticket, 1u);
    __syncthreads();
    const uint32_t work = s_ticket;
    if (work >= totalWork) break;

    // ---- locate the work item: unit, first position ts, number of positions np ------------------
    uint32_t ui, ts, np;
    uint64_t dumpBase = 0;
    if (!DUMP) {
      uint32_t lo = 0, hi = a.nUnits;
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.units[mid].tileStart <= work) lo = mid; else hi = mid; }
      ui = lo;
      const uint32_t t = work - a.units[ui].tileStart;
      ts = t * TILE + 1;
      const uint64_t lastPos = (uint64_t)a.units[ui].size + bw;      // Q20: evaluated through lastTag+bw, no clamp
      np = (uint32_t)min((uint64_t)TILE, lastPos - (ts - 1));
    } else {
      uint32_t lo = 0, hi = a.nDump;
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.dumpChunkStart[mid] <= work) lo = mid; else hi = mid; }
      const uint32_t r = lo;
      const uint32_t chunk = work - (uint32_t)a.dumpChunkStart[r];
      ui = a.dumpUnit[r];
      ts = a.dumpLeft[r] + chunk * TILE;
      np = min((uint32_t)TILE, a.dumpLen[r] - chunk * TILE);
      dumpBase = a.dumpOff[r] + (uint64_t)chunk * TILE;
    }
    const Unit U = a.units[ui];
    const int64_t cell0 = (int64_t)ts - bw;                           // position of cell 0

    // ---- A: entry ranges of every source (sample, strand) -----------------------------------------
    uint32_t myCount = 0;
    for (uint32_t j = tid; j < nRanges; j += KDE_THREADS) {
      const uint32_t s = a.srcSample[j / nStr];
      const uint32_t d = NONDIR ? (j % nStr) : U.buffer;
      uint64_t x, y;
      seg_range(a.T, s, d, U.contig, a.nTilesIdx[U.contig], cell0, (int64_t)ts + np - 1 + bw, x, y);
      ranges[j] = make_ulonglong2(x, y - x);
      myCount += (uint32_t)(y - x);
    }
    // ---- B: clear the dense cells -------------------------------------------------------------------
    for (uint32_t i = tid; i < W + 1; i += KDE_THREADS) { wdF[i] = 0; if (NONDIR) wdR[i] = 0; }
    for (uint32_t i = tid; i < TILE_BLOCKS; i += KDE_THREADS) { mask[i] = 0; hot[i] = 0; }
    const uint32_t any = __syncthreads_or(myCount != 0);
    if (!any) {                                                      // nothing within bw of this tile
      if (!DUMP) { if (tid == 0) { a.tileRunBase[work] = 0; a.tileRunCount[work] = 0; } }
      else {
        for (uint32_t i = tid; i < np; i += KDE_THREADS) { a.dumpF[dumpBase + i] = 0.0; a.dumpR[dumpBase + i] = 0.0; if (a.dumpNonzero) a.dumpNonzero[dumpBase + i] = 0; }
      }
      continue;
    }

    // ---- C: pool the samples: dense per-position weight (integer sum of non-control counts) ---------
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

    // ---- D: cold-block test: no position of a 32-position block can reach the threshold if
    //         maxTap * (sum of weights within bw of the block) < threshold ------------------------------
    const uint32_t nBlk = (np + 31) / 32;
    if (!DUMP) {
      for (uint32_t k = warp; k < nBlk; k += KDE_WARPS) {
        unsigned long long sum = 0;
        const uint32_t cEnd = min(W, 32 * k + 32 + 2 * bw);
        for (uint32_t c = 32 * k + lane; c < cEnd; c += 32) sum += (unsigned long long)wdF[c] + (NONDIR ? (unsigned long long)wdR[c] : 0ull);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
        if (lane == 0) hot[k] = (__dmul_rn((double)sum, a.kmaxBound) >= a.thr) ? 1u : 0u;
      }
    } else {
      for (uint32_t k = tid; k < nBlk; k += KDE_THREADS) hot[k] = 1u;
    }
    __syncthreads();
    if (warp == 0) {
      const uint32_t f0 = lane < nBlk ? hot[lane] : 0u;
      const uint32_t f1 = lane + 32 < nBlk ? hot[lane + 32] : 0u;
      const uint32_t m0 = __ballot_sync(0xffffffffu, f0), m1 = __ballot_sync(0xffffffffu, f1);
      if (f0) hotList[__popc(m0 & lanemask_lt())] = lane;
      if (f1) hotList[__popc(m0) + __popc(m1 & lanemask_lt())] = lane + 32;
      if (lane == 0) s_nHot = __popc(m0) + __popc(m1);
    }
    __syncthreads();
    const uint32_t nHot = s_nHot;
    if (nHot == 0) {
      if (tid == 0) { a.tileRunBase[work] = 0; a.tileRunCount[work] = 0; }
      continue;
    }
    if (!DUMP && tid == 0) atomicAdd(a.statHotTiles, 1ull);

    // ---- E: dense -> sparse sorted tag list ------------------------------------------------------------
    compact_cells(wdF, cltF, tpF, twF, W, s_warpTotals);
    if (NONDIR) compact_cells(wdR, cltR, tpR, twR, W, s_warpTotals);
    __syncthreads();

    // ---- F: exact-order evaluation of the hot blocks, one warp per block of 32 positions ---------------
    for (uint32_t h = warp; h < nHot; h += KDE_WARPS) {
      const uint32_t k = hotList[h];
      const uint32_t po = 32 * k + lane;                            // offset of this lane's position in the tile
      const bool valid = po < np;
      double sF = 0.0, sR = 0.0;
      if (valid) {
        sF = eval_exact(Ks, cltF, tpF, twF, po + bw, bw);
        if (NONDIR) sR = eval_exact(Ks, cltR, tpR, twR, po + bw, bw);
      }
      const double sc = NONDIR ? __dadd_rn(sF, sR) : sF;            // misc/peakcall.cpp:57
      if (!DUMP) {
        score[po] = sc;
        const uint32_t m = __ballot_sync(0xffffffffu, valid && sc >= a.thr);
        if (lane == 0) mask[k] = m;
      } else if (valid) {
        // a forward-strand buffer in directional mode has only F; reverseBuffer only R (src/regions.cpp:356-358)
        const bool rev = (!NONDIR) && U.buffer == 1;
        a.dumpF[dumpBase + po] = rev ? 0.0 : sF;
        a.dumpR[dumpBase + po] = NONDIR ? sR : (rev ? sF : 0.0);
        if (a.dumpNonzero) a.dumpNonzero[dumpBase + po] = (sc != 0.0) ? 1u : 0u;
      }
    }
    if (DUMP) continue;
    __syncthreads();

    // ---- G: maximal runs of above-threshold positions inside the tile ------------------------------------
    if (warp == 0) {
      uint32_t cnt2[2];
#pragma unroll
      for (int hlf = 0; hlf < 2; ++hlf) {
        const uint32_t k = lane + 32 * hlf;
        const uint32_t m = mask[k];
        const uint32_t prev = k > 0 ? (mask[k - 1] >> 31) : 0u;
        cnt2[hlf] = __popc(m & ~((m << 1) | prev));
      }
      uint32_t inc0 = cnt2[0], inc1 = cnt2[1];
#pragma unroll
      for (int o = 1; o < 32; o <<= 1) {
        const uint32_t n0 = __shfl_up_sync(0xffffffffu, inc0, o), n1 = __shfl_up_sync(0xffffffffu, inc1, o);
        if (lane >= (uint32_t)o) { inc0 += n0; inc1 += n1; }
      }
      const uint32_t tot0 = __shfl_sync(0xffffffffu, inc0, 31), tot1 = __shfl_sync(0xffffffffu, inc1, 31);
      startPref[lane] = inc0 - cnt2[0];
      startPref[lane + 32] = tot0 + inc1 - cnt2[1];
      if (lane == 0) {
        const uint32_t nRuns = tot0 + tot1;
        s_nRuns = nRuns;
        uint32_t base = 0;
        if (nRuns) base = atomicAdd(a.runCount, nRuns);
        s_base = base;
        a.tileRunBase[work] = base;
        a.tileRunCount[work] = nRuns;
      }
    }
    __syncthreads();
    const uint32_t nRuns = s_nRuns, base = s_base;
    if (nRuns == 0) continue;
    if ((uint64_t)base + nRuns > a.runCap) { if (tid == 0) atomicOr(a.flags, FLAG_RUN_OVERFLOW); continue; }
    for (uint32_t k = warp; k < TILE_BLOCKS; k += KDE_WARPS) {
      const uint32_t m = mask[k];
      const uint32_t prev = k > 0 ? (mask[k - 1] >> 31) : 0u;
      const uint32_t startm = m & ~((m << 1) | prev);
      if ((startm >> lane) & 1u) {
        const uint32_t rank = startPref[k] + __popc(startm & lanemask_lt());
        uint32_t q = 32 * k + lane;
        const uint32_t q0 = q;
        double best = score[q];
        uint32_t bestq = q, wsum = 0;
        while (q < TILE && ((mask[q >> 5] >> (q & 31)) & 1u)) {
          const double sc = score[q];
          if (sc > best) { best = sc; bestq = q; }                 // first strict maximum (misc/data.cpp:97-101)
          wsum += wdF[q + bw] + (NONDIR ? wdR[q + bw] : 0u);
          ++q;
        }
        RunRec r;
        r.left = ts + q0; r.len = q - q0; r.peak = ts + bestq; r.wsum = wsum; r.score = best;
        a.runs[base + rank] = r;
      }
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
// K2: tile-ordered runs -> candidate regions (stitch runs split at tile boundaries)
// ------------------------------------------------------------------------------------------------
struct Cand {                       // a candidate region == one Region object of the reference
  uint32_t unit, left, len, peak, wsum;
  uint32_t pad;
  double score;
};

__global__ void order_runs_kernel (const RunRec* __restrict__ runs, const uint32_t* __restrict__ tileRunBase,
                                   const uint32_t* __restrict__ tileRunCount, const unsigned long long* __restrict__ tileOrd,
                                   const Unit* __restrict__ units, uint32_t nUnits, uint32_t totalTiles,
                                   RunRec* __restrict__ ordRuns, uint32_t* __restrict__ ordUnit) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= totalTiles) return;
  const uint32_t n = tileRunCount[t];
  if (!n) return;
  uint32_t lo = 0, hi = nUnits;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (units[mid].tileStart <= t) lo = mid; else hi = mid; }
  const uint32_t b = tileRunBase[t];
  const unsigned long long o = tileOrd[t];
  for (uint32_t i = 0; i < n; ++i) { ordRuns[o + i] = runs[b + i]; ordUnit[o + i] = lo; }
}

__global__ void run_heads_kernel (const RunRec* __restrict__ r, const uint32_t* __restrict__ unit, uint32_t n, uint32_t* __restrict__ head) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  bool cont = false;
  if (j > 0) cont = unit[j] == unit[j - 1] && r[j].left == r[j - 1].left + r[j - 1].len;
  head[j] = cont ? 0u : 1u;
}

__global__ void merge_runs_kernel (const RunRec* __restrict__ r, const uint32_t* __restrict__ unit, const uint32_t* __restrict__ head,
                                   const unsigned long long* __restrict__ headScan, uint32_t n, Cand* __restrict__ cand,
                                   unsigned long long* __restrict__ unitCand) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n || !head[j]) return;
  Cand c;
  c.unit = unit[j]; c.left = r[j].left; c.len = r[j].len; c.peak = r[j].peak; c.wsum = r[j].wsum; c.score = r[j].score; c.pad = 0;
  for (uint32_t q = j + 1; q < n && !head[q]; ++q) {
    c.len += r[q].len;
    c.wsum += r[q].wsum;                                            // HitCount wraps mod 2^32 (misc/peakcall.cpp:36-40)
    if (r[q].score > c.score) { c.score = r[q].score; c.peak = r[q].peak; }
  }
  cand[headScan[j]] = c;
  atomicAdd(&unitCand[c.unit], 1ull);
}

// ------------------------------------------------------------------------------------------------
// K3: region filters and statistics
// ------------------------------------------------------------------------------------------------
__global__ void hit_filter_kernel (const Cand* __restrict__ cand, uint32_t n, double hitThr, uint32_t* __restrict__ pass) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pass[i] = ((double)cand[i].wsum >= hitThr) ? 1u : 0u;             // misc/peakcall.cpp:41 first clause
}

__global__ void compact_index_kernel (const uint32_t* __restrict__ pass, const unsigned long long* __restrict__ scan, uint32_t n, uint32_t* __restrict__ list) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (pass[i]) list[scan[i]] = i;
}

// per (surviving region, sample): Region::exptSums (misc/data.cpp:116-131) and the pieces of posMean
// (misc/data.cpp:133-148): sum of counts and sum of count*offset, offsets as UShort, both mod 2^32
struct CountArgs {
  Tags T;
  const Unit* units;
  const uint32_t* nTilesIdx;
  const Cand* cand;
  const uint32_t* list;             // surviving candidate indices
  uint32_t nList;
  int nondir;
  uint32_t* counts;                 // [nList * nSamples]
  uint32_t* offsum;                 // [nList * nSamples]
};

__global__ void region_counts_kernel (CountArgs a) {
  const uint64_t g = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n = a.T.nSamples;
  if (g >= (uint64_t)a.nList * n) return;
  const uint32_t i = (uint32_t)(g / n), s = (uint32_t)(g % n);
  const Cand c = a.cand[a.list[i]];
  const Unit U = a.units[c.unit];
  uint32_t cs = 0, os = 0;
  const uint32_t d0 = a.nondir ? 0u : U.buffer, d1 = a.nondir ? 1u : U.buffer;
  for (uint32_t d = d0; d <= d1; ++d) {
    uint64_t x, y;
    seg_range(a.T, s, d, U.contig, a.nTilesIdx[U.contig], c.left, (int64_t)c.left + c.len - 1, x, y);
    for (uint64_t e = x; e < y; ++e) {
      const uint32_t k = a.T.cntEff[e];
      cs += k;
      os += k * ((a.T.pos[e] - c.left) & 0xFFFFu);
    }
  }
  a.counts[g] = cs;
  a.offsum[g] = os;
}

// Region::posKurtosis (misc/data.cpp:164-182): sequential FP64 sums over the positions that hold tags, in
// position order; one warp per region, positions pooled in shared memory 1024 at a time
struct KurtArgs {
  Tags T;
  const Unit* units;
  const uint32_t* nTilesIdx;
  const Cand* cand;
  const uint32_t* list;
  uint32_t nList;
  int nondir;
  const uint32_t* counts;
  const uint32_t* offsum;
  double kurtThr;
  double* kurt;                     // [nList]
  uint32_t* pass;                   // [nList]
};

constexpr int KURT_CHUNK = 1024;
constexpr int KURT_WARPS = 8;

__global__ void __launch_bounds__(KURT_WARPS * 32) region_kurtosis_kernel (KurtArgs a) {
  __shared__ uint32_t cells[KURT_WARPS][KURT_CHUNK];
  const uint32_t warp = threadIdx.x >> 5, lane = lane_id();
  const uint32_t i = blockIdx.x * KURT_WARPS + warp;
  if (i >= a.nList) return;
  const uint32_t n = a.T.nSamples;
  const Cand c = a.cand[a.list[i]];
  const Unit U = a.units[c.unit];
  uint32_t count = 0, sum = 0;
  for (uint32_t s = lane; s < n; s += 32) { count += a.counts[(uint64_t)i * n + s]; sum += a.offsum[(uint64_t)i * n + s]; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) { count += __shfl_xor_sync(0xffffffffu, count, o); sum += __shfl_xor_sync(0xffffffffu, sum, o); }
  const double xbar = (double)sum / (double)count;                  // misc/data.cpp:147
  double sum2 = 0.0, sum4 = 0.0;
  uint32_t* cell = cells[warp];
  const uint32_t d0 = a.nondir ? 0u : U.buffer, d1 = a.nondir ? 1u : U.buffer;
  for (uint32_t o0 = 0; o0 < c.len; o0 += KURT_CHUNK) {
    const uint32_t oc = min((uint32_t)KURT_CHUNK, c.len - o0);
    for (uint32_t k = lane; k < KURT_CHUNK; k += 32) cell[k] = 0;
    __syncwarp();
    for (uint32_t s = 0; s < n; ++s)
      for (uint32_t d = d0; d <= d1; ++d) {
        uint64_t x, y;
        seg_range(a.T, s, d, U.contig, a.nTilesIdx[U.contig], (int64_t)c.left + o0, (int64_t)c.left + o0 + oc - 1, x, y);
        for (uint64_t e = x + lane; e < y; e += 32) atomicAdd(&cell[a.T.pos[e] - c.left - o0], a.T.cntEff[e]);
      }
    __syncwarp();
    for (uint32_t g = 0; g < oc; g += 32) {
      const uint32_t o = o0 + g + lane;
      const uint32_t pc = (g + lane < oc) ? cell[g + lane] : 0u;
      const double dd = __dsub_rn((double)(o & 0xFFFFu), xbar);     // offset counter is a UShort (Q8)
      const double d2 = __dmul_rn(dd, dd);
      const double d4 = __dmul_rn(d2, d2);                          // pow(x,4) == (x*x)*(x*x) in the reference build (Q9)
      const double t2 = __dmul_rn((double)pc, d2);
      const double t4 = __dmul_rn((double)pc, d4);
      uint32_t nz = __ballot_sync(0xffffffffu, pc != 0);
      while (nz) {
        const int b = __ffs(nz) - 1;
        sum2 = __dadd_rn(sum2, __shfl_sync(0xffffffffu, t2, b));
        sum4 = __dadd_rn(sum4, __shfl_sync(0xffffffffu, t4, b));
        nz &= nz - 1;
      }
    }
    __syncwarp();
  }
  const double k = __ddiv_rn(__dmul_rn(__dsub_rn((double)count, 1.0), sum4), __dmul_rn(sum2, sum2));   // misc/data.cpp:181
  if (lane == 0) {
    a.kurt[i] = k;
    a.pass[i] = (a.kurtThr == 0.0 || (c.len > 1 && k <= a.kurtThr)) ? 1u : 0u;   // misc/peakcall.cpp:41 second clause
  }
}

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

__global__ void corr_filter_kernel (const double* __restrict__ corr, uint32_t n, double thr, uint32_t* __restrict__ pass) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  pass[i] = (corr[i] >= thr) ? 1u : 0u;                             // NaN fails (misc/peakcall.cpp:41 third clause)
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
// gather helpers
// ------------------------------------------------------------------------------------------------
__global__ void gather_u32_kernel (const uint32_t* __restrict__ src, const uint32_t* __restrict__ list, uint32_t n, uint32_t* __restrict__ dst) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[list[i]];
}

__global__ void dump_desc_kernel (const Cand* __restrict__ cand, const uint32_t* __restrict__ list, uint32_t n,
                                  uint32_t* __restrict__ dUnit, uint32_t* __restrict__ dLeft, uint32_t* __restrict__ dLen, uint32_t* __restrict__ dChunks) {
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const Cand c = cand[list[i]];
  dUnit[i] = c.unit; dLeft[i] = c.left; dLen[i] = c.len; dChunks[i] = (c.len + TILE - 1) / TILE;
}

}  // namespace upb
