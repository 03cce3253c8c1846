// upb_pack.cuh — the compact host format of the tag arrays and its device-side decoder.
//
// The load is PCIe-bound (8 bytes per entry: position u32 + count u32).  Sorted positions and small counts compress well with
// no entropy coding at all: 16-bit position deltas inside a (sample, strand, contig) segment and 8-bit counts, with a sparse,
// index-sorted exception list for what does not fit (the first position of every segment, gaps >= 65535, counts >= 255):
// 3 bytes per entry.  The decoder is a segmented inclusive scan of the deltas (an exception restarts the running position) in
// three kernels — per-tile aggregates, one block that scans the aggregates, per-tile final scan — that writes the same pos[] /
// cnt[] arrays the plain path copies, so everything downstream is shared.  What the reference's parser guarantees (ascending
// positions inside the contig, non-zero counts) is validated afterwards on the decoded arrays exactly as for the plain path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace upb {

constexpr int PK_THREADS = 256;
constexpr int PK_ITEMS = 8;
constexpr int PK_TILE = PK_THREADS * PK_ITEMS;
constexpr uint32_t PK_DPOS_ESC = 0xFFFFu;
constexpr uint32_t PK_CNT_ESC = 0xFFu;
// the 2-byte form (upb_pack_tags2 / upb_load_tags_packed2): one u16 per entry, position delta in the low 12 bits, count in the high 4
constexpr uint32_t PK2_DPOS_MASK = 0x0FFFu, PK2_DPOS_ESC = 0x0FFFu, PK2_CNT_SHIFT = 12, PK2_CNT_ESC = 0xFu;
constexpr uint32_t FLAG_BAD_PACK = 16u;             // an escape without its exception entry

struct PkException { unsigned long long index; uint32_t pos; uint32_t cnt; };   // == upb_packed_exception

struct PkAgg { uint32_t sum; uint32_t reset; };     // running position contributed by a range; reset: the range holds an absolute position

__device__ __forceinline__ PkAgg pk_combine (const PkAgg a, const PkAgg b) {   // a then b
  PkAgg r;
  r.sum = b.reset ? b.sum : a.sum + b.sum;
  r.reset = a.reset | b.reset;
  return r;
}

// exception of entry e, or nullptr
__device__ __forceinline__ const PkException* pk_find (const PkException* __restrict__ exc, unsigned long long nExc, unsigned long long e) {
  unsigned long long lo = 0, hi = nExc;
  while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (exc[mid].index < e) lo = mid + 1; else hi = mid; }
  return (lo < nExc && exc[lo].index == e) ? exc + lo : nullptr;
}

struct PkArgs {
  const uint16_t* dpos;
  const uint8_t* cnt8;               // NULL: the 2-byte form, the count sits in the high bits of dpos
  uint32_t dposMask, dposEsc, cntEsc;   // 0xFFFF / 0xFFFF / 0xFF for the 3-byte form
  const PkException* exc;
  unsigned long long nExc;
  unsigned long long first, E;       // entries [first, first + E) are decoded (one sample's chunk when the load is pipelined)
  PkAgg* agg;                        // [nTiles]
  PkAgg* carry;                      // [nTiles] running position entering each tile
  uint32_t* pos;
  uint32_t* cnt;
  uint32_t* flags;
};

// the items of one thread: value and reset flag of up to PK_ITEMS consecutive entries
__device__ __forceinline__ PkAgg pk_thread_scan (const PkArgs& a, unsigned long long e0, uint32_t vals[PK_ITEMS], uint32_t resets[PK_ITEMS], int* nOut) {
  PkAgg t; t.sum = 0; t.reset = 0;
  int n = 0;
  const unsigned long long end = a.first + a.E;
#pragma unroll
  for (int q = 0; q < PK_ITEMS; ++q) {
    const unsigned long long e = e0 + q;
    if (e < end) {
      uint32_t v = a.dpos[e] & a.dposMask, rs = 0;
      if (v == a.dposEsc) {
        const PkException* x = pk_find(a.exc, a.nExc, e);
        if (x) { v = x->pos; rs = 1; } else { atomicOr(a.flags, FLAG_BAD_PACK); v = 0; }
      }
      PkAgg it; it.sum = v; it.reset = rs;
      t = pk_combine(t, it);
      vals[q] = t.sum; resets[q] = t.reset;
      n = q + 1;
    }
  }
  *nOut = n;
  return t;
}

// block-wide exclusive segmented scan of the per-thread aggregates; returns the aggregate of everything before this thread
// and (to thread 0's *total) of the whole tile
__device__ __forceinline__ PkAgg pk_block_exclusive (PkAgg mine, PkAgg* total) {
  __shared__ PkAgg warpAgg[PK_THREADS / 32];
  const uint32_t lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  PkAgg incl = mine;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    PkAgg up; up.sum = __shfl_up_sync(0xffffffffu, incl.sum, o); up.reset = __shfl_up_sync(0xffffffffu, incl.reset, o);
    if (lane >= (uint32_t)o) incl = pk_combine(up, incl);
  }
  if (lane == 31) warpAgg[warp] = incl;
  __syncthreads();
  PkAgg before; before.sum = 0; before.reset = 0;
  for (uint32_t w = 0; w < warp; ++w) before = pk_combine(before, warpAgg[w]);
  PkAgg prevLane; prevLane.sum = __shfl_up_sync(0xffffffffu, incl.sum, 1); prevLane.reset = __shfl_up_sync(0xffffffffu, incl.reset, 1);
  PkAgg excl = before;
  if (lane > 0) excl = pk_combine(before, prevLane);
  if (total) { PkAgg tt; tt.sum = 0; tt.reset = 0; for (uint32_t w = 0; w < PK_THREADS / 32; ++w) tt = pk_combine(tt, warpAgg[w]); *total = tt; }
  __syncthreads();
  return excl;
}

__global__ void __launch_bounds__(PK_THREADS) pk_tile_kernel (PkArgs a) {
  uint32_t vals[PK_ITEMS], resets[PK_ITEMS];
  int n;
  const unsigned long long e0 = a.first + (unsigned long long)blockIdx.x * PK_TILE + (unsigned long long)threadIdx.x * PK_ITEMS;
  const PkAgg mine = pk_thread_scan(a, e0, vals, resets, &n);
  PkAgg total;
  pk_block_exclusive(mine, &total);
  if (threadIdx.x == 0) a.agg[blockIdx.x] = total;
}

// one block: exclusive segmented scan of the tile aggregates
__global__ void __launch_bounds__(1024) pk_top_kernel (const PkAgg* __restrict__ agg, uint32_t nTiles, PkAgg* __restrict__ carry) {
  __shared__ PkAgg part[1024];
  const uint32_t per = (nTiles + 1023) / 1024;
  const uint32_t t0 = threadIdx.x * per, t1 = min(t0 + per, nTiles);
  PkAgg mine; mine.sum = 0; mine.reset = 0;
  for (uint32_t t = t0; t < t1; ++t) mine = pk_combine(mine, agg[t]);
  part[threadIdx.x] = mine;
  __syncthreads();
  if (threadIdx.x == 0) {
    PkAgg run; run.sum = 0; run.reset = 0;
    for (uint32_t i = 0; i < 1024; ++i) { const PkAgg p = part[i]; part[i] = run; run = pk_combine(run, p); }
  }
  __syncthreads();
  PkAgg run = part[threadIdx.x];
  for (uint32_t t = t0; t < t1; ++t) { carry[t] = run; run = pk_combine(run, agg[t]); }
}

__global__ void __launch_bounds__(PK_THREADS) pk_final_kernel (PkArgs a) {
  uint32_t vals[PK_ITEMS], resets[PK_ITEMS];
  int n;
  const unsigned long long e0 = a.first + (unsigned long long)blockIdx.x * PK_TILE + (unsigned long long)threadIdx.x * PK_ITEMS;
  const PkAgg mine = pk_thread_scan(a, e0, vals, resets, &n);
  const PkAgg excl = pk_combine(a.carry[blockIdx.x], pk_block_exclusive(mine, nullptr));
  for (int q = 0; q < n; ++q) {
    const unsigned long long e = e0 + q;
    a.pos[e] = resets[q] ? vals[q] : excl.sum + vals[q];
    uint32_t c = a.cnt8 ? (uint32_t)a.cnt8[e] : (uint32_t)a.dpos[e] >> PK2_CNT_SHIFT;
    if (c == a.cntEsc) {
      const PkException* x = pk_find(a.exc, a.nExc, e);
      if (x) c = x->cnt; else { atomicOr(a.flags, FLAG_BAD_PACK); c = 0; }
    }
    a.cnt[e] = c;
  }
}

// first / last position of every segment (the host keeps them for the reference-order logic)
__global__ void seg_ends_kernel (const uint32_t* __restrict__ pos, const uint64_t* __restrict__ segStart, uint32_t nSeg, uint32_t* __restrict__ firstPos, uint32_t* __restrict__ lastPos) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nSeg) return;
  const uint64_t s = segStart[g], e = segStart[g + 1];
  firstPos[g] = e > s ? pos[s] : 0u;
  lastPos[g] = e > s ? pos[e - 1] : 0u;
}

}  // namespace upb
