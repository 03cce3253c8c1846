// upb_aggr.cuh — tag counts per (strand, contig, 5' position) of a batch of alignments: the aggregation bin/convert_align does
// with one std::map per strand and contig (CountMap::add, misc/data.cpp:301-314; src/convert_align.cpp:70-90).
//
// Same two-level layout as the tag loader: a counting sort of the alignments by (strand, contig, 512-position mini-tile)
// (histogram, scan, scatter), then one warp per non-empty mini-tile pools its alignments into 512 shared-memory cells and
// writes the cells that received an alignment out in ascending position — first pass counts them, a scan places them, second
// pass emits.  "Received an alignment" is a bit per cell, not "count != 0": the map keeps a key whose count wrapped to exactly
// 0 mod 2^32 (i->second += hit.count), and never creates one for an alignment with count 0 (misc/data.cpp:302).
#pragma once

#include "upb_device.cuh"

namespace upb {

struct AggrArgs {
  const uint32_t* key;              // [n] strand * C + contig
  const uint32_t* pos;              // [n] 1-based 5' position
  const uint32_t* cnt;              // [n] or NULL (1 each)
  unsigned long long n;
  const uint32_t* sizes;            // [C]
  const uint32_t* mtBase;           // [2C+1]
  uint32_t C;
  int combine;                      // -D: every alignment counts on strand 0
  uint32_t* mtCount;                // [nMt] histogram
  uint32_t* cursor;                 // [nMt]
  const unsigned long long* mtOff;  // [nMt+1]
  uint2* ent;                       // [n] (count, offset inside the mini-tile), tile-major
  uint32_t* uniq;                   // [nMt] distinct positions per mini-tile
  const unsigned long long* outOff; // [nMt+1]
  uint32_t* outPos;
  uint32_t* outCnt;
  uint32_t nMt;
  uint32_t* flags;
  unsigned long long* tagCount;
};

template <bool SCATTER>
__global__ void aggr_sort_kernel (AggrArgs a) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.n) return;
  uint32_t k = a.key[i];
  const uint32_t p1 = a.pos[i];
  if (k >= 2 * a.C) { if (!SCATTER) atomicOr(a.flags, FLAG_BAD_POS); return; }
  const uint32_t c = k % a.C;
  if (a.combine) k = c;
  if (p1 < 1 || p1 > a.sizes[c]) { if (!SCATTER) atomicOr(a.flags, FLAG_BAD_POS); return; }
  const uint32_t mt = a.mtBase[k] + ((p1 - 1) >> MT_SHIFT);
  if (!SCATTER) { atomicAdd(&a.mtCount[mt], 1u); return; }
  const unsigned long long slot = a.mtOff[mt] + atomicAdd(&a.cursor[mt], 1u);
  a.ent[slot] = make_uint2(a.cnt ? a.cnt[i] : 1u, (p1 - 1) & ((1u << MT_SHIFT) - 1));
}

constexpr int AG_WARPS = 8;
constexpr int AG_CELLS = 1 << MT_SHIFT;

template <bool EMIT>
__global__ void __launch_bounds__(AG_WARPS * 32) aggr_reduce_kernel (AggrArgs a) {
  __shared__ uint32_t cells[AG_WARPS][AG_CELLS];
  __shared__ uint32_t present[AG_WARPS][AG_CELLS / 32];
  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t m = blockIdx.x * AG_WARPS + warp;
  if (m >= a.nMt) return;
  const unsigned long long e0 = a.mtOff[m], e1 = a.mtOff[m + 1];
  if (e0 == e1) { if (!EMIT && lane == 0) a.uniq[m] = 0; return; }
  uint32_t* cell = cells[warp];
  uint32_t* seen = present[warp];
  for (uint32_t k = lane; k < (uint32_t)AG_CELLS; k += 32) cell[k] = 0;
  if (lane < (uint32_t)AG_CELLS / 32) seen[lane] = 0;
  __syncwarp();
  unsigned long long tot = 0;
  for (unsigned long long e = e0 + lane; e < e1; e += 32) {
    const uint2 x = a.ent[e];
    if (x.x) { atomicAdd(&cell[x.y], x.x); atomicOr(&seen[x.y >> 5], 1u << (x.y & 31)); tot += x.x; }   // HitCount arithmetic wraps like the map's +=
  }
  __syncwarp();
  // position of mini-tile m inside its (strand, contig): largest dc with mtBase[dc] <= m
  uint32_t lo = 0, hi = 2 * a.C;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.mtBase[mid] <= m) lo = mid; else hi = mid; }
  const uint32_t p0 = ((m - a.mtBase[lo]) << MT_SHIFT) + 1;
  uint32_t run = 0;
  for (uint32_t k0 = 0; k0 < (uint32_t)AG_CELLS; k0 += 32) {
    const uint32_t v = cell[k0 + lane];
    const uint32_t msk = seen[k0 >> 5];
    if (EMIT && ((msk >> lane) & 1u)) {
      const unsigned long long o = a.outOff[m] + run + __popc(msk & ((1u << lane) - 1));
      a.outPos[o] = p0 + k0 + lane;
      a.outCnt[o] = v;
    }
    run += __popc(msk);
  }
  if (!EMIT) {
    if (lane == 0) a.uniq[m] = run;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
    if (lane == 0) atomicAdd(a.tagCount, tot);
  }
}

// off[dc] = number of output entries before (strand, contig) dc
__global__ void aggr_offsets_kernel (const unsigned long long* __restrict__ outOff, const uint32_t* __restrict__ mtBase, uint32_t nDc, unsigned long long* __restrict__ off) {
  const uint32_t dc = blockIdx.x * blockDim.x + threadIdx.x;
  if (dc <= nDc) off[dc] = outOff[mtBase[dc]];
}

}  // namespace upb
