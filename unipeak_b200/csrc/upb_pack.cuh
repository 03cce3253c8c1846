// upb_pack.cuh — the compact host format of the tag arrays and the device side of the load (decode + validation + the
// histogram pass of the tile-major counting sort, in one kernel).
//
// The load is PCIe-bound (plain form: 8 bytes per entry, position u32 + count u32).  Sorted positions and small counts compress
// well with no entropy coding at all: ONE u16 per entry — the position delta to the previous entry of the same
// (sample, strand, contig) segment in the low 12 bits, the count in the high 4 — with a sparse, index-sorted exception list for
// what does not fit (0xFFF / 0xF escapes: absolute position / full count).  Every entry whose index inside its segment is a
// multiple of PK_TILE (2048) MUST be a position escape, so a tile of 2048 entries decodes on its own: one block, one
// shared-memory scan, no second pass over the input and no carry between tiles (r01 decoded with three kernels and read the
// packed words twice).  On RNA-seq / ChIP-seq-like data ~0.5 % of the entries escape: 2.1 bytes per entry.
//
// load_tile_kernel works on the tiles of one LOAD GROUP — the segments of all samples on one strand and a range of contigs —
// so that a group's decode, histogram, scan and scatter run while the next group's bytes are still crossing PCIe
// (upb_api.cu: load_enqueue).  It validates what the reference's parser and asserts guarantee (positions inside the contig and
// strictly ascending, non-zero counts: misc/format.cpp:654-710, misc/peakcall.cpp:171) and what the format itself promises (an
// escape has its exception, tiles start with a position escape); the seams between tiles are checked by load_seam_kernel.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "upb_device.cuh"
#include "upb_vbyte.hpp"

namespace upb {

constexpr int PK_THREADS = 256;
constexpr int PK_ITEMS = 8;
constexpr int PK_TILE = PK_THREADS * PK_ITEMS;
static_assert(PK_TILE == (int)upb_vb::VB_TILE, "the variable-length byte form restarts its positions at the tiles of the load");
constexpr uint32_t PK2_DPOS_MASK = 0x0FFFu, PK2_DPOS_ESC = 0x0FFFu, PK2_CNT_SHIFT = 12, PK2_CNT_ESC = 0xFu;
constexpr uint32_t FLAG_BAD_PACK = 16u;             // an escape without its exception entry, or a tile that does not start with a position escape

struct PkException { unsigned long long index; uint32_t pos; uint32_t cnt; };   // == upb_packed_exception; index = entry index in the caller's arrays

// exception of source entry e, or nullptr
__device__ __forceinline__ const PkException* pk_find (const PkException* __restrict__ exc, unsigned long long nExc, unsigned long long e,
                                                      unsigned long long lo = 0, unsigned long long hi = ~0ull) {
  if (hi > nExc) hi = nExc;
  while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (exc[mid].index < e) lo = mid + 1; else hi = mid; }
  return (lo < nExc && exc[lo].index == e) ? exc + lo : nullptr;
}

struct LoadArgs {
  const uint16_t* words;            // packed form: device copy of the kept segments' words, laid out like pos / cnt (NULL: plain form, pos / cnt already hold the entries)
  const PkException* exc;
  unsigned long long nExc;
  uint32_t* pos;                    // resident sample-major arrays (compacted: only the segments this context keeps)
  uint32_t* cnt;
  const uint64_t* segStart;         // [nSeg + 1] first resident entry of segment g = (s*2+d)*C + c
  const uint64_t* srcStart;         // [nSeg] first entry of segment g in the caller's arrays
  const uint32_t* grpSeg;           // [nGrpSeg] segments of this load group
  const uint32_t* grpTile;          // [nGrpSeg + 1] tiles before each of them
  uint32_t nGrpSeg;
  const uint32_t* sizes;            // [C]
  const uint32_t* mtBase;           // [2C + 1]
  uint32_t* mtCount;                // [nMtTotal] histogram of the counting sort
  const uint32_t* winBase;          // [2C + 1]
  uint32_t* wEnt;                   // [nSamples * nWinTotal] first entry (relative to its segment) of sample s at or after each window
  uint32_t nWinTotal;
  uint32_t C;
  uint32_t* flags;
  unsigned long long* tileExc;      // [tiles of the group + 1] packed form: index of the first exception at or after each tile's first entry (load_tile_prep_kernel)
  // variable-length byte form (upb_vbyte.hpp)
  const uint8_t* vbytes;            // device copy of the kept segments' bytes, in segment order
  const uint64_t* vSrcByte;         // [nSeg + 1] first byte of segment g in the caller's byte array
  const uint64_t* vDevByte;         // [nSeg] first byte of segment g in the device copy
  const uint64_t* vSrcTile;         // [nSeg] number of segment g's first tile in the caller's tile table
  const uint64_t* vTileOff;         // [caller's tiles + 1] first byte of every tile in the caller's byte array
  unsigned long long vTiles;        // tiles in the caller's table
  struct TileDesc* tileDesc;        // [tiles of the group] what load_tile_prep_kernel found out about every tile (packed forms; NULL: plain form, the tile kernel searches itself)
};

// Per tile of a load group, worked out by one thread per tile for all tiles at once (load_tile_prep_kernel) instead of by thread 0 of
// every tile block in front of its first barrier: a binary search over the group's segments and, for the byte form, three dependent
// look-ups — ~3 us of dependent loads per block that only a handful of resident blocks per SM could hide.
struct __align__(16) TileDesc {
  uint32_t seg, tile;               // segment g = (s*2+d)*C + c, tile inside it
  uint32_t len, bad;                // byte form: bytes of the tile; bad: the tile table is inconsistent here
  unsigned long long devByte;       // byte form: first byte of the tile in the device copy
  unsigned long long pad;
};

enum { LOAD_PLAIN = 0, LOAD_WORDS = 1, LOAD_VBYTE = 2 };

// window index of sample s: every window (65536 positions) that starts after the previous entry and not after this one begins here
__device__ __forceinline__ void load_open_windows (const LoadArgs& a, uint32_t s, uint32_t dc, int64_t prevWin, uint32_t p0 /* position - 1 */, uint32_t rel) {
  const int64_t w = (int64_t)(p0 >> WIN_SHIFT);
  if (w <= prevWin) return;
  uint32_t* we = a.wEnt + (size_t)s * a.nWinTotal + a.winBase[dc];
  for (int64_t t = prevWin + 1; t <= w; ++t) we[t] = rel;
}

// Packed form: the exception list must be strictly increasing in its entry index and stay inside the caller's arrays — the look-ups are
// binary searches, which would otherwise miss (or find) an exception by accident.
__global__ void load_check_exc_kernel (const PkException* __restrict__ exc, unsigned long long nExc, unsigned long long nSrcEntries, uint32_t* __restrict__ flags) {
  const unsigned long long j = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= nExc) return;
  if (exc[j].index >= nSrcEntries || (j > 0 && exc[j - 1].index >= exc[j].index)) atomicOr(flags, FLAG_BAD_PACK);
}

// Packed form: where the exceptions of every tile start.  A tile's first entry is a position escape, so every tile has at least one
// exception and its exceptions are consecutive in the index-sorted list: one binary search per tile HERE (one thread per tile, all
// tiles of the group in flight at once) instead of a 22-step chain of dependent loads inside load_tile_kernel, where a whole block
// waited at its barrier for the threads that hold an escape (r02 capture: 44 % of the kernel's stall samples on that barrier).
__global__ void load_tile_prep_kernel (const LoadArgs a, uint32_t nTiles) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t > nTiles) return;
  if (t == nTiles) { a.tileExc[t] = a.nExc; return; }                  // (only an upper bound of the last tile's window)
  uint32_t lo = 0, hi = a.nGrpSeg;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.grpTile[mid] <= t) lo = mid; else hi = mid; }
  const uint32_t g = a.grpSeg[lo], k = t - a.grpTile[lo];
  const unsigned long long e = a.srcStart[g] + (unsigned long long)k * PK_TILE;
  unsigned long long xl = 0, xh = a.nExc;
  while (xl < xh) { const unsigned long long mid = (xl + xh) >> 1; if (a.exc[mid].index < e) xl = mid + 1; else xh = mid; }
  a.tileExc[t] = xl;
  TileDesc d;
  d.seg = g; d.tile = k; d.len = 0; d.bad = 0; d.devByte = 0; d.pad = 0;
  if (a.vbytes) {
    const unsigned long long T = a.vSrcTile[g] + k;
    unsigned long long bs = 0, be = 0;
    bool ok = T < a.vTiles;
    if (ok) { bs = a.vTileOff[T]; be = a.vTileOff[T + 1]; ok = be >= bs && be - bs <= upb_vb::VB_MAX_TILE_BYTES && bs >= a.vSrcByte[g] && be <= a.vSrcByte[g + 1]; }
    d.len = ok ? (uint32_t)(be - bs) : 0u; d.bad = ok ? 0u : 1u;
    d.devByte = a.vDevByte[g] + (ok ? bs - a.vSrcByte[g] : 0ull);
  }
  a.tileDesc[t] = d;
}

template <int MODE>
__global__ void __launch_bounds__(PK_THREADS, MODE == LOAD_VBYTE ? 6 : MODE == LOAD_WORDS ? 5 : 8) load_tile_kernel (const LoadArgs a) {
  __shared__ uint32_t warpSum[PK_THREADS / 32], warpReset[PK_THREADS / 32];
  __shared__ uint32_t lastOf[PK_THREADS];           // last decoded position of every thread (0: the thread holds no entry)
  __shared__ uint32_t s_seg, s_tile;
  __shared__ uint32_t s_len, s_bad;
  __shared__ unsigned long long s_devByte;
  if (threadIdx.x == 0) {
    if (MODE != LOAD_PLAIN) {                         // load_tile_prep_kernel has been here
      const TileDesc d = a.tileDesc[blockIdx.x];
      s_seg = d.seg; s_tile = d.tile; s_len = d.len; s_bad = d.bad; s_devByte = d.devByte;
    } else {
      uint32_t lo = 0, hi = a.nGrpSeg;                // largest j with grpTile[j] <= blockIdx.x (segments without entries have no tile and are skipped)
      while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.grpTile[mid] <= blockIdx.x) lo = mid; else hi = mid; }
      s_seg = a.grpSeg[lo]; s_tile = blockIdx.x - a.grpTile[lo];
    }
  }
  __syncthreads();
  const uint32_t g = s_seg, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t c = g % a.C, sd = g / a.C, d = sd & 1u, s = sd >> 1, dc = d * a.C + c;
  const uint64_t seg0 = a.segStart[g], segN = a.segStart[g + 1] - seg0;
  const uint64_t t0 = (uint64_t)s_tile * PK_TILE;                      // first entry of the tile, relative to the segment
  const uint32_t nHere = (uint32_t)min((uint64_t)PK_TILE, segN - t0);
  const uint32_t i0 = threadIdx.x * PK_ITEMS;
  const uint32_t size = a.sizes[c];
  uint32_t vals[PK_ITEMS], cnts[PK_ITEMS];
  uint32_t flags = 0, run = 0, reset = 0, escMask = 0;
  int n = 0;
  if constexpr (MODE == LOAD_VBYTE) {
    // ---- the tile's bytes -> (value, count) per entry, by the whole block: bit 7 marks the first byte of an entry, so a prefix count of
    // the marked bytes numbers the entries; the thread that holds an entry's first byte decodes it (upb_vbyte.hpp)
    using namespace upb_vb;
    __shared__ __align__(16) uint8_t sb[VB_MAX_TILE_BYTES + 32];
    __shared__ __align__(16) uint16_t startAt[PK_TILE + 8];              // byte index (into sb) of every entry's first byte; [nHere] = end of the tile
    __shared__ uint32_t vbWarp[PK_THREADS / 32];
    const uint8_t* const tp = a.vbytes + s_devByte;                       // the tile's first byte (a.vbytes is 256-byte aligned)
    const uint32_t len = s_len, shift = (uint32_t)(s_devByte & 15ull);
    { const uint4* src = (const uint4*)(tp - shift); uint4* dst = (uint4*)sb;   // whole aligned 16-byte words (the device copy has slack behind its last byte)
      for (uint32_t i = threadIdx.x; i < (shift + len + 15) >> 4; i += PK_THREADS) dst[i] = src[i]; }
    __syncthreads();
    const uint32_t* sbw = (const uint32_t*)sb;
    uint32_t g0, g1;
    vb_chunk16(shift, len, PK_THREADS, threadIdx.x, &g0, &g1);
    uint32_t mine = 0;
    for (uint32_t gq = g0; gq < g1; ++gq) mine += __popc(vb_starts16(sbw + 4 * gq, gq, shift, len));
    uint32_t inc = mine;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (uint32_t)o) inc += u; }
    if (lane == 31) vbWarp[warp] = inc;
    __syncthreads();
    uint32_t j = inc - mine, total = 0;                                   // entries that start before this thread's bytes
    for (uint32_t w2 = 0; w2 < (uint32_t)PK_THREADS / 32; ++w2) { if (w2 < warp) j += vbWarp[w2]; total += vbWarp[w2]; }
    if (s_bad || total != nHere || (len && !(sb[shift] & 0x80u))) flags |= FLAG_BAD_PACK;   // the tile holds exactly its entries and starts with one
    vb_list_starts(sbw, shift, len, g0, g1, j, nHere, startAt);
    if (threadIdx.x < 8) startAt[min(total, nHere) + threadIdx.x] = (uint16_t)(shift + len);   // the last entry ends with the tile
    __syncthreads();
    uint32_t ev[PK_ITEMS], ek[PK_ITEMS];
    { const uint4 s4 = ((const uint4*)startAt)[threadIdx.x];               // this thread's eight entries and the start of the next one
      const uint32_t nxt = startAt[PK_ITEMS * threadIdx.x + PK_ITEMS];
      const uint32_t at[PK_ITEMS + 1] = {s4.x & 0xFFFFu, s4.x >> 16, s4.y & 0xFFFFu, s4.y >> 16, s4.z & 0xFFFFu, s4.z >> 16, s4.w & 0xFFFFu, s4.w >> 16, nxt};
#pragma unroll
      for (int q = 0; q < PK_ITEMS; ++q) {
        ev[q] = 0; ek[q] = 1;
        if (i0 + q < nHere && i0 + q < total && !vb_get_span(sbw, at[q], at[q + 1] - at[q], &ev[q], &ek[q])) { flags |= FLAG_BAD_PACK; ev[q] = 0; ek[q] = 1; }
      } }
    static_assert(PK_ITEMS == 8, "the entry fetch above reads eight entries per thread");
    const unsigned long long src0 = a.srcStart[g] + t0;
    // the tile's exceptions: from x0 on, at most one per entry, and in front of the next tile's (the tiles of a group ascend in the caller's entry
    // index) — a window of a handful of entries instead of 2048: three dependent look-ups per escape instead of eleven, with the block waiting
    const unsigned long long x0 = a.tileExc[blockIdx.x], x1 = min(x0 + nHere, a.tileExc[blockIdx.x + 1]);
#pragma unroll
    for (int q = 0; q < PK_ITEMS; ++q) {
      const uint32_t i = i0 + q;
      if (i < nHere) {
        const uint32_t v = ev[q];
        uint32_t k = ek[q];
        if (k == 0) {                                                  // the count did not fit its byte
          const PkException* x = pk_find(a.exc, a.nExc, src0 + i, x0, x1);
          if (!x) flags |= FLAG_BAD_PACK; else k = x->cnt;
        }
        if (i == 0) { run = v; reset = 1; escMask |= 1u << q; } else run += v;   // a tile starts with an absolute position
        vals[q] = run; cnts[q] = k;
        n = q + 1;
      }
    }
  }
  if constexpr (MODE == LOAD_WORDS) {
    const unsigned long long src0 = a.srcStart[g] + t0;
    // the tile's exceptions: from x0 on, at most one per entry, and in front of the next tile's (the tiles of a group ascend in the caller's entry
    // index) — a window of a handful of entries instead of 2048: three dependent look-ups per escape instead of eleven, with the block waiting
    const unsigned long long x0 = a.tileExc[blockIdx.x], x1 = min(x0 + nHere, a.tileExc[blockIdx.x + 1]);
#pragma unroll
    for (int q = 0; q < PK_ITEMS; ++q) {
      const uint32_t i = i0 + q;
      if (i < nHere) {
        const uint32_t w = a.words[seg0 + t0 + i];                     // the device copy is laid out like the resident arrays; exceptions are indexed like the caller's
        uint32_t v = w & PK2_DPOS_MASK, k = w >> PK2_CNT_SHIFT;
        const bool escP = v == PK2_DPOS_ESC, escC = k == PK2_CNT_ESC;
        if (escP || escC) {
          const PkException* x = pk_find(a.exc, a.nExc, src0 + i, x0, x1);
          if (!x) { flags |= FLAG_BAD_PACK; if (escP) v = 0; if (escC) k = 0; }
          else { if (escP) v = x->pos; if (escC) k = x->cnt; }
        }
        if (i == 0 && !escP) flags |= FLAG_BAD_PACK;                   // a tile starts with an absolute position
        if (escP) { run = v; reset = 1; escMask |= 1u << q; } else run += v;
        vals[q] = run; cnts[q] = k;
        n = q + 1;
      }
    }
  }
  if constexpr (MODE != LOAD_PLAIN) {
    // block-wide segmented scan of (run, reset): a thread without a reset adds the running position of the threads before it
    uint32_t isum = run, ires = reset;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t us = __shfl_up_sync(0xffffffffu, isum, o), ur = __shfl_up_sync(0xffffffffu, ires, o);
      if (lane >= (uint32_t)o) { if (!ires) isum += us; ires |= ur; }
    }
    if (lane == 31) { warpSum[warp] = isum; warpReset[warp] = ires; }
    __syncthreads();
    uint32_t bs = 0;                                                   // everything before this warp
    for (uint32_t w2 = 0; w2 < warp; ++w2) { if (warpReset[w2]) bs = warpSum[w2]; else bs += warpSum[w2]; }
    uint32_t ps = __shfl_up_sync(0xffffffffu, isum, 1), pr = __shfl_up_sync(0xffffffffu, ires, 1);
    if (lane == 0) { ps = 0; pr = 0; }
    const uint32_t before = pr ? ps : bs + ps;                         // running position entering this thread
    const int firstEsc = escMask ? __ffs(escMask) - 1 : PK_ITEMS;      // entries before this thread's first escape continue the position entering the thread
#pragma unroll
    for (int q = 0; q < PK_ITEMS; ++q) if (q < n && q < firstEsc) vals[q] += before;
  } else {
#pragma unroll
    for (int q = 0; q < PK_ITEMS; ++q) {
      const uint32_t i = i0 + q;
      if (i < nHere) { vals[q] = a.pos[seg0 + t0 + i]; cnts[q] = a.cnt[seg0 + t0 + i]; n = q + 1; }
    }
  }
  lastOf[threadIdx.x] = n ? vals[n - 1] : 0u;
  __syncthreads();
  // validation, histogram, window index, and (packed form) the decoded arrays
  uint32_t prev = 0;                                                   // position of the entry before this thread's first one, 0 at the start of the tile (the seam kernel looks at that one)
  if (threadIdx.x > 0 && n) prev = lastOf[threadIdx.x - 1];
  const uint32_t gb = a.mtBase[dc];
#pragma unroll
  for (int q = 0; q < PK_ITEMS; ++q) if (q < n) {
    const uint32_t p = vals[q], k = cnts[q];
    const uint64_t rel = t0 + i0 + q;
    if (MODE != LOAD_PLAIN) { a.pos[seg0 + rel] = p; a.cnt[seg0 + rel] = k; }
    if (p < 1 || p > size) { flags |= FLAG_BAD_POS; prev = p; continue; }
    if (k == 0) flags |= FLAG_ZERO_COUNT;
    const bool first = i0 + q == 0;
    if (!first && prev >= p) flags |= FLAG_UNSORTED;
    else {
      atomicAdd(&a.mtCount[gb + ((p - 1) >> MT_SHIFT)], 1u);
      if (!first) load_open_windows(a, s, dc, (int64_t)((prev - 1) >> WIN_SHIFT), p - 1, (uint32_t)rel);
    }
    prev = p;
  }
  if (flags) atomicOr(a.flags, flags);
}

// the first entry of every tile against the entry before it (the last one of the previous tile of the same segment)
__global__ void load_seam_kernel (const LoadArgs a, uint32_t nTiles) {
  const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nTiles) return;
  uint32_t lo = 0, hi = a.nGrpSeg;
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.grpTile[mid] <= t) lo = mid; else hi = mid; }
  const uint32_t g = a.grpSeg[lo], k = t - a.grpTile[lo];
  const uint32_t c = g % a.C, sd = g / a.C, d = sd & 1u, s = sd >> 1, dc = d * a.C + c;
  const uint64_t seg0 = a.segStart[g];
  const uint64_t rel = (uint64_t)k * PK_TILE;
  const uint32_t p = a.pos[seg0 + rel];
  if (p < 1 || p > a.sizes[c]) return;                                 // already reported
  int64_t prevWin = -1;
  if (k > 0) {
    const uint32_t q = a.pos[seg0 + rel - 1];
    if (q >= p) { atomicOr(a.flags, FLAG_UNSORTED); return; }          // its histogram entry was counted by the tile kernel; the scatter pass does not run on invalid input
    if (q >= 1) prevWin = (int64_t)((q - 1) >> WIN_SHIFT);
  }
  load_open_windows(a, s, dc, prevWin, p - 1, (uint32_t)rel);
}

// window index default of the windows [w0, w1) of one (strand, contig range): "no entry at or after this window" = the segment's length
__global__ void load_win_fill_kernel (const LoadArgs a, uint32_t nSamples, uint32_t w0, uint32_t w1) {
  const uint64_t e = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t nw = w1 - w0;
  if (e >= (uint64_t)nSamples * nw) return;
  const uint32_t s = (uint32_t)(e / nw), gw = w0 + (uint32_t)(e % nw);
  uint32_t lo = 0, hi = 2 * a.C;    // largest dc with winBase[dc] <= gw (empty ranges are skipped by taking the largest)
  while (hi - lo > 1) { const uint32_t mid = (lo + hi) >> 1; if (a.winBase[mid] <= gw) lo = mid; else hi = mid; }
  const uint32_t d = lo / a.C, c = lo % a.C;
  const uint64_t g = (uint64_t)(s * 2 + d) * a.C + c;
  a.wEnt[(size_t)s * a.nWinTotal + gw] = (uint32_t)(a.segStart[g + 1] - a.segStart[g]);
}

// first / last position of every segment (the host keeps them for the reference-order logic)
__global__ void seg_ends_kernel (const uint32_t* __restrict__ pos, const uint64_t* __restrict__ segStart, uint32_t nSeg, uint32_t* __restrict__ firstPos, uint32_t* __restrict__ lastPos) {
  const uint32_t g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= nSeg) return;
  const uint64_t s = segStart[g], e = segStart[g + 1];
  firstPos[g] = e > s ? pos[s] : 0u;
  lastPos[g] = e > s ? pos[e - 1] : 0u;
}

__global__ void set_u64_kernel (unsigned long long* p, unsigned long long v) { if (blockIdx.x == 0 && threadIdx.x == 0) *p = v; }
__global__ void fill_u64_kernel (unsigned long long* p, uint32_t n, unsigned long long v) { const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) p[i] = v; }

}  // namespace upb
