// upb_api.cu — C-ABI (include/unipeak_b200.h) and host orchestration of the sm_100a kernels.
//
// There is no CPU implementation behind these entry points: every result row comes from the kernels in
// upb_device.cuh.  The host side only sizes buffers, launches, and assembles the reference-shaped
// result tables (src/regions.cpp, src/strand_shift.cpp, src/tags_in_regions.cpp drivers).
#include "../../include/unipeak_b200.h"
#include "upb_device.cuh"
#include "upb_span.cuh"
#include "upb_pack.cuh"
#include "upb_aggr.cuh"
#include "upb_introsort.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <string>
#include <vector>

using namespace upb;

// ------------------------------------------------------------------------------------------------
// context
// ------------------------------------------------------------------------------------------------
struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  template <class T> T* as () const { return (T*)p; }
};

struct HostSlab { void* p = nullptr; size_t cap = 0; };

struct upb_ctx {
  int device = 0;
  int smCount = 148;
  cudaStream_t stream = nullptr;
  cudaStream_t copyStream = nullptr;    // H2D copies of upb_load_tags run here, one sample ahead of the kernels that consume them
  cudaEvent_t copyEv = nullptr, doneEv = nullptr;
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  std::string err;
  // contigs
  std::vector<uint32_t> sizes;
  std::vector<uint32_t> nTilesIdx;      // per contig
  std::vector<uint32_t> tidxBase;       // [C+1]
  // tags
  uint32_t nSamples = 0;
  uint64_t nEntries = 0;
  bool loaded = false;
  std::vector<uint64_t> off;            // host copy
  std::vector<uint32_t> firstPos, lastPos;   // [(s*2+d)*C + c], 0 if empty
  DevBuf d_pos, d_cnt, d_off, d_segStart, d_tidx, d_tidxBase, d_sizes, d_nTilesIdx, d_small;
  DevBuf d_winBase, d_wEnt;                                          // per-sample window index of the scatter pass
  DevBuf d_pk16, d_pk8, d_pkExc, d_pkAgg, d_segEnds;                  // packed load (upb_pack.cuh)
  DevBuf d_ment, d_mtOff, d_mtBase, d_mtCount, d_mEff, d_mtW, d_blkW;   // tile-major copy and its zone maps (upb_device.cuh)
  std::vector<uint32_t> mtBase;         // [2*C+1]
  std::map<std::string, DevBuf> scratch;     // grow-only device scratch, by name
  unsigned char* h_small = nullptr;          // 4 KB of pinned host memory for small read-backs
  std::vector<HostSlab> slabPool;            // pinned host slabs of freed region tables, reused by the next call
  upb_stats stats;
  uint64_t lastRows = 0;                // rows of the last upb_call_regions still sitting in the device slab (scratch "slab")
  std::vector<uint8_t> lastControl;     // control set cntEff was built for
  bool cntEffValid = false;
  bool tidxBuilt = false;
  std::map<std::pair<const void*, size_t>, int> occCache;   // resident CTAs per SM of (kernel, dynamic shared memory)               // per-2048-tile entry offsets of the sample-major arrays: built on first use (density kernel only)
};

static thread_local std::string g_createErr;

static int fail (upb_ctx* ctx, int code, const std::string& msg) {
  if (ctx) ctx->err = msg; else g_createErr = msg;
  return code;
}

#define CU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { \
  return fail(ctx, UPB_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e__)); } } while (0)

static int ensure (upb_ctx* ctx, DevBuf& b, size_t bytes) {
  if (bytes <= b.cap) return UPB_OK;
  if (b.p) { cudaFree(b.p); b.p = nullptr; b.cap = 0; }
  size_t want = bytes + bytes / 4 + 256;
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) { want = bytes; e = cudaMalloc(&b.p, want); }
  if (e != cudaSuccess) { b.p = nullptr; return fail(ctx, UPB_ERR_NOMEM, std::string("cudaMalloc(") + std::to_string(bytes) + "): " + cudaGetErrorString(e)); }
  b.cap = want;
  return UPB_OK;
}
#define ENSURE(buf, bytes) do { int r__ = ensure(ctx, buf, bytes); if (r__) return r__; } while (0)

static void release (DevBuf& b) { if (b.p) cudaFree(b.p); b.p = nullptr; b.cap = 0; }

extern "C" int upb_device_count (void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

extern "C" void upb_destroy (upb_ctx* ctx);

extern "C" int upb_create (int device, upb_ctx** out) {
  if (!out) return fail(nullptr, UPB_ERR_BAD_ARG, "upb_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0) return fail(nullptr, UPB_ERR_NO_DEVICE, std::string("no CUDA device (there is no CPU path): ") + cudaGetErrorString(e));
  if (device < 0 || device >= n) return fail(nullptr, UPB_ERR_BAD_ARG, "upb_create: device index out of range");
  upb_ctx* ctx = new upb_ctx;
  ctx->device = device;
  if ((e = cudaSetDevice(device)) != cudaSuccess) { delete ctx; return fail(nullptr, UPB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return fail(nullptr, UPB_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e)); }
  ctx->smCount = prop.multiProcessorCount;
  if ((e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking)) != cudaSuccess) { delete ctx; return fail(nullptr, UPB_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
  for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
  cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->copyEv, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->doneEv, cudaEventDisableTiming);
  if ((e = cudaMallocHost((void**)&ctx->h_small, 4096)) != cudaSuccess) { upb_destroy(ctx); return fail(nullptr, UPB_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); }
  memset(&ctx->stats, 0, sizeof ctx->stats);
  *out = ctx;
  return UPB_OK;
}

extern "C" void upb_destroy (upb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  DevBuf* all[] = {&ctx->d_winBase, &ctx->d_wEnt, &ctx->d_pk16, &ctx->d_pk8, &ctx->d_pkExc, &ctx->d_pkAgg, &ctx->d_segEnds, &ctx->d_mtW, &ctx->d_blkW, &ctx->d_ment, &ctx->d_mtOff, &ctx->d_mtBase, &ctx->d_mtCount, &ctx->d_mEff, &ctx->d_pos, &ctx->d_cnt, &ctx->d_off, &ctx->d_segStart, &ctx->d_tidx, &ctx->d_tidxBase, &ctx->d_sizes,
    &ctx->d_nTilesIdx, &ctx->d_small};
  for (DevBuf* b : all) release(*b);
  for (auto& kv : ctx->scratch) release(kv.second);
  for (HostSlab& h : ctx->slabPool) if (h.p) cudaFreeHost(h.p);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->copyEv) cudaEventDestroy(ctx->copyEv);
  if (ctx->doneEv) cudaEventDestroy(ctx->doneEv);
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  delete ctx;
}

extern "C" const char* upb_last_error (const upb_ctx* ctx) { return ctx ? ctx->err.c_str() : g_createErr.c_str(); }

extern "C" int upb_host_alloc (void** out, size_t bytes) {
  if (!out) return UPB_ERR_BAD_ARG;
  cudaError_t e = cudaMallocHost(out, bytes ? bytes : 1);
  if (e != cudaSuccess) { *out = nullptr; return fail(nullptr, UPB_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); }
  return UPB_OK;
}
extern "C" void upb_host_free (void* p) { if (p) cudaFreeHost(p); }

extern "C" int upb_host_register (void* p, size_t bytes, void** device_ptr) {
  if (!p || !bytes) return UPB_ERR_BAD_ARG;
  cudaError_t e = cudaHostRegister(p, bytes, cudaHostRegisterPortable | cudaHostRegisterMapped);
  if (e != cudaSuccess) return fail(nullptr, UPB_ERR_CUDA, std::string("cudaHostRegister: ") + cudaGetErrorString(e));
  if (device_ptr) {
    e = cudaHostGetDevicePointer(device_ptr, p, 0);
    if (e != cudaSuccess) { cudaHostUnregister(p); return fail(nullptr, UPB_ERR_CUDA, std::string("cudaHostGetDevicePointer: ") + cudaGetErrorString(e)); }
  }
  return UPB_OK;
}
extern "C" void upb_host_unregister (void* p) { if (p) cudaHostUnregister(p); }

extern "C" void* upb_stream (const upb_ctx* ctx) { return ctx ? (void*)ctx->stream : nullptr; }

extern "C" int upb_last_stats (const upb_ctx* ctx, upb_stats* out) {
  if (!ctx || !out) return UPB_ERR_BAD_ARG;
  *out = ctx->stats;
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// Kernel taps — misc/kernel.cpp:12-35, same expression order, on the host
// ------------------------------------------------------------------------------------------------
extern "C" int upb_kernel_taps (uint32_t bandwidth, double sum, double* taps) {
  if (bandwidth == 0 || !taps) return UPB_ERR_BAD_ARG;
  const int bw = (int)bandwidth;
  volatile double acc = 0;   // volatile: keep every rounding step (no contraction, no reassociation)
  int n = 0;
  for (int i = -bw; i <= bw; ++i) {
    volatile double x = (double)i / (double)bw;
    volatile double xx = x * x;
    volatile double om = 1 - xx;
    volatile double th = 3 * om;
    taps[n] = th / 4;
    acc = acc + taps[n];
    ++n;
  }
  volatile double scale = sum / acc;
  for (int j = 0; j < n; ++j) { volatile double v = taps[j] * scale; taps[j] = v; }
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// inputs
// ------------------------------------------------------------------------------------------------
extern "C" int upb_set_contigs (upb_ctx* ctx, uint32_t n_contigs, const uint32_t* sizes) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (n_contigs == 0 || !sizes) return fail(ctx, UPB_ERR_BAD_ARG, "upb_set_contigs: empty contig table");
  CU(cudaSetDevice(ctx->device));
  ctx->sizes.assign(sizes, sizes + n_contigs);
  ctx->nTilesIdx.resize(n_contigs);
  ctx->tidxBase.resize(n_contigs + 1);
  uint64_t run = 0;
  for (uint32_t c = 0; c < n_contigs; ++c) {
    ctx->nTilesIdx[c] = (uint32_t)(((uint64_t)sizes[c] + TILE - 1) / TILE);
    ctx->tidxBase[c] = (uint32_t)run;
    run += (uint64_t)ctx->nTilesIdx[c] + 3;
    if (run > 0xFFFFFFFFull) return fail(ctx, UPB_ERR_UNSUPPORTED, "genome too large for the tile index");
  }
  ctx->tidxBase[n_contigs] = (uint32_t)run;
  ENSURE(ctx->d_sizes, n_contigs * 4);
  ENSURE(ctx->d_nTilesIdx, n_contigs * 4);
  ENSURE(ctx->d_tidxBase, (n_contigs + 1) * 4);
  CU(cudaMemcpyAsync(ctx->d_sizes.p, ctx->sizes.data(), n_contigs * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_nTilesIdx.p, ctx->nTilesIdx.data(), n_contigs * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_tidxBase.p, ctx->tidxBase.data(), (n_contigs + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->loaded = false;
  return UPB_OK;
}

static Tags make_tags (const upb_ctx* ctx, bool useEff) {
  Tags T;
  T.pos = ctx->d_pos.as<uint32_t>();
  T.cnt = ctx->d_cnt.as<uint32_t>();
  T.ment = ctx->d_ment.as<uint2>();
  T.mtOff = ctx->d_mtOff.as<unsigned long long>();
  T.mtBase = ctx->d_mtBase.as<uint32_t>();
  T.mEff = useEff ? ctx->d_mEff.as<uint32_t>() : nullptr;
  T.mtW = ctx->d_mtW.as<unsigned long long>();
  T.blkW = ctx->d_blkW.as<float>();
  T.off = ctx->d_off.as<uint64_t>();
  T.tidx = ctx->d_tidx.as<uint32_t>();
  T.tidxBase = ctx->d_tidxBase.as<uint32_t>();
  T.tidxStride = ctx->tidxBase.back();
  T.C = (uint32_t)ctx->sizes.size();
  T.nSamples = ctx->nSamples;
  return T;
}

static int scan_u32 (upb_ctx* ctx, const uint32_t* d_in, uint64_t n, unsigned long long* d_out, unsigned long long* d_total, uint32_t* launches);

// Both load paths: validate the offsets, bring pos[] / cnt[] to the device SAMPLE BY SAMPLE on the copy stream (plain: two
// copies; packed: the 3-byte form, decoded by upb_pack.cuh) with the validation + histogram pass of the counting sort running
// one sample behind on the compute stream, then the scan, the scatter pass and the zone maps.
struct PackedSrc { const uint16_t* dpos; const uint8_t* cnt8; uint64_t nExc; const upb_packed_exception* exc; bool twoByte; };   // twoByte: cnt8 unused

static int load_tags_impl (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt, const PackedSrc* pk) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (ctx->sizes.empty()) return fail(ctx, UPB_ERR_NOT_LOADED, "upb_load_tags: call upb_set_contigs first");
  if (n_samples == 0 || !off) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: no samples");
  CU(cudaSetDevice(ctx->device));
  const uint32_t C = (uint32_t)ctx->sizes.size();
  const size_t nOff = (size_t)n_samples * 2 * (C + 1);
  const uint32_t nSeg = n_samples * 2 * C;
  // offsets must describe one contiguous, monotone packing (what TagData / the CLI front end produce)
  std::vector<uint64_t> segStart(nSeg + 1);
  uint64_t expect = off[0];
  if (expect != 0) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: off[0] must be 0");
  for (uint32_t sd = 0; sd < n_samples * 2; ++sd) {
    const uint64_t* o = off + (size_t)sd * (C + 1);
    if (o[0] != expect) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: offset blocks must be contiguous");
    for (uint32_t c = 0; c < C; ++c) {
      if (o[c + 1] < o[c]) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: offsets must be non-decreasing");
      if (o[c + 1] - o[c] > 0xFFFFFFFFull) return fail(ctx, UPB_ERR_UNSUPPORTED, "upb_load_tags: more than 2^32 entries in one segment");
      segStart[(size_t)sd * C + c] = o[c];
    }
    expect = o[C];
  }
  const uint64_t E = expect;
  segStart[nSeg] = E;
  if (E && !pk && (!pos || !cnt)) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: pos/cnt are NULL");
  if (E && pk && (!pk->dpos || (!pk->twoByte && !pk->cnt8) || (pk->nExc && !pk->exc))) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_packed: NULL argument");
  if (n_samples > (1u << 20)) return fail(ctx, UPB_ERR_UNSUPPORTED, "more than 2^20 samples");
  ctx->loaded = false;
  ctx->nSamples = n_samples;
  ctx->nEntries = E;
  ctx->off.assign(off, off + nOff);
  ctx->firstPos.assign(nSeg, 0);
  ctx->lastPos.assign(nSeg, 0);
  if (!pk)
    for (uint32_t g = 0; g < nSeg; ++g)
      if (segStart[g + 1] > segStart[g]) { ctx->firstPos[g] = pos[segStart[g]]; ctx->lastPos[g] = pos[segStart[g + 1] - 1]; }

  ENSURE(ctx->d_pos, (E ? E : 1) * 4);
  ENSURE(ctx->d_cnt, (E ? E : 1) * 4);
  ENSURE(ctx->d_off, nOff * 8);
  ENSURE(ctx->d_segStart, ((size_t)nSeg + 1) * 8);
  ENSURE(ctx->d_small, 4096);
  ctx->mtBase.assign(2 * (size_t)C + 1, 0);
  { uint64_t run = 0;
    for (uint32_t d = 0; d < 2; ++d) for (uint32_t c = 0; c < C; ++c) { ctx->mtBase[(size_t)d * C + c] = (uint32_t)run; run += ((uint64_t)(ctx->sizes[c] - 1) >> MT_SHIFT) + 1; }
    if (run > 0xFFFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "genome too large for the mini-tile index");
    ctx->mtBase[2 * (size_t)C] = (uint32_t)run; }
  const uint32_t nMtTotal = ctx->mtBase.back();
  std::vector<uint32_t> winBase(2 * (size_t)C + 1, 0);
  { uint32_t run = 0;
    for (size_t dc = 0; dc < 2 * (size_t)C; ++dc) { winBase[dc] = run; run += (ctx->mtBase[dc + 1] - ctx->mtBase[dc] + WIN_MT - 1) / WIN_MT; }
    winBase[2 * (size_t)C] = run; }
  const uint32_t nWinTotal = winBase.back();
  ENSURE(ctx->d_mtBase, (2 * (size_t)C + 1) * 4);
  ENSURE(ctx->d_winBase, (2 * (size_t)C + 1) * 4);
  ENSURE(ctx->d_wEnt, ((size_t)n_samples * nWinTotal + 2) * 4);
  ENSURE(ctx->d_mtCount, ((size_t)nMtTotal + 1) * 4);
  ENSURE(ctx->d_mtOff, ((size_t)nMtTotal + 2) * 8);
  ENSURE(ctx->d_ment, (E ? E : 1) * 8);
  ENSURE(ctx->d_mtW, ((size_t)nMtTotal + 1) * 8);
  ENSURE(ctx->d_blkW, ((size_t)nMtTotal + 1) * BLK_PER_MT * 4);
  uint64_t maxChunk = 0;
  for (uint32_t s = 0; s < n_samples; ++s) maxChunk = std::max(maxChunk, segStart[(size_t)(s + 1) * 2 * C] - segStart[(size_t)s * 2 * C]);
  if (pk) {
    ENSURE(ctx->d_pk16, (E ? E : 1) * 2);
    if (!pk->twoByte) ENSURE(ctx->d_pk8, (E ? E : 1));
    ENSURE(ctx->d_pkExc, (pk->nExc ? pk->nExc : 1) * sizeof(PkException));
    ENSURE(ctx->d_pkAgg, 2 * (maxChunk / PK_TILE + 2) * sizeof(PkAgg));
    ENSURE(ctx->d_segEnds, 2 * ((size_t)nSeg + 1) * 4);
    if (pk->nExc) CU(cudaMemcpyAsync(ctx->d_pkExc.p, pk->exc, pk->nExc * sizeof(PkException), cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaMemcpyAsync(ctx->d_off.p, off, nOff * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_segStart.p, segStart.data(), ((size_t)nSeg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_small.p, 0, 4096, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_mtBase.p, ctx->mtBase.data(), (2 * (size_t)C + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_mtCount.p, 0, ((size_t)nMtTotal + 1) * 4, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_winBase.p, winBase.data(), (2 * (size_t)C + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  // the copy stream starts after whatever still reads the old arrays on the compute stream
  CU(cudaEventRecord(ctx->doneEv, ctx->stream));
  CU(cudaStreamWaitEvent(ctx->copyStream, ctx->doneEv, 0));

  MtArgs ma;
  ma.pos = ctx->d_pos.as<uint32_t>(); ma.cnt = ctx->d_cnt.as<uint32_t>(); ma.segStart = ctx->d_segStart.as<uint64_t>(); ma.mtBase = ctx->d_mtBase.as<uint32_t>();
  ma.mtCount = ctx->d_mtCount.as<uint32_t>(); ma.mtOff = ctx->d_mtOff.as<unsigned long long>(); ma.ment = ctx->d_ment.as<uint2>();
  ma.sizes = ctx->d_sizes.as<uint32_t>(); ma.flags = ctx->d_small.as<uint32_t>();
  ma.C = C; ma.nSeg = nSeg;
  ma.mtW = ctx->d_mtW.as<unsigned long long>(); ma.blkW = ctx->d_blkW.as<float>();
  ma.winBase = ctx->d_winBase.as<uint32_t>(); ma.wEnt = ctx->d_wEnt.as<uint32_t>(); ma.nWinTotal = nWinTotal;
  ma.first = 0; ma.nEntries = 0;
  win_fill_kernel<<<(unsigned)(((uint64_t)n_samples * nWinTotal + 255) / 256), 256, 0, ctx->stream>>>(ma, n_samples);
  // tile-major copy, pass 1 (validation + per-mini-tile histogram), one sample behind the copies
  for (uint32_t s = 0; s < n_samples; ++s) {
    const uint64_t first = segStart[(size_t)s * 2 * C], n = segStart[(size_t)(s + 1) * 2 * C] - first;
    if (!n) continue;
    if (!pk) {
      CU(cudaMemcpyAsync(ctx->d_pos.as<uint32_t>() + first, pos + first, n * 4, cudaMemcpyHostToDevice, ctx->copyStream));
      CU(cudaMemcpyAsync(ctx->d_cnt.as<uint32_t>() + first, cnt + first, n * 4, cudaMemcpyHostToDevice, ctx->copyStream));
    } else {
      CU(cudaMemcpyAsync(ctx->d_pk16.as<uint16_t>() + first, pk->dpos + first, n * 2, cudaMemcpyHostToDevice, ctx->copyStream));
      if (!pk->twoByte) CU(cudaMemcpyAsync(ctx->d_pk8.as<uint8_t>() + first, pk->cnt8 + first, n, cudaMemcpyHostToDevice, ctx->copyStream));
    }
    CU(cudaEventRecord(ctx->copyEv, ctx->copyStream));
    CU(cudaStreamWaitEvent(ctx->stream, ctx->copyEv, 0));
    if (pk) {
      PkArgs pa;
      pa.dpos = ctx->d_pk16.as<uint16_t>(); pa.cnt8 = pk->twoByte ? nullptr : ctx->d_pk8.as<uint8_t>(); pa.exc = ctx->d_pkExc.as<PkException>(); pa.nExc = pk->nExc;
      pa.dposMask = pk->twoByte ? PK2_DPOS_MASK : 0xFFFFu; pa.dposEsc = pk->twoByte ? PK2_DPOS_ESC : PK_DPOS_ESC; pa.cntEsc = pk->twoByte ? PK2_CNT_ESC : PK_CNT_ESC;
      pa.first = first; pa.E = n; pa.pos = ctx->d_pos.as<uint32_t>(); pa.cnt = ctx->d_cnt.as<uint32_t>(); pa.flags = ctx->d_small.as<uint32_t>();
      const uint32_t nTiles = (uint32_t)((n + PK_TILE - 1) / PK_TILE);
      pa.agg = ctx->d_pkAgg.as<PkAgg>(); pa.carry = pa.agg + (maxChunk / PK_TILE + 2);
      pk_tile_kernel<<<nTiles, PK_THREADS, 0, ctx->stream>>>(pa);
      pk_top_kernel<<<1, 1024, 0, ctx->stream>>>(pa.agg, nTiles, pa.carry);
      pk_final_kernel<<<nTiles, PK_THREADS, 0, ctx->stream>>>(pa);
    }
    ma.first = first; ma.nEntries = n;
    mt_build_kernel<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ma);
    CU(cudaGetLastError());
  }
  {
    uint32_t l = 0;
    unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 16;
    int r = scan_u32(ctx, ma.mtCount, nMtTotal, ctx->d_mtOff.as<unsigned long long>(), d_tot, &l);
    if (r) return r;
    CU(cudaMemcpyAsync(ctx->d_mtOff.as<unsigned long long>() + nMtTotal, d_tot, 8, cudaMemcpyDeviceToDevice, ctx->stream));
    ma.first = 0; ma.nEntries = E;
    if (E) mt_scatter_kernel<<<nWinTotal, 256, 0, ctx->stream>>>(ma, n_samples);
    CU(cudaGetLastError());
  }
  uint32_t flags = 0;
  CU(cudaMemcpyAsync(&flags, ctx->d_small.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (pk) {   // first / last position of every segment come back from the decoded arrays
    uint32_t* ends = ctx->d_segEnds.as<uint32_t>();
    seg_ends_kernel<<<(nSeg + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_pos.as<uint32_t>(), ctx->d_segStart.as<uint64_t>(), nSeg, ends, ends + nSeg);
    CU(cudaMemcpyAsync(ctx->firstPos.data(), ends, (size_t)nSeg * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(ctx->lastPos.data(), ends + nSeg, (size_t)nSeg * 4, cudaMemcpyDeviceToHost, ctx->stream));
  }
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->cntEffValid = false;
  ctx->tidxBuilt = false;
  if (flags & FLAG_BAD_PACK) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_packed: an escape code without its exception entry");
  if (flags & FLAG_UNSORTED) return fail(ctx, UPB_ERR_UNSORTED, "upb_load_tags: positions must be strictly ascending within each (sample, strand, contig) — the reference asserts on unsorted input (misc/peakcall.cpp:171)");
  if (flags & FLAG_BAD_POS) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: position outside 1..contig size (the reference's parser drops those, misc/format.cpp:670-676)");
  if (flags & FLAG_ZERO_COUNT) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: zero count (the reference's parser never delivers those, misc/format.cpp:710)");
  ctx->loaded = true;
  ctx->stats.n_entries = E;
  return UPB_OK;
}

extern "C" int upb_load_tags (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt) {
  return load_tags_impl(ctx, n_samples, off, pos, cnt, nullptr);
}

extern "C" int upb_load_tags_packed (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint16_t* dpos, const uint8_t* cnt8,
                                     uint64_t n_exceptions, const upb_packed_exception* exceptions) {
  PackedSrc pk; pk.dpos = dpos; pk.cnt8 = cnt8; pk.nExc = n_exceptions; pk.exc = exceptions; pk.twoByte = false;
  return load_tags_impl(ctx, n_samples, off, nullptr, nullptr, &pk);
}

extern "C" int upb_load_tags_packed2 (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint16_t* words,
                                      uint64_t n_exceptions, const upb_packed_exception* exceptions) {
  PackedSrc pk; pk.dpos = words; pk.cnt8 = nullptr; pk.nExc = n_exceptions; pk.exc = exceptions; pk.twoByte = true;
  return load_tags_impl(ctx, n_samples, off, nullptr, nullptr, &pk);
}

// host-only encoder of the packed form (what a front end does while it parses)
// one encoder for both packed forms: dposEsc / cntEsc are the escape codes, cntShift != 0 puts the count into the high bits of the u16
static int pack_tags_impl (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                           uint16_t* dpos_out, uint8_t* cnt8_out, upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out,
                           uint32_t dposEsc, uint32_t cntEsc, uint32_t cntShift) {
  if (!off || !n_exc_out || (exc_cap && !exc_out)) return UPB_ERR_BAD_ARG;
  uint64_t nExc = 0;
  for (uint32_t sd = 0; sd < n_samples * 2; ++sd)
    for (uint32_t c = 0; c < n_contigs; ++c) {
      const uint64_t a = off[(size_t)sd * (n_contigs + 1) + c], b = off[(size_t)sd * (n_contigs + 1) + c + 1];
      uint32_t prev = 0;
      for (uint64_t e = a; e < b; ++e) {
        const uint32_t p = pos[e], k = cnt[e];
        const bool escP = e == a || p < prev || p - prev >= dposEsc;   // unsorted input is kept verbatim: the load then reports it
        const bool escC = k >= cntEsc;
        if (escP || escC) {
          if (nExc < exc_cap) { exc_out[nExc].index = e; exc_out[nExc].pos = escP ? p : 0u; exc_out[nExc].cnt = escC ? k : 0u; }
          ++nExc;
        }
        const uint32_t dv = escP ? dposEsc : p - prev, cv = escC ? cntEsc : k;
        if (cntShift) { if (dpos_out) dpos_out[e] = (uint16_t)(dv | (cv << cntShift)); }
        else {
          if (dpos_out) dpos_out[e] = (uint16_t)dv;
          if (cnt8_out) cnt8_out[e] = (uint8_t)cv;
        }
        prev = p;
      }
    }
  *n_exc_out = nExc;
  return nExc <= exc_cap ? UPB_OK : UPB_ERR_NOMEM;
}

extern "C" int upb_pack_tags (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                              uint16_t* dpos_out, uint8_t* cnt8_out, upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out) {
  return pack_tags_impl(n_samples, n_contigs, off, pos, cnt, dpos_out, cnt8_out, exc_out, exc_cap, n_exc_out, PK_DPOS_ESC, PK_CNT_ESC, 0);
}

extern "C" int upb_pack_tags2 (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                               uint16_t* words_out, upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out) {
  return pack_tags_impl(n_samples, n_contigs, off, pos, cnt, words_out, nullptr, exc_out, exc_cap, n_exc_out, PK2_DPOS_ESC, PK2_CNT_ESC, PK2_CNT_SHIFT);
}

// ------------------------------------------------------------------------------------------------
// device-wide exclusive scan (uint32 in, uint64 out); total -> *d_total
// ------------------------------------------------------------------------------------------------
#define SCR(name, T, count) ({ DevBuf& b__ = ctx->scratch[name]; int r__ = ensure(ctx, b__, (size_t)(count) * sizeof(T) + 16); if (r__) return r__; (T*)b__.p; })

static int scan_u32 (upb_ctx* ctx, const uint32_t* d_in, uint64_t n, unsigned long long* d_out, unsigned long long* d_total, uint32_t* launches) {
  const uint32_t nBlocks = (uint32_t)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  unsigned long long* part = SCR("scanPartial", unsigned long long, (size_t)nBlocks + 1);
  if (nBlocks == 0) { CU(cudaMemsetAsync(d_total, 0, 8, ctx->stream)); return UPB_OK; }
  scan_partial_kernel<<<nBlocks, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, part);
  scan_top_kernel<<<1, SCAN_THREADS, 0, ctx->stream>>>(part, nBlocks, d_total);
  scan_final_kernel<<<nBlocks, SCAN_THREADS, 0, ctx->stream>>>(d_in, n, part, d_out);
  CU(cudaGetLastError());
  *launches += 3;
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// the region pipeline shared by bin/regions and bin/strand_shift
// ------------------------------------------------------------------------------------------------
struct PipeCfg {
  uint32_t bw;
  double thr, kurtThr, corrThr, hitThr, background;
  bool nondir;
  bool needCorr;                     // compute strandCorr(0) for regions passing the earlier filters
  const uint8_t* control;            // may be NULL
  bool deviceOnly = false;           // leave the rows in the device slab (multi-GPU ranks: upb_scatter_rows reads them there)
  bool allowLeftEdge = false;        // replay the reference's left-edge behaviour (Q2) instead of refusing such inputs
  uint32_t nCoeffs = 0;              // 0, or the number of non-control samples (-z)
  const double* coeffs = nullptr;    // already normalised as src/regions.cpp:216-229
};

// column layout of one result slab (device copy and pinned host copy share it)
struct SlabLayout {
  size_t oContig, oLeft, oLen, oPeak, oWsum, oCandIndex, oPeakScore, oKurt, oCorr, oCounts, oBuffer, devBytes;
  size_t oLabel, oDropped, oTags, oUnitCand, oUnitRows, oUnitLast, totalBytes;   // host-only tail
};

static SlabLayout slab_layout (size_t R, uint32_t n, uint32_t C) {
  SlabLayout L;
  const size_t Rp = (R + 7) & ~(size_t)7;                            // keeps every column 8-byte aligned
  size_t o = 0;
  L.oCandIndex = o; o += Rp * 8;
  L.oPeakScore = o; o += Rp * 8;
  L.oKurt = o;      o += Rp * 8;
  L.oCorr = o;      o += Rp * 8;
  L.oContig = o;    o += Rp * 4;
  L.oLeft = o;      o += Rp * 4;
  L.oLen = o;       o += Rp * 4;
  L.oPeak = o;      o += Rp * 4;
  L.oWsum = o;      o += Rp * 4;
  L.oCounts = o;    o += ((R * n + 7) & ~(size_t)7) * 4;
  L.oBuffer = o;    o += Rp;
  L.devBytes = o;
  L.oLabel = o;     o += Rp;
  L.oDropped = o;   o += Rp;
  L.oTags = o;      o += (size_t)n * 8;
  L.oUnitCand = o;  o += 2 * (size_t)C * 8;
  L.oUnitRows = o;  o += 2 * (size_t)C * 8;
  L.oUnitLast = o;  o += ((2 * (size_t)C + 1) & ~(size_t)1) * 4;
  L.totalBytes = o + 64;
  return L;
}

struct PipeOut {
  std::vector<Unit> units;
  std::vector<uint64_t> unitCand;    // [nUnits]
  uint64_t nCand = 0;
  uint64_t R = 0;                    // accepted regions, canonical order (buffer, contig, left)
  std::vector<uint32_t> unitEdgeStop; // [nUnits] last position the left-edge replay owns in this unit, 0 if none
  HostSlab slab;                     // pinned; columns per `layout`
  SlabLayout layout;
  template <class T> T* col (size_t off) const { return (T*)((unsigned char*)slab.p + off); }
};

static int acquire_slab (upb_ctx* ctx, size_t bytes, HostSlab& out) {
  size_t best = (size_t)-1;
  for (size_t i = 0; i < ctx->slabPool.size(); ++i)
    if (ctx->slabPool[i].cap >= bytes && (best == (size_t)-1 || ctx->slabPool[i].cap < ctx->slabPool[best].cap)) best = i;
  if (best != (size_t)-1) { out = ctx->slabPool[best]; ctx->slabPool.erase(ctx->slabPool.begin() + best); return UPB_OK; }
  // drop the smaller ones: a pool that only ever grows would pin memory for ever
  for (HostSlab& h : ctx->slabPool) if (h.p) cudaFreeHost(h.p);
  ctx->slabPool.clear();
  const size_t want = bytes + bytes / 4 + 4096;
  cudaError_t e = cudaMallocHost(&out.p, want);
  if (e != cudaSuccess) { out.p = nullptr; out.cap = 0; return fail(ctx, UPB_ERR_NOMEM, std::string("cudaMallocHost(") + std::to_string(want) + "): " + cudaGetErrorString(e)); }
  out.cap = want;
  return UPB_OK;
}

static int check_left_edge (upb_ctx* ctx, uint32_t bw) {
  // every add() call moves the window, control samples included (src/regions.cpp:356-359)
  const uint32_t nSeg = (uint32_t)ctx->firstPos.size();
  for (uint32_t g = 0; g < nSeg; ++g)
    if (ctx->firstPos[g] != 0 && ctx->firstPos[g] <= bw)
      return fail(ctx, UPB_ERR_LEFT_EDGE, "a tag lies within `bandwidth` positions of a contig start: the reference's left-edge path (misc/peakcall.cpp:176-184, SURVEY Q2) is not emulated");
  return UPB_OK;
}

struct Prep {
  std::vector<Unit> units;
  std::vector<double> taps;
  std::vector<uint32_t> affected;    // units with a tag within `bandwidth` of the contig start: replayed by left_edge_kernel
  double maxCoeff = 0;               // -z: largest coefficient (bounds the weight of a position)
  const double* d_coeffBySample = nullptr;
  uint32_t totalSpans = 0;
  uint32_t nSrc = 0;
  bool haveControl = false;
  Tags T;
  DensArgs dens;                     // template for the density kernel
};

// common set-up: units, taps, source list, control filter
static int prepare (upb_ctx* ctx, const PipeCfg& cfg, Prep& pr, uint32_t* launches, bool allowLeftEdge = false) {
  const uint32_t C = (uint32_t)ctx->sizes.size();
  const uint32_t n = ctx->nSamples;
  if (!ctx->loaded) return fail(ctx, UPB_ERR_NOT_LOADED, "no tags loaded (upb_set_contigs + upb_load_tags first)");
  if (cfg.bw == 0 || cfg.bw > (uint32_t)MAX_BW_FAST) return fail(ctx, UPB_ERR_UNSUPPORTED, "bandwidth must be in 1..448");
  if (!(cfg.thr > 0)) return fail(ctx, UPB_ERR_BAD_ARG, "region threshold must be > 0 (the reference's leap handling relies on it, SURVEY Q10)");
  if (!(cfg.background > 0)) return fail(ctx, UPB_ERR_BAD_ARG, "background must be > 0");
  int r;
  if (!allowLeftEdge && (r = check_left_edge(ctx, cfg.bw))) return r;

  std::vector<uint32_t> src;
  std::vector<uint8_t> control(n, 0);
  bool haveControl = false;
  for (uint32_t s = 0; s < n; ++s) { if (cfg.control && cfg.control[s]) { control[s] = 1; haveControl = true; } else src.push_back(s); }
  pr.haveControl = haveControl;
  pr.nSrc = (uint32_t)src.size();
  std::vector<double> coeffBySample(n, 0.0), coeffBySrc(src.size(), 0.0);
  pr.maxCoeff = 0; pr.d_coeffBySample = nullptr;
  if (cfg.nCoeffs) {
    if (cfg.nCoeffs != src.size() || !cfg.coeffs) return fail(ctx, UPB_ERR_BAD_ARG, "n_coeffs must equal the number of non-control samples (src/regions.cpp:221)");
    for (size_t k = 0; k < src.size(); ++k) {
      if (!(cfg.coeffs[k] >= 0)) return fail(ctx, UPB_ERR_BAD_ARG, "scaling coefficients must be >= 0");
      coeffBySample[src[k]] = cfg.coeffs[k]; coeffBySrc[k] = cfg.coeffs[k];
      pr.maxCoeff = std::max(pr.maxCoeff, cfg.coeffs[k]);
    }
  }

  // units: one per (buffer, contig) that holds at least one non-control tag
  pr.units.clear();
  pr.affected.clear();
  uint64_t spanStart = 0;
  const uint32_t nBuf = cfg.nondir ? 1u : 2u;
  for (uint32_t b = 0; b < nBuf; ++b)
    for (uint32_t c = 0; c < C; ++c) {
      bool any = false;
      for (uint32_t s : src) {
        if (cfg.nondir) any = any || ctx->firstPos[(size_t)(s * 2 + 0) * C + c] || ctx->firstPos[(size_t)(s * 2 + 1) * C + c];
        else any = any || ctx->firstPos[(size_t)(s * 2 + b) * C + c];
      }
      if (!any) continue;
      Unit u;
      u.contig = c; u.buffer = b; u.size = ctx->sizes[c];
      u.nSpans = (uint32_t)(((uint64_t)ctx->sizes[c] + cfg.bw + SPAN_LEN - 1) / SPAN_LEN);
      u.spanStart = (uint32_t)spanStart;
      spanStart += u.nSpans;
      if (spanStart > 0xFFFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many spans");
      // Q2: every add() moves the window, control samples included (src/regions.cpp:356-359)
      uint32_t first = 0xFFFFFFFFu, last = 0;
      for (uint32_t s = 0; s < n; ++s) for (uint32_t d = (cfg.nondir ? 0u : b); d <= (cfg.nondir ? 1u : b); ++d) {
        const uint32_t fp = ctx->firstPos[(size_t)(s * 2 + d) * C + c], lp = ctx->lastPos[(size_t)(s * 2 + d) * C + c];
        if (fp) { first = std::min(first, fp); last = std::max(last, lp); }
      }
      if (first <= cfg.bw) {
        if (last <= cfg.bw) return fail(ctx, UPB_ERR_LEFT_EDGE, "every tag of a contig lies within `bandwidth` of its start: the reference then leaks window cells into the next contig (misc/peakcall.cpp:176-184,224-231), which is not emulated");
        pr.affected.push_back((uint32_t)pr.units.size());
      }
      pr.units.push_back(u);
    }
  pr.totalSpans = (uint32_t)spanStart;

  pr.taps.assign(2 * (size_t)cfg.bw + 1, 0.0);
  upb_kernel_taps(cfg.bw, 1 / cfg.background, pr.taps.data());      // src/regions.cpp:232
  double* d_taps = SCR("taps", double, pr.taps.size());
  uint32_t* d_src = SCR("src", uint32_t, src.size() + 1);
  Unit* d_units = SCR("units", Unit, pr.units.size() + 1);
  uint8_t* d_control = SCR("control", uint8_t, n + 1);
  ENSURE(ctx->d_small, 4096);
  CU(cudaMemcpyAsync(d_taps, pr.taps.data(), pr.taps.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (!src.empty()) CU(cudaMemcpyAsync(d_src, src.data(), src.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!pr.units.empty()) CU(cudaMemcpyAsync(d_units, pr.units.data(), pr.units.size() * sizeof(Unit), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_control, control.data(), n, cudaMemcpyHostToDevice, ctx->stream));
  double* d_coefSrc = nullptr;
  if (cfg.nCoeffs) {
    double* d_coefSample = SCR("coefSample", double, n);
    d_coefSrc = SCR("coefSrc", double, src.size() + 1);
    CU(cudaMemcpyAsync(d_coefSample, coeffBySample.data(), (size_t)n * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(d_coefSrc, coeffBySrc.data(), src.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));                             // the host vectors go out of scope
    pr.d_coeffBySample = d_coefSample;
  }
  CU(cudaMemsetAsync(ctx->d_small.p, 0, 4096, ctx->stream));

  // Q4: zero the control tags the reference forgets
  if (haveControl && !(ctx->cntEffValid && ctx->lastControl == control)) {
    ENSURE(ctx->d_mEff, (ctx->nEntries ? ctx->nEntries : 1) * 4);
    const uint32_t nMtTotal = ctx->mtBase.back();
    if (ctx->nEntries) {
      mt_control_kernel<<<(nMtTotal + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_ment.as<uint2>(), ctx->d_mtOff.as<unsigned long long>(), nMtTotal, d_control,
                                                                        ctx->d_mEff.as<uint32_t>());
      CU(cudaGetLastError());
      *launches += 1;
    }
    ctx->lastControl = control;
    ctx->cntEffValid = true;
  }
  pr.T = make_tags(ctx, haveControl);

  DensArgs& da = pr.dens;
  memset(&da, 0, sizeof da);
  da.T = pr.T;
  da.units = d_units;
  da.nTilesIdx = ctx->d_nTilesIdx.as<uint32_t>();
  da.srcSample = d_src;
  da.nSrc = (uint32_t)src.size();
  da.bw = cfg.bw;
  da.taps = d_taps;
  da.srcCoeff = d_coefSrc;
  // d_small layout (4 KB): [0] flags u32, [1] ticket u32, [2] runCount u32, [3] multiCount u32, [4..11] span stats u64 x4, [16..] scan totals u64
  da.ticket = ctx->d_small.as<uint32_t>() + 1;
  return UPB_OK;
}

static int ensure_tidx (upb_ctx* ctx, uint32_t* launches) {
  if (ctx->tidxBuilt) return UPB_OK;
  const uint32_t C = (uint32_t)ctx->sizes.size(), n = ctx->nSamples;
  const uint64_t stride = ctx->tidxBase.back();
  const uint64_t idxTotal = stride * n * 2;
  ENSURE(ctx->d_tidx, idxTotal * 4);
  IndexArgs ia;
  ia.pos = ctx->d_pos.as<uint32_t>(); ia.cnt = ctx->d_cnt.as<uint32_t>(); ia.segStart = ctx->d_segStart.as<uint64_t>();
  ia.sizes = ctx->d_sizes.as<uint32_t>(); ia.tidxBase = ctx->d_tidxBase.as<uint32_t>(); ia.tidx = ctx->d_tidx.as<uint32_t>();
  ia.tidxStride = stride; ia.C = C; ia.nSeg = n * 2 * C; ia.nEntries = ctx->nEntries;
  ia.flags = ctx->d_small.as<uint32_t>() + 15;                        // validation already happened in mt_build_kernel
  index_fill_kernel<<<(unsigned)((idxTotal + 255) / 256), 256, 0, ctx->stream>>>(ia, idxTotal);
  if (ctx->nEntries) index_build_kernel<<<(unsigned)((ctx->nEntries + 255) / 256), 256, 0, ctx->stream>>>(ia);
  CU(cudaGetLastError());
  *launches += 2;
  ctx->tidxBuilt = true;
  return UPB_OK;
}

// largest dynamic shared memory size each kernel has been opted into, per device — the attribute belongs to the function, not to
// a context, and setting a smaller value would lower it under another context's feet: only ever raised
static std::map<std::pair<int, const void*>, size_t> g_smemRaised;
static std::mutex g_smemMutex;

template <class K>
static int cached_occupancy (upb_ctx* ctx, K kern, int threads, size_t smem, int* perSm) {
  const std::pair<const void*, size_t> key((const void*)kern, smem);
  if (smem > 48 * 1024) {
    std::lock_guard<std::mutex> lock(g_smemMutex);
    size_t& raised = g_smemRaised[std::make_pair(ctx->device, (const void*)kern)];
    if (smem > raised) { CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); raised = smem; }
  }
  auto it = ctx->occCache.find(key);
  if (it != ctx->occCache.end()) { *perSm = it->second; return UPB_OK; }
  int v = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&v, kern, threads, smem));
  if (v < 1) v = 1;
  ctx->occCache[key] = v;
  *perSm = v;
  return UPB_OK;
}

static int launch_density (upb_ctx* ctx, DensArgs& da, bool nondir, uint32_t totalWork, uint32_t* launches) {
  { int rr = ensure_tidx(ctx, launches); if (rr) return rr; da.T.tidx = ctx->d_tidx.as<uint32_t>(); }
  const DensSmem L = dens_smem_layout(da.bw, da.nSrc * (nondir ? 2u : 1u), nondir, da.srcCoeff != nullptr);
  const size_t smem = L.total;
  if (smem > 227 * 1024) return fail(ctx, UPB_ERR_UNSUPPORTED, "bandwidth / sample count need more shared memory than one SM has");
  void (*kern)(const DensArgs) = nondir ? density_kernel<true> : density_kernel<false>;
  int perSm = 0;
  { int rr = cached_occupancy(ctx, kern, KDE_THREADS, smem, &perSm); if (rr) return rr; }
  const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)ctx->smCount * perSm, std::max<uint32_t>(totalWork, 1u));
  kern<<<grid, KDE_THREADS, smem, ctx->stream>>>(da);
  CU(cudaGetLastError());
  *launches += 1;
  return UPB_OK;
}

// exact F/R densities for `n` position ranges (descriptor arrays dUnit/dLeft/dLen/dChunks already filled on the
// device) and, if asked, strandCorr(shift) for shifts 0..nShifts-1 into scratch "corr"
static int dump_and_corr (upb_ctx* ctx, const DensArgs& base, bool nondir, uint32_t n, uint32_t nShifts, uint32_t* launches) {
  unsigned long long* dOff = SCR("dOff", unsigned long long, (size_t)n + 1);
  unsigned long long* dChunkStart = SCR("dChunkStart", unsigned long long, (size_t)n + 1);
  uint32_t* dUnit = ctx->scratch["dUnit"].as<uint32_t>(); uint32_t* dLeft = ctx->scratch["dLeft"].as<uint32_t>();
  uint32_t* dLen = ctx->scratch["dLen"].as<uint32_t>(); uint32_t* dChunks = ctx->scratch["dChunks"].as<uint32_t>();
  unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 16;
  int r;
  if ((r = scan_u32(ctx, dLen, n, dOff, d_tot, launches))) return r;
  if ((r = scan_u32(ctx, dChunks, n, dChunkStart, d_tot + 1, launches))) return r;
  unsigned long long* tot = (unsigned long long*)ctx->h_small;
  CU(cudaMemcpyAsync(tot, d_tot, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const unsigned long long totLen = tot[0], totChunks = tot[1];
  // append the grand totals as entry n of the two prefix arrays
  CU(cudaMemcpyAsync(dOff + n, d_tot, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dChunkStart + n, d_tot + 1, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  if (totChunks > 0xFFFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many positions to evaluate in one batch");
  double* F = SCR("F", double, (size_t)totLen + 1);
  double* R = SCR("R", double, (size_t)totLen + 1);
  DensArgs da = base;
  da.dumpUnit = dUnit; da.dumpLeft = dLeft; da.dumpLen = dLen;
  da.dumpOff = (const uint64_t*)dOff; da.dumpChunkStart = (const uint64_t*)dChunkStart; da.nDump = n;
  da.dumpF = F; da.dumpR = R;
  CU(cudaMemsetAsync(da.ticket, 0, 4, ctx->stream));
  if ((r = launch_density(ctx, da, nondir, (uint32_t)totChunks, launches))) return r;
  if (nShifts) {
    double* corr = SCR("corr", double, (size_t)n * nShifts);
    CorrArgs ca;
    ca.F = F; ca.R = R; ca.off = (const uint64_t*)dOff; ca.len = dLen; ca.n = n; ca.nShifts = nShifts; ca.corr = corr;
    const uint64_t warps = (uint64_t)n * nShifts;
    const uint64_t blocks = (warps * 32 + 255) / 256;
    strand_corr_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ca);
    CU(cudaGetLastError());
    *launches += 1;
  }
  return UPB_OK;
}

static int launch_span (upb_ctx* ctx, const SpanArgs& sa, bool nondir, uint32_t* launches) {
  const bool coeff = sa.coeff != nullptr;
  const SpanSmem L = span_smem_layout(sa.bw, sa.nSamples, nondir, coeff);
  const size_t smem = L.total;
  if (smem > 227 * 1024) return fail(ctx, UPB_ERR_UNSUPPORTED, "bandwidth / sample count need more shared memory than one SM has");
  void (*kern)(const SpanArgs) = coeff ? (nondir ? span_kernel<true, true> : span_kernel<false, true>) : (nondir ? span_kernel<true, false> : span_kernel<false, false>);
  int perSm = 0;
  { int rr = cached_occupancy(ctx, kern, SP_THREADS, smem, &perSm); if (rr) return rr; }
  if (!coeff && perSm < SP_MIN_BLOCKS) {                              // shared memory admits fewer CTAs than the 64-register build is sized for: take the build without that cap
    kern = nondir ? span_kernel<true, false, SP_MIN_BLOCKS - 1> : span_kernel<false, false, SP_MIN_BLOCKS - 1>;
    int rr = cached_occupancy(ctx, kern, SP_THREADS, smem, &perSm); if (rr) return rr;
  }
  const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)ctx->smCount * perSm, std::max<uint32_t>(sa.totalSpans, 1u));
  kern<<<grid, SP_THREADS, smem, ctx->stream>>>(sa);
  CU(cudaGetLastError());
  *launches += 1;
  return UPB_OK;
}

static void finish_stats (upb_ctx* ctx, uint32_t launches) {
  cudaEventRecord(ctx->ev[3], ctx->stream);
  cudaEventSynchronize(ctx->ev[3]);
  float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]);
  ctx->stats.ms_total = ms;
  ctx->stats.launches = launches;
}

static int run_pipeline (upb_ctx* ctx, const PipeCfg& cfg, Prep& pr, PipeOut& out) {
  CU(cudaSetDevice(ctx->device));
  uint32_t launches = 0;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  int r = prepare(ctx, cfg, pr, &launches, cfg.allowLeftEdge);
  if (r) return r;
  const uint32_t n = ctx->nSamples, C = (uint32_t)ctx->sizes.size();
  const uint32_t nUnits = (uint32_t)pr.units.size();
  out.units = pr.units;
  out.unitCand.assign(nUnits, 0);
  out.unitEdgeStop.assign(nUnits, 0);
  out.nCand = 0; out.R = 0;
  ctx->lastRows = 0;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  ctx->stats.n_entries = ctx->nEntries;
  ctx->stats.kde_tiles = pr.totalSpans;
  out.layout = slab_layout(0, n, C);
  if (nUnits == 0 || pr.nSrc == 0) {
    if ((r = acquire_slab(ctx, out.layout.totalBytes, out.slab))) return r;
    memset(out.slab.p, 0, out.layout.totalBytes);
    finish_stats(ctx, launches);
    return UPB_OK;
  }
  uint32_t* small32 = ctx->d_small.as<uint32_t>();
  unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 16;
  uint32_t* h32 = (uint32_t*)ctx->h_small;
  Unit* d_units = ctx->scratch["units"].as<Unit>();

  // ---- K1: the fused span kernel ---------------------------------------------------------------------
  SpanArgs sa;
  memset(&sa, 0, sizeof sa);
  sa.T = pr.T; sa.units = d_units; sa.nTilesIdx = ctx->d_nTilesIdx.as<uint32_t>();
  sa.control = ctx->scratch["control"].as<uint8_t>(); sa.nSamples = n; sa.hasControl = pr.haveControl ? 1 : 0;
  sa.nUnits = nUnits; sa.totalSpans = pr.totalSpans; sa.bw = cfg.bw;
  sa.taps = ctx->scratch["taps"].as<double>();
  sa.thr = cfg.thr;
  const double kmax = pr.taps[cfg.bw];                                // K(0) = 0.75 * scale
  HotArgs ha;
  memset(&ha, 0, sizeof ha);
  {
    // cold-block bound: a tag j blocks (of 32 positions) away from block B is at least 32(|j|-1)+1 positions from every position of
    // B (0 for j = 0); blocks further than J = (bw-2)/32+1 away cannot reach it (the taps at +-bw are exactly 0, misc/kernel.cpp:12-14)
    const int bwI = (int)cfg.bw, J = bwI >= 2 ? (bwI - 2) / 32 + 1 : 0;
    std::vector<float> hf(J + 1);
    for (int j = 0; j <= J; ++j) {
      const int m = j == 0 ? 0 : 32 * (j - 1) + 1;
      const double hc = pr.taps[bwI - m] * (1.0 + pr.maxCoeff) * (1.0 + 1e-9);   // -z: a count c weighs at most c*(coeff+1)
      hf[j] = std::nextafter((float)(hc * (1.0 + 1e-5)), INFINITY);              // float sum of <= 58 terms: 1e-5 covers its rounding
    }
    float* d_hf = SCR("hotCoefF", float, J + 1);
    CU(cudaMemcpyAsync(d_hf, hf.data(), (size_t)(J + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));   // pageable source: staged before the call returns
    ha.hotCoefF = d_hf; ha.hotJ = (uint32_t)J;
    ha.thrLow = std::nextafter((float)cfg.thr, -INFINITY);
  }
  const double flatBound = kmax * (1.0 + pr.maxCoeff) * (1.0 + 1e-9);  // largest tap: maxTap * window weight bounds every score (zone-map prefilter)
  sa.coeff = pr.d_coeffBySample;
  sa.cN = kmax / ((double)cfg.bw * (double)cfg.bw);
  sa.dS = kmax * 1e-11;                                               // >= 100x the worst-case rounding difference between the two forms
  sa.flags = small32; sa.ticket = small32 + 1; sa.runCount = small32 + 2; sa.runReal = small32 + 3; sa.stat = (unsigned long long*)(small32 + 4);
  sa.spanRunCount = SCR("spanRunCount", uint32_t, pr.totalSpans);
  WarmRec* warmList = SCR("warmList", WarmRec, pr.totalSpans);
  uint2* hotList = SCR("hotList", uint2, pr.totalSpans);
  uint32_t* hotMask = SCR("hotMask", uint32_t, (size_t)pr.totalSpans * SPAN_WORDS);
  sa.hotList = hotList; sa.nHot = small32 + 14; sa.hotMask = hotMask;
  ha.T = pr.T; ha.units = d_units; ha.warmList = warmList; ha.nWarm = small32 + 12; ha.bw = cfg.bw; ha.nondir = cfg.nondir ? 1 : 0;
  ha.hotMask = hotMask; ha.hotList = hotList; ha.nHot = small32 + 14; ha.spanRunCount = sa.spanRunCount; ha.stat = sa.stat;
  WarmArgs wa;
  wa.T = pr.T; wa.units = d_units; wa.nUnits = nUnits; wa.totalSpans = pr.totalSpans; wa.nondir = cfg.nondir ? 1 : 0;
  wa.nAdj = ((31 + 2 * cfg.bw) >> MT_SHIFT) + 2;
  wa.warmList = warmList; wa.nWarm = small32 + 12; wa.spanRunCount = sa.spanRunCount; wa.bw = cfg.bw;

  // ---- K0: left-edge replay for the units that need it ------------------------------------
  const uint32_t nAff = (uint32_t)pr.affected.size();
  const uint32_t EDGE_REG_CAP = 1u << 18, EDGE_ROW_CAP = 1024;
  EdgeArgs eg;
  memset(&eg, 0, sizeof eg);
  uint32_t* d_affected = nullptr;
  std::vector<uint32_t> edgeStop(nAff, 0), edgeCand(nAff, 0), edgeRows(nAff, 0);
  if (nAff) {
    const uint32_t w = 2 * cfg.bw + 1;
    d_affected = SCR("edgeAffected", uint32_t, nAff);
    CU(cudaMemcpyAsync(d_affected, pr.affected.data(), (size_t)nAff * 4, cudaMemcpyHostToDevice, ctx->stream));
    eg.T = pr.T; eg.units = d_units; eg.affected = d_affected; eg.nAffected = nAff; eg.control = ctx->scratch["control"].as<uint8_t>();
    eg.coeffBySample = pr.d_coeffBySample; eg.taps = ctx->scratch["taps"].as<double>(); eg.bw = cfg.bw; eg.nondir = cfg.nondir ? 1 : 0;
    eg.thr = cfg.thr; eg.kurtThr = cfg.kurtThr; eg.corrThr = cfg.nondir ? cfg.corrThr : -1.0; eg.hitThr = cfg.hitThr;
    eg.cursors = SCR("edgeCursors", unsigned long long, (size_t)nAff * 2 * n);
    eg.ringF = SCR("edgeRingF", double, (size_t)nAff * w); eg.ringR = SCR("edgeRingR", double, (size_t)nAff * w);
    eg.ringHits = SCR("edgeRingHits", uint32_t, (size_t)nAff * w * n); eg.ringHas = SCR("edgeRingHas", uint8_t, (size_t)nAff * w);
    eg.regPc = SCR("edgeRegPc", uint32_t, (size_t)nAff * EDGE_REG_CAP); eg.regHas = SCR("edgeRegHas", uint8_t, (size_t)nAff * EDGE_REG_CAP);
    eg.regF = SCR("edgeRegF", double, (size_t)nAff * EDGE_REG_CAP); eg.regR = SCR("edgeRegR", double, (size_t)nAff * EDGE_REG_CAP);
    eg.regSums = SCR("edgeRegSums", uint32_t, (size_t)nAff * n); eg.callCounts = SCR("edgeCall", uint32_t, (size_t)nAff * n);
    eg.regCap = EDGE_REG_CAP; eg.rowCap = EDGE_ROW_CAP; eg.maxPositions = 1ull << 23;
    uint32_t* eo = SCR("edgeOut", uint32_t, (size_t)3 * nAff + 4);
    CU(cudaMemsetAsync(eo, 0, ((size_t)3 * nAff + 4) * 4, ctx->stream));
    eg.stop = eo; eg.nCand = eo + nAff; eg.nRows = eo + 2 * nAff; eg.flags = eo + 3 * nAff;
    eg.rows = SCR("edgeRows", EdgeRow, (size_t)nAff * EDGE_ROW_CAP);
    eg.rowCounts = SCR("edgeRowCounts", uint32_t, (size_t)nAff * EDGE_ROW_CAP * n);
    left_edge_kernel<<<nAff, 32, 0, ctx->stream>>>(eg);
    CU(cudaGetLastError());
    launches += 1;
    std::vector<uint32_t> eoh((size_t)3 * nAff + 4);
    CU(cudaMemcpyAsync(eoh.data(), eo, eoh.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    if (eoh[3 * nAff]) return fail(ctx, UPB_ERR_UNSUPPORTED, "left-edge replay (SURVEY Q2) ran out of its scratch or position budget");
    for (uint32_t k = 0; k < nAff; ++k) { edgeStop[k] = eoh[k]; edgeCand[k] = eoh[nAff + k]; edgeRows[k] = eoh[2 * nAff + k]; out.unitEdgeStop[pr.affected[k]] = eoh[k]; }
  }
  uint64_t edgeRowsTotal = 0;
  for (uint32_t k = 0; k < nAff; ++k) edgeRowsTotal += edgeRows[k];

  size_t runCap = std::max<size_t>(1u << 18, (size_t)(ctx->nEntries / 64));
  { const size_t have = ctx->scratch["runs"].cap / sizeof(SpanRun); if (have > runCap) runCap = have; }
  uint32_t nRuns = 0, runSlots = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    sa.runs = SCR("runs", SpanRun, runCap);
    sa.runCap = (uint32_t)std::min<size_t>(runCap, 0xFFFFFFF0u);
    CU(cudaMemsetAsync(ctx->d_small.p, 0, 256, ctx->stream));
    wa.flatBound = flatBound; wa.thr = cfg.thr;
    warm_span_kernel<<<(pr.totalSpans + 255) / 256, 256, 0, ctx->stream>>>(wa);
    { int perSm = 0;
      if ((r = cached_occupancy(ctx, hot_span_kernel, HOT_THREADS, 0, &perSm))) return r;
      hot_span_kernel<<<ctx->smCount * perSm, HOT_THREADS, 0, ctx->stream>>>(ha); }
    launches += 2;
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if ((r = launch_span(ctx, sa, cfg.nondir, &launches))) return r;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    CU(cudaMemcpyAsync(h32, ctx->d_small.p, 64, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    nRuns = h32[3];                                                   // runs written
    runSlots = h32[2];                                                // slots handed out (a warp reserves RUN_CHUNK at a time)
    if (!(h32[0] & FLAG_RUN_OVERFLOW)) break;
    if (attempt == 1) return fail(ctx, UPB_ERR_NOMEM, "run buffer overflow after resize");
    runCap = (size_t)runSlots + (1u << 20);                           // the counter holds the need; chunks may fall differently next time
  }
  ctx->stats.warm_spans = h32[12];
  { unsigned long long st[4]; memcpy(st, h32 + 4, 32); ctx->stats.kde_hot_tiles = st[1]; ctx->stats.hot_blocks = st[0]; ctx->stats.exact_evals = st[2]; ctx->stats.spans_with_tags = st[3]; }
  { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->stats.ms_kde = ms; }
  if (nRuns == 0 && edgeRowsTotal == 0) {
    for (uint32_t k = 0; k < nAff; ++k) { out.unitCand[pr.affected[k]] = edgeCand[k]; out.nCand += edgeCand[k]; }
    if ((r = acquire_slab(ctx, out.layout.totalBytes, out.slab))) return r;
    memset(out.slab.p, 0, out.layout.totalBytes);
    finish_stats(ctx, launches);
    return UPB_OK;
  }
  unsigned long long* h64 = (unsigned long long*)ctx->h_small;
  uint32_t nCand = 0, n1 = 0;
  uint64_t R = 0;
  Cand2* cand = nullptr; uint32_t* list1 = nullptr; unsigned long long* unitFirst = nullptr;
  uint32_t* statCounts = nullptr; double* statKurt = nullptr; uint32_t* pass2 = nullptr; unsigned long long* pass2Scan = nullptr; double* rowCorr = nullptr;
  std::vector<uint32_t> edgeDrop(nAff, 0);
  if (nRuns) {
  // ---- K2: (span, rank) order, stitch runs cut by piece / span boundaries ------------------------------
  unsigned long long* spanOrd = SCR("spanOrd", unsigned long long, pr.totalSpans);
  if ((r = scan_u32(ctx, sa.spanRunCount, pr.totalSpans, spanOrd, d_tot, &launches))) return r;
  uint32_t* ordIdx = SCR("ordIdx", uint32_t, nRuns);
  uint32_t* head = SCR("head", uint32_t, nRuns);
  unsigned long long* headScan = SCR("headScan", unsigned long long, nRuns);
  const unsigned gRuns = (nRuns + 255) / 256;
  run_order_kernel<<<(runSlots + 255) / 256, 256, 0, ctx->stream>>>(sa.runs, runSlots, spanOrd, ordIdx);
  run_heads2_kernel<<<gRuns, 256, 0, ctx->stream>>>(sa.runs, ordIdx, nRuns, d_units, nUnits, head);
  CU(cudaGetLastError());
  launches += 2;
  if ((r = scan_u32(ctx, head, nRuns, headScan, d_tot, &launches))) return r;
  cand = SCR("cand", Cand2, nRuns);                                   // nCand <= nRuns
  unsigned long long* unitCand = SCR("unitCand", unsigned long long, 2 * (size_t)nUnits + 2);
  unitFirst = unitCand + nUnits + 1;
  CU(cudaMemsetAsync(unitCand, 0, (2 * (size_t)nUnits + 2) * 8, ctx->stream));
  merge_runs2_kernel<<<gRuns, 256, 0, ctx->stream>>>(sa.runs, ordIdx, head, headScan, nRuns, d_units, nUnits, cand, unitCand);
  unit_first_kernel<<<1, 32, 0, ctx->stream>>>(unitCand, nUnits, unitFirst);
  CU(cudaGetLastError());
  launches += 2;
  // the number of candidates stays on the device for now (one read-back fewer): what follows runs over nRuns >= nCand slots and
  // looks the count up itself
  const unsigned long long* d_nCand = d_tot + 5;
  CU(cudaMemcpyAsync(d_tot + 5, d_tot, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  const unsigned gCand = gRuns;

  // ---- K3: hit filter (misc/peakcall.cpp:41), then per-region statistics for the survivors ------------------
  uint32_t* pass = SCR("pass", uint32_t, nRuns);
  unsigned long long* passScan = SCR("passScan", unsigned long long, nRuns);
  list1 = SCR("list1", uint32_t, nRuns);
  cand_hit_kernel<<<gCand, 256, 0, ctx->stream>>>(cand, d_nCand, nRuns, cfg.hitThr, pass);
  launches += 1;
  uint32_t* d_drop = nullptr;
  if (nAff) {                                                         // what the replay owns is not a Region of the reference
    d_drop = SCR("edgeDrop", uint32_t, nAff);
    CU(cudaMemsetAsync(d_drop, 0, (size_t)nAff * 4, ctx->stream));
    edge_drop_kernel<<<gCand, 256, 0, ctx->stream>>>(d_affected, eg.stop, nAff, cand, d_nCand, pass, d_drop);
    launches += 1;
  }
  if ((r = scan_u32(ctx, pass, nRuns, passScan, d_tot, &launches))) return r;
  compact_index_kernel<<<gCand, 256, 0, ctx->stream>>>(pass, passScan, nRuns, list1);
  // pooled position cells of the survivors live back to back in one scratch array: offsets by a scan of their lengths
  uint32_t* candLen = SCR("candLen", uint32_t, nRuns);
  unsigned long long* cellOff = SCR("cellOff", unsigned long long, nRuns);
  cand_len_kernel<<<gCand, 256, 0, ctx->stream>>>(cand, pass, nRuns, candLen);
  if ((r = scan_u32(ctx, candLen, nRuns, cellOff, d_tot + 2, &launches))) return r;
  CU(cudaGetLastError());
  launches += 2;
  CU(cudaMemcpyAsync(h64, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(h64 + 2, d_tot + 2, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(h64 + 5, d_tot + 5, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(h64 + 8, unitCand, std::min<size_t>((size_t)nUnits, 400) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  if (nAff) CU(cudaMemcpyAsync(edgeDrop.data(), d_drop, (size_t)nAff * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  nCand = (uint32_t)h64[5];
  out.nCand = nCand;
  ctx->stats.n_candidates = nCand;
  ctx->stats.n_stitched = nRuns - nCand;
  n1 = (uint32_t)h64[0];
  const unsigned long long totalCells = h64[2];
  if (nUnits <= 400) for (uint32_t u = 0; u < nUnits; ++u) out.unitCand[u] = h64[8 + u];
  else { CU(cudaMemcpy(out.unitCand.data(), unitCand, (size_t)nUnits * 8, cudaMemcpyDeviceToHost)); }

  if (n1) {
    const unsigned g1 = (n1 + 255) / 256;
    statCounts = SCR("statCounts", uint32_t, (size_t)n1 * n);
    statKurt = SCR("statKurt", double, n1);
    pass2 = SCR("pass2", uint32_t, n1);
    pass2Scan = SCR("pass2Scan", unsigned long long, n1);
    StatArgs st;
    st.T = pr.T; st.units = d_units; st.cand = cand; st.list = list1; st.nList = n1;
    st.nondir = cfg.nondir ? 1 : 0; st.kurtThr = cfg.kurtThr; st.counts = statCounts; st.kurt = statKurt; st.pass = pass2;
    st.cellOff = cellOff; st.cells = SCR("statCells", uint32_t, (size_t)totalCells + 1);
    const size_t groupBytes = ((size_t)(ST_CHUNK + n) * 4 + 15) / 16 * 16;
    const size_t statSmem = (size_t)ST_WARPS * ST_PER_WARP * groupBytes;
    if (statSmem > 200 * 1024) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many samples for the region statistics kernel");
    st.ticket = small32 + 13;
    CU(cudaMemsetAsync(st.ticket, 0, 4, ctx->stream));
    int perSm = 0;
    if ((r = cached_occupancy(ctx, region_pool_kernel, ST_WARPS * 32, statSmem, &perSm))) return r;
    const uint32_t perBlock = ST_WARPS * ST_PER_WARP;
    const uint32_t gridS = std::min<uint32_t>((n1 + perBlock - 1) / perBlock, (uint32_t)(ctx->smCount * std::max(perSm, 1)));
    region_pool_kernel<<<gridS, ST_WARPS * 32, statSmem, ctx->stream>>>(st);
    region_moment_kernel<<<g1, 256, 0, ctx->stream>>>(st);
    CU(cudaGetLastError());
    launches += 2;

    // strand correlation for the regions that got this far (misc/peakcall.cpp:41 evaluates it last)
    if (cfg.needCorr) {
      if ((r = scan_u32(ctx, pass2, n1, pass2Scan, d_tot, &launches))) return r;
      uint32_t* list2 = SCR("list2", uint32_t, n1);
      compact_index_kernel<<<g1, 256, 0, ctx->stream>>>(pass2, pass2Scan, n1, list2);
      launches += 1;
      CU(cudaMemcpyAsync(h64, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      const uint32_t n2 = (uint32_t)h64[0];
      rowCorr = SCR("rowCorr", double, n1);
      CU(cudaMemsetAsync(rowCorr, 0xFF, (size_t)n1 * 8, ctx->stream));   // NaN where not computed
      if (n2) {
        uint32_t* dUnit = SCR("dUnit", uint32_t, n2); uint32_t* dLeft = SCR("dLeft", uint32_t, n2);
        uint32_t* dLen = SCR("dLen", uint32_t, n2); uint32_t* dChunks = SCR("dChunks", uint32_t, n2);
        dump_desc2_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(cand, list1, list2, n2, dUnit, dLeft, dLen, dChunks);
        CU(cudaGetLastError());
        launches += 1;
        if ((r = dump_and_corr(ctx, pr.dens, cfg.nondir, n2, 1, &launches))) return r;
        cand_corr_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(list2, n2, ctx->scratch["corr"].as<double>(), cfg.corrThr, pass2, rowCorr);
        CU(cudaGetLastError());
        launches += 1;
      }
    }
    if ((r = scan_u32(ctx, pass2, n1, pass2Scan, d_tot, &launches))) return r;
    CU(cudaMemcpyAsync(h64, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    R = h64[0];
  }

  }  // nRuns

  // ---- accepted rows -> result slab -------------------------------------------------------------------------
  const uint64_t Rnormal = R;
  R += edgeRowsTotal;
  out.R = R;
  ctx->lastRows = R;
  out.layout = slab_layout(R, n, C);
  const SlabLayout& SL = out.layout;
  if ((r = acquire_slab(ctx, SL.totalBytes, out.slab))) return r;
  unsigned char* hs = (unsigned char*)out.slab.p;
  memset(hs + SL.devBytes, 0, SL.totalBytes - SL.devBytes);
  // replay bookkeeping: its candidates replace the dropped ones, its rows precede the unit's other rows
  long long* d_shift = nullptr;
  if (nAff) {
    std::vector<long long> shift(2 * (size_t)nUnits, 0);
    long long runRows = 0;
    size_t k = 0;
    for (uint32_t u = 0; u < nUnits; ++u) {
      if (k < nAff && pr.affected[k] == u) { runRows += edgeRows[k]; shift[nUnits + u] = (long long)edgeCand[k] - (long long)edgeDrop[k]; ++k; }
      shift[u] = runRows;
    }
    for (k = 0; k < nAff; ++k) {
      const uint32_t u = pr.affected[k];
      out.unitCand[u] = (uint64_t)((long long)out.unitCand[u] + shift[nUnits + u]);
      out.nCand = (uint64_t)((long long)out.nCand + shift[nUnits + u]);
    }
    d_shift = SCR("edgeShift", long long, 2 * (size_t)nUnits);
    CU(cudaMemcpyAsync(d_shift, shift.data(), shift.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
  }
  if (R) {
    unsigned char* ds = SCR("slab", unsigned char, SL.devBytes);
    unsigned long long* d_tir = SCR("tir", unsigned long long, n);
    CU(cudaMemsetAsync(d_tir, 0, (size_t)n * 8, ctx->stream));
    EmitArgs ea;
    ea.cand = cand; ea.list = list1; ea.pass = pass2; ea.passScan = pass2Scan; ea.nList = n1; ea.units = d_units; ea.unitFirst = unitFirst;
    ea.unitRowShift = d_shift; ea.unitCandShift = d_shift ? d_shift + nUnits : nullptr;
    ea.statCounts = statCounts; ea.statKurt = statKurt; ea.rowCorr = rowCorr; ea.nSamples = n;
    ea.buffer = ds + SL.oBuffer; ea.contig = (uint32_t*)(ds + SL.oContig); ea.left = (uint32_t*)(ds + SL.oLeft); ea.len = (uint32_t*)(ds + SL.oLen);
    ea.peak = (uint32_t*)(ds + SL.oPeak); ea.wsum = (uint32_t*)(ds + SL.oWsum); ea.candIndex = (unsigned long long*)(ds + SL.oCandIndex);
    ea.peakScore = (double*)(ds + SL.oPeakScore); ea.kurt = (double*)(ds + SL.oKurt); ea.corr = (double*)(ds + SL.oCorr);
    ea.counts = (uint32_t*)(ds + SL.oCounts);
    if (n1) { emit_rows_kernel<<<(n1 + 255) / 256, 256, 0, ctx->stream>>>(ea); launches += 1; }
    if (edgeRowsTotal) {
      // first table row of each unit's replay rows = accepted rows of span_kernel in earlier units + replay rows of earlier units
      std::vector<unsigned long long> rowBase(nAff, 0);
      std::vector<uint32_t> l1(n1), p2(n1);
      std::vector<Cand2> hc;                                          // rare path: a few small host look-ups are fine
      unsigned long long before = 0;
      if (n1) {
        CU(cudaMemcpyAsync(l1.data(), list1, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaMemcpyAsync(p2.data(), pass2, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        hc.resize(nCand);
        CU(cudaMemcpyAsync(hc.data(), cand, (size_t)nCand * sizeof(Cand2), cudaMemcpyDeviceToHost, ctx->stream));
        CU(cudaStreamSynchronize(ctx->stream));
      }
      size_t j = 0; unsigned long long acc = 0;
      for (uint32_t k = 0; k < nAff; ++k) {
        const uint32_t u = pr.affected[k];
        while (j < n1 && hc[l1[j]].unit < u) { acc += p2[j]; ++j; }
        rowBase[k] = acc + before;
        before += edgeRows[k];
      }
      unsigned long long* d_rowBase = SCR("edgeRowBase", unsigned long long, nAff);
      CU(cudaMemcpyAsync(d_rowBase, rowBase.data(), (size_t)nAff * 8, cudaMemcpyHostToDevice, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      EdgeEmitArgs ee;
      ee.rows = eg.rows; ee.rowCounts = eg.rowCounts; ee.rowCap = EDGE_ROW_CAP; ee.nSamples = n; ee.nAffected = nAff; ee.affected = d_affected; ee.nRows = eg.nRows;
      ee.rowBase = d_rowBase; ee.units = d_units;
      ee.buffer = ea.buffer; ee.contig = ea.contig; ee.left = ea.left; ee.len = ea.len; ee.peak = ea.peak; ee.candIndex = ea.candIndex;
      ee.peakScore = ea.peakScore; ee.kurt = ea.kurt; ee.corr = ea.corr; ee.counts = ea.counts; ee.wsum = ea.wsum;
      edge_emit_kernel<<<dim3((EDGE_ROW_CAP + 127) / 128, nAff), 128, 0, ctx->stream>>>(ee);
      launches += 1;
    }
    tags_in_regions_kernel<<<std::min<unsigned>(ctx->smCount * 4, (unsigned)((R * n + 255) / 256)), 256, (size_t)n * 8, ctx->stream>>>(ea.counts, R, n, d_tir);
    CU(cudaGetLastError());
    launches += 1;
    if (!cfg.deviceOnly) CU(cudaMemcpyAsync(hs, ds, SL.devBytes, cudaMemcpyDeviceToHost, ctx->stream));
    else {                                                            // only what the per-unit row counts need (5 bytes per row)
      const size_t Rp = (R + 7) & ~(size_t)7;
      CU(cudaMemcpyAsync(hs + SL.oContig, ds + SL.oContig, Rp * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaMemcpyAsync(hs + SL.oBuffer, ds + SL.oBuffer, Rp, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU(cudaMemcpyAsync(hs + SL.oTags, d_tir, (size_t)n * 8, cudaMemcpyDeviceToHost, ctx->stream));
  }
  (void)Rnormal;
  finish_stats(ctx, launches);
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// reference emission order, labels and dropped tail (src/regions.cpp:311-391)
// ------------------------------------------------------------------------------------------------
// For one sample: is its '-' track consumed during pass 1 of the directional merge loop?  The stream is
// '+' contigs in table order, then '-' contigs.  Pass 1 visits contigs in order and consumes whatever
// sits at the head of the stream for that contig regardless of strand (src/regions.cpp:341), so the
// '-' track is reached in pass 1 exactly when its first contig is not before the last '+' contig.
static bool early_reverse (uint32_t C, const uint32_t* firstPos /* [2*C] of this sample */) {
  int lastPlus = -1, firstMinus = -1;
  for (uint32_t c = 0; c < C; ++c) { if (firstPos[c]) lastPlus = (int)c; }
  for (uint32_t c = 0; c < C; ++c) { if (firstPos[C + c]) { firstMinus = (int)c; break; } }
  if (firstMinus < 0) return false;
  return firstMinus >= lastPlus;     // lastPlus == -1 (no '+' tags at all) also lands here
}

static int reference_order_impl (uint32_t C, uint32_t n, int directional, uint32_t bw, const uint32_t* firstPos, const uint32_t* lastPos,
                                 upb_region_table* t, std::vector<uint64_t>* permOut, bool* identity);

extern "C" int upb_reference_order (uint32_t C, uint32_t n, int directional, uint32_t bw, const uint32_t* firstPos, const uint32_t* lastPos,
                                    upb_region_table* t, uint64_t* perm) {
  if (!t || !perm || !firstPos || !lastPos) return UPB_ERR_BAD_ARG;
  std::vector<uint64_t> p;
  bool identity = true;
  const int r = reference_order_impl(C, n, directional, bw, firstPos, lastPos, t, &p, &identity);
  if (r) return r;
  if (identity) { for (uint64_t i = 0; i < t->n_rows; ++i) perm[i] = i; }
  else memcpy(perm, p.data(), t->n_rows * 8);
  return UPB_OK;
}

static int reference_order_impl (uint32_t C, uint32_t n, int directional, uint32_t bw, const uint32_t* firstPos, const uint32_t* lastPos,
                                 upb_region_table* t, std::vector<uint64_t>* permOut, bool* identity) {
  *identity = true;
  const uint64_t R = t->n_rows;
  // last add() position per (buffer, contig): every sample counts, controls included
  std::vector<uint32_t> lastTag(2 * (size_t)C, 0);
  for (uint32_t s = 0; s < n; ++s)
    for (uint32_t d = 0; d < 2; ++d)
      for (uint32_t c = 0; c < C; ++c) {
        const uint32_t lp = lastPos[((size_t)s * 2 + d) * C + c];
        const uint32_t b = directional ? d : 0;
        if (lp > lastTag[(size_t)b * C + c]) lastTag[(size_t)b * C + c] = lp;
      }
  // a region is finalised by add() calls iff some later tag of its buffer/contig lies beyond right+bw+1;
  // otherwise only flushContig() closes it (misc/peakcall.cpp:224-231)
  auto tail = [&] (uint64_t i) -> bool {
    const uint64_t right = (uint64_t)t->left[i] + t->len[i] - 1;
    const uint32_t L = lastTag[(size_t)t->buffer[i] * C + t->contig[i]];
    return right + bw + 2 > (uint64_t)L;
  };
  memset(t->dropped, 0, R);

  // which samples read their '-' track during pass 1 (Q7)
  bool anyEarly = false;
  uint32_t nWithTags = 0;
  if (directional) {
    for (uint32_t s = 0; s < n; ++s) {
      bool has = false;
      for (uint32_t k = 0; k < 2 * C; ++k) has = has || firstPos[(size_t)s * 2 * C + k];
      if (!has) continue;
      ++nWithTags;
      if (early_reverse(C, firstPos + (size_t)s * 2 * C)) anyEarly = true;
    }
    if (anyEarly && nWithTags > 1) return UPB_ERR_REFERENCE_ABORTS;
  }

  // canonical rows are sorted by (buffer, contig, left): first row of the reverse buffer by binary search
  uint64_t firstRev = R;
  { uint64_t lo = 0, hi = R; while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (t->buffer[mid] == 0) lo = mid + 1; else hi = mid; } firstRev = lo; }
  memset(t->forward_label, 1, firstRev);
  memset(t->forward_label + firstRev, 0, R - firstRev);

  // Q5: the first Region object of reverseBuffer is constructed forward (misc/peakcall.cpp:103); if the
  // first candidate of the reverse buffer is accepted it is printed left-right
  if (directional && t->unit_candidates) {
    for (uint32_t c = 0; c < C; ++c) {
      if (t->unit_candidates[(size_t)C + c] == 0) continue;
      for (uint64_t i = firstRev; i < R && t->contig[i] <= c; ++i)
        if (t->contig[i] == c && t->cand_index[i] == 0) t->forward_label[i] = 1;
      break;
    }
  }

  if (!directional || !anyEarly) {
    // rows are written in canonical order ('+' regions by contig, then '-' regions); only what the very last
    // flushContig() finalises is lost (Q6): the tail of the last contig of the last buffer — a suffix of the table
    for (uint64_t i = R; i-- > 0;) {
      if (t->contig[i] != C - 1 || (directional && t->buffer[i] != 1) || !tail(i)) break;
      t->dropped[i] = 1;
    }
    return UPB_OK;                                                    // canonical order; dropped rows already sit at the end
  }
  // single sample whose '-' track is read in pass 1: per contig iteration the '+' tags come first, then
  // the '-' tags; regions closed by add() are written at once, the two flushes at the end of the
  // iteration hand over the '+' tail and then the '-' tail; pass 2 then writes what is pending (nothing is lost)
  std::vector<uint64_t> order;
  order.reserve(R);
  std::vector<std::vector<uint64_t>> byUnit(2 * (size_t)C);
  for (uint64_t i = 0; i < R; ++i) byUnit[(size_t)t->buffer[i] * C + t->contig[i]].push_back(i);
  for (uint32_t c = 0; c < C; ++c) {
    for (int b = 0; b < 2; ++b) for (uint64_t i : byUnit[(size_t)b * C + c]) if (!tail(i)) order.push_back(i);
    for (int b = 0; b < 2; ++b) for (uint64_t i : byUnit[(size_t)b * C + c]) if (tail(i)) order.push_back(i);
  }
  *identity = false;
  permOut->assign(order.begin(), order.end());
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// result table: columns point into one pinned slab that goes back to the context's pool when freed
// ------------------------------------------------------------------------------------------------
extern "C" void upb_free_region_table (upb_region_table* t) {
  if (!t) return;
  if (t->_slab) {
    upb_ctx* ctx = (upb_ctx*)t->_ctx;
    HostSlab h; h.p = t->_slab; h.cap = (size_t)t->_slab_bytes;
    if (ctx) ctx->slabPool.push_back(h); else cudaFreeHost(h.p);
  }
  memset(t, 0, sizeof *t);
}

static void fill_table (upb_ctx* ctx, const PipeOut& po, upb_region_table* t) {
  const uint32_t n = ctx->nSamples, C = (uint32_t)ctx->sizes.size();
  const SlabLayout& L = po.layout;
  memset(t, 0, sizeof *t);
  t->n_rows = po.R; t->n_samples = n;
  t->buffer = po.col<uint8_t>(L.oBuffer); t->forward_label = po.col<uint8_t>(L.oLabel); t->dropped = po.col<uint8_t>(L.oDropped);
  t->contig = po.col<uint32_t>(L.oContig); t->left = po.col<uint32_t>(L.oLeft); t->len = po.col<uint32_t>(L.oLen); t->peak = po.col<uint32_t>(L.oPeak);
  t->hit_sum = po.col<uint32_t>(L.oWsum);
  t->cand_index = po.col<uint64_t>(L.oCandIndex); t->peak_score = po.col<double>(L.oPeakScore); t->kurtosis = po.col<double>(L.oKurt);
  t->corr = po.col<double>(L.oCorr); t->counts = po.col<uint32_t>(L.oCounts); t->tags_in_regions = po.col<uint64_t>(L.oTags);
  t->unit_candidates = po.col<uint64_t>(L.oUnitCand); t->unit_last_tag = po.col<uint32_t>(L.oUnitLast); t->unit_rows = po.col<uint64_t>(L.oUnitRows);
  t->_slab = po.slab.p; t->_slab_bytes = po.slab.cap; t->_ctx = ctx;
  // rows are sorted by (buffer, contig, left): unit boundaries by binary search instead of a pass over the rows
  {
    auto lower = [&] (uint32_t b, uint32_t c) -> uint64_t {           // first row with (buffer, contig) >= (b, c)
      uint64_t lo = 0, hi = po.R;
      while (lo < hi) { const uint64_t mid = (lo + hi) >> 1; if (t->buffer[mid] < b || (t->buffer[mid] == b && t->contig[mid] < c)) lo = mid + 1; else hi = mid; }
      return lo;
    };
    uint64_t prev = 0;
    for (uint32_t b = 0; b < 2; ++b) for (uint32_t c = 0; c < C; ++c) {
      const uint64_t nxt = (b == 1 && c == C - 1) ? po.R : lower(c + 1 == C ? b + 1 : b, c + 1 == C ? 0 : c + 1);
      t->unit_rows[(size_t)b * C + c] = nxt - prev;
      prev = nxt;
    }
    const uint64_t firstRev = lower(1, 0);
    memset(t->forward_label, 1, firstRev);
    memset(t->forward_label + firstRev, 0, po.R - firstRev);
  }
  for (size_t u = 0; u < po.units.size(); ++u) t->unit_candidates[(size_t)po.units[u].buffer * C + po.units[u].contig] = po.unitCand[u];
  t->n_passed = po.R;
  t->n_rejected = po.nCand - po.R;
}

extern "C" int upb_call_regions (upb_ctx* ctx, const upb_region_params* p, upb_region_table* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !out) return fail(ctx, UPB_ERR_BAD_ARG, "upb_call_regions: NULL argument");
  memset(out, 0, sizeof *out);
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  PipeCfg cfg;
  cfg.bw = p->bandwidth; cfg.thr = p->region_threshold; cfg.kurtThr = p->kurtosis_threshold; cfg.corrThr = p->corr_threshold;
  cfg.hitThr = p->hit_threshold; cfg.background = p->background; cfg.nondir = !p->directional; cfg.control = p->control;
  cfg.nCoeffs = p->n_coeffs; cfg.coeffs = p->coeffs;
  cfg.needCorr = cfg.nondir;        // -D: the filter (if corr_threshold > -1) and the optional correlation column
  cfg.allowLeftEdge = true;
  cfg.deviceOnly = p->reference_order == 2;
  PipeOut po;
  Prep pr;
  int r = run_pipeline(ctx, cfg, pr, po);
  if (r) return r;
  // canonical order: (buffer, contig, left) — units are in that order and candidates are position-ordered
  fill_table(ctx, po, out);
  const uint32_t C = (uint32_t)ctx->sizes.size(), n = ctx->nSamples;
  for (uint32_t s = 0; s < n; ++s) for (uint32_t d = 0; d < 2; ++d) for (uint32_t c = 0; c < C; ++c) {
    const uint32_t lp = ctx->lastPos[((size_t)s * 2 + d) * C + c];
    const uint32_t b = p->directional ? d : 0;
    if (lp > out->unit_last_tag[(size_t)b * C + c]) out->unit_last_tag[(size_t)b * C + c] = lp;
  }
  if (p->reference_order == 1) {
    std::vector<uint64_t> perm;
    bool identity = true;
    r = reference_order_impl(C, n, p->directional, p->bandwidth, ctx->firstPos.data(), ctx->lastPos.data(), out, &perm, &identity);
    if (r) { upb_free_region_table(out); return fail(ctx, r, "the reference's merge loop would abort on this input (several directional samples whose '-' track is read during pass 1; SURVEY Q7)"); }
    const uint64_t R = out->n_rows;
    if (!identity) {
      auto permute = [&] (auto* col, size_t width) {
        typedef typename std::remove_pointer<decltype(col)>::type T;
        std::vector<T> tmp(col, col + R * width);
        for (uint64_t k = 0; k < R; ++k) for (size_t w = 0; w < width; ++w) col[k * width + w] = tmp[perm[k] * width + w];
      };
      permute(out->buffer, 1); permute(out->forward_label, 1); permute(out->dropped, 1); permute(out->contig, 1); permute(out->left, 1);
      permute(out->len, 1); permute(out->peak, 1); permute(out->hit_sum, 1); permute(out->cand_index, 1); permute(out->peak_score, 1);
      permute(out->kurtosis, 1); permute(out->corr, 1); permute(out->counts, n);
    }
  }
  ctx->stats.ms_call = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// several GPUs: one table out of the ranks' tables
// ------------------------------------------------------------------------------------------------
extern "C" int upb_slab_layout (uint64_t n_rows, uint32_t n_samples, uint64_t* off) {
  if (!off) return UPB_ERR_BAD_ARG;
  const SlabLayout L = slab_layout((size_t)n_rows, n_samples, 1);
  off[UPB_COL_CAND_INDEX] = L.oCandIndex; off[UPB_COL_PEAK_SCORE] = L.oPeakScore; off[UPB_COL_KURTOSIS] = L.oKurt; off[UPB_COL_CORR] = L.oCorr;
  off[UPB_COL_CONTIG] = L.oContig; off[UPB_COL_LEFT] = L.oLeft; off[UPB_COL_LEN] = L.oLen; off[UPB_COL_PEAK] = L.oPeak; off[UPB_COL_HIT_SUM] = L.oWsum;
  off[UPB_COL_COUNTS] = L.oCounts; off[UPB_COL_BUFFER] = L.oBuffer; off[UPB_COL_END] = L.devBytes;
  return UPB_OK;
}

struct ScatterArgs {
  const unsigned char* src; unsigned char* dst;
  SlabLayout ls, ld;
  unsigned long long R;
  uint32_t n, C;
  const unsigned long long* unitDst;   // [2C] first table-wide row of each (buffer, contig)
  const unsigned long long* unitSrc;   // [2C] first local row of each (buffer, contig)
};

__global__ void scatter_rows_kernel (ScatterArgs a) {
  const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= a.R) return;
  const uint32_t b = a.src[a.ls.oBuffer + i];
  const uint32_t c = ((const uint32_t*)(a.src + a.ls.oContig))[i];
  const size_t u = (size_t)b * a.C + c;
  const unsigned long long o = a.unitDst[u] + (i - a.unitSrc[u]);
#define UPB_MOVE(T, off) ((T*)(a.dst + a.ld.off))[o] = ((const T*)(a.src + a.ls.off))[i]
  UPB_MOVE(unsigned long long, oCandIndex); UPB_MOVE(double, oPeakScore); UPB_MOVE(double, oKurt); UPB_MOVE(double, oCorr);
  UPB_MOVE(uint32_t, oContig); UPB_MOVE(uint32_t, oLeft); UPB_MOVE(uint32_t, oLen); UPB_MOVE(uint32_t, oPeak); UPB_MOVE(uint32_t, oWsum);
  UPB_MOVE(uint8_t, oBuffer);
#undef UPB_MOVE
}

// the counts matrix, one thread per element: consecutive threads write consecutive words (the destination may be mapped host
// memory, where only coalesced stores travel well over PCIe)
__global__ void scatter_counts_kernel (ScatterArgs a) {
  const unsigned long long total = a.R * a.n;
  const uint32_t* sc = (const uint32_t*)(a.src + a.ls.oCounts);
  uint32_t* dc = (uint32_t*)(a.dst + a.ld.oCounts);
  for (unsigned long long g = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; g < total; g += (unsigned long long)gridDim.x * blockDim.x) {
    const unsigned long long i = g / a.n;
    const uint32_t b = a.src[a.ls.oBuffer + i];
    const uint32_t c = ((const uint32_t*)(a.src + a.ls.oContig))[i];
    const size_t u = (size_t)b * a.C + c;
    dc[(a.unitDst[u] + (i - a.unitSrc[u])) * a.n + (g - i * a.n)] = sc[g];
  }
}

extern "C" int upb_scatter_rows (upb_ctx* ctx, uint64_t total_rows, const uint64_t* unit_rows_total, const uint64_t* unit_rows_local, void* d_slab) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!unit_rows_total || !unit_rows_local || !d_slab) return fail(ctx, UPB_ERR_BAD_ARG, "upb_scatter_rows: NULL argument");
  CU(cudaSetDevice(ctx->device));
  const uint32_t C = (uint32_t)ctx->sizes.size(), n = ctx->nSamples;
  const uint64_t R = ctx->lastRows;
  if (R == 0) return UPB_OK;
  ScatterArgs sa;
  sa.src = ctx->scratch["slab"].as<unsigned char>(); sa.dst = (unsigned char*)d_slab;
  sa.ls = slab_layout(R, n, C); sa.ld = slab_layout(total_rows, n, C);
  sa.R = R; sa.n = n; sa.C = C;
  // local first row of each unit: rows are in canonical order, so count them from the buffer / contig columns
  std::vector<unsigned long long> h(4 * (size_t)C, 0);
  { unsigned long long run = 0; for (size_t u = 0; u < 2 * (size_t)C; ++u) { h[u] = run; run += unit_rows_total[u]; } if (run != total_rows) return fail(ctx, UPB_ERR_BAD_ARG, "upb_scatter_rows: unit_rows_total does not add up to total_rows"); }
  { unsigned long long run = 0; for (size_t u = 0; u < 2 * (size_t)C; ++u) { h[2 * (size_t)C + u] = run; run += unit_rows_local[u]; } if (run != R) return fail(ctx, UPB_ERR_BAD_ARG, "upb_scatter_rows: unit_rows_local does not match the last upb_call_regions on this context"); }
  unsigned long long* d_u = SCR("scatterUnits", unsigned long long, 4 * (size_t)C);
  CU(cudaMemcpyAsync(d_u, h.data(), 4 * (size_t)C * 8, cudaMemcpyHostToDevice, ctx->stream));
  sa.unitDst = d_u; sa.unitSrc = d_u + 2 * (size_t)C;
  scatter_rows_kernel<<<(unsigned)((R + 255) / 256), 256, 0, ctx->stream>>>(sa);
  scatter_counts_kernel<<<(unsigned)std::min<uint64_t>((R * n + 255) / 256, (uint64_t)ctx->smCount * 16), 256, 0, ctx->stream>>>(sa);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ctx->stream));
  return UPB_OK;
}

extern "C" int upb_table_from_slab (void* h_slab, uint64_t n_rows, uint32_t n_samples, upb_region_table* t) {
  if (!h_slab || !t) return UPB_ERR_BAD_ARG;
  const SlabLayout L = slab_layout((size_t)n_rows, n_samples, 1);
  unsigned char* b = (unsigned char*)h_slab;
  memset(t, 0, sizeof *t);
  t->n_rows = n_rows; t->n_samples = n_samples;
  t->buffer = b + L.oBuffer; t->contig = (uint32_t*)(b + L.oContig); t->left = (uint32_t*)(b + L.oLeft); t->len = (uint32_t*)(b + L.oLen);
  t->peak = (uint32_t*)(b + L.oPeak); t->hit_sum = (uint32_t*)(b + L.oWsum); t->cand_index = (uint64_t*)(b + L.oCandIndex);
  t->peak_score = (double*)(b + L.oPeakScore); t->kurtosis = (double*)(b + L.oKurt); t->corr = (double*)(b + L.oCorr); t->counts = (uint32_t*)(b + L.oCounts);
  t->n_passed = n_rows;
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// bin/strand_shift — src/strand_shift.cpp:139-259
// ------------------------------------------------------------------------------------------------
extern "C" void upb_free_shift_report (upb_shift_report* r) {
  if (!r) return;
  free(r->shift_freq); free(r->tested_contig); free(r->tested_left); free(r->tested_len); free(r->tested_best_shift); free(r->tested_best_corr);
  memset(r, 0, sizeof *r);
}

template <class T> static T* alloc_arr (size_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }

struct SumIdx { uint32_t sum; uint32_t idx; };
static bool compareRegionCounts (const SumIdx& i, const SumIdx& j) { return i.sum > j.sum; }   // src/strand_shift.cpp:26-29

// the pipeline configuration of src/strand_shift.cpp:139-190: one non-directional buffer, no correlation filter (:142)
static PipeCfg shift_cfg (const upb_shift_params* p) {
  PipeCfg cfg;
  cfg.bw = p->bandwidth; cfg.thr = p->region_threshold; cfg.kurtThr = p->kurtosis_threshold; cfg.corrThr = -1;   // :142
  cfg.hitThr = (double)p->hit_threshold; cfg.background = p->background; cfg.nondir = true; cfg.control = nullptr; cfg.needCorr = false;
  cfg.allowLeftEdge = true;          // regions at a contig start come from the replay of the reference's left-edge behaviour (Q2)
  return cfg;
}

// :206-219 for a batch of regions that are long enough: strandCorr(shift) for every shift 0..max_shift on the device, then the
// first strict maximum per region.  `dens` is the density template of a prepare() / run_pipeline() with shift_cfg.
static int shift_sweep_batch (upb_ctx* ctx, const DensArgs& dens, const std::vector<uint32_t>& unitOfContig, uint32_t nb, const uint32_t* contig,
                              const uint32_t* left, const uint32_t* len, uint32_t nShifts, uint32_t* bestShift, double* bestCorr, uint32_t* launches) {
  if (!nb) return UPB_OK;
  std::vector<uint32_t> hUnit(nb), hChunks(nb);
  for (uint32_t k = 0; k < nb; ++k) { hUnit[k] = unitOfContig[contig[k]]; hChunks[k] = (len[k] + TILE - 1) / TILE; }
  uint32_t* dUnit = SCR("dUnit", uint32_t, nb); uint32_t* dLeft = SCR("dLeft", uint32_t, nb);
  uint32_t* dLen = SCR("dLen", uint32_t, nb); uint32_t* dChunks = SCR("dChunks", uint32_t, nb);
  CU(cudaMemcpyAsync(dUnit, hUnit.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dLeft, left, (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dLen, len, (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(dChunks, hChunks.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));                           // the host vectors go out of scope below
  int r;
  if ((r = dump_and_corr(ctx, dens, true, nb, nShifts, launches))) return r;
  uint32_t* dBestShift = SCR("bestShift", uint32_t, nb);
  double* dBestCorr = SCR("bestCorr", double, nb);
  best_shift_kernel<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(ctx->scratch["corr"].as<double>(), nb, nShifts, dBestShift, dBestCorr);
  CU(cudaGetLastError());
  *launches += 1;
  CU(cudaMemcpyAsync(bestShift, dBestShift, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(bestCorr, dBestCorr, (size_t)nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  return UPB_OK;
}

// :239-258 smooth the histogram of maxima with Kernel(5) and take the mode (151 numbers: host)
static uint32_t shift_mode (const uint64_t* shiftFreq, size_t m, uint32_t minShift) {
  double modeKernel[11];
  upb_kernel_taps(5, 1, modeKernel);
  std::vector<double> density(m, 0.0);
  for (size_t i = 0; i < m; ++i)
    for (int j = 0; j < 11; ++j) {
      const long k = (long)i - 5 + j;
      if (k >= 0 && k < (long)m) { volatile double term = (double)shiftFreq[i] * modeKernel[j]; density[k] = density[k] + term; }
    }
  double bestDensity = 0;
  uint32_t best = 0;
  if (m >= 5) for (size_t i = minShift; i < m - 5; ++i) if (density[i] > bestDensity) { best = (uint32_t)i; bestDensity = density[i]; }
  return best;
}

extern "C" int upb_strand_shift (upb_ctx* ctx, const upb_shift_params* p, upb_shift_report* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !out) return fail(ctx, UPB_ERR_BAD_ARG, "upb_strand_shift: NULL argument");
  memset(out, 0, sizeof *out);
  const PipeCfg cfg = shift_cfg(p);
  PipeOut po;
  Prep pr;
  int r = run_pipeline(ctx, cfg, pr, po);
  if (r) return r;
  struct SlabGuard { upb_ctx* c; HostSlab h; ~SlabGuard () { if (h.p) c->slabPool.push_back(h); } } guard{ctx, po.slab};
  const size_t R = po.R;
  const uint32_t* rContig = po.col<uint32_t>(po.layout.oContig);
  const uint32_t* rLeft = po.col<uint32_t>(po.layout.oLeft);
  const uint32_t* rLen = po.col<uint32_t>(po.layout.oLen);
  const uint32_t* rSum = po.col<uint32_t>(po.layout.oWsum);
  std::vector<uint32_t> unitOfContig(ctx->sizes.size(), 0);
  for (size_t u = 0; u < po.units.size(); ++u) unitOfContig[po.units[u].contig] = (uint32_t)u;
  out->n_regions_found = R;
  out->n_freq = p->max_shift + 1;
  out->shift_freq = alloc_arr<uint64_t>(out->n_freq);
  const uint16_t toTest = (uint16_t)p->n_regions_to_test;            // UShort (Q12)
  const size_t cap = toTest ? toTest : 1;
  out->tested_contig = alloc_arr<uint32_t>(cap); out->tested_left = alloc_arr<uint32_t>(cap); out->tested_len = alloc_arr<uint32_t>(cap);
  out->tested_best_shift = alloc_arr<uint32_t>(cap); out->tested_best_corr = alloc_arr<double>(cap);

  // :198 std::sort by Region::sum() descending — same comparator outcomes, same libstdc++ algorithm, hence the same
  // permutation as the reference's sort of Region pointers; only as much of it as the examination below needs (upb_introsort.hpp)
  std::vector<SumIdx> unsorted(R), order;
  for (size_t i = 0; i < R; ++i) { unsorted[i].sum = rSum[i]; unsorted[i].idx = (uint32_t)i; }   // no controls: sum() == pooled weight
  size_t sortedEnd = 0;
  auto sort_more = [&] (size_t want) {
    order = unsorted;
    sortedEnd = sort_prefix_like_std_sort(order.begin(), order.end(), want, compareRegionCounts);
  };
  sort_more(std::min<size_t>(R, 2 * (size_t)toTest + 256));

  // :201-226 — regions are examined in sorted order until enough qualify; the device sweeps a batch at a time
  uint16_t tested = 0;
  size_t cursor = 0;
  uint32_t launches = ctx->stats.launches;
  const uint32_t nShifts = p->max_shift + 1;
  while (tested < toTest && cursor < R) {
    std::vector<uint32_t> batch;                                     // row indices
    const size_t want = (size_t)(toTest - tested) + (size_t)(toTest - tested) / 4 + 8;
    while (cursor < R && batch.size() < want) {
      if (cursor >= sortedEnd) sort_more(std::min<size_t>(R, std::max<size_t>(2 * sortedEnd, cursor + 1024)));
      const uint32_t i = order[cursor].idx;
      if (rLen[i] > 2 * p->max_shift + 3) {                            // :206 large enough to try all shifts
        // a region the left-edge replay produced has displaced densities (Q2); its shift sweep is not replayed
        if (rLeft[i] <= po.unitEdgeStop[unitOfContig[rContig[i]]])
          return fail(ctx, UPB_ERR_LEFT_EDGE, "a region within `bandwidth` of a contig start is long enough for the shift sweep: the reference's left-edge densities (SURVEY Q2) are not emulated there");
        batch.push_back(i);
      }
      ++cursor;
    }
    if (batch.empty()) break;
    const uint32_t nb = (uint32_t)batch.size();
    std::vector<uint32_t> hContig(nb), hLeft(nb), hLen(nb), bs(nb);
    std::vector<double> bc(nb);
    for (uint32_t k = 0; k < nb; ++k) { const uint32_t i = batch[k]; hContig[k] = rContig[i]; hLeft[k] = rLeft[i]; hLen[k] = rLen[i]; }
    if ((r = shift_sweep_batch(ctx, pr.dens, unitOfContig, nb, hContig.data(), hLeft.data(), hLen.data(), nShifts, bs.data(), bc.data(), &launches))) return r;
    for (uint32_t k = 0; k < nb && tested < toTest; ++k) {
      if (bc[k] >= p->corr_threshold) {                              // :220 disqualify regions under threshold
        const uint32_t i = batch[k];
        ++out->shift_freq[bs[k]];
        out->n_tags_in_regions += rSum[i];
        out->tested_contig[tested] = rContig[i]; out->tested_left[tested] = rLeft[i]; out->tested_len[tested] = rLen[i];
        out->tested_best_shift[tested] = bs[k]; out->tested_best_corr[tested] = bc[k];
        ++tested;
      }
    }
  }
  ctx->stats.launches = launches;
  out->n_regions_tested = tested;
  if (tested == 0) return fail(ctx, UPB_ERR_NO_REGIONS, "no regions qualified with given settings");   // :228-230
  out->best_shift = shift_mode(out->shift_freq, out->n_freq, p->min_shift);
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// bin/strand_shift over contig-sharded ranks: the same steps as upb_strand_shift, cut where the ranks have to
// exchange something.  Protocol (unipeak_b200/multi.py::strand_shift_sharded runs it over torch.distributed):
//   1. every rank: upb_shift_regions   — the Region vector of :143-190 for its contigs, canonical order (contig, left)
//   2. all-gather of the lists (16 B per region); concatenated by contig they are the reference's vector, on every rank
//   3. every rank: upb_shift_examine_order on the global sums — the head of the reference's std::sort, hence the same order everywhere
//   4. per batch of regions to examine: each rank sweeps the ones on its contigs (upb_shift_sweep), all-gather of (shift, corr)
//   5. every rank applies :220-225 in examination order and smooths the histogram (upb_shift_mode)
// ------------------------------------------------------------------------------------------------
extern "C" void upb_free_shift_regions (upb_shift_region_list* l) {
  if (!l) return;
  free(l->contig); free(l->left); free(l->len); free(l->sum); free(l->replayed);
  memset(l, 0, sizeof *l);
}

extern "C" int upb_shift_regions (upb_ctx* ctx, const upb_shift_params* p, upb_shift_region_list* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !out) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_regions: NULL argument");
  memset(out, 0, sizeof *out);
  const PipeCfg cfg = shift_cfg(p);
  PipeOut po;
  Prep pr;
  int r = run_pipeline(ctx, cfg, pr, po);
  if (r) return r;
  struct SlabGuard { upb_ctx* c; HostSlab h; ~SlabGuard () { if (h.p) c->slabPool.push_back(h); } } guard{ctx, po.slab};
  const size_t R = po.R;
  std::vector<uint32_t> unitOfContig(ctx->sizes.size(), 0);
  for (size_t u = 0; u < po.units.size(); ++u) unitOfContig[po.units[u].contig] = (uint32_t)u;
  out->n = R;
  out->contig = alloc_arr<uint32_t>(R); out->left = alloc_arr<uint32_t>(R); out->len = alloc_arr<uint32_t>(R); out->sum = alloc_arr<uint32_t>(R);
  out->replayed = alloc_arr<uint8_t>(R);
  if (!out->contig || !out->left || !out->len || !out->sum || !out->replayed) { upb_free_shift_regions(out); return fail(ctx, UPB_ERR_NOMEM, "upb_shift_regions: out of host memory"); }
  if (R) {
    memcpy(out->contig, po.col<uint32_t>(po.layout.oContig), R * 4);
    memcpy(out->left, po.col<uint32_t>(po.layout.oLeft), R * 4);
    memcpy(out->len, po.col<uint32_t>(po.layout.oLen), R * 4);
    memcpy(out->sum, po.col<uint32_t>(po.layout.oWsum), R * 4);
    for (size_t i = 0; i < R; ++i) out->replayed[i] = out->left[i] <= po.unitEdgeStop[unitOfContig[out->contig[i]]] ? 1 : 0;
  }
  return UPB_OK;
}

extern "C" int upb_shift_examine_order (uint64_t n, const uint32_t* sums, uint64_t want, uint32_t* order_out, uint64_t* sorted_end) {
  if ((n && (!sums || !order_out)) || !sorted_end || n > 0xFFFFFFFFull) return UPB_ERR_BAD_ARG;
  std::vector<SumIdx> order((size_t)n);
  for (size_t i = 0; i < (size_t)n; ++i) { order[i].sum = sums[i]; order[i].idx = (uint32_t)i; }
  *sorted_end = sort_prefix_like_std_sort(order.begin(), order.end(), (size_t)std::min<uint64_t>(want, n), compareRegionCounts);
  for (size_t i = 0; i < (size_t)n; ++i) order_out[i] = order[i].idx;
  return UPB_OK;
}

extern "C" int upb_shift_sweep (upb_ctx* ctx, const upb_shift_params* p, uint64_t n, const uint32_t* contig, const uint32_t* left, const uint32_t* len,
                                uint32_t* best_shift_out, double* best_corr_out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || (n && (!contig || !left || !len || !best_shift_out || !best_corr_out))) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_sweep: NULL argument");
  if (n > 0x7FFFFFFFull) return fail(ctx, UPB_ERR_UNSUPPORTED, "upb_shift_sweep: too many regions in one batch");
  if (!n) return UPB_OK;
  CU(cudaSetDevice(ctx->device));
  const PipeCfg cfg = shift_cfg(p);
  Prep pr;
  uint32_t launches = 0;
  int r = prepare(ctx, cfg, pr, &launches, true);
  if (r) return r;
  const uint32_t C = (uint32_t)ctx->sizes.size();
  std::vector<uint32_t> unitOfContig(C, 0xFFFFFFFFu);
  for (size_t u = 0; u < pr.units.size(); ++u) unitOfContig[pr.units[u].contig] = (uint32_t)u;
  for (uint64_t k = 0; k < n; ++k) {
    if (contig[k] >= C || unitOfContig[contig[k]] == 0xFFFFFFFFu) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_sweep: a region lies on a contig without tags on this context");
    if (len[k] <= 2 * p->max_shift + 3) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_sweep: a region is too short to try all shifts (src/strand_shift.cpp:206)");
    if (left[k] < 1 || (uint64_t)left[k] + len[k] - 1 > (uint64_t)ctx->sizes[contig[k]] + p->bandwidth) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_sweep: a region lies outside its contig");
  }
  r = shift_sweep_batch(ctx, pr.dens, unitOfContig, (uint32_t)n, contig, left, len, p->max_shift + 1, best_shift_out, best_corr_out, &launches);
  ctx->stats.launches = launches;
  return r;
}

extern "C" int upb_shift_mode (const uint64_t* shift_freq, uint32_t n_freq, uint32_t min_shift, uint32_t* best_out) {
  if (!shift_freq || !best_out) return UPB_ERR_BAD_ARG;
  *best_out = shift_mode(shift_freq, n_freq, min_shift);
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// bin/regions -w: every position with score != 0 of one (buffer, contig) — misc/peakcall.cpp:81-84
// ------------------------------------------------------------------------------------------------
extern "C" void upb_free_profile (upb_profile* p) { if (!p) return; free(p->pos); free(p->score); memset(p, 0, sizeof *p); }

extern "C" int upb_profile_contig (upb_ctx* ctx, const upb_region_params* p, int buffer, uint32_t contig, upb_profile* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !out) return fail(ctx, UPB_ERR_BAD_ARG, "upb_profile_contig: NULL argument");
  memset(out, 0, sizeof *out);
  const uint32_t C = (uint32_t)ctx->sizes.size();
  if (contig >= C || buffer < 0 || buffer > 1 || (!p->directional && buffer != 0)) return fail(ctx, UPB_ERR_BAD_ARG, "upb_profile_contig: bad buffer/contig");
  CU(cudaSetDevice(ctx->device));
  PipeCfg cfg;
  cfg.bw = p->bandwidth; cfg.thr = p->region_threshold; cfg.kurtThr = p->kurtosis_threshold; cfg.corrThr = p->corr_threshold;
  cfg.hitThr = p->hit_threshold; cfg.background = p->background; cfg.nondir = !p->directional; cfg.control = p->control; cfg.needCorr = false;
  cfg.nCoeffs = p->n_coeffs; cfg.coeffs = p->coeffs;
  Prep pr;
  uint32_t launches = 0;
  int r = prepare(ctx, cfg, pr, &launches);
  if (r) return r;
  int unit = -1;
  for (size_t u = 0; u < pr.units.size(); ++u) if (pr.units[u].contig == contig && pr.units[u].buffer == (uint32_t)buffer) unit = (int)u;
  if (unit < 0) return UPB_OK;      // no non-control tags: the reference writes nothing for it
  // evaluate the contig in slabs of 8M positions; the reference evaluates firstTag-bw .. lastTag+bw (Q20)
  const uint64_t last = (uint64_t)ctx->sizes[contig] + cfg.bw;
  const uint32_t SLAB = 8u << 20;
  std::vector<uint32_t> posOut;
  std::vector<double> scoreOut;
  for (uint64_t start = 1; start <= last; start += SLAB) {
    const uint32_t len = (uint32_t)std::min<uint64_t>(SLAB, last - start + 1);
    const uint32_t desc[4] = {(uint32_t)unit, (uint32_t)start, len, (len + TILE - 1) / TILE};
    uint32_t* dUnit = SCR("dUnit", uint32_t, 1); uint32_t* dLeft = SCR("dLeft", uint32_t, 1);
    uint32_t* dLen = SCR("dLen", uint32_t, 1); uint32_t* dChunks = SCR("dChunks", uint32_t, 1);
    CU(cudaMemcpyAsync(dUnit, &desc[0], 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dLeft, &desc[1], 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dLen, &desc[2], 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(dChunks, &desc[3], 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((r = dump_and_corr(ctx, pr.dens, cfg.nondir, 1, 0, &launches))) return r;
    // compact the non-zero positions on the device; only those cross PCIe
    uint32_t* d_flag = SCR("profFlag", uint32_t, len);
    unsigned long long* d_poff = SCR("profOff", unsigned long long, len);
    const double* dF = ctx->scratch["F"].as<double>();
    const double* dR = ctx->scratch["R"].as<double>();
    unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 16;
    profile_flag_kernel<<<(len + 255) / 256, 256, 0, ctx->stream>>>(dF, dR, len, d_flag);
    if ((r = scan_u32(ctx, d_flag, len, d_poff, d_tot, &launches))) return r;
    unsigned long long* h64 = (unsigned long long*)ctx->h_small;
    CU(cudaMemcpyAsync(h64, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    const size_t nz = (size_t)h64[0];
    if (nz) {
      uint32_t* d_pp = SCR("profPos", uint32_t, nz);
      double* d_ps = SCR("profScore", double, nz);
      profile_emit_kernel<<<(len + 255) / 256, 256, 0, ctx->stream>>>(dF, dR, len, (uint32_t)start, d_flag, d_poff, d_pp, d_ps);
      CU(cudaGetLastError());
      const size_t at = posOut.size();
      posOut.resize(at + nz); scoreOut.resize(at + nz);
      CU(cudaMemcpyAsync(posOut.data() + at, d_pp, nz * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaMemcpyAsync(scoreOut.data() + at, d_ps, nz * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
  }
  out->n = posOut.size();
  out->pos = alloc_arr<uint32_t>(out->n);
  out->score = alloc_arr<double>(out->n);
  if (!out->pos || !out->score) { upb_free_profile(out); return fail(ctx, UPB_ERR_NOMEM, "out of host memory for the profile"); }
  memcpy(out->pos, posOut.data(), out->n * 4);
  memcpy(out->score, scoreOut.data(), out->n * 8);
  return UPB_OK;
}

extern "C" int upb_count_tags_in_regions (upb_ctx* ctx, int directional, uint32_t extension, uint64_t n_regions,
                                          const upb_region_key* regions, uint32_t* counts_out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!ctx->loaded) return fail(ctx, UPB_ERR_NOT_LOADED, "no tags loaded (upb_set_contigs + upb_load_tags first)");
  if (n_regions && (!regions || !counts_out)) return fail(ctx, UPB_ERR_BAD_ARG, "upb_count_tags_in_regions: NULL argument");
  CU(cudaSetDevice(ctx->device));
  const uint32_t C = (uint32_t)ctx->sizes.size(), n = ctx->nSamples;
  if (C >= (1u << 28)) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many contigs");
  if (n_regions == 0) return UPB_OK;
  // src/tags_in_regions.cpp:141-172: strand from the coordinate order, optional extension (Pos arithmetic wraps like the reference's)
  std::vector<TirRegion> tr(n_regions);
  for (uint64_t k = 0; k < n_regions; ++k) {
    const upb_region_key& q = regions[k];
    if (q.contig >= C) return fail(ctx, UPB_ERR_BAD_ARG, "upb_count_tags_in_regions: contig not in table");
    const bool forward = q.end >= q.start;
    if (!forward && !directional) return fail(ctx, UPB_ERR_REVERSE_NONDIR, "reverse regions in non-directional analysis");   // :167-170
    const uint32_t lo = forward ? q.start : q.end;
    tr[k].contig = q.contig;
    tr[k].left = lo < extension ? 0u : lo - extension;
    tr[k].right = (forward ? q.end : q.start) + extension;
    tr[k].strand = forward ? 0u : 1u;
  }
  TirArgs ta;
  ta.T = make_tags(ctx, false);
  TirRegion* d_regions = SCR("tirRegions", TirRegion, n_regions);
  ta.regions = d_regions; ta.nRegions = n_regions; ta.directional = directional ? 1 : 0;
  ta.hits = SCR("tirHits", uint32_t, n_regions * n);
  ta.E = SCR("tirE", unsigned long long, n_regions * n);
  ta.M = SCR("tirM", unsigned long long, n_regions * n);
  CU(cudaMemcpyAsync(d_regions, tr.data(), n_regions * sizeof(TirRegion), cudaMemcpyHostToDevice, ctx->stream));
  const uint64_t total = n_regions * n;
  const unsigned blocks = (unsigned)((total + 255) / 256);
  ta.firstDead = SCR("tirDead", unsigned long long, n);
  CU(cudaMemsetAsync(ta.firstDead, 0xFF, (size_t)n * 8, ctx->stream));
  tir_eval_kernel<<<blocks, 256, 0, ctx->stream>>>(ta, 0);
  tir_scan_kernel<<<n, 1024, 0, ctx->stream>>>(ta.E, ta.M, n_regions);
  if (directional) tir_dead_kernel<<<blocks, 256, 0, ctx->stream>>>(ta);
  tir_eval_kernel<<<blocks, 256, 0, ctx->stream>>>(ta, 1);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(counts_out, ta.hits, total * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stats.launches = directional ? 4 : 3;
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// bin/convert_align: tag counts per (strand, contig, 5' position)
// ------------------------------------------------------------------------------------------------
extern "C" void upb_free_tag_counts (upb_tag_counts* t) {
  if (!t) return;
  free(t->off); free(t->pos); free(t->cnt);
  memset(t, 0, sizeof *t);
}

extern "C" int upb_aggregate_tags (upb_ctx* ctx, uint64_t n, const uint32_t* key, const uint32_t* pos, const uint32_t* cnt, int combine_strands,
                                   upb_tag_counts* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!out || (n && (!key || !pos))) return fail(ctx, UPB_ERR_BAD_ARG, "upb_aggregate_tags: NULL argument");
  if (ctx->sizes.empty()) return fail(ctx, UPB_ERR_NOT_LOADED, "upb_aggregate_tags: call upb_set_contigs first");
  memset(out, 0, sizeof *out);
  CU(cudaSetDevice(ctx->device));
  const uint32_t C = (uint32_t)ctx->sizes.size();
  std::vector<uint32_t> mtBase(2 * (size_t)C + 1, 0);
  { uint64_t run = 0;
    for (uint32_t d = 0; d < 2; ++d) for (uint32_t c = 0; c < C; ++c) { mtBase[(size_t)d * C + c] = (uint32_t)run; run += ((uint64_t)(ctx->sizes[c] - 1) >> MT_SHIFT) + 1; }
    if (run > 0xFFFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "genome too large for the mini-tile index");
    mtBase[2 * (size_t)C] = (uint32_t)run; }
  const uint32_t nMt = mtBase.back();
  out->off = (uint64_t*)calloc(2 * (size_t)C + 1, 8);
  if (!out->off) return fail(ctx, UPB_ERR_NOMEM, "out of host memory");
  if (n == 0) return UPB_OK;
  AggrArgs a;
  memset(&a, 0, sizeof a);
  uint32_t* d_key = SCR("agKey", uint32_t, n); uint32_t* d_pos = SCR("agPos", uint32_t, n);
  uint32_t* d_cnt = cnt ? SCR("agCnt", uint32_t, n) : nullptr;
  uint32_t* d_mtBase = SCR("agMtBase", uint32_t, 2 * (size_t)C + 1);
  uint32_t* d_count = SCR("agCount", uint32_t, (size_t)nMt + 1); uint32_t* d_cursor = SCR("agCursor", uint32_t, (size_t)nMt + 1);
  unsigned long long* d_mtOff = SCR("agMtOff", unsigned long long, (size_t)nMt + 2);
  uint2* d_ent = SCR("agEnt", uint2, n);
  uint32_t* d_uniq = SCR("agUniq", uint32_t, (size_t)nMt + 1);
  unsigned long long* d_outOff = SCR("agOutOff", unsigned long long, (size_t)nMt + 2);
  unsigned long long* d_off = SCR("agOff", unsigned long long, 2 * (size_t)C + 2);
  ENSURE(ctx->d_small, 4096);
  CU(cudaMemcpyAsync(d_key, key, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_pos, pos, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (cnt) CU(cudaMemcpyAsync(d_cnt, cnt, n * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(d_mtBase, mtBase.data(), (2 * (size_t)C + 1) * 4, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(d_count, 0, ((size_t)nMt + 1) * 4, ctx->stream));
  CU(cudaMemsetAsync(d_cursor, 0, ((size_t)nMt + 1) * 4, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_small.p, 0, 4096, ctx->stream));
  a.key = d_key; a.pos = d_pos; a.cnt = d_cnt; a.n = n; a.sizes = ctx->d_sizes.as<uint32_t>(); a.mtBase = d_mtBase; a.C = C;
  a.combine = combine_strands ? 1 : 0; a.mtCount = d_count; a.cursor = d_cursor; a.mtOff = d_mtOff; a.ent = d_ent; a.uniq = d_uniq;
  a.outOff = d_outOff; a.nMt = nMt; a.flags = ctx->d_small.as<uint32_t>();
  unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 16;
  a.tagCount = d_tot + 4;
  uint32_t launches = 0;
  const unsigned gN = (unsigned)((n + 255) / 256), gM = (nMt + AG_WARPS - 1) / AG_WARPS;
  aggr_sort_kernel<false><<<gN, 256, 0, ctx->stream>>>(a);
  int r = scan_u32(ctx, d_count, nMt, d_mtOff, d_tot, &launches);
  if (r) return r;
  CU(cudaMemcpyAsync(d_mtOff + nMt, d_tot, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  aggr_sort_kernel<true><<<gN, 256, 0, ctx->stream>>>(a);
  aggr_reduce_kernel<false><<<gM, AG_WARPS * 32, 0, ctx->stream>>>(a);
  if ((r = scan_u32(ctx, d_uniq, nMt, d_outOff, d_tot + 1, &launches))) return r;
  CU(cudaMemcpyAsync(d_outOff + nMt, d_tot + 1, 8, cudaMemcpyDeviceToDevice, ctx->stream));
  unsigned long long* h64 = (unsigned long long*)ctx->h_small;
  CU(cudaMemcpyAsync(h64, d_tot, 64, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(h64 + 8, ctx->d_small.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  if (((uint32_t*)(h64 + 8))[0] & FLAG_BAD_POS) return fail(ctx, UPB_ERR_BAD_ARG, "upb_aggregate_tags: strand/contig key or position out of range");
  const uint64_t nOut = h64[1];
  out->n = nOut; out->tag_count = h64[4];
  out->pos = (uint32_t*)malloc((nOut ? nOut : 1) * 4); out->cnt = (uint32_t*)malloc((nOut ? nOut : 1) * 4);
  if (!out->pos || !out->cnt) { upb_free_tag_counts(out); return fail(ctx, UPB_ERR_NOMEM, "out of host memory"); }
  a.outPos = SCR("agOutPos", uint32_t, nOut + 1); a.outCnt = SCR("agOutCnt", uint32_t, nOut + 1);
  aggr_reduce_kernel<true><<<gM, AG_WARPS * 32, 0, ctx->stream>>>(a);
  aggr_offsets_kernel<<<(2 * C + 1 + 255) / 256, 256, 0, ctx->stream>>>(d_outOff, d_mtBase, 2 * C, d_off);
  CU(cudaGetLastError());
  CU(cudaMemcpyAsync(out->pos, a.outPos, nOut * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(out->cnt, a.outCnt, nOut * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(out->off, d_off, (2 * (size_t)C + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->stats.launches = launches + 6;
  return UPB_OK;
}

