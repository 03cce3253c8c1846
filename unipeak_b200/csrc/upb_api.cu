// upb_api.cu — C-ABI (include/unipeak_b200.h) and host orchestration of the sm_100a kernels.
//
// There is no CPU implementation behind these entry points: every result row comes from the kernels in
// upb_device.cuh.  The host side only sizes buffers, launches, and assembles the reference-shaped
// result tables (src/regions.cpp, src/strand_shift.cpp, src/tags_in_regions.cpp drivers).
#include "../../include/unipeak_b200.h"
#include "upb_device.cuh"
#include "upb_span.cuh"
#include "upb_pack.cuh"
#include "upb_vbyte.hpp"
#include "upb_aggr.cuh"
#include "upb_introsort.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <memory>
#include <mutex>
#include <atomic>
#include <string>
#include <thread>
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

struct LoadGroup {                      // (strand d, contigs c0 .. c1-1): one unit of the pipelined load
  uint32_t d, c0, c1;
  uint64_t entries, mentBase;           // its entries and where they start in the tile-major copy
  uint32_t m0, m1, w0, w1;              // its mini-tiles and scatter windows
  uint32_t seg0, nSeg, tileTab0, nTiles;   // its non-empty segments / tile table (d_grpSeg, d_grpTile)
};

struct upb_ctx {
  int device = 0;
  int smCount = 148;
  cudaStream_t stream = nullptr;
  // A region call works through its units in groups; consecutive groups alternate between two LANES — `stream` and `laneStream`, each with its
  // own set of per-group scratch arrays and its own GroupCtr — so that one group's latency-bound tail (order / stitch / pooling / rows over PCIe)
  // runs beside the next group's span kernel.  `cur` / `lane` are what the helpers (scan_dev, launch_span, SCR) enqueue on / allocate for.
  cudaStream_t laneStream = nullptr;
  cudaStream_t cur = nullptr;
  int lane = 0;
  cudaEvent_t laneSetupEv = nullptr, laneDoneEv = nullptr, lanePoolEv = nullptr;
  size_t lastN1 = 0;                    // hit-filter survivors of the last region call (decides whether the next call splits its tail)
  std::vector<cudaEvent_t> placeEv;     // per group: its rows have their place (row_place_kernel done)
  std::vector<cudaEvent_t> spanEv;      // per group: before / after its span kernel
  cudaStream_t copyStream = nullptr;    // H2D copies of the load run here, one load group ahead of the kernels that consume them ...
  cudaStream_t copyStream2 = nullptr;   // second H2D queue: the copies of a group alternate between the two, so the next descriptor is always ready
  cudaStream_t loadStream = nullptr;    // ... on this one (decode, validation, counting sort, zone maps), while `stream` may already run the region pipeline of earlier groups
  cudaEvent_t loadDoneEv = nullptr;
  std::vector<cudaEvent_t> groupEv, copyEvs, copyEvs2;   // per load group: index built / bytes on the device
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
  std::vector<uint64_t> off;            // offsets of the RESIDENT arrays in the caller's layout [(s*2+d)*(C+1)+c]
  std::vector<uint64_t> segStart, srcStart;   // [nSeg+1] / [nSeg] first entry of segment g on the device / in the caller's arrays
  std::vector<uint8_t> keptContig;      // [C] contigs whose segments this context holds (a rank of a multi-GPU job keeps the ones it owns)
  std::vector<LoadGroup> groups;        // load groups of the current data, in transfer order
  bool loadPending = false;             // upb_load_tags_async returned, nobody has waited yet
  bool loadStrandMajor = true;
  bool loadCopied = false;              // the current data came over PCIe (not upb_reindex)
  bool lastPosKnown = false;            // packed loads learn the last position of a segment from the device
  HostSlab h_ends;                      // pinned: first / last positions and the validation flags come back here
  std::vector<uint32_t> firstPos, lastPos;   // [(s*2+d)*C + c], 0 if empty
  DevBuf d_pos, d_cnt, d_off, d_segStart, d_tidx, d_tidxBase, d_sizes, d_nTilesIdx, d_small;
  DevBuf d_winBase, d_wEnt;                                          // per-sample window index of the scatter pass
  DevBuf d_pk16, d_pkExc, d_segEnds, d_srcStart, d_grpSeg, d_grpTile, d_loadFlags;   // load (upb_pack.cuh)
  DevBuf d_tileExc, d_tileDesc;        // packed load: first exception of every tile of the load group being decoded, and where the tile is (load_tile_prep_kernel)
  DevBuf d_vbytes, d_vSrcByte, d_vDevByte, d_vSrcTile, d_vTileOff;   // variable-length byte form (upb_vbyte.hpp)
  DevBuf d_ment, d_mtOff, d_mtBase, d_mtCount, d_mEff, d_mtW, d_blkW;   // tile-major copy and its zone maps (upb_device.cuh)
  std::vector<uint32_t> mtBase;         // [2*C+1]
  std::map<std::string, DevBuf> scratch;     // grow-only device scratch, by name
  unsigned char* h_small = nullptr;          // 4 KB of pinned host memory for small read-backs
  std::vector<HostSlab> slabPool;            // pinned host slabs of freed region tables, reused by the next call
  upb_stats stats;
  // what the region pipeline remembers from call to call: capacities of its arrays (a call runs without reading a size back when
  // the previous call's numbers, plus a quarter, still hold; otherwise it sizes them itself)
  size_t needRunCap = 0, capN1 = 0, capCells = 0, capRows = 0;
  HostSlab h_res;                            // pinned block the counters of a call come back in
  HostSlab h_param;                          // pinned block the small per-call tables are staged in (read by param_copy_kernel)
  bool profiling = false;                    // upb_set_profiling: one CUDA event between the stages of the pipeline
  std::vector<cudaEvent_t> stageEv;
  size_t stageUsed = 0;
  std::vector<const char*> stageNames;
  std::vector<float> stageMs;
  std::vector<uint8_t> lastControl;     // control set cntEff was built for
  uint64_t callNo = 0;                  // region calls of this context so far (trace output)
  std::vector<std::pair<const char*, std::chrono::steady_clock::time_point>> hostMarks;   // when the host passed the stage marks of the last call
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
  const std::chrono::steady_clock::time_point tc0 = std::chrono::steady_clock::now();
  if ((e = cudaSetDevice(device)) != cudaSuccess) { delete ctx; return fail(nullptr, UPB_ERR_CUDA, std::string("cudaSetDevice: ") + cudaGetErrorString(e)); }
  cudaDeviceProp prop;
  if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) { delete ctx; return fail(nullptr, UPB_ERR_CUDA, std::string("cudaGetDeviceProperties: ") + cudaGetErrorString(e)); }
  ctx->smCount = prop.multiProcessorCount;
  // The region pipeline is a chain of ~27 small dependent kernels per group of units; when it runs behind a load (upb_load_tags_async) the
  // index build of the groups still arriving keeps the device full of blocks, and every link of the chain then queues behind them.  Priority
  // UPB_PRIO: 0 none, 1 (default) the call stream above the load stream, 2 the other way round
  int prLo = 0, prHi = 0;
  cudaDeviceGetStreamPriorityRange(&prLo, &prHi);                     // lowest, highest (numerically smaller = higher)
  const int mode = getenv("UPB_PRIO") ? atoi(getenv("UPB_PRIO")) : 1;
  if ((e = cudaStreamCreateWithPriority(&ctx->stream, cudaStreamNonBlocking, mode == 1 ? prHi : prLo)) != cudaSuccess) { delete ctx; return fail(nullptr, UPB_ERR_CUDA, std::string("cudaStreamCreate: ") + cudaGetErrorString(e)); }
  ctx->cur = ctx->stream;
  cudaStreamCreateWithPriority(&ctx->laneStream, cudaStreamNonBlocking, mode == 1 ? prHi : prLo);
  cudaEventCreateWithFlags(&ctx->laneSetupEv, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->laneDoneEv, cudaEventDisableTiming);
  cudaEventCreateWithFlags(&ctx->lanePoolEv, cudaEventDisableTiming);
  for (int i = 0; i < 4; ++i) cudaEventCreate(&ctx->ev[i]);
  cudaStreamCreateWithFlags(&ctx->copyStream, cudaStreamNonBlocking);
  cudaStreamCreateWithPriority(&ctx->loadStream, cudaStreamNonBlocking, mode == 2 ? prHi : prLo);
  cudaStreamCreateWithFlags(&ctx->copyStream2, cudaStreamNonBlocking);
  cudaEventCreateWithFlags(&ctx->loadDoneEv, cudaEventDisableTiming);
  if ((e = cudaMallocHost((void**)&ctx->h_small, 4096)) != cudaSuccess) { upb_destroy(ctx); return fail(nullptr, UPB_ERR_NOMEM, std::string("cudaMallocHost: ") + cudaGetErrorString(e)); }
  memset(&ctx->stats, 0, sizeof ctx->stats);
  if (getenv("UPB_TRACE") && *getenv("UPB_TRACE"))
    fprintf(stderr, "[upb] create: device %d, CUDA context + streams + pinned block in %.1f ms\n", device, std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - tc0).count());
  *out = ctx;
  return UPB_OK;
}

extern "C" void upb_destroy (upb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  DevBuf* all[] = {&ctx->d_vbytes, &ctx->d_vSrcByte, &ctx->d_vDevByte, &ctx->d_vSrcTile, &ctx->d_vTileOff, &ctx->d_tileExc, &ctx->d_tileDesc, &ctx->d_winBase, &ctx->d_wEnt, &ctx->d_pk16, &ctx->d_pkExc, &ctx->d_segEnds, &ctx->d_srcStart, &ctx->d_grpSeg, &ctx->d_grpTile, &ctx->d_loadFlags, &ctx->d_mtW, &ctx->d_blkW, &ctx->d_ment, &ctx->d_mtOff, &ctx->d_mtBase, &ctx->d_mtCount, &ctx->d_mEff, &ctx->d_pos, &ctx->d_cnt, &ctx->d_off, &ctx->d_segStart, &ctx->d_tidx, &ctx->d_tidxBase, &ctx->d_sizes,
    &ctx->d_nTilesIdx, &ctx->d_small};
  for (DevBuf* b : all) release(*b);
  for (auto& kv : ctx->scratch) release(kv.second);
  for (HostSlab& h : ctx->slabPool) if (h.p) cudaFreeHost(h.p);
  if (ctx->h_small) cudaFreeHost(ctx->h_small);
  if (ctx->h_res.p) cudaFreeHost(ctx->h_res.p);
  if (ctx->h_param.p) cudaFreeHost(ctx->h_param.p);
  for (cudaEvent_t e : ctx->stageEv) cudaEventDestroy(e);
  for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
  if (ctx->loadDoneEv) cudaEventDestroy(ctx->loadDoneEv);
  for (cudaEvent_t e : ctx->groupEv) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->copyEvs) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->copyEvs2) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->placeEv) cudaEventDestroy(e);
  for (cudaEvent_t e : ctx->spanEv) cudaEventDestroy(e);
  if (ctx->laneSetupEv) cudaEventDestroy(ctx->laneSetupEv);
  if (ctx->laneDoneEv) cudaEventDestroy(ctx->laneDoneEv);
  if (ctx->lanePoolEv) cudaEventDestroy(ctx->lanePoolEv);
  if (ctx->laneStream) cudaStreamDestroy(ctx->laneStream);
  if (ctx->copyStream2) cudaStreamDestroy(ctx->copyStream2);
  if (ctx->h_ends.p) cudaFreeHost(ctx->h_ends.p);
  if (ctx->copyStream) cudaStreamDestroy(ctx->copyStream);
  if (ctx->loadStream) cudaStreamDestroy(ctx->loadStream);
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
  for (uint32_t c = 0; c < n_contigs; ++c) if (sizes[c] == 0) return fail(ctx, UPB_ERR_BAD_ARG, "upb_set_contigs: a contig of size 0 (it cannot hold a tag; the index arithmetic assumes size >= 1)");
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

static int ensure_host (upb_ctx* ctx, HostSlab& h, size_t bytes);

// ------------------------------------------------------------------------------------------------
// the load: host arrays -> resident sample-major arrays + tile-major copy + zone maps
//
// Work is cut into LOAD GROUPS — (strand, range of consecutive contigs this context keeps), sized to ~1/8 of the entries of a
// strand — and pipelined group by group over three streams: the H2D copies of group k+1 (copy stream) run while group k is
// decoded, validated, histogrammed (load_tile_kernel), scanned and scattered into the tile-major copy with its zone maps
// (mt_scatter_kernel) on the load stream, and — when the load was started with upb_load_tags_async — while the region pipeline
// of an earlier group already runs on the context's main stream (run_pipeline_once waits on the groups' events, not on the host).
// r01 pipelined sample by sample, which left the scatter, the call and the table copy (~5 ms) exposed after the last byte.
// ------------------------------------------------------------------------------------------------
static int scan_load (upb_ctx* ctx, const uint32_t* d_in, uint64_t n, unsigned long long* d_out, unsigned long long bias, unsigned long long* d_total) {
  const uint64_t maxTiles = (n + SCAN_BLOCK - 1) / SCAN_BLOCK;
  DevBuf& b = ctx->scratch["loadScanPartial"];                        // not the pipeline's: the two run on different streams
  { int r = ensure(ctx, b, (maxTiles + 2) * 8); if (r) return r; }
  unsigned long long* part = b.as<unsigned long long>();
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(maxTiles, (uint64_t)ctx->smCount * 4));
  const DevN nn = devn_imm(n);
  scanv_partial_kernel<<<grid, SCAN_THREADS, 0, ctx->loadStream>>>(LoadU32{d_in}, nn, part);
  scanv_top_kernel<<<1, SCAN_THREADS, 0, ctx->loadStream>>>(part, nn, d_total);
  scanv_final_kernel<<<grid, SCAN_THREADS, 0, ctx->loadStream>>>(LoadU32{d_in}, nn, part, d_out, bias);
  CU(cudaGetLastError());
  return UPB_OK;
}

static int load_finish (upb_ctx* ctx);

// enqueue the whole load (resident == false: copies from `src`; true: rebuild the index from the arrays already on the device)
static int load_enqueue (upb_ctx* ctx, const upb_tag_source* src, bool resident) {
  const uint32_t C = (uint32_t)ctx->sizes.size();
  const uint32_t n_samples = resident ? ctx->nSamples : src->n_samples;
  const uint32_t nSeg = n_samples * 2 * C;
  const bool vbyte = !resident && src->vbytes != nullptr;
  const bool packed = !resident && !vbyte && src->words != nullptr;
  cudaStream_t ls = ctx->loadStream, cs = ctx->copyStream;
  std::vector<uint64_t> vSrcByte, vDevByte, vSrcTile;                  // vbyte form: per segment
  if (vbyte) {
    if (src->words || !src->vbyte_off || !src->vtile_off) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_vbyte: byte_off / tile_off are NULL (or two packed forms at once)");
    vSrcByte.assign((size_t)nSeg + 1, 0); vDevByte.assign((size_t)nSeg + 1, 0); vSrcTile.assign((size_t)nSeg + 1, 0);
    uint64_t tiles = 0, devAt = 0, expectB = 0;
    for (uint32_t sd = 0; sd < n_samples * 2; ++sd) {
      const uint64_t* o = src->off + (size_t)sd * (C + 1);
      const uint64_t* bo = src->vbyte_off + (size_t)sd * (C + 1);
      if (bo[0] != expectB) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_vbyte: byte offset blocks must be contiguous and start at 0");
      for (uint32_t c = 0; c < C; ++c) {
        if (o[c + 1] < o[c] || bo[c + 1] < bo[c]) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_vbyte: offsets must be non-decreasing");
        const uint64_t len = o[c + 1] - o[c], nb = bo[c + 1] - bo[c], nt = (len + PK_TILE - 1) / PK_TILE;
        if (nb < len || nb > len * upb_vb::VB_MAX_ENTRY) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_vbyte: a segment's bytes cannot hold its entries");
        const size_t g = (size_t)sd * C + c;
        vSrcByte[g] = bo[c]; vDevByte[g] = devAt; vSrcTile[g] = tiles;
        if (nt && (tiles + nt > src->n_vtiles || src->vtile_off[tiles] != bo[c])) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_vbyte: the tile table does not match the segments");
        tiles += nt;
        if (!src->keep_contig || src->keep_contig[c]) devAt += nb;
      }
      expectB = bo[C];
    }
    vSrcByte[nSeg] = expectB; vDevByte[nSeg] = devAt; vSrcTile[nSeg] = tiles;
    if (tiles != src->n_vtiles || src->vtile_off[tiles] != expectB) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_vbyte: the tile table does not match the segments");
  }
  if (!resident) {
    const uint64_t* off = src->off;
    const size_t nOff = (size_t)n_samples * 2 * (C + 1);
    // offsets must describe one contiguous, monotone packing (what TagData / the CLI front end produce)
    ctx->srcStart.assign(nSeg, 0);
    ctx->segStart.assign((size_t)nSeg + 1, 0);
    uint64_t expect = off[0], run = 0;
    if (expect != 0) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: off[0] must be 0");
    for (uint32_t sd = 0; sd < n_samples * 2; ++sd) {
      const uint64_t* o = off + (size_t)sd * (C + 1);
      if (o[0] != expect) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: offset blocks must be contiguous");
      for (uint32_t c = 0; c < C; ++c) {
        if (o[c + 1] < o[c]) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: offsets must be non-decreasing");
        if (o[c + 1] - o[c] > 0xFFFFFFFFull) return fail(ctx, UPB_ERR_UNSUPPORTED, "upb_load_tags: more than 2^32 entries in one segment");
        const size_t g = (size_t)sd * C + c;
        ctx->srcStart[g] = o[c];
        ctx->segStart[g] = run;
        if (!src->keep_contig || src->keep_contig[c]) run += o[c + 1] - o[c];
      }
      expect = o[C];
    }
    ctx->segStart[nSeg] = run;
    const uint64_t Esrc = expect, E = run;
    if (Esrc && !packed && !vbyte && (!src->pos || !src->cnt)) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: pos/cnt are NULL");
    if ((packed || vbyte) && src->n_exceptions && !src->exceptions) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_packed2 / _vbyte: NULL exception list");
    ctx->nSamples = n_samples;
    ctx->nEntries = E;
    // device-side offsets in the layout of the caller's `off` (density kernel, tags_in_regions)
    ctx->off.assign(nOff, 0);
    for (uint32_t sd = 0; sd < n_samples * 2; ++sd) {
      for (uint32_t c = 0; c < C; ++c) ctx->off[(size_t)sd * (C + 1) + c] = ctx->segStart[(size_t)sd * C + c];
      ctx->off[(size_t)sd * (C + 1) + C] = ctx->segStart[(size_t)(sd + 1) * C];
    }
    ctx->firstPos.assign(nSeg, 0);
    ctx->lastPos.assign(nSeg, 0);
    ctx->lastPosKnown = !packed && !vbyte;
    for (uint32_t g = 0; g < nSeg; ++g) {
      const uint64_t len = ctx->segStart[g + 1] - ctx->segStart[g];
      if (!len) continue;
      const uint64_t a = ctx->srcStart[g];
      if (vbyte) {                                                      // the first entry of a segment opens a tile: its value is the position itself
        const uint8_t* b = src->vbytes + vSrcByte[g];
        const uint64_t nb = vSrcByte[g + 1] - vSrcByte[g];
        uint32_t nCont = 0, v = 0, k = 0;
        while (nCont <= upb_vb::VB_MAX_ENTRY && 1 + nCont < nb && !(b[1 + nCont] & 0x80u)) ++nCont;
        ctx->firstPos[g] = ((b[0] & 0x80u) && upb_vb::vb_get(b, nCont, &v, &k) && v) ? v : 0xFFFFFFFFu;   // malformed: the device reports it
      }
      else if (!packed) { ctx->firstPos[g] = src->pos[a]; ctx->lastPos[g] = src->pos[a + len - 1]; }
      else {                                                            // the first entry of a segment is a position escape: its exception holds the position
        const upb_packed_exception* x = std::lower_bound(src->exceptions, src->exceptions + src->n_exceptions, a,
                                                         [] (const upb_packed_exception& e, uint64_t v) { return e.index < v; });
        ctx->firstPos[g] = (x != src->exceptions + src->n_exceptions && x->index == a && x->pos) ? x->pos : 0xFFFFFFFFu;   // malformed: the device reports it
      }
    }
  }
  const uint64_t E = ctx->nEntries;
  const size_t nOff = (size_t)n_samples * 2 * (C + 1);

  // ---- index geometry ---------------------------------------------------------------------------------------------------
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

  // ---- load groups: (strand, run of kept contigs), cut to ~1/8 of a strand's entries -----------------------------------------
  std::vector<uint64_t> perDC(2 * (size_t)C, 0);
  std::vector<uint8_t> kept(C, 0);
  for (uint32_t c = 0; c < C; ++c) kept[c] = resident ? ctx->keptContig[c] : (uint8_t)(!src->keep_contig || src->keep_contig[c]);
  ctx->keptContig = kept;
  for (uint32_t s = 0; s < n_samples; ++s) for (uint32_t d = 0; d < 2; ++d) for (uint32_t c = 0; c < C; ++c) {
    const size_t g = ((size_t)s * 2 + d) * C + c;
    perDC[(size_t)d * C + c] += ctx->segStart[g + 1] - ctx->segStart[g];
  }
  // Contig ranges of decreasing size: every (sample, range) is one copy, and a copy costs ~5 us on top of its bytes (r02: 416 copies of
  // 1.5 MB reached 46 GB/s), so the ranges are large where it does not matter — early, while the device has nothing else to do — and
  // small at the end, where the last range's index build and region calls trail the last byte.  UPB_LOAD_GROUPS=uniform: 13 equal ranges
  struct Range { uint32_t c0, c1; };
  std::vector<Range> ranges;
  if (getenv("UPB_LOAD_GROUPS") && !strcmp(getenv("UPB_LOAD_GROUPS"), "uniform")) {
    const uint64_t target = std::max<uint64_t>(E / 16 + 1, 1u << 20);  // entries (both strands) per contig range
    uint32_t c = 0;
    while (c < C) {
      if (!kept[c]) { ++c; continue; }
      Range r; r.c0 = c;
      uint64_t acc = 0;
      while (c < C && kept[c] && (acc == 0 || acc < target)) { acc += perDC[c] + perDC[(size_t)C + c]; ++c; }
      r.c1 = c;
      ranges.push_back(r);
    }
  } else {
    // (UPB_LOAD_CUM=a,b,...: other cumulative shares, for experiments)
    double cum[16] = {0.26, 0.48, 0.66, 0.80, 0.90, 0.96, 1.0};
    int nCut = 7;
    if (const char* e = getenv("UPB_LOAD_CUM"); e && *e) {
      nCut = 0;
      for (const char* q = e; *q && nCut < 15;) { char* end = nullptr; const double v = strtod(q, &end); if (end == q) break; cum[nCut++] = v; q = *end ? end + 1 : end; }
      if (nCut == 0 || cum[nCut - 1] < 1.0) cum[nCut++] = 1.0;
    }
    int part = 0;
    uint64_t acc = 0, inRange = 0;
    uint32_t c = 0;
    while (c < C) {
      if (!kept[c]) { ++c; continue; }
      Range r; r.c0 = c;
      inRange = 0;
      // a range ends at a contig that is not kept (its segments are not contiguous across it), or when the running total reaches its share
      while (c < C && kept[c]) {
        const uint64_t e = perDC[c] + perDC[(size_t)C + c];
        acc += e; inRange += e; ++c;
        if ((double)acc >= cum[part] * (double)E && inRange >= (1u << 20)) { while (part < nCut - 1 && (double)acc >= cum[part] * (double)E) ++part; break; }
      }
      r.c1 = c;
      ranges.push_back(r);
    }
  }
  const bool strandMajor = resident ? ctx->loadStrandMajor : src->directional_hint != 0;
  ctx->loadStrandMajor = strandMajor;
  ctx->groups.clear();
  auto addGroup = [&] (uint32_t d, const Range& r) { LoadGroup g; memset(&g, 0, sizeof g); g.d = d; g.c0 = r.c0; g.c1 = r.c1; ctx->groups.push_back(g); };
  if (strandMajor) { for (uint32_t d = 0; d < 2; ++d) for (const Range& r : ranges) addGroup(d, r); }
  else { for (const Range& r : ranges) for (uint32_t d = 0; d < 2; ++d) addGroup(d, r); }
  // position of every group in the tile-major copy (mini-tile order = strand, contig): entries of everything before it
  { std::vector<uint64_t> before(2 * (size_t)C + 1, 0);
    for (size_t dc = 0; dc < 2 * (size_t)C; ++dc) before[dc + 1] = before[dc] + perDC[dc];
    for (LoadGroup& g : ctx->groups) {
      g.mentBase = before[(size_t)g.d * C + g.c0];
      g.entries = before[(size_t)g.d * C + g.c1] - g.mentBase;
      g.m0 = ctx->mtBase[(size_t)g.d * C + g.c0]; g.m1 = ctx->mtBase[(size_t)g.d * C + g.c1];
      g.w0 = winBase[(size_t)g.d * C + g.c0]; g.w1 = winBase[(size_t)g.d * C + g.c1];
    } }
  // tiles of every group: its segments in (sample, contig) order
  std::vector<uint32_t> grpSeg, grpTile;
  for (LoadGroup& g : ctx->groups) {
    g.seg0 = (uint32_t)grpSeg.size();
    uint32_t tiles = 0;
    g.tileTab0 = (uint32_t)grpTile.size();
    for (uint32_t s = 0; s < n_samples; ++s) for (uint32_t c = g.c0; c < g.c1; ++c) {
      const uint32_t sg = (s * 2 + g.d) * C + c;
      const uint64_t len = ctx->segStart[(size_t)sg + 1] - ctx->segStart[sg];
      if (!len) continue;
      grpSeg.push_back(sg); grpTile.push_back(tiles);
      tiles += (uint32_t)((len + PK_TILE - 1) / PK_TILE);
    }
    grpTile.push_back(tiles);
    g.nSeg = (uint32_t)grpSeg.size() - g.seg0;
    g.nTiles = tiles;
  }
  while (ctx->groupEv.size() < ctx->groups.size()) {
    cudaEvent_t e = nullptr;
    CU(cudaEventCreate(&e));
    ctx->groupEv.push_back(e);
  }
  while (ctx->copyEvs.size() < ctx->groups.size() + 1) {                // + 1: the start of the transfers (UPB_TRACE timeline)
    cudaEvent_t e = nullptr;
    CU(cudaEventCreate(&e));
    ctx->copyEvs.push_back(e);
    CU(cudaEventCreate(&e));
    ctx->copyEvs2.push_back(e);
  }

  // ---- device arrays ------------------------------------------------------------------------------------------------------
  ENSURE(ctx->d_pos, (E ? E : 1) * 4);
  ENSURE(ctx->d_cnt, (E ? E : 1) * 4);
  ENSURE(ctx->d_off, nOff * 8);
  ENSURE(ctx->d_segStart, ((size_t)nSeg + 1) * 8);
  ENSURE(ctx->d_srcStart, ((size_t)nSeg + 1) * 8);
  ENSURE(ctx->d_small, 4096);
  ENSURE(ctx->d_loadFlags, 256);
  ENSURE(ctx->d_mtBase, (2 * (size_t)C + 1) * 4);
  ENSURE(ctx->d_winBase, (2 * (size_t)C + 1) * 4);
  ENSURE(ctx->d_wEnt, ((size_t)n_samples * nWinTotal + 2) * 4);
  ENSURE(ctx->d_mtCount, ((size_t)nMtTotal + 1) * 4);
  ENSURE(ctx->d_mtOff, ((size_t)nMtTotal + 2) * 8);
  ENSURE(ctx->d_ment, (E ? E : 1) * 8);
  ENSURE(ctx->d_mtW, ((size_t)nMtTotal + 1) * 8);
  ENSURE(ctx->d_blkW, ((size_t)nMtTotal + 1) * BLK_PER_MT * 4);
  ENSURE(ctx->d_grpSeg, (grpSeg.size() + 1) * 4);
  ENSURE(ctx->d_grpTile, (grpTile.size() + 1) * 4);
  ENSURE(ctx->d_segEnds, 2 * ((size_t)nSeg + 1) * 4);
  if (packed || vbyte) {
    uint32_t maxTiles = 0;
    for (const LoadGroup& g : ctx->groups) maxTiles = std::max(maxTiles, g.nTiles);
    ENSURE(ctx->d_tileExc, ((size_t)maxTiles + 2) * 8);
    ENSURE(ctx->d_tileDesc, ((size_t)maxTiles + 2) * sizeof(TileDesc));
    if (packed) ENSURE(ctx->d_pk16, (E ? E : 1) * 2);
    ENSURE(ctx->d_pkExc, (src->n_exceptions ? src->n_exceptions : 1) * sizeof(PkException));
  }
  if (vbyte) {
    ENSURE(ctx->d_vbytes, vDevByte[nSeg] + 64);                         // (+ slack: the decoder moves whole 16-byte words)
    ENSURE(ctx->d_vSrcByte, ((size_t)nSeg + 1) * 8);
    ENSURE(ctx->d_vDevByte, ((size_t)nSeg + 1) * 8);
    ENSURE(ctx->d_vSrcTile, ((size_t)nSeg + 1) * 8);
    ENSURE(ctx->d_vTileOff, (src->n_vtiles + 1) * 8);
  }
  { int r = ensure_host(ctx, ctx->h_ends, 2 * ((size_t)nSeg + 1) * 4 + 64); if (r) return r; }
  // small tables: staged through pageable memory by the runtime before the call returns
  CU(cudaMemcpyAsync(ctx->d_off.p, ctx->off.data(), nOff * 8, cudaMemcpyHostToDevice, ls));
  CU(cudaMemcpyAsync(ctx->d_segStart.p, ctx->segStart.data(), ((size_t)nSeg + 1) * 8, cudaMemcpyHostToDevice, ls));
  CU(cudaMemcpyAsync(ctx->d_srcStart.p, ctx->srcStart.data(), (size_t)nSeg * 8, cudaMemcpyHostToDevice, ls));
  CU(cudaMemcpyAsync(ctx->d_mtBase.p, ctx->mtBase.data(), (2 * (size_t)C + 1) * 4, cudaMemcpyHostToDevice, ls));
  CU(cudaMemcpyAsync(ctx->d_winBase.p, winBase.data(), (2 * (size_t)C + 1) * 4, cudaMemcpyHostToDevice, ls));
  if (!grpSeg.empty()) CU(cudaMemcpyAsync(ctx->d_grpSeg.p, grpSeg.data(), grpSeg.size() * 4, cudaMemcpyHostToDevice, ls));
  CU(cudaMemcpyAsync(ctx->d_grpTile.p, grpTile.data(), grpTile.size() * 4, cudaMemcpyHostToDevice, ls));
  if (vbyte) {
    CU(cudaMemcpyAsync(ctx->d_vSrcByte.p, vSrcByte.data(), ((size_t)nSeg + 1) * 8, cudaMemcpyHostToDevice, ls));
    CU(cudaMemcpyAsync(ctx->d_vDevByte.p, vDevByte.data(), ((size_t)nSeg + 1) * 8, cudaMemcpyHostToDevice, ls));
    CU(cudaMemcpyAsync(ctx->d_vSrcTile.p, vSrcTile.data(), ((size_t)nSeg + 1) * 8, cudaMemcpyHostToDevice, ls));
    CU(cudaMemcpyAsync(ctx->d_vTileOff.p, src->vtile_off, (src->n_vtiles + 1) * 8, cudaMemcpyHostToDevice, ls));
  }
  CU(cudaMemsetAsync(ctx->d_loadFlags.p, 0, 256, ls));
  CU(cudaMemsetAsync(ctx->d_mtCount.p, 0, ((size_t)nMtTotal + 1) * 4, ls));
  CU(cudaMemsetAsync(ctx->d_mtOff.p, 0, ((size_t)nMtTotal + 2) * 8, ls));
  // ranges of contigs this context does not keep stay empty, but mtOff stays a prefix array over ALL mini-tiles (mtOff[m + 1] - mtOff[m] is a count
  // for every m: mt_control_kernel walks them all): a range that is not kept repeats the number of entries before it
  { uint64_t before = 0;
    for (size_t dc = 0; dc < 2 * (size_t)C; ++dc) {
      const uint32_t m0 = ctx->mtBase[dc], m1 = ctx->mtBase[dc + 1];
      if (!kept[dc % C] && m1 > m0) fill_u64_kernel<<<(m1 - m0 + 255) / 256, 256, 0, ls>>>(ctx->d_mtOff.as<unsigned long long>() + m0, m1 - m0, before);
      before += perDC[dc];
    }
    fill_u64_kernel<<<1, 32, 0, ls>>>(ctx->d_mtOff.as<unsigned long long>() + nMtTotal, 2, before); }
  CU(cudaStreamSynchronize(ls));                                       // the host vectors above go out of scope; nothing large has been queued yet
  CU(cudaEventRecord(ctx->copyEvs[ctx->groups.size()], cs));
  ctx->loadCopied = !resident;
  if ((packed || vbyte) && src->n_exceptions) {
    CU(cudaMemcpyAsync(ctx->d_pkExc.p, src->exceptions, src->n_exceptions * sizeof(PkException), cudaMemcpyHostToDevice, cs));
    load_check_exc_kernel<<<(unsigned)((src->n_exceptions + 255) / 256), 256, 0, cs>>>(ctx->d_pkExc.as<PkException>(), src->n_exceptions, src->off[(size_t)n_samples * 2 * (C + 1) - 1], ctx->d_loadFlags.as<uint32_t>());
  }

  LoadArgs la;
  memset(&la, 0, sizeof la);
  la.words = packed ? ctx->d_pk16.as<uint16_t>() : nullptr;
  la.exc = ctx->d_pkExc.as<PkException>(); la.nExc = (packed || vbyte) ? src->n_exceptions : 0;
  if (vbyte) {
    la.vbytes = ctx->d_vbytes.as<uint8_t>(); la.vSrcByte = ctx->d_vSrcByte.as<uint64_t>(); la.vDevByte = ctx->d_vDevByte.as<uint64_t>();
    la.vSrcTile = ctx->d_vSrcTile.as<uint64_t>(); la.vTileOff = ctx->d_vTileOff.as<uint64_t>(); la.vTiles = src->n_vtiles;
  }
  la.pos = ctx->d_pos.as<uint32_t>(); la.cnt = ctx->d_cnt.as<uint32_t>();
  la.segStart = ctx->d_segStart.as<uint64_t>(); la.srcStart = ctx->d_srcStart.as<uint64_t>();
  la.sizes = ctx->d_sizes.as<uint32_t>(); la.mtBase = ctx->d_mtBase.as<uint32_t>(); la.mtCount = ctx->d_mtCount.as<uint32_t>();
  la.winBase = ctx->d_winBase.as<uint32_t>(); la.wEnt = ctx->d_wEnt.as<uint32_t>(); la.nWinTotal = nWinTotal; la.C = C;
  la.flags = ctx->d_loadFlags.as<uint32_t>();
  la.tileExc = (packed || vbyte) ? ctx->d_tileExc.as<unsigned long long>() : nullptr;
  la.tileDesc = (packed || vbyte) ? ctx->d_tileDesc.as<TileDesc>() : nullptr;
  MtArgs ma;
  memset(&ma, 0, sizeof ma);
  ma.pos = la.pos; ma.cnt = la.cnt; ma.segStart = la.segStart; ma.mtBase = la.mtBase;
  ma.mtCount = la.mtCount; ma.mtOff = ctx->d_mtOff.as<unsigned long long>(); ma.ment = ctx->d_ment.as<uint2>();
  ma.sizes = la.sizes; ma.flags = la.flags; ma.C = C; ma.nSeg = nSeg;
  ma.mtW = ctx->d_mtW.as<unsigned long long>(); ma.blkW = ctx->d_blkW.as<float>();
  ma.winBase = la.winBase; ma.wEnt = la.wEnt; ma.nWinTotal = nWinTotal;
  unsigned long long* d_scanTot = (unsigned long long*)(ctx->d_loadFlags.as<unsigned char>() + 64);

  for (size_t gi = 0; gi < ctx->groups.size(); ++gi) {
    const LoadGroup& g = ctx->groups[gi];
    if (!resident && g.entries) {
      // one copy per (sample, array): the segments of contigs c0 .. c1-1 of one (sample, strand) are contiguous on both sides
      for (uint32_t s = 0; s < n_samples; ++s) {
        const size_t sg0 = ((size_t)s * 2 + g.d) * C + g.c0, sg1 = ((size_t)s * 2 + g.d) * C + g.c1;
        const uint64_t dev0 = ctx->segStart[sg0], cntE = ctx->segStart[sg1] - dev0, src0 = ctx->srcStart[sg0];
        if (!cntE) continue;
        cudaStream_t q = cs;                                             // one queue: alternating the copies of a group over two queues was slower (r02: 18.5 vs 16.7 ms per load + call)
        if (vbyte) CU(cudaMemcpyAsync(ctx->d_vbytes.as<uint8_t>() + vDevByte[sg0], src->vbytes + vSrcByte[sg0], vSrcByte[sg1] - vSrcByte[sg0], cudaMemcpyHostToDevice, q));
        else if (packed) CU(cudaMemcpyAsync(ctx->d_pk16.as<uint16_t>() + dev0, src->words + src0, cntE * 2, cudaMemcpyHostToDevice, q));
        else {
          CU(cudaMemcpyAsync(ctx->d_pos.as<uint32_t>() + dev0, src->pos + src0, cntE * 4, cudaMemcpyHostToDevice, q));
          CU(cudaMemcpyAsync(ctx->d_cnt.as<uint32_t>() + dev0, src->cnt + src0, cntE * 4, cudaMemcpyHostToDevice, q));
        }
      }
      CU(cudaEventRecord(ctx->copyEvs[gi], cs));
      CU(cudaStreamWaitEvent(ls, ctx->copyEvs[gi], 0));
    }
    la.grpSeg = ctx->d_grpSeg.as<uint32_t>() + g.seg0; la.grpTile = ctx->d_grpTile.as<uint32_t>() + g.tileTab0; la.nGrpSeg = g.nSeg;
    if (g.w1 > g.w0) load_win_fill_kernel<<<(unsigned)(((uint64_t)n_samples * (g.w1 - g.w0) + 255) / 256), 256, 0, ls>>>(la, n_samples, g.w0, g.w1);
    if (g.nTiles) {
      if (packed || vbyte) load_tile_prep_kernel<<<g.nTiles / 256 + 1, 256, 0, ls>>>(la, g.nTiles);
      if (vbyte) load_tile_kernel<LOAD_VBYTE><<<g.nTiles, PK_THREADS, 0, ls>>>(la);
      else if (packed) load_tile_kernel<LOAD_WORDS><<<g.nTiles, PK_THREADS, 0, ls>>>(la);
      else load_tile_kernel<LOAD_PLAIN><<<g.nTiles, PK_THREADS, 0, ls>>>(la);
      load_seam_kernel<<<(g.nTiles + 255) / 256, 256, 0, ls>>>(la, g.nTiles);
    }
    if (g.m1 > g.m0) {
      int r = scan_load(ctx, ma.mtCount + g.m0, g.m1 - g.m0, ctx->d_mtOff.as<unsigned long long>() + g.m0, g.mentBase, d_scanTot);
      if (r) return r;
      set_u64_kernel<<<1, 1, 0, ls>>>(ctx->d_mtOff.as<unsigned long long>() + g.m1, g.mentBase + g.entries);
      mt_scatter_kernel<<<g.w1 - g.w0, n_samples > 8 ? 512 : 256, 0, ls>>>(ma, n_samples, g.w0);   // one warp per sample at a time
    }
    CU(cudaGetLastError());
    CU(cudaEventRecord(ctx->groupEv[gi], ls));
  }
  // first / last position of every segment (the reference-order logic and the left-edge test want them) and the validation flags
  uint32_t* ends = ctx->d_segEnds.as<uint32_t>();
  seg_ends_kernel<<<(nSeg + 255) / 256, 256, 0, ls>>>(ctx->d_pos.as<uint32_t>(), ctx->d_segStart.as<uint64_t>(), nSeg, ends, ends + nSeg);
  CU(cudaGetLastError());
  uint32_t* hEnds = (uint32_t*)ctx->h_ends.p;
  CU(cudaMemcpyAsync(hEnds, ends, 2 * (size_t)nSeg * 4, cudaMemcpyDeviceToHost, ls));
  CU(cudaMemcpyAsync(hEnds + 2 * (size_t)nSeg, ctx->d_loadFlags.p, 4, cudaMemcpyDeviceToHost, ls));
  CU(cudaEventRecord(ctx->loadDoneEv, ls));
  ctx->loaded = false;
  ctx->loadPending = true;
  ctx->cntEffValid = false;
  ctx->tidxBuilt = false;
  return UPB_OK;
}

// wait for a pending load and take over what it reports; errors of the input surface here
static int load_finish (upb_ctx* ctx) {
  if (!ctx->loadPending) return UPB_OK;
  CU(cudaSetDevice(ctx->device));
  ctx->loadPending = false;
  CU(cudaEventSynchronize(ctx->loadDoneEv));
  if (getenv("UPB_TRACE") && *getenv("UPB_TRACE")) {                                            // timeline of the pipelined load: when each group's bytes were on the device, when its index was built
    const size_t G = ctx->groups.size();
    fprintf(stderr, "[upb] load: %zu groups, %llu entries\n", G, (unsigned long long)ctx->nEntries);
    for (size_t k = 0; k < G; ++k) {
      float a = -1, b = -1;
      if (ctx->loadCopied && ctx->groups[k].entries) cudaEventElapsedTime(&a, ctx->copyEvs[G], ctx->copyEvs[k]);
      cudaEventElapsedTime(&b, ctx->copyEvs[G], ctx->groupEv[k]);
      fprintf(stderr, "[upb]   group %2zu strand %u contigs %2u..%2u  %9llu entries  copied at %7.3f ms  indexed at %7.3f ms\n", k, ctx->groups[k].d, ctx->groups[k].c0,
              ctx->groups[k].c1 - 1, (unsigned long long)ctx->groups[k].entries, a, b);
    }
  }
  const uint32_t nSeg = (uint32_t)ctx->firstPos.size();
  const uint32_t* hEnds = (const uint32_t*)ctx->h_ends.p;
  const uint32_t flags = hEnds[2 * (size_t)nSeg];
  if (flags & FLAG_BAD_PACK) return fail(ctx, UPB_ERR_BAD_ARG, "packed load: malformed input (2-byte form: an escape code without its exception entry, or a tile of a segment that does not start with a position escape; byte form: a malformed entry, a tile that does not hold exactly its entries, or a count escape without its exception entry)");
  if (flags & FLAG_UNSORTED) return fail(ctx, UPB_ERR_UNSORTED, "upb_load_tags: positions must be strictly ascending within each (sample, strand, contig) — the reference asserts on unsorted input (misc/peakcall.cpp:171)");
  if (flags & FLAG_BAD_POS) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: position outside 1..contig size (the reference's parser drops those, misc/format.cpp:670-676)");
  if (flags & FLAG_ZERO_COUNT) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: zero count (the reference's parser never delivers those, misc/format.cpp:710)");
  for (uint32_t g = 0; g < nSeg; ++g) { ctx->firstPos[g] = hEnds[g]; ctx->lastPos[g] = hEnds[nSeg + g]; }
  ctx->lastPosKnown = true;
  ctx->loaded = true;
  ctx->stats.n_entries = ctx->nEntries;
  return UPB_OK;
}

extern "C" int upb_load_tags_async (upb_ctx* ctx, const upb_tag_source* src) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (ctx->sizes.empty()) return fail(ctx, UPB_ERR_NOT_LOADED, "upb_load_tags: call upb_set_contigs first");
  if (!src || src->n_samples == 0 || !src->off) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: no samples");
  if (src->n_samples > (1u << 20) || (uint64_t)src->n_samples * 2 * ctx->sizes.size() > 0x7FFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many (sample, strand, contig) segments");
  CU(cudaSetDevice(ctx->device));
  if (ctx->loadPending) { ctx->loadPending = false; cudaEventSynchronize(ctx->loadDoneEv); }   // a load nobody waited for is simply replaced
  ctx->loaded = false;
  const int r = load_enqueue(ctx, src, false);
  if (r) { ctx->loadPending = false; cudaStreamSynchronize(ctx->copyStream); cudaStreamSynchronize(ctx->copyStream2); cudaStreamSynchronize(ctx->loadStream); }
  return r;
}

// when the last byte of the last load was on the device and when its index was complete, in ms after the first copy was queued
// (CUDA events of the load; valid once the load has been waited for).  Instrumentation only (bench.py: per-rank load phases).
extern "C" int upb_last_load_ms (const upb_ctx* ctx, float* copied_ms, float* indexed_ms) {
  if (!ctx || !copied_ms || !indexed_ms) return UPB_ERR_BAD_ARG;
  *copied_ms = *indexed_ms = -1.f;
  const size_t G = ctx->groups.size();
  if (ctx->loadPending || !G || ctx->copyEvs.size() <= G || ctx->groupEv.size() < G) return UPB_ERR_NOT_LOADED;
  if (ctx->loadCopied) for (size_t k = G; k-- > 0;) if (ctx->groups[k].entries) { cudaEventElapsedTime(copied_ms, ctx->copyEvs[G], ctx->copyEvs[k]); break; }
  cudaEventElapsedTime(indexed_ms, ctx->copyEvs[G], ctx->groupEv[G - 1]);
  return UPB_OK;
}

extern "C" int upb_load_wait (upb_ctx* ctx) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!ctx->loadPending) return ctx->loaded ? UPB_OK : fail(ctx, UPB_ERR_NOT_LOADED, "no tags loaded (upb_set_contigs + upb_load_tags first)");
  return load_finish(ctx);
}

extern "C" int upb_reindex (upb_ctx* ctx) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  { int r = load_finish(ctx); if (r) return r; }
  if (!ctx->loaded) return fail(ctx, UPB_ERR_NOT_LOADED, "no tags loaded (upb_set_contigs + upb_load_tags first)");
  CU(cudaSetDevice(ctx->device));
  int r = load_enqueue(ctx, nullptr, true);
  if (r) return r;
  return load_finish(ctx);
}

extern "C" int upb_load_tags (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt) {
  upb_tag_source s;
  memset(&s, 0, sizeof s);
  s.n_samples = n_samples; s.off = off; s.pos = pos; s.cnt = cnt; s.directional_hint = 1;
  const int r = upb_load_tags_async(ctx, &s);
  return r ? r : load_finish(ctx);
}

extern "C" int upb_load_tags_packed2 (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint16_t* words,
                                      uint64_t n_exceptions, const upb_packed_exception* exceptions) {
  if (ctx && !words) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_packed2: NULL argument");
  upb_tag_source s;
  memset(&s, 0, sizeof s);
  s.n_samples = n_samples; s.off = off; s.words = words; s.n_exceptions = n_exceptions; s.exceptions = exceptions; s.directional_hint = 1;
  const int r = upb_load_tags_async(ctx, &s);
  return r ? r : load_finish(ctx);
}

// host-only encoder of the 2-byte form (what a front end does while it parses): position delta in the low 12 bits, count in the
// high 4, escapes 0xFFF / 0xF with the full values in the exception list.  A position escape opens every segment and every
// PK_TILE-th entry of it (upb_pack.cuh).
extern "C" int upb_load_tags_vbyte (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint8_t* bytes, const uint64_t* byte_off,
                                    uint64_t n_tiles, const uint64_t* tile_off, uint64_t n_exceptions, const upb_packed_exception* exceptions) {
  if (ctx && (!bytes || !byte_off || !tile_off)) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags_vbyte: NULL argument");
  upb_tag_source s;
  memset(&s, 0, sizeof s);
  s.n_samples = n_samples; s.off = off; s.vbytes = bytes; s.vbyte_off = byte_off; s.vtile_off = tile_off; s.n_vtiles = n_tiles;
  s.n_exceptions = n_exceptions; s.exceptions = exceptions; s.directional_hint = 1;
  const int r = upb_load_tags_async(ctx, &s);
  return r ? r : load_finish(ctx);
}

extern "C" int upb_pack_tags2 (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                               uint16_t* words_out, upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out) {
  if (!off || !n_exc_out || (exc_cap && !exc_out)) return UPB_ERR_BAD_ARG;
  // one task per (sample, strand): first the exception counts (so that every task knows where its exceptions go), then the words
  const uint32_t nTasks = n_samples * 2;
  std::vector<uint64_t> nExcOf(nTasks, 0);
  auto run = [&] (uint32_t sd, bool write, uint64_t excBase) {
    uint64_t nExc = 0;
    for (uint32_t c = 0; c < n_contigs; ++c) {
      const uint64_t a = off[(size_t)sd * (n_contigs + 1) + c], b = off[(size_t)sd * (n_contigs + 1) + c + 1];
      uint32_t prev = 0;
      for (uint64_t e = a; e < b; ++e) {
        const uint32_t p = pos[e], k = cnt[e];
        const bool escP = ((e - a) & (PK_TILE - 1)) == 0 || p < prev || p - prev >= PK2_DPOS_ESC;   // unsorted input is kept verbatim: the load then reports it
        const bool escC = k >= PK2_CNT_ESC;
        if (escP || escC) {
          if (write && excBase + nExc < exc_cap) { upb_packed_exception& x = exc_out[excBase + nExc]; x.index = e; x.pos = escP ? p : 0u; x.cnt = escC ? k : 0u; }
          ++nExc;
        }
        if (write && words_out) words_out[e] = (uint16_t)((escP ? PK2_DPOS_ESC : p - prev) | ((escC ? PK2_CNT_ESC : k) << PK2_CNT_SHIFT));
        prev = p;
      }
    }
    if (!write) nExcOf[sd] = nExc;
  };
  const unsigned nThreads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), nTasks));
  auto parallel = [&] (bool write, const std::vector<uint64_t>& base) {
    std::vector<std::thread> th;
    std::atomic<uint32_t> next(0);
    for (unsigned t = 0; t < nThreads; ++t) th.emplace_back([&] { for (uint32_t sd; (sd = next.fetch_add(1)) < nTasks;) run(sd, write, base.empty() ? 0 : base[sd]); });
    for (std::thread& t : th) t.join();
  };
  parallel(false, std::vector<uint64_t>());
  std::vector<uint64_t> base(nTasks, 0);
  uint64_t total = 0;
  for (uint32_t sd = 0; sd < nTasks; ++sd) { base[sd] = total; total += nExcOf[sd]; }
  *n_exc_out = total;
  if (total > exc_cap || (!words_out && !exc_cap)) return total <= exc_cap ? UPB_OK : UPB_ERR_NOMEM;
  parallel(true, base);
  return UPB_OK;
}

// host-only encoder of the variable-length byte form (upb_vbyte.hpp): one task per (sample, strand).  A call without buffers only adds up
// the sizes; a call with buffers encodes every task ONCE into a buffer of its own and then moves the tasks' bytes, tile offsets and exceptions
// to their places (a second pass of plain copies: the sizes that place them are only known when every task is through)
extern "C" int upb_pack_tags_vbyte (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                                    uint8_t* bytes_out, uint64_t bytes_cap, uint64_t* n_bytes_out, uint64_t* byte_off_out,
                                    uint64_t* tile_off_out, uint64_t tile_cap, uint64_t* n_tiles_out,
                                    upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out) {
  using namespace upb_vb;
  if (!off || !n_bytes_out || !n_tiles_out || !n_exc_out) return UPB_ERR_BAD_ARG;
  const uint32_t nTasks = n_samples * 2, C1 = n_contigs + 1;
  const bool encode = bytes_out && byte_off_out && tile_off_out && bytes_cap > 0;
  // (bytes: worst-case sized but never initialised — only the pages the code really fills are ever touched)
  struct Task { std::unique_ptr<uint8_t[]> bytes; std::vector<uint64_t> tileRel, segBytes; std::vector<upb_packed_exception> exc; uint64_t nBytes = 0, nTiles = 0, nExc = 0; };
  std::vector<Task> tasks(nTasks);
  auto run = [&] (uint32_t sd) {
    Task& T = tasks[sd];
    const uint64_t a0 = off[(size_t)sd * C1], b0 = off[(size_t)sd * C1 + n_contigs];
    T.segBytes.assign(n_contigs, 0);
    if (encode) { T.bytes.reset(new uint8_t[(size_t)(b0 - a0) * VB_MAX_ENTRY + 8]); T.tileRel.reserve((size_t)((b0 - a0) / VB_TILE) + n_contigs + 1); }
    uint64_t at = 0;
    for (uint32_t c = 0; c < n_contigs; ++c) {
      const uint64_t a = off[(size_t)sd * C1 + c], b = off[(size_t)sd * C1 + c + 1];
      const uint64_t at0 = at;
      uint32_t prev = 0;
      for (uint64_t e = a; e < b; ++e) {
        const uint32_t p = pos[e], k = cnt[e];
        const bool head = ((e - a) & (VB_TILE - 1)) == 0;
        const uint32_t v = head ? p : (p > prev ? p - prev : 0u);       // unsorted input decodes to a repeated position: the load then reports it
        bool esc = false;
        const uint32_t nb = vb_size(v, k, &esc);
        if (encode) {
          if (head) T.tileRel.push_back(at);
          vb_put(T.bytes.get() + at, v, k);
          if (esc) { upb_packed_exception x; x.index = e; x.pos = 0u; x.cnt = k; T.exc.push_back(x); }
        }
        if (head) ++T.nTiles;
        if (esc) ++T.nExc;
        at += nb;
        prev = p;
      }
      T.segBytes[c] = at - at0;
    }
    T.nBytes = at;
  };
  const unsigned nThreads = std::max(1u, std::min<unsigned>(std::thread::hardware_concurrency(), nTasks));
  auto parallel = [&] (const std::function<void(uint32_t)>& f) {
    std::vector<std::thread> th;
    std::atomic<uint32_t> next(0);
    for (unsigned t = 0; t < nThreads; ++t) th.emplace_back([&] { for (uint32_t sd; (sd = next.fetch_add(1)) < nTasks;) f(sd); });
    for (std::thread& t : th) t.join();
  };
  parallel(run);
  std::vector<uint64_t> byteBase(nTasks, 0), tileBase(nTasks, 0), excBase(nTasks, 0);
  uint64_t nBytes = 0, nTiles = 0, nExc = 0;
  for (uint32_t sd = 0; sd < nTasks; ++sd) { byteBase[sd] = nBytes; tileBase[sd] = nTiles; excBase[sd] = nExc; nBytes += tasks[sd].nBytes; nTiles += tasks[sd].nTiles; nExc += tasks[sd].nExc; }
  *n_bytes_out = nBytes; *n_tiles_out = nTiles; *n_exc_out = nExc;
  if (nBytes > bytes_cap || nTiles + 1 > tile_cap || nExc > exc_cap) return UPB_ERR_NOMEM;
  if (!encode || (nExc && !exc_out)) return UPB_ERR_BAD_ARG;
  parallel([&] (uint32_t sd) {
    const Task& T = tasks[sd];
    if (T.nBytes) memcpy(bytes_out + byteBase[sd], T.bytes.get(), (size_t)T.nBytes);
    for (size_t i = 0; i < T.tileRel.size(); ++i) tile_off_out[tileBase[sd] + i] = byteBase[sd] + T.tileRel[i];
    for (size_t i = 0; i < T.exc.size(); ++i) exc_out[excBase[sd] + i] = T.exc[i];
    uint64_t at = byteBase[sd];
    for (uint32_t c = 0; c < n_contigs; ++c) { byte_off_out[(size_t)sd * C1 + c] = at; at += T.segBytes[c]; }
    byte_off_out[(size_t)sd * C1 + n_contigs] = at;
  });
  tile_off_out[nTiles] = nBytes;
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// device-wide exclusive scan (uint32 in, uint64 out); total -> *d_total
// ------------------------------------------------------------------------------------------------
#define SCR(name, T, count) ({ DevBuf& b__ = ctx->scratch[ctx->lane ? std::string(name) + "~lane1" : std::string(name)]; int r__ = ensure(ctx, b__, (size_t)(count) * sizeof(T) + 16); if (r__) return r__; (T*)b__.p; })

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
// small per-call tables (taps, units, sample lists ...) reach the device through a KERNEL that reads them from pinned, mapped host
// memory — not through cudaMemcpyAsync: an H2D copy, however small, queues on the copy engine behind every byte of a load that is
// still in flight (upb_load_tags_async), and the call that was meant to work through the load's groups as they arrive would start
// only when the last of them is on the device (measured: load + call 22.0 ms against 15.7 + 2.4 apart).
// ------------------------------------------------------------------------------------------------
struct ParamCopyArgs { const unsigned char* src; unsigned char* dst[8]; uint32_t off[8]; uint32_t bytes[8]; uint32_t n; };

__global__ void param_copy_kernel (ParamCopyArgs a) {
  for (uint32_t j = 0; j < a.n; ++j) {
    const unsigned char* s = a.src + a.off[j];
    unsigned char* d = a.dst[j];
    const uint32_t words = a.bytes[j] >> 2;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) ((uint32_t*)d)[i] = ((const uint32_t*)s)[i];
    if (blockIdx.x == 0) for (uint32_t i = (words << 2) + threadIdx.x; i < a.bytes[j]; i += blockDim.x) d[i] = s[i];
  }
}

struct ParamStage { ParamCopyArgs a; size_t used; };
static void stage_begin (upb_ctx* ctx, ParamStage& ps) { memset(&ps, 0, sizeof ps); ps.a.src = (const unsigned char*)ctx->h_param.p; }
static int stage_add (upb_ctx* ctx, ParamStage& ps, void* d_dst, const void* h_src, size_t bytes);
static int stage_flush (upb_ctx* ctx, ParamStage& ps) {
  if (!ps.a.n) return UPB_OK;
  param_copy_kernel<<<4, 256, 0, ctx->stream>>>(ps.a);
  CU(cudaGetLastError());
  ps.a.n = 0;
  return UPB_OK;
}
static int stage_add (upb_ctx* ctx, ParamStage& ps, void* d_dst, const void* h_src, size_t bytes) {
  if (!bytes) return UPB_OK;
  if (ps.a.n == 8) { const int r = stage_flush(ctx, ps); if (r) return r; }
  const size_t at = (ps.used + 15) & ~(size_t)15;
  if (at + bytes > ctx->h_param.cap) return fail(ctx, UPB_ERR_NOMEM, "parameter staging block too small");
  memcpy((unsigned char*)ctx->h_param.p + at, h_src, bytes);
  ps.a.dst[ps.a.n] = (unsigned char*)d_dst; ps.a.off[ps.a.n] = (uint32_t)at; ps.a.bytes[ps.a.n] = (uint32_t)bytes; ++ps.a.n;
  ps.used = at + bytes;
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
  std::vector<uint64_t> unitRows;    // [nUnits] accepted rows
  std::vector<uint64_t> tir;         // [nSamples] tags in regions
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

struct Prep {
  std::vector<Unit> units;
  std::vector<double> taps;
  std::vector<uint32_t> affected;    // units with a tag within `bandwidth` of the contig start, or inheriting stale window cells: replayed by left_edge_kernel
  std::vector<uint32_t> chainNext;   // [affected] the affected unit that inherits this one's window (every tag of this one lies within `bandwidth`), else 0xFFFFFFFF
  std::vector<uint8_t> chainHead;    // [affected] 0: replayed by its predecessor's thread
  double maxCoeff = 0;               // -z: largest coefficient (bounds the weight of a position)
  const float* d_hotCoefF = nullptr; // hot_span_kernel's envelope coefficients
  uint32_t hotJ = 0;
  const double* d_coeffBySample = nullptr;
  uint32_t totalSpans = 0;
  uint32_t nSrc = 0;
  bool haveControl = false;
  Tags T;
  DensArgs dens;                     // template for the density kernel
};

// common set-up: units, taps, source list, control filter
static int prepare (upb_ctx* ctx, const PipeCfg& cfg, Prep& pr, uint32_t* launches, bool allowPending = false) {
  const uint32_t C = (uint32_t)ctx->sizes.size();
  const uint32_t n = ctx->nSamples;
  if (ctx->loadPending && !allowPending) { const int rl = load_finish(ctx); if (rl) return rl; }
  if (!ctx->loaded && !(allowPending && ctx->loadPending)) return fail(ctx, UPB_ERR_NOT_LOADED, "no tags loaded (upb_set_contigs + upb_load_tags first)");
  if (cfg.bw == 0 || cfg.bw > (uint32_t)MAX_BW_FAST) return fail(ctx, UPB_ERR_UNSUPPORTED, "bandwidth must be in 1..448");
  if (!(cfg.thr > 0)) return fail(ctx, UPB_ERR_BAD_ARG, "region threshold must be > 0 (the reference's leap handling relies on it, SURVEY Q10)");
  if (!(cfg.background > 0)) return fail(ctx, UPB_ERR_BAD_ARG, "background must be > 0");
  int r;

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
  pr.chainNext.clear(); pr.chainHead.clear();
  for (uint32_t b = 0; b < nBuf; ++b) {
    int32_t leakFrom = -1;                                            // affected index of the unit whose flush leaves stale cells in this buffer's window
    for (uint32_t c = 0; c < C; ++c) {
      bool any = false;
      for (uint32_t s : src) {
        if (cfg.nondir) any = any || ctx->firstPos[(size_t)(s * 2 + 0) * C + c] || ctx->firstPos[(size_t)(s * 2 + 1) * C + c];
        else any = any || ctx->firstPos[(size_t)(s * 2 + b) * C + c];
      }
      // Q2: every add() moves the window, control samples included (src/regions.cpp:356-359)
      uint32_t first = 0xFFFFFFFFu, last = 0;
      for (uint32_t s = 0; s < n; ++s) for (uint32_t d = (cfg.nondir ? 0u : b); d <= (cfg.nondir ? 1u : b); ++d) {
        const uint32_t fp = ctx->firstPos[(size_t)(s * 2 + d) * C + c], lp = ctx->lastPos[(size_t)(s * 2 + d) * C + c];
        if (fp) { first = std::min(first, fp); last = std::max(last, lp); }
      }
      // a contig that only control samples touch makes no unit — unless the buffer arrives with stale cells, which its add() calls then report
      if (!any && !(leakFrom >= 0 && first != 0xFFFFFFFFu)) continue;
      Unit u;
      u.contig = c; u.buffer = b; u.size = ctx->sizes[c];
      u.nSpans = (uint32_t)(((uint64_t)ctx->sizes[c] + cfg.bw + SPAN_LEN - 1) / SPAN_LEN);
      u.spanStart = (uint32_t)spanStart;
      spanStart += u.nSpans;
      if (spanStart > 0xFFFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many spans");
      if (first <= cfg.bw || leakFrom >= 0) {
        const int32_t k = (int32_t)pr.affected.size();
        pr.affected.push_back((uint32_t)pr.units.size());
        pr.chainNext.push_back(0xFFFFFFFFu);
        pr.chainHead.push_back(leakFrom < 0 ? 1 : 0);
        if (leakFrom >= 0) pr.chainNext[leakFrom] = (uint32_t)k;
        // every tag within `bandwidth` of the start: the flush cannot drain the window (misc/peakcall.cpp:176-184,224-231) and the next contig inherits it
        leakFrom = last <= cfg.bw ? k : -1;
      }
      pr.units.push_back(u);
    }
  }
  pr.totalSpans = (uint32_t)spanStart;

  pr.taps.assign(2 * (size_t)cfg.bw + 1, 0.0);
  upb_kernel_taps(cfg.bw, 1 / cfg.background, pr.taps.data());      // src/regions.cpp:232
  double* d_taps = SCR("taps", double, pr.taps.size());
  uint32_t* d_src = SCR("src", uint32_t, src.size() + 1);
  Unit* d_units = SCR("units", Unit, pr.units.size() + 1);
  uint8_t* d_control = SCR("control", uint8_t, n + 1);
  ENSURE(ctx->d_small, 4096);
  // cold-block bound of hot_span_kernel: a tag j blocks (of 32 positions) away from block B is at least 32(|j|-1)+1 positions from every
  // position of B (0 for j = 0); blocks further than J = (bw-2)/32+1 away cannot reach it (the taps at +-bw are exactly 0, misc/kernel.cpp:12-14)
  const int bwI = (int)cfg.bw, J = bwI >= 2 ? (bwI - 2) / 32 + 1 : 0;
  std::vector<float> hf(J + 1);
  for (int j = 0; j <= J; ++j) {
    const int m = j == 0 ? 0 : 32 * (j - 1) + 1;
    const double hc = pr.taps[bwI - m] * (1.0 + pr.maxCoeff) * (1.0 + 1e-9);   // -z: a count c weighs at most c*(coeff+1)
    hf[j] = std::nextafter((float)(hc * (1.0 + 1e-5)), INFINITY);              // float sum of <= 58 terms: 1e-5 covers its rounding
  }
  float* d_hf = SCR("hotCoefF", float, J + 1);
  pr.d_hotCoefF = d_hf; pr.hotJ = (uint32_t)J;
  double* d_coefSrc = nullptr;
  double* d_coefSample = nullptr;
  if (cfg.nCoeffs) { d_coefSample = SCR("coefSample", double, n); d_coefSrc = SCR("coefSrc", double, src.size() + 1); pr.d_coeffBySample = d_coefSample; }
  { const int rh = ensure_host(ctx, ctx->h_param, pr.taps.size() * 8 + src.size() * 12 + pr.units.size() * sizeof(Unit) + (size_t)n * 9 + (size_t)(J + 1) * 4 + 1024); if (rh) return rh; }
  ParamStage ps;
  stage_begin(ctx, ps);
  int rs = 0;
  if ((rs = stage_add(ctx, ps, d_taps, pr.taps.data(), pr.taps.size() * 8))) return rs;
  if ((rs = stage_add(ctx, ps, d_src, src.data(), src.size() * 4))) return rs;
  if ((rs = stage_add(ctx, ps, d_units, pr.units.data(), pr.units.size() * sizeof(Unit)))) return rs;
  if ((rs = stage_add(ctx, ps, d_control, control.data(), n))) return rs;
  if ((rs = stage_add(ctx, ps, d_hf, hf.data(), (size_t)(J + 1) * 4))) return rs;
  if (cfg.nCoeffs) {
    if ((rs = stage_add(ctx, ps, d_coefSample, coeffBySample.data(), (size_t)n * 8))) return rs;
    if ((rs = stage_add(ctx, ps, d_coefSrc, coeffBySrc.data(), src.size() * 8))) return rs;
  }
  if ((rs = stage_flush(ctx, ps))) return rs;
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
struct VecPatch { uint32_t k, src, len; };     // range k of the dump takes its scores from the replay's vector pool at `src`
__global__ void vec_patch_kernel (const VecPatch* __restrict__ patches, const unsigned long long* __restrict__ dOff, const double* __restrict__ srcF, const double* __restrict__ srcR,
                                  double* __restrict__ F, double* __restrict__ R) {
  const VecPatch q = patches[blockIdx.x];
  const unsigned long long at = dOff[q.k];
  for (uint32_t i = threadIdx.x; i < q.len; i += blockDim.x) { F[at + i] = srcF[q.src + i]; R[at + i] = srcR[q.src + i]; }
}

static int dump_and_corr (upb_ctx* ctx, const DensArgs& base, bool nondir, uint32_t n, uint32_t nShifts, uint32_t* launches,
                          const std::vector<VecPatch>* patches = nullptr, const double* patchF = nullptr, const double* patchR = nullptr) {
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
  if (patches && !patches->empty()) {                                 // regions of the left-edge replay: the reference's displaced scores, not the density
    VecPatch* d_p = SCR("vecPatches", VecPatch, patches->size());
    CU(cudaMemcpyAsync(d_p, patches->data(), patches->size() * sizeof(VecPatch), cudaMemcpyHostToDevice, ctx->stream));
    vec_patch_kernel<<<(unsigned)patches->size(), 256, 0, ctx->stream>>>(d_p, dOff, patchF, patchR, F, R);
    CU(cudaGetLastError());
    CU(cudaStreamSynchronize(ctx->stream));                           // `patches` is the caller's
    *launches += 1;
  }
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
  // eight warps per CTA where their private windows fit one SM's shared memory; fewer where they do not (-z with -D at bandwidth
  // >= ~290: two u32 and two f64 windows of 512 + 2 bandwidth cells per warp)
  const size_t smemMax = 227 * 1024;
  if (L.warp0 + L.warpStride > smemMax) return fail(ctx, UPB_ERR_UNSUPPORTED, "bandwidth / sample count need more shared memory than one SM has");
  const uint32_t nWarps = (uint32_t)std::min<size_t>(SP_WARPS, (smemMax - L.warp0) / L.warpStride);
  const int threads = (int)nWarps * 32;
  const size_t smem = (size_t)L.warp0 + (size_t)L.warpStride * nWarps;
  void (*kern)(const SpanArgs) = coeff ? (nondir ? span_kernel<true, true> : span_kernel<false, true>) : (nondir ? span_kernel<true, false> : span_kernel<false, false>);
  int perSm = 0;
  { int rr = cached_occupancy(ctx, kern, threads, smem, &perSm); if (rr) return rr; }
  if (!coeff && perSm < SP_MIN_BLOCKS) {                              // shared memory admits fewer CTAs than the 64-register build is sized for: take the build without that cap
    kern = nondir ? span_kernel<true, false, SP_MIN_BLOCKS - 1> : span_kernel<false, false, SP_MIN_BLOCKS - 1>;
    int rr = cached_occupancy(ctx, kern, threads, smem, &perSm); if (rr) return rr;
  }
  // B1 through the bulk-copy engine (cp.async.bulk + mbarrier, SASS UBLKCP) where it was measured to win: directional calls at bandwidth >= 100, whose
  // chunks (piece + 2 x bandwidth of halo, all samples) are long enough to amortise arming a barrier (profiles/r02_b1_bulk_ab.txt: 0.539 -> 0.497 ms at
  // bandwidth 200, level at 50, behind at 10 / 20 and on the two-sample -D workload).  UPB_B1=bulk|ldg overrides (A/B runs)
  bool bulk = !coeff && !nondir && sa.bw >= 100;
  if (const char* e = getenv("UPB_B1")) { if (!strcmp(e, "bulk")) bulk = !coeff && !nondir; else if (!strcmp(e, "ldg")) bulk = false; }
  if (bulk) kern = perSm < SP_MIN_BLOCKS ? span_kernel<false, false, SP_MIN_BLOCKS - 1, true> : span_kernel<false, false, SP_MIN_BLOCKS, true>;
  if (bulk) { int rr = cached_occupancy(ctx, kern, threads, smem, &perSm); if (rr) return rr; }
  const uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)ctx->smCount * perSm, std::max<uint32_t>(sa.totalSpans, 1u));
  kern<<<grid, threads, smem, ctx->cur>>>(sa);
  CU(cudaGetLastError());
  *launches += 1;
  return UPB_OK;
}

// ---- per-stage device timing (upb_set_profiling): one event between the stages of the pipeline, read after the call ----------
static void stage_reset (upb_ctx* ctx) { ctx->stageUsed = 0; ctx->stageNames.clear(); ctx->hostMarks.clear(); }
static void stage_mark (upb_ctx* ctx, const char* name) {
  ctx->hostMarks.push_back(std::make_pair(name, std::chrono::steady_clock::now()));   // host side of the same points (UPB_TRACE_SLOW)
  if (!ctx->profiling) return;
  if (ctx->stageUsed == ctx->stageEv.size()) { cudaEvent_t e = nullptr; if (cudaEventCreate(&e) != cudaSuccess) return; ctx->stageEv.push_back(e); }
  cudaEventRecord(ctx->stageEv[ctx->stageUsed++], ctx->cur);
  ctx->stageNames.push_back(name);
}
static void stage_collect (upb_ctx* ctx) {                            // after the stream has been synchronised
  ctx->stageMs.clear();
  for (size_t k = 0; k + 1 < ctx->stageUsed; ++k) { float ms = 0; cudaEventElapsedTime(&ms, ctx->stageEv[k], ctx->stageEv[k + 1]); ctx->stageMs.push_back(ms); }
}

extern "C" int upb_set_profiling (upb_ctx* ctx, int on) { if (!ctx) return UPB_ERR_BAD_ARG; ctx->profiling = on != 0; return UPB_OK; }

extern "C" int upb_last_stage_times (const upb_ctx* ctx, uint32_t cap, const char** names_out, float* ms_out, uint32_t* n_out) {
  if (!ctx || !n_out) return UPB_ERR_BAD_ARG;
  const uint32_t n = (uint32_t)ctx->stageMs.size();
  *n_out = n;
  for (uint32_t k = 0; k < n && k < cap; ++k) { if (names_out) names_out[k] = ctx->stageNames[k + 1]; if (ms_out) ms_out[k] = ctx->stageMs[k]; }
  return UPB_OK;
}

// device-wide exclusive scan whose element count may live on the device (upb_device.cuh: scanv_*); `cap` bounds it
template <class Load>
static int scan_dev (upb_ctx* ctx, Load ld, DevN n, uint64_t cap, unsigned long long* d_out, unsigned long long* d_total, uint32_t* launches) {
  const uint64_t maxTiles = (cap + SCAN_BLOCK - 1) / SCAN_BLOCK;
  unsigned long long* part = SCR("scanPartial", unsigned long long, maxTiles + 1);
  const unsigned grid = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(maxTiles, (uint64_t)ctx->smCount * 4));
  scanv_partial_kernel<<<grid, SCAN_THREADS, 0, ctx->cur>>>(ld, n, part);
  scanv_top_kernel<<<1, SCAN_THREADS, 0, ctx->cur>>>(part, n, d_total);
  scanv_final_kernel<<<grid, SCAN_THREADS, 0, ctx->cur>>>(ld, n, part, d_out, 0ull);
  CU(cudaGetLastError());
  *launches += 3;
  return UPB_OK;
}

static int ensure_host (upb_ctx* ctx, HostSlab& h, size_t bytes) {
  if (bytes <= h.cap) return UPB_OK;
  if (h.p) { cudaFreeHost(h.p); h.p = nullptr; h.cap = 0; }
  const size_t want = bytes + bytes / 2 + 4096;
  cudaError_t e = cudaMallocHost(&h.p, want);
  if (e != cudaSuccess) { h.p = nullptr; return fail(ctx, UPB_ERR_NOMEM, std::string("cudaMallocHost(") + std::to_string(want) + "): " + cudaGetErrorString(e)); }
  h.cap = want;
  return UPB_OK;
}

// A table the ranks of a multi-GPU job share (upb_call_regions_shared): columns in mapped host memory every rank writes its rows
// into, and the exchange area through which row_place_kernel learns the other ranks' row counts.
struct SharedDst {
  unsigned char* slab;              // table columns, laid out by slab_layout(rowCap, n, C)
  uint64_t rowCap;
  XchgArgs x;
};

// One pass over the units: every kernel of the region pipeline is enqueued without reading a size back (sizing == false), or with
// two read-backs that size the per-survivor arrays and the table exactly (sizing == true: the first call of a context, and the
// repeat of a call whose numbers outgrew what the previous call left behind).  *retry: run again with sizing.
// ---- left-edge replay (upb_edge.cuh): one launch and one read-back; rare (a tag within `bandwidth` of a contig start) -------------
constexpr uint32_t EDGE_REG_CAP = 1u << 18, EDGE_ROW_CAP = 1024;
struct EdgeRec {                     // optional recordings
  int profUnit = -1;                 // unit whose -w profile points are wanted
  uint32_t vecMinLen = 0xFFFFFFFFu;  // accepted rows longer than this keep their forward / reverse score vectors (the shift sweep)
};
struct EdgeRun {
  EdgeArgs eg;                       // device side, valid until the scratch arrays are reused
  uint32_t* d_affected = nullptr;
  std::vector<uint32_t> out;         // [3 * nAff + 8]: stop, nCand, nRows per affected unit; flags; profile points; vector pool use
  uint32_t profN = 0;
};
static int run_edge_replay (upb_ctx* ctx, const PipeCfg& cfg, const Prep& pr, cudaStream_t st_, const EdgeRec& rec, EdgeRun& er, uint32_t* launches) {
  const uint32_t nAff = (uint32_t)pr.affected.size(), n = ctx->nSamples, w = 2 * cfg.bw + 1;
  EdgeArgs& eg = er.eg;
  memset(&eg, 0, sizeof eg);
  er.d_affected = SCR("edgeAffected", uint32_t, nAff);
  uint32_t* d_chainNext = SCR("edgeChainNext", uint32_t, nAff);
  uint8_t* d_chainHead = SCR("edgeChainHead", uint8_t, nAff);
  CU(cudaMemcpyAsync(er.d_affected, pr.affected.data(), (size_t)nAff * 4, cudaMemcpyHostToDevice, st_));
  CU(cudaMemcpyAsync(d_chainNext, pr.chainNext.data(), (size_t)nAff * 4, cudaMemcpyHostToDevice, st_));
  CU(cudaMemcpyAsync(d_chainHead, pr.chainHead.data(), (size_t)nAff, cudaMemcpyHostToDevice, st_));
  eg.T = pr.T; eg.units = ctx->scratch["units"].as<Unit>(); eg.affected = er.d_affected; eg.chainNext = d_chainNext; eg.chainHead = d_chainHead;
  eg.nAffected = nAff; eg.control = ctx->scratch["control"].as<uint8_t>();
  eg.coeffBySample = pr.d_coeffBySample; eg.taps = ctx->scratch["taps"].as<double>(); eg.bw = cfg.bw; eg.nondir = cfg.nondir ? 1 : 0;
  eg.thr = cfg.thr; eg.kurtThr = cfg.kurtThr; eg.corrThr = cfg.nondir ? cfg.corrThr : -1.0; eg.hitThr = cfg.hitThr;
  eg.cursors = SCR("edgeCursors", unsigned long long, (size_t)nAff * 2 * n);
  eg.ringF = SCR("edgeRingF", double, (size_t)nAff * w); eg.ringR = SCR("edgeRingR", double, (size_t)nAff * w);
  eg.ringHits = SCR("edgeRingHits", uint32_t, (size_t)nAff * w * n); eg.ringHas = SCR("edgeRingHas", uint8_t, (size_t)nAff * w);
  eg.regPc = SCR("edgeRegPc", uint32_t, (size_t)nAff * EDGE_REG_CAP); eg.regHas = SCR("edgeRegHas", uint8_t, (size_t)nAff * EDGE_REG_CAP);
  eg.regF = SCR("edgeRegF", double, (size_t)nAff * EDGE_REG_CAP); eg.regR = SCR("edgeRegR", double, (size_t)nAff * EDGE_REG_CAP);
  eg.regSums = SCR("edgeRegSums", uint32_t, (size_t)nAff * n); eg.callCounts = SCR("edgeCall", uint32_t, (size_t)nAff * n);
  eg.regCap = EDGE_REG_CAP; eg.rowCap = EDGE_ROW_CAP; eg.maxPositions = 1ull << 23;
  uint32_t* eo = SCR("edgeOut", uint32_t, (size_t)3 * nAff + 8);
  CU(cudaMemsetAsync(eo, 0, ((size_t)3 * nAff + 8) * 4, st_));
  eg.stop = eo; eg.nCand = eo + nAff; eg.nRows = eo + 2 * nAff; eg.flags = eo + 3 * nAff;
  eg.profN = eo + 3 * nAff + 1; eg.vecUsed = eo + 3 * nAff + 2;
  eg.rows = SCR("edgeRows", EdgeRow, (size_t)nAff * EDGE_ROW_CAP);
  eg.rowCounts = SCR("edgeRowCounts", uint32_t, (size_t)nAff * EDGE_ROW_CAP * n);
  eg.profAff = 0xFFFFFFFFu; eg.vecMinLen = rec.vecMinLen;
  if (rec.profUnit >= 0) {
    for (uint32_t k = 0; k < nAff; ++k) if (pr.affected[k] == (uint32_t)rec.profUnit) eg.profAff = k;
    eg.profCap = (uint32_t)eg.maxPositions + 4 * w;                   // the budget bounds the positions a replay processes
    eg.profPos = SCR("edgeProfPos", uint32_t, eg.profCap); eg.profScore = SCR("edgeProfScore", double, eg.profCap);
  }
  if (rec.vecMinLen != 0xFFFFFFFFu) {
    eg.vecCap = nAff * EDGE_REG_CAP > (1u << 24) ? (1u << 24) : nAff * EDGE_REG_CAP;
    eg.vecF = SCR("edgeVecF", double, eg.vecCap); eg.vecR = SCR("edgeVecR", double, eg.vecCap);
  }
  left_edge_kernel<<<nAff, 32, 0, st_>>>(eg);
  CU(cudaGetLastError());
  *launches += 1;
  er.out.resize((size_t)3 * nAff + 8);
  CU(cudaMemcpyAsync(er.out.data(), eo, er.out.size() * 4, cudaMemcpyDeviceToHost, st_));
  CU(cudaStreamSynchronize(st_));
  if (er.out[3 * nAff]) return fail(ctx, UPB_ERR_UNSUPPORTED, "left-edge replay (SURVEY Q2) ran out of its scratch or position budget");
  er.profN = er.out[3 * nAff + 1];
  return UPB_OK;
}

static int run_pipeline_once (upb_ctx* ctx, const PipeCfg& cfg, Prep& pr, PipeOut& out, bool sizing, bool* retry, const SharedDst* shared) {
  CU(cudaSetDevice(ctx->device));
  *retry = false;
  uint32_t launches = 0;
  typedef std::chrono::steady_clock Clk;
  const Clk::time_point tA = Clk::now();
  ++ctx->callNo;
  double waitMid = 0;                                                 // ms the host spent blocked on the device before the final read-back
  stage_reset(ctx);
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  // A load that is still on its way (upb_load_tags_async): work through its groups as they arrive — unless this call needs a
  // read-back in the middle (sizing, -D correlation), every mini-tile at once (controls) or the first positions' neighbourhood
  // (left-edge replay), in which case it simply waits for the load.
  // (a shared-table call too, once the context's arrays are sized: its groups count their rows as they go, the last one publishes and places
  //  them all — row_place_kernel, countOnly — and the rows of the earlier groups are emitted behind it)
  bool pipelined = ctx->loadPending && !sizing && !cfg.needCorr && (!shared || (ctx->capN1 != 0 && ctx->capCells != 0));
  if (pipelined && shared && getenv("UPB_SHARED_PIPELINE") && !atoi(getenv("UPB_SHARED_PIPELINE"))) pipelined = false;
  if (pipelined) {
    if (cfg.control) for (uint32_t s = 0; s < ctx->nSamples; ++s) if (cfg.control[s]) pipelined = false;
    for (uint32_t fp : ctx->firstPos) if (fp != 0 && fp <= cfg.bw) pipelined = false;
  }
  int r = prepare(ctx, cfg, pr, &launches, pipelined);
  if (r) return r;
  const Clk::time_point tB = Clk::now();
  const uint32_t n = ctx->nSamples, C = (uint32_t)ctx->sizes.size();
  const uint32_t nUnits = (uint32_t)pr.units.size();
  out.units = pr.units;
  out.unitCand.assign(nUnits, 0);
  out.unitEdgeStop.assign(nUnits, 0);
  out.nCand = 0; out.R = 0;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  ctx->stats.n_entries = ctx->nEntries;
  ctx->stats.kde_tiles = pr.totalSpans;
  cudaStream_t st_ = ctx->stream;
  GroupCtr* g = (GroupCtr*)ctx->d_small.p;
  CallCtr* call = (CallCtr*)((unsigned char*)ctx->d_small.p + CALLCTR_OFFSET);
  const bool empty = nUnits == 0 || pr.nSrc == 0;
  if (empty && !shared) {
    out.layout = slab_layout(0, n, C);
    if ((r = acquire_slab(ctx, out.layout.totalBytes, out.slab))) return r;
    memset(out.slab.p, 0, out.layout.totalBytes);
    cudaEventRecord(ctx->ev[3], st_); cudaEventSynchronize(ctx->ev[3]);
    ctx->stats.launches = launches;
    if (ctx->loadPending) return load_finish(ctx);
    return UPB_OK;
  }
  Unit* d_units = ctx->scratch["units"].as<Unit>();
  stage_mark(ctx, "begin");

  // ---- K1: the fused span kernel and the two zone-map passes in front of it --------------------------------
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
  ha.hotCoefF = pr.d_hotCoefF; ha.hotJ = pr.hotJ;
  ha.thrLow = std::nextafter((float)cfg.thr, -INFINITY);
  const double flatBound = kmax * (1.0 + pr.maxCoeff) * (1.0 + 1e-9);  // largest tap: maxTap * window weight bounds every score (zone-map prefilter)
  sa.coeff = pr.d_coeffBySample;
  sa.cN = kmax / ((double)cfg.bw * (double)cfg.bw);
  sa.dS = kmax * 1e-11;                                               // >= 100x the worst-case rounding difference between the two forms
  sa.flags = &g->flags; sa.ticket = &g->ticket; sa.runCount = &g->runCount; sa.runReal = &g->runReal; sa.stat = call->stat;
  sa.spanRunCount = SCR("spanRunCount", uint32_t, pr.totalSpans);
  WarmRec* warmList = SCR("warmList", WarmRec, pr.totalSpans);
  uint2* hotList = SCR("hotList", uint2, pr.totalSpans);
  uint32_t* hotMask = SCR("hotMask", uint32_t, (size_t)pr.totalSpans * SPAN_WORDS);
  sa.hotList = hotList; sa.nHot = &g->nHot; sa.hotMask = hotMask;
  ha.T = pr.T; ha.units = d_units; ha.warmList = warmList; ha.nWarm = &g->nWarm; ha.bw = cfg.bw; ha.nondir = cfg.nondir ? 1 : 0;
  ha.hotMask = hotMask; ha.hotList = hotList; ha.nHot = &g->nHot; ha.spanRunCount = sa.spanRunCount; ha.stat = call->stat;
  WarmArgs wa;
  wa.T = pr.T; wa.units = d_units; wa.nUnits = nUnits; wa.totalSpans = pr.totalSpans; wa.nondir = cfg.nondir ? 1 : 0;
  wa.nAdj = ((31 + 2 * cfg.bw) >> MT_SHIFT) + 2;
  wa.warmList = warmList; wa.nWarm = &g->nWarm; wa.spanRunCount = sa.spanRunCount; wa.bw = cfg.bw;
  wa.flatBound = flatBound; wa.thr = cfg.thr;
  wa.loadFlags = pipelined ? ctx->d_loadFlags.as<uint32_t>() : nullptr;

  // ---- K0: left-edge replay for the units that need it (rare: one small kernel and one read-back) ------------
  const uint32_t nAff = (uint32_t)pr.affected.size();
  EdgeArgs eg;
  memset(&eg, 0, sizeof eg);
  uint32_t* d_affected = nullptr;
  std::vector<uint32_t> edgeCand(nAff, 0), edgeRows(nAff, 0);
  uint32_t* d_edgeRowsOfUnit = nullptr; int32_t* d_affOfUnit = nullptr; uint32_t* d_drop = nullptr;
  uint64_t edgeRowsTotal = 0;
  if (nAff) {
    EdgeRun er;
    if ((r = run_edge_replay(ctx, cfg, pr, st_, EdgeRec(), er, &launches))) return r;
    eg = er.eg; d_affected = er.d_affected;
    const std::vector<uint32_t>& eoh = er.out;
    std::vector<uint32_t> rowsOfUnit(nUnits, 0);
    std::vector<int32_t> affOfUnit(nUnits, -1);
    for (uint32_t k = 0; k < nAff; ++k) {
      edgeCand[k] = eoh[nAff + k]; edgeRows[k] = eoh[2 * nAff + k]; out.unitEdgeStop[pr.affected[k]] = eoh[k];
      rowsOfUnit[pr.affected[k]] = edgeRows[k]; affOfUnit[pr.affected[k]] = (int32_t)k; edgeRowsTotal += edgeRows[k];
    }
    d_edgeRowsOfUnit = SCR("edgeRowsOfUnit", uint32_t, nUnits);
    d_affOfUnit = SCR("edgeAffOfUnit", int32_t, nUnits);
    d_drop = SCR("edgeDrop", uint32_t, nAff);
    CU(cudaMemcpyAsync(d_edgeRowsOfUnit, rowsOfUnit.data(), (size_t)nUnits * 4, cudaMemcpyHostToDevice, st_));
    CU(cudaMemcpyAsync(d_affOfUnit, affOfUnit.data(), (size_t)nUnits * 4, cudaMemcpyHostToDevice, st_));
    CU(cudaMemsetAsync(d_drop, 0, (size_t)nAff * 4, st_));
    CU(cudaStreamSynchronize(st_));                                   // the host vectors go out of scope
    stage_mark(ctx, "left_edge_kernel");
  }

  auto hmark = [&] (const char* name) { ctx->hostMarks.push_back(std::make_pair(name, Clk::now())); };
  hmark("h:setup");
  // ---- arrays sized by capacity ----------------------------------------------------------------------------------
  size_t runCap = std::max<size_t>(std::max<size_t>(1u << 18, (size_t)(ctx->nEntries / 64)), ctx->needRunCap);
  // (what the array already holds — minus SCR's 16 bytes of slack: counting them as room made the request outgrow the array whenever
  //  cap mod sizeof(SpanRun) < 16, and every run-sized array was then freed and reallocated 25 % larger, call after call, until a
  //  capacity happened to leave more than 16 bytes over: 2 - 20 ms stalls in the first calls of a context, r02 N = 4 trace)
  { const size_t cap = ctx->scratch["runs"].cap; const size_t have = cap > 16 ? (cap - 16) / sizeof(SpanRun) : 0; if (have > runCap) runCap = have; }
  runCap = std::min<size_t>(runCap, 0xFFFFFFF0u);
  sa.runCap = (uint32_t)runCap;                                        // (the run-sized arrays are per lane: allocated at the head of the group loop)
  unsigned long long* spanOrd = SCR("spanOrd", unsigned long long, pr.totalSpans);
  unsigned long long* unitCand = SCR("unitCand", unsigned long long, 6 * (size_t)nUnits + 8);
  unsigned long long* unitFirst = unitCand + nUnits;
  unsigned long long* normalFirst = unitFirst + nUnits;               // nUnits + 1
  unsigned long long* unitRowsDev = normalFirst + nUnits + 1;
  unsigned long long* edgeRowBase = unitRowsDev + nUnits;
  long long* rowShift = (long long*)(edgeRowBase + nUnits);
  unsigned long long* d_tir = SCR("tir", unsigned long long, n);
  CU(cudaMemsetAsync(unitCand, 0, (6 * (size_t)nUnits + 8) * 8, st_));
  CU(cudaMemsetAsync(d_tir, 0, (size_t)n * 8, st_));
  (void)SCR("scanPartial", unsigned long long, std::max<size_t>(runCap, pr.totalSpans) / SCAN_BLOCK + 2);   // largest scan of the call: no regrowth (= free + malloc) in between
  const unsigned GS = (unsigned)ctx->smCount * 8;                     // grid of the grid-stride kernels

  // pinned block the counters come back in: GroupCtr, CallCtr, tir[n], unitCand[nUnits], unitRows[nUnits], edgeDrop[nAff]
  const size_t resBytes = 1024 + (size_t)n * 8 + 2 * (size_t)nUnits * 8 + (size_t)nAff * 4 + 64;
  if ((r = ensure_host(ctx, ctx->h_res, resBytes))) return r;
  unsigned char* hres = (unsigned char*)ctx->h_res.p;
  GroupCtr* hg = (GroupCtr*)hres;
  CallCtr* hcall = (CallCtr*)(hres + 512);
  unsigned long long* hTir = (unsigned long long*)(hres + 1024);
  unsigned long long* hUnitCand = hTir + n;
  unsigned long long* hUnitRows = hUnitCand + nUnits;
  uint32_t* hDrop = (uint32_t*)(hUnitRows + nUnits);
  auto read_group = [&] () -> int {                                    // sizing read-back of the group counters
    const Clk::time_point w0 = Clk::now();
    CU(cudaMemcpyAsync(hg, g, sizeof(GroupCtr), cudaMemcpyDeviceToHost, st_));
    CU(cudaStreamSynchronize(st_));
    waitMid += std::chrono::duration<double, std::milli>(Clk::now() - w0).count();
    return UPB_OK;
  };

  hmark("h:scratch");
  // ---- the groups of units: one group (all units), or one per load group of a pending load ------------------------------
  struct PGroup { uint32_t u0, u1, s0, s1; int ev0, ev1; };
  std::vector<PGroup> pgroups;
  if (!pipelined) { PGroup pg; pg.u0 = 0; pg.u1 = nUnits; pg.s0 = 0; pg.s1 = pr.totalSpans; pg.ev0 = pg.ev1 = -1; pgroups.push_back(pg); }
  else {
    // units are sorted by (buffer, contig); a directional unit needs the load group (its strand, its contig range), a -D unit both strands'
    std::vector<std::pair<uint32_t, uint32_t>> order;                  // (sort key, load group index), strand 0 groups only for -D
    for (uint32_t k = 0; k < ctx->groups.size(); ++k) {
      const LoadGroup& lg = ctx->groups[k];
      if (cfg.nondir && lg.d != 0) continue;
      order.push_back(std::make_pair((cfg.nondir ? 0u : lg.d) * C + lg.c0, k));
    }
    std::sort(order.begin(), order.end());
    uint32_t u = 0;
    for (const auto& ok : order) {
      const LoadGroup& lg = ctx->groups[ok.second];
      const uint32_t b = cfg.nondir ? 0u : lg.d;
      PGroup pg;
      while (u < nUnits && (pr.units[u].buffer < b || (pr.units[u].buffer == b && pr.units[u].contig < lg.c0))) ++u;   // (cannot skip any: every unit's contig is kept)
      pg.u0 = u;
      while (u < nUnits && pr.units[u].buffer == b && pr.units[u].contig < lg.c1) ++u;
      pg.u1 = u;
      if (pg.u1 == pg.u0) continue;
      pg.s0 = pr.units[pg.u0].spanStart; pg.s1 = pr.units[pg.u1 - 1].spanStart + pr.units[pg.u1 - 1].nSpans;
      pg.ev0 = (int)ok.second; pg.ev1 = -1;
      if (cfg.nondir) for (uint32_t k = 0; k < ctx->groups.size(); ++k) if (ctx->groups[k].d == 1 && ctx->groups[k].c0 == lg.c0) pg.ev1 = (int)k;
      pgroups.push_back(pg);
    }
    // The index is built in fine groups (the copy -> index latency stays short), regions are called in coarser ones: every group of
    // units costs ~0.5 ms of launch latency whatever its size (~27 small kernels), so 26 of them would trail the load by more than
    // they hide.  Per buffer: 35 / 30 / 20 / 15 % of the entries — the last group, the only one that runs after the last byte, is small.
    if (pgroups.size() > 4) {
      std::vector<PGroup> merged;
      size_t a = 0;
      while (a < pgroups.size()) {
        const uint32_t b = pr.units[pgroups[a].u0].buffer;
        size_t e = a;
        uint64_t total = 0;
        while (e < pgroups.size() && pr.units[pgroups[e].u0].buffer == b) { total += ctx->groups[pgroups[e].ev0].entries + 1; ++e; }
        const double cut[4] = {0.35, 0.65, 0.85, 1.0};
        uint64_t acc = 0;
        int part = 0;
        size_t start = a;
        for (size_t k = a; k < e; ++k) {
          acc += ctx->groups[pgroups[k].ev0].entries + 1;
          if ((double)acc >= cut[part] * (double)total || k + 1 == e) {
            PGroup pg = pgroups[start];
            pg.u1 = pgroups[k].u1; pg.s1 = pgroups[k].s1;
            for (size_t q = start; q <= k; ++q) { pg.ev0 = std::max(pg.ev0, pgroups[q].ev0); pg.ev1 = std::max(pg.ev1, pgroups[q].ev1); }   // load groups complete in order on one stream: the latest one's event covers the others
            merged.push_back(pg);
            start = k + 1;
            while (part < 3 && (double)acc >= cut[part] * (double)total) ++part;
          }
        }
        a = e;
      }
      pgroups.swap(merged);
    }
    if (pgroups.empty() || pgroups.back().u1 != nUnits || pgroups.front().u0 != 0) {   // not expected; fall back to waiting for the whole load
      pipelined = false;
      const int rl = load_finish(ctx); if (rl) return rl;
      pgroups.clear();
      PGroup pg; pg.u0 = 0; pg.u1 = nUnits; pg.s0 = 0; pg.s1 = pr.totalSpans; pg.ev0 = pg.ev1 = -1; pgroups.push_back(pg);
    }
  }
  // Two lanes (see upb_ctx::laneStream).  A resident call is cut into `wantGroups` groups of units of about equal span counts, dealt to the lanes in
  // turn; a call behind a load deals the groups it already has.  Not for calls with a read-back in the middle (sizing, -D correlation), the
  // left-edge replay or per-stage profiling (events on one stream).  With a shared table the groups before the last only COUNT their rows
  // (row_place_kernel, countOnly); the last group's launch publishes the rank's counts once, waits for the other ranks and places every
  // unit, and the earlier groups' rows are emitted behind it.
  // OFF by default — measured on C3 (tools/lanes_ab.py, profiles/r02_lanes_ab.jsonl): two groups 2.32 -> 2.22 ms on the whole workload (326 M
  // entries) in a tight loop of calls, 1 - 3 % inside bench.py's timed loop, level at bandwidth 200, three and more groups slower (every group
  // pays the chain's ~27 launches); on a rank's shard of a 2 / 4 / 8-GPU job (163 / 81 / 40 M entries) two groups LOSE 4 - 7 %; behind a load
  // (e2e) 12.51 ms with two lanes against 12.45 with one.  The pooling kernels of two groups share the SMs and the emits share the PCIe link,
  // so only launch gaps and one group's emit hide behind the other group's kernels — and the span kernel's event bracket then holds a share of
  // the other lane's kernels (0.82 -> 0.93 ms).  UPB_CALL_GROUPS=n (n >= 2) turns the lanes on with n groups per resident call.
  int wantGroups = 1;
  if (const char* e = getenv("UPB_CALL_GROUPS")) wantGroups = std::max(1, std::min(16, atoi(e)));
  bool twoLanes = wantGroups > 1 && !sizing && !cfg.needCorr && nAff == 0 && !ctx->profiling && !empty && ctx->laneStream != nullptr;
  if (shared && (ctx->capN1 == 0 || ctx->capCells == 0)) twoLanes = false;   // (reads its sizes back in the middle of the call)
  if (twoLanes && !pipelined) {
    if (nUnits < 2 || pr.totalSpans < (1u << 16)) twoLanes = false;    // small inputs: one chain of launches is all there is to wait for
    else {
      pgroups.clear();
      uint32_t u = 0;
      for (int k = 0; k < wantGroups && u < nUnits; ++k) {
        const uint64_t upTo = (uint64_t)pr.totalSpans * (uint64_t)(k + 1) / (uint64_t)wantGroups;
        PGroup pg; pg.u0 = u; pg.ev0 = pg.ev1 = -1;
        while (u < nUnits && (u == pg.u0 || k + 1 == wantGroups || (uint64_t)pr.units[u].spanStart + pr.units[u].nSpans / 2 < upTo)) ++u;
        pg.u1 = u; pg.s0 = pr.units[pg.u0].spanStart; pg.s1 = pr.units[pg.u1 - 1].spanStart + pr.units[pg.u1 - 1].nSpans;
        pgroups.push_back(pg);
      }
      if (pgroups.size() < 2) twoLanes = false;
    }
  }
  if (twoLanes && pgroups.size() < 2) twoLanes = false;
  // (on every way out: the helpers are back on lane 0, and — an error may leave in the middle of the groups — `stream` is behind whatever the other lane still runs)
  bool laneBusy = twoLanes;                                              // the second stream has been given work in this call
  struct LaneReset { upb_ctx* c; const bool* two; ~LaneReset () { c->lane = 0; c->cur = c->stream; if (*two && cudaEventRecord(c->laneDoneEv, c->laneStream) == cudaSuccess) cudaStreamWaitEvent(c->stream, c->laneDoneEv, 0); } } laneReset{ctx, &laneBusy};
  // Split tail: the statistics + emit half of a large resident call (pooling, moments, final scan, row placement, rows over PCIe) runs in two halves of
  // the units on the two streams — the second half's pooling starts when the first half's is through, so the first half's moments, scan and 9 MB of
  // rows crossing PCIe run beside it.  Unlike the two lanes above the front of the call (prefilters, span kernel, order / stitch / filter) runs once.
  // Measured (tools/lanes_ab.py, groups = 0; profiles/r02_split_tail_ab.jsonl): C3 at bandwidth 50 2.34 -> 2.22 ms, at 10 2.80 -> 2.74, but 1.52 -> 1.59
  // at bandwidth 200 (82 K regions) and 3 - 10 % slower on a 2 / 4 / 8-GPU rank's shard: nine more launches against half of a shorter emit.  Hence
  // only for calls whose predecessor had >= 150 K hit-filter survivors.  UPB_SPLIT_TAIL=0 never, =2 whenever possible (the GPU tests force it).
  const int splitMode = getenv("UPB_SPLIT_TAIL") ? atoi(getenv("UPB_SPLIT_TAIL")) : 1;   // 0 never, 1 large calls (by the last call's survivors), 2 whenever possible (tests)
  const bool splitTail = splitMode != 0 && !twoLanes && !pipelined && !sizing && !cfg.needCorr && !shared && nAff == 0 && !ctx->profiling && !empty &&
                         ctx->laneStream != nullptr && pgroups.size() == 1 && nUnits >= 2 && (splitMode == 2 || ctx->lastN1 >= 150000);
  while (ctx->placeEv.size() < pgroups.size()) { cudaEvent_t e = nullptr; CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming)); ctx->placeEv.push_back(e); }
  while (ctx->spanEv.size() < 2 * pgroups.size()) { cudaEvent_t e = nullptr; CU(cudaEventCreate(&e)); ctx->spanEv.push_back(e); }
  GroupCtr* const gLane[2] = {g, (GroupCtr*)((unsigned char*)ctx->d_small.p + LANE1_GROUPCTR_OFFSET)};
  if (twoLanes) { CU(cudaEventRecord(ctx->laneSetupEv, ctx->stream)); CU(cudaStreamWaitEvent(ctx->laneStream, ctx->laneSetupEv, 0)); }
  const bool trace = getenv("UPB_TRACE") && *getenv("UPB_TRACE");
  if (trace) fprintf(stderr, "[upb] call: %s, %zu group(s) of units on %d lane(s), sizing %d\n", pipelined ? "pipelined behind the load" : "resident", pgroups.size(), twoLanes ? 2 : 1, (int)sizing);
  uint64_t rowCap = 0;
  unsigned char* hs = nullptr;                                         // table columns (mapped host memory)
  if (!empty && !shared && !sizing) {                                  // the table: sized from the previous call (sizing: once this call's row count is known)
    rowCap = ctx->capRows;
    out.layout = slab_layout(rowCap, n, C);
    if ((r = acquire_slab(ctx, out.layout.totalBytes, out.slab))) return r;
    hs = (unsigned char*)out.slab.p;
  }
  if (shared) { rowCap = shared->rowCap; hs = shared->slab; out.layout = slab_layout(rowCap, n, C); }
  struct Deferred { EmitArgs ea; GroupEndArgs ge; cudaStream_t st; };
  std::vector<Deferred> deferred;
  if (!empty) for (size_t gi = 0; gi < pgroups.size(); ++gi) {
    const uint32_t u0 = pgroups[gi].u0, u1 = pgroups[gi].u1, s0 = pgroups[gi].s0, s1 = pgroups[gi].s1;
    const int lane = twoLanes ? (int)(gi & 1) : 0;
    ctx->lane = lane; ctx->cur = lane ? ctx->laneStream : ctx->stream;
    cudaStream_t st_ = ctx->cur;                                         // (shadows: everything below is this lane's)
    GroupCtr* g = gLane[lane];
    sa.flags = &g->flags; sa.ticket = &g->ticket; sa.runCount = &g->runCount; sa.runReal = &g->runReal; sa.nHot = &g->nHot;
    ha.nWarm = &g->nWarm; ha.nHot = &g->nHot; wa.nWarm = &g->nWarm;
    WarmRec* warmList = SCR("warmList", WarmRec, pr.totalSpans);
    uint2* hotList = SCR("hotList", uint2, pr.totalSpans);
    sa.hotList = hotList; ha.warmList = warmList; ha.hotList = hotList; wa.warmList = warmList;
    sa.runs = SCR("runs", SpanRun, runCap);
    uint32_t* ordIdx = SCR("ordIdx", uint32_t, runCap);
    uint32_t* head = SCR("head", uint32_t, runCap);
    unsigned long long* headScan = SCR("headScan", unsigned long long, runCap);
    Cand2* cand = SCR("cand", Cand2, runCap);
    uint32_t* pass = SCR("pass", uint32_t, runCap);
    uint32_t* candLen = SCR("candLen", uint32_t, runCap);
    unsigned long long* pvScan = SCR("pvScan", unsigned long long, runCap);
    uint32_t* list1 = SCR("list1", uint32_t, runCap);
    (void)SCR("scanPartial", unsigned long long, std::max<size_t>(runCap, pr.totalSpans) / SCAN_BLOCK + 2);
    if (pgroups[gi].ev0 >= 0) CU(cudaStreamWaitEvent(st_, ctx->groupEv[pgroups[gi].ev0], 0));
    if (pgroups[gi].ev1 >= 0) CU(cudaStreamWaitEvent(st_, ctx->groupEv[pgroups[gi].ev1], 0));
    hmark("h:events");
    CU(cudaMemsetAsync(g, 0, sizeof(GroupCtr), st_));
    hmark("h:memset g");
    wa.spanBegin = s0; wa.spanEnd = s1;
    warm_span_kernel<<<(s1 - s0 + 255) / 256, 256, 0, st_>>>(wa);
    hmark("h:warm launch");
    { int perSm = 0;
      if ((r = cached_occupancy(ctx, hot_span_kernel, HOT_THREADS, 0, &perSm))) return r;
      hot_span_kernel<<<ctx->smCount * perSm, HOT_THREADS, 0, st_>>>(ha); }
    hmark("h:hot launch");
    launches += 2;
    stage_mark(ctx, "warm_span_kernel+hot_span_kernel");
    CU(cudaEventRecord(ctx->spanEv[2 * gi], st_));
    if ((r = launch_span(ctx, sa, cfg.nondir, &launches))) return r;
    CU(cudaEventRecord(ctx->spanEv[2 * gi + 1], st_));
    stage_mark(ctx, "span_kernel");

    // ---- K2: (span, rank) order, stitch runs cut by piece / span boundaries, hit filter ------------------------------
    if ((r = scan_dev(ctx, LoadU32{sa.spanRunCount + s0}, devn_imm(s1 - s0), s1 - s0, spanOrd + s0, &g->totSpanRuns, &launches))) return r;
    run_order_kernel<<<GS, 256, 0, st_>>>(sa.runs, g, (uint32_t)runCap, spanOrd, ordIdx);
    run_heads2_kernel<<<GS, 256, 0, st_>>>(sa.runs, ordIdx, g, d_units, nUnits, head);
    if ((r = scan_dev(ctx, LoadU32{head}, devn_u32(&g->runReal), runCap, headScan, &g->nCand, &launches))) return r;
    merge_runs2_kernel<<<GS, 256, 0, st_>>>(sa.runs, ordIdx, head, headScan, g, d_units, nUnits, cfg.hitThr, cand, pass, candLen, unitCand);
    unit_first_kernel<<<1, 32, 0, st_>>>(unitCand, u0, u1, unitFirst);
    launches += 4;
    if (nAff) {                                                         // what the replay owns is not a Region of the reference
      edge_drop_kernel<<<GS, 256, 0, st_>>>(d_affected, eg.stop, nAff, cand, g, pass, candLen, d_drop);
      launches += 1;
    }
    if ((r = scan_dev(ctx, LoadPacked{pass, candLen}, devn_u64(&g->nCand), runCap, pvScan, &g->pv, &launches))) return r;
    if (splitTail) {
      uint32_t uMid = u0 + 1;
      while (uMid + 1 < u1 && (uint64_t)pr.units[uMid].spanStart < (uint64_t)s0 + (s1 - s0) / 2) ++uMid;
      const size_t n1Cap = std::min<size_t>(ctx->capN1, runCap), cellCap = ctx->capCells;
      uint32_t* statCells = SCR("statCells", uint32_t, cellCap + 1);   // (pooled cells: one array, the halves own disjoint stretches of it)
      GroupCtr* gT[2] = {(GroupCtr*)((unsigned char*)ctx->d_small.p + TAIL_GROUPCTR_OFFSET), (GroupCtr*)((unsigned char*)ctx->d_small.p + TAIL_GROUPCTR_OFFSET + 256)};
      uint32_t* listT[2] = {list1, nullptr};
      ctx->lane = 1; listT[1] = SCR("list1", uint32_t, runCap); ctx->lane = 0;
      compact_pv_split_kernel<<<GS, 256, 0, st_>>>(pass, pvScan, g, gT[0], gT[1], unitFirst, uMid, (uint32_t)n1Cap, cellCap, listT[0], listT[1]);
      launches += 1;
      CU(cudaEventRecord(ctx->laneSetupEv, st_));
      CU(cudaStreamWaitEvent(ctx->laneStream, ctx->laneSetupEv, 0));
      laneBusy = true;
      const size_t groupBytes = ((size_t)(ST_CHUNK + n) * 4 + 15) / 16 * 16;
      const size_t statSmem = (size_t)ST_WARPS * ST_PER_WARP * groupBytes;
      if (statSmem > 200 * 1024) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many samples for the region statistics kernel");
      int poolPerSm = 0;
      if ((r = cached_occupancy(ctx, region_pool_kernel, ST_WARPS * 32, statSmem, &poolPerSm))) return r;
      const SlabLayout& SL = out.layout;
      for (int t = 0; t < 2; ++t) {
        ctx->lane = t; ctx->cur = t ? ctx->laneStream : ctx->stream;
        cudaStream_t ts = ctx->cur;
        uint32_t* statCounts = SCR("statCounts", uint32_t, n1Cap * n);
        double* statKurt = SCR("statKurt", double, n1Cap);
        uint32_t* pass2 = SCR("pass2", uint32_t, n1Cap);
        unsigned long long* pass2Scan = SCR("pass2Scan", unsigned long long, n1Cap);
        (void)SCR("scanPartial", unsigned long long, std::max<size_t>(runCap, pr.totalSpans) / SCAN_BLOCK + 2);
        StatArgs sg;
        sg.T = pr.T; sg.units = d_units; sg.cand = cand; sg.list = listT[t]; sg.g = gT[t];
        sg.nondir = cfg.nondir ? 1 : 0; sg.kurtThr = cfg.kurtThr; sg.counts = statCounts; sg.kurt = statKurt; sg.pass = pass2;
        sg.pvScan = pvScan; sg.cells = statCells; sg.cellCap = cellCap; sg.ticket = &gT[t]->poolTicket;
        if (t == 1) CU(cudaStreamWaitEvent(ts, ctx->lanePoolEv, 0));     // the halves' pooling kernels one after the other: side by side they only share the SMs
        // (the pooling kernel at full occupancy owns every register of an SM: the second half runs one CTA per SM short, so that the first half's
        //  emit kernels get on the SMs beside it instead of behind it)
        const int less = (t == 1 && poolPerSm > 1) ? (getenv("UPB_SPLIT_POOL_LESS") ? atoi(getenv("UPB_SPLIT_POOL_LESS")) : 1) : 0;
        region_pool_kernel<<<(unsigned)(ctx->smCount * std::max(poolPerSm - less, 1)), ST_WARPS * 32, statSmem, ts>>>(sg);
        if (t == 0) CU(cudaEventRecord(ctx->lanePoolEv, ts));
        region_moment_kernel<<<GS, 256, 0, ts>>>(sg);
        CU(cudaGetLastError());
        launches += 2;
        if ((r = scan_dev(ctx, LoadU32{pass2}, devn_u64(&gT[t]->n1), n1Cap, pass2Scan, &gT[t]->rows, &launches))) return r;
        PlaceArgs pa;
        memset(&pa, 0, sizeof pa);
        pa.cand = cand; pa.list = listT[t]; pa.pass2Scan = pass2Scan; pa.g = gT[t]; pa.call = call; pa.units = d_units;
        pa.u0 = t ? uMid : u0; pa.u1 = t ? u1 : uMid;
        pa.normalFirst = normalFirst; pa.rowShift = rowShift; pa.edgeRowBase = edgeRowBase; pa.unitRows = unitRowsDev;
        if (t == 1) CU(cudaStreamWaitEvent(ts, ctx->placeEv[0], 0));      // the second half's rows follow the first half's
        row_place_kernel<<<1, 256, 0, ts>>>(pa);
        if (t == 0) CU(cudaEventRecord(ctx->placeEv[0], ts));
        EmitArgs ea;
        memset(&ea, 0, sizeof ea);
        ea.cand = cand; ea.list = listT[t]; ea.pass = pass2; ea.passScan = pass2Scan; ea.g = gT[t]; ea.units = d_units; ea.unitFirst = unitFirst;
        ea.rowShift = rowShift; ea.statCounts = statCounts; ea.statKurt = statKurt; ea.nSamples = n; ea.rowCap = rowCap;
        ea.buffer = hs + SL.oBuffer; ea.contig = (uint32_t*)(hs + SL.oContig); ea.left = (uint32_t*)(hs + SL.oLeft); ea.len = (uint32_t*)(hs + SL.oLen);
        ea.peak = (uint32_t*)(hs + SL.oPeak); ea.wsum = (uint32_t*)(hs + SL.oWsum); ea.candIndex = (unsigned long long*)(hs + SL.oCandIndex);
        ea.peakScore = (double*)(hs + SL.oPeakScore); ea.kurt = (double*)(hs + SL.oKurt); ea.corr = (double*)(hs + SL.oCorr);
        ea.counts = (uint32_t*)(hs + SL.oCounts);
        emit_rows_kernel<<<GS, 256, 0, ts>>>(ea);
        emit_counts_kernel<<<GS, 256, 0, ts>>>(ea);
        GroupEndArgs ge;
        memset(&ge, 0, sizeof ge);
        ge.statCounts = statCounts; ge.pass = pass2; ge.g = gT[t]; ge.call = call; ge.nSamples = n; ge.tir = d_tir;
        group_end_kernel<<<(unsigned)ctx->smCount * 2, 256, (size_t)n * 8, ts>>>(ge);
        CU(cudaGetLastError());
        launches += 4;
      }
      ctx->lane = 0; ctx->cur = ctx->stream;
      CU(cudaEventRecord(ctx->laneDoneEv, ctx->laneStream));
      CU(cudaStreamWaitEvent(ctx->stream, ctx->laneDoneEv, 0));
      continue;
    }
    // per-survivor arrays: sized from the previous call, or (sizing) from this one's numbers
    // (a shared call cannot be repeated with larger arrays — the other ranks would wait for it — so it reads its sizes back unless an earlier
    //  call of this context has already sized the per-survivor arrays, margin included; outgrowing those is then an error, like outgrowing the table)
    const bool readSizes = sizing || cfg.needCorr || (shared && (ctx->capN1 == 0 || ctx->capCells == 0));
    if (readSizes) {
      compact_pv_kernel<<<GS, 256, 0, st_>>>(pass, pvScan, g, (uint32_t)runCap, ~0ull, list1);
      launches += 1;
      CU(cudaGetLastError());
      if ((r = read_group())) return r;
      if (hg->flags & FLAG_RUN_OVERFLOW) { ctx->needRunCap = (size_t)hg->runCount + (1u << 20); *retry = true; return UPB_OK; }
      const size_t n1 = (size_t)(hg->pv >> PACK_SHIFT), cells = (size_t)(hg->pv & PACK_MASK);
      ctx->capN1 = std::max(ctx->capN1, n1 + n1 / 4 + 1024);
      ctx->capCells = std::max(ctx->capCells, cells + cells / 4 + 4096);
    }
    const size_t n1Cap = std::min<size_t>(ctx->capN1, runCap), cellCap = ctx->capCells;
    uint32_t* statCounts = SCR("statCounts", uint32_t, n1Cap * n);
    double* statKurt = SCR("statKurt", double, n1Cap);
    uint32_t* pass2 = SCR("pass2", uint32_t, n1Cap);
    unsigned long long* pass2Scan = SCR("pass2Scan", unsigned long long, n1Cap);
    uint32_t* statCells = SCR("statCells", uint32_t, cellCap + 1);
    if (!readSizes) {
      compact_pv_kernel<<<GS, 256, 0, st_>>>(pass, pvScan, g, (uint32_t)n1Cap, cellCap, list1);
      launches += 1;
    }
    stage_mark(ctx, "order+stitch+hit filter");

    // ---- K3: per-region statistics of the survivors ---------------------------------------------------------------------
    StatArgs sg;
    sg.T = pr.T; sg.units = d_units; sg.cand = cand; sg.list = list1; sg.g = g;
    sg.nondir = cfg.nondir ? 1 : 0; sg.kurtThr = cfg.kurtThr; sg.counts = statCounts; sg.kurt = statKurt; sg.pass = pass2;
    sg.pvScan = pvScan; sg.cells = statCells; sg.cellCap = cellCap; sg.ticket = &g->poolTicket;
    {
      const size_t groupBytes = ((size_t)(ST_CHUNK + n) * 4 + 15) / 16 * 16;
      const size_t statSmem = (size_t)ST_WARPS * ST_PER_WARP * groupBytes;
      if (statSmem > 200 * 1024) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many samples for the region statistics kernel");
      int perSm = 0;
      if ((r = cached_occupancy(ctx, region_pool_kernel, ST_WARPS * 32, statSmem, &perSm))) return r;
      region_pool_kernel<<<(unsigned)(ctx->smCount * std::max(perSm, 1)), ST_WARPS * 32, statSmem, st_>>>(sg);
      stage_mark(ctx, "region_pool_kernel");
      region_moment_kernel<<<GS, 256, 0, st_>>>(sg);
      CU(cudaGetLastError());
      launches += 2;
      stage_mark(ctx, "region_moment_kernel");
    }

    // strand correlation for the regions that got this far (misc/peakcall.cpp:41 evaluates it last); -D only, sizes read back
    double* rowCorr = nullptr;
    if (cfg.needCorr) {
      const uint32_t n1 = (uint32_t)(hg->pv >> PACK_SHIFT);
      rowCorr = SCR("rowCorr", double, n1Cap);
      if (n1) {
        unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 16;
        if ((r = scan_u32(ctx, pass2, n1, pass2Scan, d_tot, &launches))) return r;
        uint32_t* list2 = SCR("list2", uint32_t, n1);
        compact_index_kernel<<<(n1 + 255) / 256, 256, 0, st_>>>(pass2, pass2Scan, n1, list2);
        launches += 1;
        unsigned long long* h64 = (unsigned long long*)ctx->h_small;
        CU(cudaMemcpyAsync(h64, d_tot, 8, cudaMemcpyDeviceToHost, st_));
        CU(cudaStreamSynchronize(st_));
        const uint32_t n2 = (uint32_t)h64[0];
        CU(cudaMemsetAsync(rowCorr, 0xFF, (size_t)n1 * 8, st_));      // NaN where not computed
        if (n2) {
          uint32_t* dUnit = SCR("dUnit", uint32_t, n2); uint32_t* dLeft = SCR("dLeft", uint32_t, n2);
          uint32_t* dLen = SCR("dLen", uint32_t, n2); uint32_t* dChunks = SCR("dChunks", uint32_t, n2);
          dump_desc2_kernel<<<(n2 + 255) / 256, 256, 0, st_>>>(cand, list1, list2, n2, dUnit, dLeft, dLen, dChunks);
          CU(cudaGetLastError());
          launches += 1;
          if ((r = dump_and_corr(ctx, pr.dens, cfg.nondir, n2, 1, &launches))) return r;
          cand_corr_kernel<<<(n2 + 255) / 256, 256, 0, st_>>>(list2, n2, ctx->scratch["corr"].as<double>(), cfg.corrThr, pass2, rowCorr);
          CU(cudaGetLastError());
          launches += 1;
        }
      }
      stage_mark(ctx, "strand correlation");
    }
    if ((r = scan_dev(ctx, LoadU32{pass2}, devn_u64(&g->n1), n1Cap, pass2Scan, &g->rows, &launches))) return r;

    // ---- where the rows go, then the rows themselves, straight into the (mapped) host table ---------------------------------
    PlaceArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.cand = cand; pa.list = list1; pa.pass2Scan = pass2Scan; pa.g = g; pa.call = call; pa.units = d_units; pa.u0 = u0; pa.u1 = u1;
    pa.edgeRows = d_edgeRowsOfUnit; pa.normalFirst = normalFirst; pa.rowShift = rowShift; pa.edgeRowBase = edgeRowBase; pa.unitRows = unitRowsDev;
    pa.shared = shared ? 1 : 0;
    const bool lastGroup = gi + 1 == pgroups.size();
    const bool sharedGroups = shared && pgroups.size() > 1;             // several groups into a shared table: only the last one publishes and places
    if (shared) { pa.x = shared->x; pa.countOnly = (sharedGroups && !lastGroup) ? 1 : 0; pa.allU0 = sharedGroups ? 0u : u0; pa.allU1 = sharedGroups ? nUnits : u1; }
    if (twoLanes && gi > 0) CU(cudaStreamWaitEvent(st_, ctx->placeEv[gi - 1], 0));   // rows of a group follow the rows of the groups before it
    row_place_kernel<<<1, 256, 0, st_>>>(pa);
    if (twoLanes || sharedGroups) CU(cudaEventRecord(ctx->placeEv[gi], st_));
    launches += 1;
    if (!shared && sizing) {
      if ((r = read_group())) return r;
      const size_t R = (size_t)hg->rows + edgeRowsTotal;
      ctx->capRows = std::max(ctx->capRows, R + R / 4 + 1024);
      rowCap = ctx->capRows;
      out.layout = slab_layout(rowCap, n, C);
      if ((r = acquire_slab(ctx, out.layout.totalBytes, out.slab))) return r;
      hs = (unsigned char*)out.slab.p;
    }
    const SlabLayout& SL = out.layout;
    EmitArgs ea;
    memset(&ea, 0, sizeof ea);
    ea.cand = cand; ea.list = list1; ea.pass = pass2; ea.passScan = pass2Scan; ea.g = g; ea.units = d_units; ea.unitFirst = unitFirst;
    ea.rowShift = rowShift; ea.affOfUnit = d_affOfUnit; ea.edgeCand = eg.nCand; ea.edgeDrop = d_drop;
    ea.statCounts = statCounts; ea.statKurt = statKurt; ea.rowCorr = rowCorr; ea.nSamples = n; ea.rowCap = rowCap;
    ea.buffer = hs + SL.oBuffer; ea.contig = (uint32_t*)(hs + SL.oContig); ea.left = (uint32_t*)(hs + SL.oLeft); ea.len = (uint32_t*)(hs + SL.oLen);
    ea.peak = (uint32_t*)(hs + SL.oPeak); ea.wsum = (uint32_t*)(hs + SL.oWsum); ea.candIndex = (unsigned long long*)(hs + SL.oCandIndex);
    ea.peakScore = (double*)(hs + SL.oPeakScore); ea.kurt = (double*)(hs + SL.oKurt); ea.corr = (double*)(hs + SL.oCorr);
    ea.counts = (uint32_t*)(hs + SL.oCounts);
    GroupEndArgs ge;
    memset(&ge, 0, sizeof ge);
    ge.statCounts = statCounts; ge.pass = pass2; ge.g = g; ge.call = call; ge.nSamples = n;
    ge.edgeRowCounts = eg.rowCounts; ge.edgeNRows = eg.nRows; ge.edgeRowCap = EDGE_ROW_CAP; ge.nAffected = edgeRowsTotal ? nAff : 0; ge.edgeRowsTotal = edgeRowsTotal;
    ge.tir = d_tir;
    if (sharedGroups && !lastGroup) {                              // its rows have no place yet: emitted behind the last group's row_place_kernel
      Deferred df; df.ea = ea; df.ge = ge; df.st = st_;
      deferred.push_back(df);
      continue;
    }
    emit_rows_kernel<<<GS, 256, 0, st_>>>(ea);
    emit_counts_kernel<<<GS, 256, 0, st_>>>(ea);
    launches += 2;
    if (edgeRowsTotal) {
      EdgeEmitArgs ee;
      memset(&ee, 0, sizeof ee);
      ee.rows = eg.rows; ee.rowCounts = eg.rowCounts; ee.rowCap = EDGE_ROW_CAP; ee.nSamples = n; ee.nAffected = nAff; ee.affected = d_affected; ee.nRows = eg.nRows;
      ee.edgeRowBase = edgeRowBase; ee.units = d_units; ee.tableCap = rowCap; ee.g = g;
      ee.buffer = ea.buffer; ee.contig = ea.contig; ee.left = ea.left; ee.len = ea.len; ee.peak = ea.peak; ee.candIndex = ea.candIndex;
      ee.peakScore = ea.peakScore; ee.kurt = ea.kurt; ee.corr = ea.corr; ee.counts = ea.counts; ee.wsum = ea.wsum;
      edge_emit_kernel<<<dim3((EDGE_ROW_CAP + 127) / 128, nAff), 128, 0, st_>>>(ee);
      launches += 1;
    }
    group_end_kernel<<<(unsigned)ctx->smCount * 2, 256, (size_t)n * 8, st_>>>(ge);
    CU(cudaGetLastError());
    launches += 1;
    stage_mark(ctx, "place+emit rows (to host)+totals");
    if (trace && gi < ctx->copyEvs2.size()) cudaEventRecord(ctx->copyEvs2[gi], st_);
  }
  for (const Deferred& df : deferred) {
    CU(cudaStreamWaitEvent(df.st, ctx->placeEv[pgroups.size() - 1], 0));
    emit_rows_kernel<<<GS, 256, 0, df.st>>>(df.ea);
    emit_counts_kernel<<<GS, 256, 0, df.st>>>(df.ea);
    group_end_kernel<<<(unsigned)ctx->smCount * 2, 256, (size_t)n * 8, df.st>>>(df.ge);
    CU(cudaGetLastError());
    launches += 3;
  }
  ctx->lane = 0; ctx->cur = ctx->stream;
  if (twoLanes) { CU(cudaEventRecord(ctx->laneDoneEv, ctx->laneStream)); CU(cudaStreamWaitEvent(ctx->stream, ctx->laneDoneEv, 0)); }
  if (empty) {                                                          // a rank of a shared table with nothing loaded still takes part in the exchange
    CU(cudaMemsetAsync(g, 0, sizeof(GroupCtr), st_));
    PlaceArgs pa;
    memset(&pa, 0, sizeof pa);
    pa.g = g; pa.call = call; pa.units = d_units; pa.normalFirst = normalFirst; pa.rowShift = rowShift; pa.edgeRowBase = edgeRowBase; pa.unitRows = unitRowsDev;
    pa.shared = 1; pa.x = shared->x;
    row_place_kernel<<<1, 256, 0, st_>>>(pa);
    launches += 1;
  }

  // ---- the one read-back: counters, totals, per-unit numbers ----------------------------------------------------------
  CU(cudaMemcpyAsync(hg, g, sizeof(GroupCtr), cudaMemcpyDeviceToHost, st_));
  CU(cudaMemcpyAsync(hcall, call, sizeof(CallCtr), cudaMemcpyDeviceToHost, st_));
  CU(cudaMemcpyAsync(hTir, d_tir, (size_t)n * 8, cudaMemcpyDeviceToHost, st_));
  if (nUnits) {
    CU(cudaMemcpyAsync(hUnitCand, unitCand, (size_t)nUnits * 8, cudaMemcpyDeviceToHost, st_));
    CU(cudaMemcpyAsync(hUnitRows, unitRowsDev, (size_t)nUnits * 8, cudaMemcpyDeviceToHost, st_));
  }
  if (nAff) CU(cudaMemcpyAsync(hDrop, d_drop, (size_t)nAff * 4, cudaMemcpyDeviceToHost, st_));
  CU(cudaEventRecord(ctx->ev[3], st_));
  if (!shared) memset(hs + out.layout.devBytes, 0, out.layout.totalBytes - out.layout.devBytes);   // host-only tail of the slab, while the device works
  const Clk::time_point tC = Clk::now();
  CU(cudaEventSynchronize(ctx->ev[3]));
  const Clk::time_point tD = Clk::now();
  { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); ctx->stats.ms_total = ms; }
  { auto ms = [] (Clk::time_point a, Clk::time_point b) { return std::chrono::duration<float, std::milli>(b - a).count(); };
    ctx->stats.ms_host_prepare = ms(tA, tB); ctx->stats.ms_host_enqueue = ms(tB, tC) - (float)waitMid; ctx->stats.ms_host_wait = (float)waitMid + ms(tC, tD);
    if (ctx->stats.ms_host_enqueue > 2.f && getenv("UPB_TRACE_SLOW")) {   // a call whose launches took unusually long to issue: where
      fprintf(stderr, "[upb] slow enqueue on device %d, call %llu%s: %.3f ms (mid-call waits %.3f ms);", ctx->device, (unsigned long long)ctx->callNo, shared ? " (shared)" : "", ctx->stats.ms_host_enqueue, waitMid);
      Clk::time_point prev = tB;
      for (const auto& hm : ctx->hostMarks) { fprintf(stderr, " %s +%.3f", hm.first, ms(prev, hm.second)); prev = hm.second; }
      fprintf(stderr, " end +%.3f\n", ms(prev, tC));
    } }
  if (!empty) {                                                          // span kernel time: the groups' launches added up (two lanes: each bracket may hold a share of the other lane's kernels)
    float sum = 0;
    for (size_t gi = 0; gi < pgroups.size(); ++gi) { float ms = 0; cudaEventElapsedTime(&ms, ctx->spanEv[2 * gi], ctx->spanEv[2 * gi + 1]); sum += ms; }
    ctx->stats.ms_kde = sum;
  }
  stage_collect(ctx);
  ctx->stats.launches = launches;
  if (pipelined) { const int rl = load_finish(ctx); if (rl) return rl; }   // what the load had to say about the input
  if (trace && pipelined) for (size_t gi = 0; gi < pgroups.size() && gi < ctx->copyEvs2.size(); ++gi) {
    float a = -1, b = -1;
    cudaEventElapsedTime(&a, ctx->copyEvs[ctx->groups.size()], ctx->groupEv[pgroups[gi].ev0]);
    cudaEventElapsedTime(&b, ctx->copyEvs[ctx->groups.size()], ctx->copyEvs2[gi]);
    fprintf(stderr, "[upb]   units %2u..%2u  index ready at %7.3f ms  regions called at %7.3f ms\n", pgroups[gi].u0, pgroups[gi].u1 - 1, a, b);
  }
  const uint32_t flags = hcall->flags | hg->flags;
  if (flags & FLAG_XCHG_TIMEOUT) return fail(ctx, UPB_ERR_CUDA, "shared table: another rank never published its row counts (timeout)");
  if (flags & (FLAG_RUN_OVERFLOW | FLAG_CAP_OVERFLOW)) {
    if (shared) {
      ctx->capN1 = 0; ctx->capCells = 0;                                // the next shared call of this context reads its sizes back instead of trusting the old ones
      return fail(ctx, UPB_ERR_NOMEM, "shared table: this call needs more room than the table or the arrays sized by the previous call have (call upb_call_regions once per context before sharing)");
    }
    if (out.slab.p) { ctx->slabPool.push_back(out.slab); out.slab = HostSlab(); }
    if (flags & FLAG_RUN_OVERFLOW) ctx->needRunCap = (size_t)std::max<unsigned long long>(hcall->needSlots, hg->runCount) + (1u << 20);
    *retry = true;
    return UPB_OK;
  }
  ctx->stats.warm_spans = hcall->warm;
  ctx->stats.hot_blocks = hcall->stat[0]; ctx->stats.kde_hot_tiles = hcall->stat[1]; ctx->stats.exact_evals = hcall->stat[2]; ctx->stats.spans_with_tags = hcall->stat[3];
  out.R = hcall->rows;
  out.nCand = hcall->cands;
  ctx->stats.n_stitched = hcall->runs - hcall->cands;
  for (uint32_t u = 0; u < nUnits; ++u) out.unitCand[u] = hUnitCand[u];
  out.unitRows.assign(hUnitRows, hUnitRows + nUnits);
  for (uint32_t k = 0; k < nAff; ++k) {                                 // the replay's candidates replace the ones it dropped
    const long long d = (long long)edgeCand[k] - (long long)hDrop[k];
    out.unitCand[pr.affected[k]] = (uint64_t)((long long)out.unitCand[pr.affected[k]] + d);
    out.nCand = (uint64_t)((long long)out.nCand + d);
  }
  ctx->stats.n_candidates = out.nCand;
  {
    // what the next call of this context may assume without asking
    const size_t n1 = (size_t)hcall->n1, cells = (size_t)hcall->cells, R = (size_t)out.R;
    ctx->lastN1 = n1;
    ctx->capN1 = std::max(ctx->capN1, n1 + n1 / 4 + 1024);
    ctx->capCells = std::max(ctx->capCells, cells + cells / 4 + 4096);
    if (!shared) {
      memcpy(hs + out.layout.oTags, hTir, (size_t)n * 8);
      ctx->capRows = std::max(ctx->capRows, R + R / 4 + 1024);
    }
  }
  out.tir.assign(hTir, hTir + n);
  ctx->stats.ms_host_post = std::chrono::duration<float, std::milli>(Clk::now() - tD).count();
  return UPB_OK;
}

static int run_pipeline (upb_ctx* ctx, const PipeCfg& cfg, Prep& pr, PipeOut& out, const SharedDst* shared = nullptr) {
  // the first call of a context sizes its arrays from its own numbers; later calls run without a read-back and are repeated with
  // sizing only when the numbers outgrew the arrays
  bool sizing = !shared && (ctx->capN1 == 0 || ctx->capRows == 0);
  for (int attempt = 0; attempt < 4; ++attempt) {
    bool retry = false;
    const int r = run_pipeline_once(ctx, cfg, pr, out, sizing, &retry, shared);
    if (r) { if (out.slab.p) { ctx->slabPool.push_back(out.slab); out.slab = HostSlab(); } return r; }
    if (!retry) return UPB_OK;
    sizing = true;
  }
  return fail(ctx, UPB_ERR_NOMEM, "region pipeline: buffers still too small after resizing");
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

// rows of a table into the order `perm` gives (row k of the result = old row perm[k])
static void permute_table (upb_region_table* out, const std::vector<uint64_t>& perm, uint32_t n) {
  const uint64_t R = out->n_rows;
  auto permute = [&] (auto* col, size_t width) {
    typedef typename std::remove_pointer<decltype(col)>::type T;
    std::vector<T> tmp(col, col + R * width);
    for (uint64_t k = 0; k < R; ++k) for (size_t w = 0; w < width; ++w) col[k * width + w] = tmp[perm[k] * width + w];
  };
  permute(out->buffer, 1); permute(out->forward_label, 1); permute(out->dropped, 1); permute(out->contig, 1); permute(out->left, 1);
  permute(out->len, 1); permute(out->peak, 1); permute(out->hit_sum, 1); permute(out->cand_index, 1); permute(out->peak_score, 1);
  permute(out->kurtosis, 1); permute(out->corr, 1); permute(out->counts, n);
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
    if (r) { upb_free_region_table(out); return fail(ctx, r, "several directional samples and a '-' track that the reference's merge loop reads during pass 1: it then aborts (assert) or pools that sample's '-' tags separately, neither is reproduced (SURVEY Q7); reference_order = 0 returns the canonical table"); }
    if (!identity) permute_table(out, perm, n);
  }
  ctx->stats.ms_call = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// several GPUs: one table every rank writes its rows into ("final region list — gather rows to rank 0")
//
// One rank per GPU (threads of one process: upb_multi_*; or processes: bench.py under torchrun), contigs sharded, no data-path
// exchange.  The table lives in ONE page-locked host segment that every rank maps.  Each rank's row_place_kernel publishes its
// per-(buffer, contig) row counts there and waits — on the device — for the other ranks' counts, so its emit_rows_kernel can write
// its rows straight to their place in the table-wide order over its own PCIe link.  Rank 0's host then only waits for the other
// ranks' "done" words.  No collective and no host round trip inside a step (r01: NCCL all-reduce + .cpu() + scatter + all-reduce).
//
// segment layout:  ready[world] u64 | done[world] u64 | counts[2][world][2C] u64 | per-rank blocks | table slab (slab_layout)
// per-rank block:  nCand u64 | unitCand[2C] u64 | tir[n] u64 | lastTag[2C] u32 | firstPos[n*2*C] u32 | lastPos[n*2*C] u32
// ------------------------------------------------------------------------------------------------
struct SharedLayout { size_t oReady, oDone, oCounts, oRank, rankBytes, oSlab, total; };

static SharedLayout shared_layout (uint32_t C, uint32_t n, uint32_t world, uint64_t rowCap) {
  SharedLayout L;
  size_t o = 0;
  L.oReady = o; o += (size_t)world * 8;
  L.oDone = o;  o += (size_t)world * 8;
  o = (o + 255) & ~(size_t)255;
  L.oCounts = o; o += 2 * (size_t)world * 2 * C * 8;
  o = (o + 255) & ~(size_t)255;
  L.rankBytes = ((1 + 2 * (size_t)C + n) * 8 + (2 * (size_t)C + 2 * (size_t)n * 2 * C) * 4 + 255) & ~(size_t)255;
  L.oRank = o; o += (size_t)world * L.rankBytes;
  o = (o + 4095) & ~(size_t)4095;
  L.oSlab = o; o += slab_layout((size_t)rowCap, n, C).totalBytes;
  L.total = (o + 4095) & ~(size_t)4095;
  return L;
}

extern "C" uint64_t upb_shared_table_bytes (uint32_t n_contigs, uint32_t n_samples, uint32_t world, uint64_t row_capacity) {
  return shared_layout(n_contigs, n_samples, world, row_capacity).total;
}

extern "C" int upb_shared_table_init (const upb_shared_table* t) {
  if (!t || !t->host || !t->world) return UPB_ERR_BAD_ARG;
  const SharedLayout L = shared_layout(t->n_contigs, t->n_samples, t->world, t->row_capacity);
  if (t->bytes < L.total) return UPB_ERR_BAD_ARG;
  memset(t->host, 0, L.oSlab);
  return UPB_OK;
}

static void fill_table (upb_ctx* ctx, const PipeOut& po, upb_region_table* t);

extern "C" int upb_call_regions_shared (upb_ctx* ctx, const upb_region_params* p, const upb_shared_table* t, uint64_t step, upb_region_table* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !t || !out || !t->host || !t->device || !t->world || t->rank >= t->world || step == 0) return fail(ctx, UPB_ERR_BAD_ARG, "upb_call_regions_shared: bad argument");
  memset(out, 0, sizeof *out);
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  const uint32_t C = (uint32_t)ctx->sizes.size(), n = ctx->nSamples;
  if (C != t->n_contigs || n != t->n_samples) return fail(ctx, UPB_ERR_BAD_ARG, "upb_call_regions_shared: the table was laid out for another contig table / sample count");
  const SharedLayout L = shared_layout(C, n, t->world, t->row_capacity);
  if (t->bytes < L.total) return fail(ctx, UPB_ERR_BAD_ARG, "upb_call_regions_shared: segment smaller than upb_shared_table_bytes()");
  unsigned char* hb = (unsigned char*)t->host;
  unsigned char* db = (unsigned char*)t->device;
  PipeCfg cfg;
  cfg.bw = p->bandwidth; cfg.thr = p->region_threshold; cfg.kurtThr = p->kurtosis_threshold; cfg.corrThr = p->corr_threshold;
  cfg.hitThr = p->hit_threshold; cfg.background = p->background; cfg.nondir = !p->directional; cfg.control = p->control;
  cfg.nCoeffs = p->n_coeffs; cfg.coeffs = p->coeffs;
  cfg.needCorr = cfg.nondir;
  SharedDst sd;
  sd.slab = db + L.oSlab; sd.rowCap = t->row_capacity;
  sd.x.counts = (volatile unsigned long long*)(db + L.oCounts + (size_t)(step & 1) * t->world * 2 * C * 8);
  sd.x.ready = (volatile unsigned long long*)(db + L.oReady);
  sd.x.rank = t->rank; sd.x.world = t->world; sd.x.C = C; sd.x.step = step;
  PipeOut po;
  Prep pr;
  int r = run_pipeline(ctx, cfg, pr, po, &sd);
  const std::chrono::steady_clock::time_point tAsm = std::chrono::steady_clock::now();
  unsigned long long* done = (unsigned long long*)(hb + L.oDone);
  if (r) { __atomic_store_n(&done[t->rank], ~0ull, __ATOMIC_RELEASE); return r; }   // poison: rank 0 reports the step as failed instead of waiting
  // this rank's contribution to what rank 0 assembles
  {
    unsigned char* rb = hb + L.oRank + (size_t)t->rank * L.rankBytes;
    unsigned long long* w = (unsigned long long*)rb;
    w[0] = po.nCand;
    unsigned long long* uc = w + 1; unsigned long long* tir = uc + 2 * (size_t)C;
    uint32_t* lastTag = (uint32_t*)(tir + n); uint32_t* fp = lastTag + 2 * (size_t)C; uint32_t* lp = fp + (size_t)n * 2 * C;
    memset(uc, 0, 2 * (size_t)C * 8);
    for (size_t u = 0; u < po.units.size(); ++u) uc[(size_t)po.units[u].buffer * C + po.units[u].contig] = po.unitCand[u];
    for (uint32_t s = 0; s < n; ++s) tir[s] = s < po.tir.size() ? po.tir[s] : 0;
    memset(lastTag, 0, 2 * (size_t)C * 4);
    for (uint32_t s = 0; s < n; ++s) for (uint32_t d = 0; d < 2; ++d) for (uint32_t c = 0; c < C; ++c) {
      const uint32_t lpv = ctx->lastPos[((size_t)s * 2 + d) * C + c];
      const uint32_t b = p->directional ? d : 0;
      if (lpv > lastTag[(size_t)b * C + c]) lastTag[(size_t)b * C + c] = lpv;
    }
    memcpy(fp, ctx->firstPos.data(), (size_t)n * 2 * C * 4);
    memcpy(lp, ctx->lastPos.data(), (size_t)n * 2 * C * 4);
  }
  out->n_samples = n;
  out->n_rows = po.R;                                                  // this rank's rows
  __atomic_store_n(&done[t->rank], step, __ATOMIC_RELEASE);
  if (t->rank != 0) { ctx->stats.ms_call = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count(); return UPB_OK; }

  // rank 0: every rank's rows have landed once its `done` word says so
  for (uint32_t rk = 0; rk < t->world; ++rk) {
    const std::chrono::steady_clock::time_point w0 = std::chrono::steady_clock::now();
    for (uint64_t spin = 0;; ++spin) {
      const unsigned long long v = __atomic_load_n(&done[rk], __ATOMIC_ACQUIRE);
      if (v == ~0ull) return fail(ctx, UPB_ERR_CUDA, "shared table: a rank reported a failed step");
      if (v >= step) break;
      if ((spin & 1023) == 1023 && std::chrono::duration<double>(std::chrono::steady_clock::now() - w0).count() > 30.0)
        return fail(ctx, UPB_ERR_CUDA, "shared table: timed out waiting for a rank");
    }
  }
  const SlabLayout SL = slab_layout((size_t)t->row_capacity, n, C);
  unsigned char* hs = hb + L.oSlab;
  const unsigned long long* counts = (const unsigned long long*)(hb + L.oCounts + (size_t)(step & 1) * t->world * 2 * C * 8);
  uint64_t R = 0;
  memset(hs + SL.devBytes, 0, SL.totalBytes - SL.devBytes);
  memset(out, 0, sizeof *out);
  out->n_samples = n;
  out->buffer = hs + SL.oBuffer; out->forward_label = hs + SL.oLabel; out->dropped = hs + SL.oDropped;
  out->contig = (uint32_t*)(hs + SL.oContig); out->left = (uint32_t*)(hs + SL.oLeft); out->len = (uint32_t*)(hs + SL.oLen); out->peak = (uint32_t*)(hs + SL.oPeak);
  out->hit_sum = (uint32_t*)(hs + SL.oWsum); out->cand_index = (uint64_t*)(hs + SL.oCandIndex); out->peak_score = (double*)(hs + SL.oPeakScore);
  out->kurtosis = (double*)(hs + SL.oKurt); out->corr = (double*)(hs + SL.oCorr); out->counts = (uint32_t*)(hs + SL.oCounts);
  out->tags_in_regions = (uint64_t*)(hs + SL.oTags); out->unit_candidates = (uint64_t*)(hs + SL.oUnitCand);
  out->unit_last_tag = (uint32_t*)(hs + SL.oUnitLast); out->unit_rows = (uint64_t*)(hs + SL.oUnitRows);
  std::vector<uint32_t> firstPos((size_t)n * 2 * C, 0), lastPos((size_t)n * 2 * C, 0);
  uint64_t nCand = 0;
  for (uint32_t rk = 0; rk < t->world; ++rk) {
    const unsigned char* rb = hb + L.oRank + (size_t)rk * L.rankBytes;
    const unsigned long long* w = (const unsigned long long*)rb;
    nCand += w[0];
    const unsigned long long* uc = w + 1; const unsigned long long* tir = uc + 2 * (size_t)C;
    const uint32_t* lastTag = (const uint32_t*)(tir + n); const uint32_t* fp = lastTag + 2 * (size_t)C; const uint32_t* lp = fp + (size_t)n * 2 * C;
    for (size_t k = 0; k < 2 * (size_t)C; ++k) {
      out->unit_candidates[k] += uc[k];
      out->unit_rows[k] += counts[(size_t)rk * 2 * C + k];
      if (lastTag[k] > out->unit_last_tag[k]) out->unit_last_tag[k] = lastTag[k];
    }
    for (uint32_t s = 0; s < n; ++s) out->tags_in_regions[s] += tir[s];
    for (size_t k = 0; k < (size_t)n * 2 * C; ++k) { if (fp[k]) firstPos[k] = fp[k]; if (lp[k]) lastPos[k] = lp[k]; }
  }
  for (size_t k = 0; k < 2 * (size_t)C; ++k) R += out->unit_rows[k];
  if (R > t->row_capacity) return fail(ctx, UPB_ERR_NOMEM, "shared table: more rows than the table has room for");
  // A contig whose every tag lies within `bandwidth` of its start hands its window to the next contig of the buffer (SURVEY Q2,
  // upb_edge.cuh "Chains"): a rank replays that only if it holds both contigs
  for (uint32_t b = 0; b < (p->directional ? 2u : 1u); ++b) {
    int leakRank = -1;
    for (uint32_t c = 0; c < C; ++c) {
      uint32_t last = 0; int rankOf = -1;
      for (uint32_t rk = 0; rk < t->world; ++rk) {
        const unsigned char* rb = hb + L.oRank + (size_t)rk * L.rankBytes;
        const uint32_t* lp = (const uint32_t*)((const unsigned long long*)rb + 1 + 2 * (size_t)C + n) + 2 * (size_t)C + (size_t)n * 2 * C;
        for (uint32_t s = 0; s < n; ++s) for (uint32_t d = (p->directional ? b : 0u); d <= (p->directional ? b : 1u); ++d) {
          const uint32_t v = lp[((size_t)s * 2 + d) * C + c];
          if (v) { last = std::max(last, v); rankOf = (int)rk; }
        }
      }
      if (rankOf < 0) continue;
      if (leakRank >= 0 && leakRank != rankOf)
        return fail(ctx, UPB_ERR_UNSUPPORTED, "every tag of a contig lies within `bandwidth` of its start and the next contig of that buffer is loaded on another rank: the reference's window travels between them (SURVEY Q2); load both on one rank");
      leakRank = last <= p->bandwidth ? rankOf : -1;
    }
  }
  out->n_rows = R; out->n_passed = R; out->n_rejected = nCand - R;
  { uint64_t firstRev = 0; for (size_t k = 0; k < C; ++k) firstRev += out->unit_rows[k];
    if (!p->directional) firstRev = R;
    memset(out->forward_label, 1, firstRev); memset(out->forward_label + firstRev, 0, R - firstRev); }
  if (p->reference_order == 1) {
    std::vector<uint64_t> perm;
    bool identity = true;
    r = reference_order_impl(C, n, p->directional, p->bandwidth, firstPos.data(), lastPos.data(), out, &perm, &identity);
    if (r) return fail(ctx, r, "several directional samples and a '-' track that the reference's merge loop reads during pass 1: it then aborts (assert) or pools that sample's '-' tags separately, neither is reproduced (SURVEY Q7); reference_order = 0 returns the canonical table");
    if (!identity) permute_table(out, perm, n);
  }
  ctx->stats.ms_call = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - t0).count();
  ctx->stats.ms_host_assemble = std::chrono::duration<float, std::milli>(std::chrono::steady_clock::now() - tAsm).count();
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
  return cfg;
}

// :206-219 for a batch of regions that are long enough: strandCorr(shift) for every shift 0..max_shift on the device, then the
// first strict maximum per region.  `dens` is the density template of a prepare() / run_pipeline() with shift_cfg.
static int shift_sweep_batch (upb_ctx* ctx, const PipeCfg& cfg, const Prep& pr, const std::vector<uint32_t>& unitOfContig, uint32_t nb, const uint32_t* contig,
                              const uint32_t* left, const uint32_t* len, uint32_t nShifts, uint32_t* bestShift, double* bestCorr, uint32_t* launches) {
  if (!nb) return UPB_OK;
  const DensArgs& dens = pr.dens;
  // regions that the left-edge replay produced are swept over the score vectors the reference's Region holds for them —
  // the displaced contents of its unshifted window — which only the replay knows: run it again, keeping the vectors of the long rows
  std::vector<VecPatch> patches;
  EdgeRun er;
  {
    std::vector<int32_t> affOfUnit(pr.units.size(), -1);
    for (size_t k = 0; k < pr.affected.size(); ++k) affOfUnit[pr.affected[k]] = (int32_t)k;
    bool any = false;
    for (uint32_t k = 0; k < nb && !any; ++k) any = affOfUnit[unitOfContig[contig[k]]] >= 0;
    if (any) {
      EdgeRec rec; rec.vecMinLen = 2 * (nShifts - 1) + 3;
      int r = run_edge_replay(ctx, cfg, pr, ctx->stream, rec, er, launches);
      if (r) return r;
      const uint32_t nAff = (uint32_t)pr.affected.size();
      if (er.out[3 * nAff + 2] > er.eg.vecCap) return fail(ctx, UPB_ERR_UNSUPPORTED, "left-edge replay (SURVEY Q2): score-vector pool too small");
      std::vector<EdgeRow> rows;
      for (uint32_t k = 0; k < nb; ++k) {
        const int32_t a = affOfUnit[unitOfContig[contig[k]]];
        if (a < 0 || left[k] > er.out[a]) continue;                   // beyond the position the replay hands over at
        const uint32_t nRows = std::min<uint32_t>(er.out[2 * nAff + a], EDGE_ROW_CAP);
        rows.resize(nRows);
        if (nRows) CU(cudaMemcpy(rows.data(), er.eg.rows + (size_t)a * EDGE_ROW_CAP, (size_t)nRows * sizeof(EdgeRow), cudaMemcpyDeviceToHost));
        bool found = false;
        for (const EdgeRow& row : rows) if (row.left == left[k] && row.len == len[k] && row.vecOff) { patches.push_back(VecPatch{k, row.vecOff - 1, len[k]}); found = true; break; }
        if (!found) return fail(ctx, UPB_ERR_BAD_ARG, "shift sweep: a region at a contig start is not one the left-edge replay (SURVEY Q2) produces");
      }
    }
  }
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
  if ((r = dump_and_corr(ctx, dens, true, nb, nShifts, launches, &patches, er.eg.vecF, er.eg.vecR))) return r;
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
  if (!out->shift_freq || !out->tested_contig || !out->tested_left || !out->tested_len || !out->tested_best_shift || !out->tested_best_corr) {
    upb_free_shift_report(out);
    return fail(ctx, UPB_ERR_NOMEM, "upb_strand_shift: out of host memory");
  }

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
        batch.push_back(i);                                            // (a region of the left-edge replay is swept over the replay's own score vectors, shift_sweep_batch)
      }
      ++cursor;
    }
    if (batch.empty()) break;
    const uint32_t nb = (uint32_t)batch.size();
    std::vector<uint32_t> hContig(nb), hLeft(nb), hLen(nb), bs(nb);
    std::vector<double> bc(nb);
    for (uint32_t k = 0; k < nb; ++k) { const uint32_t i = batch[k]; hContig[k] = rContig[i]; hLeft[k] = rLeft[i]; hLen[k] = rLen[i]; }
    if ((r = shift_sweep_batch(ctx, cfg, pr, unitOfContig, nb, hContig.data(), hLeft.data(), hLen.data(), nShifts, bs.data(), bc.data(), &launches))) { upb_free_shift_report(out); return r; }
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
  int r = prepare(ctx, cfg, pr, &launches);
  if (r) return r;
  const uint32_t C = (uint32_t)ctx->sizes.size();
  std::vector<uint32_t> unitOfContig(C, 0xFFFFFFFFu);
  for (size_t u = 0; u < pr.units.size(); ++u) unitOfContig[pr.units[u].contig] = (uint32_t)u;
  for (uint64_t k = 0; k < n; ++k) {
    if (contig[k] >= C || unitOfContig[contig[k]] == 0xFFFFFFFFu) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_sweep: a region lies on a contig without tags on this context");
    if (len[k] <= 2 * p->max_shift + 3) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_sweep: a region is too short to try all shifts (src/strand_shift.cpp:206)");
    if (left[k] < 1 || (uint64_t)left[k] + len[k] - 1 > (uint64_t)ctx->sizes[contig[k]] + p->bandwidth) return fail(ctx, UPB_ERR_BAD_ARG, "upb_shift_sweep: a region lies outside its contig");
  }
  r = shift_sweep_batch(ctx, cfg, pr, unitOfContig, (uint32_t)n, contig, left, len, p->max_shift + 1, best_shift_out, best_corr_out, &launches);
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
  std::vector<uint32_t> posOut;
  std::vector<double> scoreOut;
  // a unit with a tag within `bandwidth` of the contig start (or one that inherits stale window cells): what the reference writes for its first
  // positions are the displaced scores of its unshifted window — they come from the replay, up to the position it hands over at
  uint64_t first = 1;
  for (size_t k = 0; k < pr.affected.size(); ++k) if (pr.affected[k] == (uint32_t)unit) {
    EdgeRec rec; rec.profUnit = unit;
    EdgeRun er;
    if ((r = run_edge_replay(ctx, cfg, pr, ctx->stream, rec, er, &launches))) return r;
    if (er.profN) {
      posOut.resize(er.profN); scoreOut.resize(er.profN);
      CU(cudaMemcpyAsync(posOut.data(), er.eg.profPos, (size_t)er.profN * 4, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaMemcpyAsync(scoreOut.data(), er.eg.profScore, (size_t)er.profN * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
    first = (uint64_t)er.out[k] + 1;                                  // 2^32 if the replay owns the whole unit
  }
  // evaluate the contig in slabs of 8M positions; the reference evaluates firstTag-bw .. lastTag+bw (Q20)
  const uint64_t last = (uint64_t)ctx->sizes[contig] + cfg.bw;
  const uint32_t SLAB = 8u << 20;
  for (uint64_t start = first; start <= last; start += SLAB) {
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
  { const int rl = load_finish(ctx); if (rl) return rl; }
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
  if (((uint32_t*)(h64 + 8))[0] & FLAG_BAD_POS) { upb_free_tag_counts(out); return fail(ctx, UPB_ERR_BAD_ARG, "upb_aggregate_tags: strand/contig key or position out of range"); }
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

