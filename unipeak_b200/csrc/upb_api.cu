// upb_api.cu — C-ABI (include/unipeak_b200.h) and host orchestration of the sm_100a kernels.
//
// There is no CPU implementation behind these entry points: every result row comes from the kernels in
// upb_device.cuh.  The host side only sizes buffers, launches, and assembles the reference-shaped
// result tables (src/regions.cpp, src/strand_shift.cpp, src/tags_in_regions.cpp drivers).
#include "../../include/unipeak_b200.h"
#include "upb_device.cuh"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
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

struct upb_ctx {
  int device = 0;
  int smCount = 148;
  cudaStream_t stream = nullptr;
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
  DevBuf d_pos, d_cnt, d_cntEff, d_off, d_segStart, d_tidx, d_tidxBase, d_sizes, d_nTilesIdx;
  // scratch (grow-only)
  DevBuf d_units, d_taps, d_src, d_control, d_small, d_runs, d_tileBase, d_tileCount, d_tileOrd, d_scanPartial,
         d_ordRuns, d_ordUnit, d_head, d_headScan, d_cand, d_pass, d_passScan, d_list1, d_counts, d_offsum, d_kurt,
         d_pass2, d_dUnit, d_dLeft, d_dLen, d_dChunks, d_dOff, d_dChunkStart, d_F, d_R, d_corr, d_candOut, d_bestShift, d_bestCorr,
         d_pass2Scan, d_list2;
  upb_stats stats;
  std::vector<uint8_t> lastControl;     // control set cntEff was built for
  bool cntEffValid = false;
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
  memset(&ctx->stats, 0, sizeof ctx->stats);
  *out = ctx;
  return UPB_OK;
}

extern "C" void upb_destroy (upb_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  DevBuf* all[] = {&ctx->d_pos, &ctx->d_cnt, &ctx->d_cntEff, &ctx->d_off, &ctx->d_segStart, &ctx->d_tidx, &ctx->d_tidxBase, &ctx->d_sizes,
    &ctx->d_nTilesIdx, &ctx->d_units, &ctx->d_taps, &ctx->d_src, &ctx->d_control, &ctx->d_small, &ctx->d_runs, &ctx->d_tileBase, &ctx->d_tileCount,
    &ctx->d_tileOrd, &ctx->d_scanPartial, &ctx->d_ordRuns, &ctx->d_ordUnit, &ctx->d_head, &ctx->d_headScan, &ctx->d_cand, &ctx->d_pass,
    &ctx->d_passScan, &ctx->d_list1, &ctx->d_counts, &ctx->d_offsum, &ctx->d_kurt, &ctx->d_pass2, &ctx->d_dUnit, &ctx->d_dLeft, &ctx->d_dLen,
    &ctx->d_dChunks, &ctx->d_dOff, &ctx->d_dChunkStart, &ctx->d_F, &ctx->d_R, &ctx->d_corr, &ctx->d_candOut, &ctx->d_bestShift, &ctx->d_bestCorr,
    &ctx->d_pass2Scan, &ctx->d_list2};
  for (DevBuf* b : all) release(*b);
  for (int i = 0; i < 4; ++i) if (ctx->ev[i]) cudaEventDestroy(ctx->ev[i]);
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
  T.cntEff = useEff ? ctx->d_cntEff.as<uint32_t>() : ctx->d_cnt.as<uint32_t>();
  T.off = ctx->d_off.as<uint64_t>();
  T.tidx = ctx->d_tidx.as<uint32_t>();
  T.tidxBase = ctx->d_tidxBase.as<uint32_t>();
  T.tidxStride = ctx->tidxBase.back();
  T.C = (uint32_t)ctx->sizes.size();
  T.nSamples = ctx->nSamples;
  return T;
}

extern "C" int upb_load_tags (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt) {
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
  if (E && (!pos || !cnt)) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: pos/cnt are NULL");
  ctx->nSamples = n_samples;
  ctx->nEntries = E;
  ctx->off.assign(off, off + nOff);
  ctx->firstPos.assign(nSeg, 0);
  ctx->lastPos.assign(nSeg, 0);
  for (uint32_t g = 0; g < nSeg; ++g)
    if (segStart[g + 1] > segStart[g]) { ctx->firstPos[g] = pos[segStart[g]]; ctx->lastPos[g] = pos[segStart[g + 1] - 1]; }

  ENSURE(ctx->d_pos, (E ? E : 1) * 4);
  ENSURE(ctx->d_cnt, (E ? E : 1) * 4);
  ENSURE(ctx->d_off, nOff * 8);
  ENSURE(ctx->d_segStart, ((size_t)nSeg + 1) * 8);
  const uint64_t stride = ctx->tidxBase.back();
  const uint64_t idxTotal = stride * n_samples * 2;
  ENSURE(ctx->d_tidx, idxTotal * 4);
  ENSURE(ctx->d_small, 4096);
  if (E) {
    CU(cudaMemcpyAsync(ctx->d_pos.p, pos, E * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_cnt.p, cnt, E * 4, cudaMemcpyHostToDevice, ctx->stream));
  }
  CU(cudaMemcpyAsync(ctx->d_off.p, off, nOff * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_segStart.p, segStart.data(), ((size_t)nSeg + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_small.p, 0, 4096, ctx->stream));

  IndexArgs ia;
  ia.pos = ctx->d_pos.as<uint32_t>(); ia.cnt = ctx->d_cnt.as<uint32_t>(); ia.segStart = ctx->d_segStart.as<uint64_t>();
  ia.sizes = ctx->d_sizes.as<uint32_t>(); ia.tidxBase = ctx->d_tidxBase.as<uint32_t>(); ia.tidx = ctx->d_tidx.as<uint32_t>();
  ia.tidxStride = stride; ia.C = C; ia.nSeg = nSeg; ia.nEntries = E; ia.flags = ctx->d_small.as<uint32_t>();
  {
    const uint64_t blocks = (idxTotal + 255) / 256;
    index_fill_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ia, idxTotal);
    CU(cudaGetLastError());
  }
  if (E) {
    const uint64_t blocks = (E + 255) / 256;
    index_build_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ia);
    CU(cudaGetLastError());
  }
  uint32_t flags = 0;
  CU(cudaMemcpyAsync(&flags, ctx->d_small.p, 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  ctx->loaded = false;
  ctx->cntEffValid = false;
  if (flags & FLAG_UNSORTED) return fail(ctx, UPB_ERR_UNSORTED, "upb_load_tags: positions must be strictly ascending within each (sample, strand, contig) — the reference asserts on unsorted input (misc/peakcall.cpp:171)");
  if (flags & FLAG_BAD_POS) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: position outside 1..contig size (the reference's parser drops those, misc/format.cpp:670-676)");
  if (flags & FLAG_ZERO_COUNT) return fail(ctx, UPB_ERR_BAD_ARG, "upb_load_tags: zero count (the reference's parser never delivers those, misc/format.cpp:710)");
  ctx->loaded = true;
  ctx->stats.n_entries = E;
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// device-wide exclusive scan (uint32 in, uint64 out); total -> *d_total
// ------------------------------------------------------------------------------------------------
static int scan_u32 (upb_ctx* ctx, const uint32_t* d_in, uint64_t n, unsigned long long* d_out, unsigned long long* d_total, uint32_t* launches) {
  const uint32_t nBlocks = (uint32_t)((n + SCAN_BLOCK - 1) / SCAN_BLOCK);
  ENSURE(ctx->d_scanPartial, ((size_t)nBlocks + 1) * 8);
  unsigned long long* part = ctx->d_scanPartial.as<unsigned long long>();
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
struct HostRegion {
  uint32_t unit, buffer, contig, left, len, peak, wsum;
  uint64_t candIndex;
  double score, kurt, corr;
  std::vector<uint32_t> counts;
};

struct PipeCfg {
  uint32_t bw;
  double thr, kurtThr, corrThr, hitThr, background;
  bool nondir;
  bool needCorr;                     // compute strandCorr(0) for regions passing the earlier filters
  const uint8_t* control;            // may be NULL
};

struct PipeOut {
  std::vector<Unit> units;
  std::vector<HostRegion> regions;   // accepted, canonical order
  std::vector<uint64_t> unitCand;    // [nUnits]
  uint64_t nCand = 0;
};

static int check_left_edge (upb_ctx* ctx, uint32_t bw) {
  // every add() call moves the window, control samples included (src/regions.cpp:356-359)
  const uint32_t nSeg = (uint32_t)ctx->firstPos.size();
  for (uint32_t g = 0; g < nSeg; ++g)
    if (ctx->firstPos[g] != 0 && ctx->firstPos[g] <= bw)
      return fail(ctx, UPB_ERR_LEFT_EDGE, "a tag lies within `bandwidth` positions of a contig start: the reference's left-edge path (misc/peakcall.cpp:176-184, SURVEY Q2) is not emulated");
  return UPB_OK;
}

static int launch_kde (upb_ctx* ctx, KdeArgs& ka, bool nondir, bool dump, uint32_t totalWork, uint32_t* launches) {
  const KdeSmem L = kde_smem_layout(ka.bw, ka.nSrc * (nondir ? 2u : 1u), nondir);
  const size_t smem = L.total;
  if (smem > 227 * 1024) return fail(ctx, UPB_ERR_UNSUPPORTED, "bandwidth / sample count need more shared memory than one SM has");
  void (*kern)(const KdeArgs) = nullptr;
  if (nondir) kern = dump ? kde_kernel<true, true> : kde_kernel<true, false>;
  else kern = dump ? kde_kernel<false, true> : kde_kernel<false, false>;
  CU(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int perSm = 0;
  CU(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&perSm, kern, KDE_THREADS, smem));
  if (perSm < 1) perSm = 1;
  uint32_t grid = (uint32_t)std::min<uint64_t>((uint64_t)ctx->smCount * perSm, std::max<uint32_t>(totalWork, 1u));
  kern<<<grid, KDE_THREADS, smem, ctx->stream>>>(ka);
  CU(cudaGetLastError());
  *launches += 1;
  return UPB_OK;
}

// dump exact F/R densities for `n` position ranges (device descriptor arrays already filled) and, if asked,
// strandCorr(shift) for shifts 0..nShifts-1
static int dump_and_corr (upb_ctx* ctx, KdeArgs base, bool nondir, uint32_t n, uint32_t nShifts, uint32_t* launches,
                          std::vector<double>* corrOut /* n*nShifts */) {
  // chunk prefix and scratch offsets
  ENSURE(ctx->d_dOff, ((size_t)n + 1) * 8);
  ENSURE(ctx->d_dChunkStart, ((size_t)n + 1) * 8);
  unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 8;
  int r;
  if ((r = scan_u32(ctx, ctx->d_dLen.as<uint32_t>(), n, ctx->d_dOff.as<unsigned long long>(), d_tot, launches))) return r;
  if ((r = scan_u32(ctx, ctx->d_dChunks.as<uint32_t>(), n, ctx->d_dChunkStart.as<unsigned long long>(), d_tot + 1, launches))) return r;
  unsigned long long tot[2];
  CU(cudaMemcpyAsync(tot, d_tot, 16, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  // append the grand totals as entry n of the two prefix arrays
  CU(cudaMemcpyAsync(ctx->d_dOff.as<unsigned long long>() + n, &tot[0], 8, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_dChunkStart.as<unsigned long long>() + n, &tot[1], 8, cudaMemcpyHostToDevice, ctx->stream));
  if (tot[1] > 0xFFFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many positions to dump in one batch");
  ENSURE(ctx->d_F, (size_t)(tot[0] + 1) * 8);
  ENSURE(ctx->d_R, (size_t)(tot[0] + 1) * 8);
  KdeArgs ka = base;
  ka.dumpUnit = ctx->d_dUnit.as<uint32_t>(); ka.dumpLeft = ctx->d_dLeft.as<uint32_t>(); ka.dumpLen = ctx->d_dLen.as<uint32_t>();
  ka.dumpOff = (const uint64_t*)ctx->d_dOff.p; ka.dumpChunkStart = (const uint64_t*)ctx->d_dChunkStart.p; ka.nDump = n;
  ka.dumpF = ctx->d_F.as<double>(); ka.dumpR = ctx->d_R.as<double>(); ka.dumpNonzero = nullptr;
  CU(cudaMemsetAsync(ka.ticket, 0, 4, ctx->stream));
  if ((r = launch_kde(ctx, ka, nondir, true, (uint32_t)tot[1], launches))) return r;
  if (nShifts) {
    ENSURE(ctx->d_corr, (size_t)n * nShifts * 8);
    CorrArgs ca;
    ca.F = ctx->d_F.as<double>(); ca.R = ctx->d_R.as<double>(); ca.off = (const uint64_t*)ctx->d_dOff.p; ca.len = ctx->d_dLen.as<uint32_t>();
    ca.n = n; ca.nShifts = nShifts; ca.corr = ctx->d_corr.as<double>();
    const uint64_t warps = (uint64_t)n * nShifts;
    const uint64_t blocks = (warps * 32 + 255) / 256;
    strand_corr_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ca);
    CU(cudaGetLastError());
    *launches += 1;
    if (corrOut) {
      corrOut->resize((size_t)n * nShifts);
      CU(cudaMemcpyAsync(corrOut->data(), ctx->d_corr.p, (size_t)n * nShifts * 8, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
    }
  }
  return UPB_OK;
}

struct KdeBase {
  KdeArgs ka;
  std::vector<Unit> units;
  uint32_t totalTiles;
  bool haveControl;
};

// common set-up: units, taps, source list, control filter
static int prepare_kde (upb_ctx* ctx, const PipeCfg& cfg, KdeBase& kb, uint32_t* launches) {
  const uint32_t C = (uint32_t)ctx->sizes.size();
  const uint32_t n = ctx->nSamples;
  if (!ctx->loaded) return fail(ctx, UPB_ERR_NOT_LOADED, "no tags loaded (upb_set_contigs + upb_load_tags first)");
  if (cfg.bw == 0 || cfg.bw > (uint32_t)MAX_BW) return fail(ctx, UPB_ERR_UNSUPPORTED, "bandwidth must be in 1..1024");
  if (!(cfg.thr > 0)) return fail(ctx, UPB_ERR_BAD_ARG, "region threshold must be > 0 (the reference's leap handling relies on it, SURVEY Q10)");
  if (!(cfg.background > 0)) return fail(ctx, UPB_ERR_BAD_ARG, "background must be > 0");
  int r;
  if ((r = check_left_edge(ctx, cfg.bw))) return r;

  std::vector<uint32_t> src;
  std::vector<uint8_t> control(n, 0);
  bool haveControl = false;
  for (uint32_t s = 0; s < n; ++s) { if (cfg.control && cfg.control[s]) { control[s] = 1; haveControl = true; } else src.push_back(s); }
  kb.haveControl = haveControl;

  // units: one per (buffer, contig) that holds at least one non-control tag
  kb.units.clear();
  uint32_t tileStart = 0;
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
      u.nTiles = (uint32_t)(((uint64_t)ctx->sizes[c] + cfg.bw + TILE - 1) / TILE);
      u.tileStart = tileStart;
      if ((uint64_t)tileStart + u.nTiles > 0xFFFFFFF0ull) return fail(ctx, UPB_ERR_UNSUPPORTED, "too many tiles");
      tileStart += u.nTiles;
      kb.units.push_back(u);
    }
  kb.totalTiles = tileStart;

  std::vector<double> taps(2 * (size_t)cfg.bw + 1);
  upb_kernel_taps(cfg.bw, 1 / cfg.background, taps.data());          // src/regions.cpp:232
  ENSURE(ctx->d_taps, taps.size() * 8);
  ENSURE(ctx->d_src, (src.size() + 1) * 4);
  ENSURE(ctx->d_units, (kb.units.size() + 1) * sizeof(Unit));
  ENSURE(ctx->d_control, n + 1);
  ENSURE(ctx->d_small, 4096);
  CU(cudaMemcpyAsync(ctx->d_taps.p, taps.data(), taps.size() * 8, cudaMemcpyHostToDevice, ctx->stream));
  if (!src.empty()) CU(cudaMemcpyAsync(ctx->d_src.p, src.data(), src.size() * 4, cudaMemcpyHostToDevice, ctx->stream));
  if (!kb.units.empty()) CU(cudaMemcpyAsync(ctx->d_units.p, kb.units.data(), kb.units.size() * sizeof(Unit), cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemcpyAsync(ctx->d_control.p, control.data(), n, cudaMemcpyHostToDevice, ctx->stream));
  CU(cudaMemsetAsync(ctx->d_small.p, 0, 4096, ctx->stream));

  // Q4: zero the control tags the reference forgets
  if (haveControl && !(ctx->cntEffValid && ctx->lastControl == control)) {
    ENSURE(ctx->d_cntEff, (ctx->nEntries ? ctx->nEntries : 1) * 4);
    if (ctx->nEntries) {
      ControlArgs ca;
      ca.T = make_tags(ctx, false);
      ca.nTilesIdx = ctx->d_nTilesIdx.as<uint32_t>(); ca.segStart = ctx->d_segStart.as<uint64_t>();
      ca.control = ctx->d_control.as<uint8_t>(); ca.srcSample = ctx->d_src.as<uint32_t>(); ca.nSrc = (uint32_t)src.size();
      ca.cntEff = ctx->d_cntEff.as<uint32_t>(); ca.nSeg = n * 2 * C; ca.nEntries = ctx->nEntries;
      const uint64_t blocks = (ctx->nEntries + 255) / 256;
      control_filter_kernel<<<(unsigned)blocks, 256, 0, ctx->stream>>>(ca);
      CU(cudaGetLastError());
      *launches += 1;
    }
    ctx->lastControl = control;
    ctx->cntEffValid = true;
  }

  KdeArgs& ka = kb.ka;
  memset(&ka, 0, sizeof ka);
  ka.T = make_tags(ctx, haveControl);
  ka.units = ctx->d_units.as<Unit>();
  ka.nTilesIdx = ctx->d_nTilesIdx.as<uint32_t>();
  ka.nUnits = (uint32_t)kb.units.size();
  ka.totalTiles = kb.totalTiles;
  ka.srcSample = ctx->d_src.as<uint32_t>();
  ka.nSrc = (uint32_t)src.size();
  ka.bw = cfg.bw;
  ka.taps = ctx->d_taps.as<double>();
  ka.thr = cfg.thr;
  ka.kmaxBound = taps[cfg.bw] * (1.0 + 1e-9);
  ka.nondir = cfg.nondir ? 1 : 0;
  // d_small layout (4 KB): [0] flags u32, [1] ticket u32, [2] runCount u32, [4..5] hot tiles u64, [8..] scan totals u64
  ka.flags = ctx->d_small.as<uint32_t>();
  ka.ticket = ctx->d_small.as<uint32_t>() + 1;
  ka.runCount = ctx->d_small.as<uint32_t>() + 2;
  ka.statHotTiles = (unsigned long long*)(ctx->d_small.as<uint32_t>() + 4);
  return UPB_OK;
}

static int run_pipeline (upb_ctx* ctx, const PipeCfg& cfg, PipeOut& out) {
  CU(cudaSetDevice(ctx->device));
  uint32_t launches = 0;
  KdeBase kb;
  CU(cudaEventRecord(ctx->ev[0], ctx->stream));
  int r = prepare_kde(ctx, cfg, kb, &launches);
  if (r) return r;
  const uint32_t n = ctx->nSamples;
  out.units = kb.units;
  out.unitCand.assign(kb.units.size(), 0);
  out.regions.clear();
  out.nCand = 0;
  memset(&ctx->stats, 0, sizeof ctx->stats);
  ctx->stats.n_entries = ctx->nEntries;
  ctx->stats.kde_tiles = kb.totalTiles;
  if (kb.units.empty() || kb.ka.nSrc == 0) { ctx->stats.launches = launches; return UPB_OK; }
  KdeArgs ka = kb.ka;
  const uint32_t totalTiles = kb.totalTiles;

  // ---- K1 -----------------------------------------------------------------------------------------
  ENSURE(ctx->d_tileBase, (size_t)totalTiles * 4);
  ENSURE(ctx->d_tileCount, (size_t)totalTiles * 4);
  size_t runCap = std::max<size_t>(1u << 20, (size_t)(ctx->nEntries / 8));
  if (ctx->d_runs.cap / sizeof(RunRec) > runCap) runCap = ctx->d_runs.cap / sizeof(RunRec);
  uint32_t nRuns = 0;
  unsigned long long hotTiles = 0;
  for (int attempt = 0; attempt < 2; ++attempt) {
    ENSURE(ctx->d_runs, runCap * sizeof(RunRec));
    ka.runs = ctx->d_runs.as<RunRec>();
    ka.runCap = (uint32_t)std::min<size_t>(runCap, 0xFFFFFFFFu);
    ka.tileRunBase = ctx->d_tileBase.as<uint32_t>();
    ka.tileRunCount = ctx->d_tileCount.as<uint32_t>();
    CU(cudaMemsetAsync(ctx->d_small.p, 0, 64, ctx->stream));
    CU(cudaEventRecord(ctx->ev[1], ctx->stream));
    if ((r = launch_kde(ctx, ka, cfg.nondir, false, totalTiles, &launches))) return r;
    CU(cudaEventRecord(ctx->ev[2], ctx->stream));
    uint32_t small[8];
    CU(cudaMemcpyAsync(small, ctx->d_small.p, 32, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    nRuns = small[2];
    memcpy(&hotTiles, &small[4], 8);
    if (!(small[0] & FLAG_RUN_OVERFLOW)) break;
    if (attempt == 1) return fail(ctx, UPB_ERR_NOMEM, "run buffer overflow after resize");
    runCap = (size_t)nRuns + 1024;                                    // the counter holds the exact need
  }
  ctx->stats.kde_hot_tiles = hotTiles;
  { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[1], ctx->ev[2]); ctx->stats.ms_kde = ms; }
  if (nRuns == 0) {
    CU(cudaEventRecord(ctx->ev[3], ctx->stream)); CU(cudaEventSynchronize(ctx->ev[3]));
    float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); ctx->stats.ms_total = ms; ctx->stats.launches = launches;
    return UPB_OK;
  }

  // ---- K2: order by tile, stitch runs into candidate regions ---------------------------------------
  unsigned long long* d_tot = ctx->d_small.as<unsigned long long>() + 8;
  ENSURE(ctx->d_tileOrd, (size_t)totalTiles * 8);
  if ((r = scan_u32(ctx, ctx->d_tileCount.as<uint32_t>(), totalTiles, ctx->d_tileOrd.as<unsigned long long>(), d_tot, &launches))) return r;
  ENSURE(ctx->d_ordRuns, (size_t)nRuns * sizeof(RunRec));
  ENSURE(ctx->d_ordUnit, (size_t)nRuns * 4);
  ENSURE(ctx->d_head, (size_t)nRuns * 4);
  ENSURE(ctx->d_headScan, (size_t)nRuns * 8);
  order_runs_kernel<<<(totalTiles + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_runs.as<RunRec>(), ctx->d_tileBase.as<uint32_t>(), ctx->d_tileCount.as<uint32_t>(),
      ctx->d_tileOrd.as<unsigned long long>(), ctx->d_units.as<Unit>(), (uint32_t)kb.units.size(), totalTiles, ctx->d_ordRuns.as<RunRec>(), ctx->d_ordUnit.as<uint32_t>());
  run_heads_kernel<<<(nRuns + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_ordRuns.as<RunRec>(), ctx->d_ordUnit.as<uint32_t>(), nRuns, ctx->d_head.as<uint32_t>());
  CU(cudaGetLastError());
  launches += 2;
  if ((r = scan_u32(ctx, ctx->d_head.as<uint32_t>(), nRuns, ctx->d_headScan.as<unsigned long long>(), d_tot, &launches))) return r;
  ENSURE(ctx->d_cand, (size_t)nRuns * sizeof(Cand));                // nCand <= nRuns
  ENSURE(ctx->d_candOut, (kb.units.size() + 1) * 8);
  CU(cudaMemsetAsync(ctx->d_candOut.p, 0, (kb.units.size() + 1) * 8, ctx->stream));
  merge_runs_kernel<<<(nRuns + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_ordRuns.as<RunRec>(), ctx->d_ordUnit.as<uint32_t>(), ctx->d_head.as<uint32_t>(),
      ctx->d_headScan.as<unsigned long long>(), nRuns, ctx->d_cand.as<Cand>(), ctx->d_candOut.as<unsigned long long>());
  CU(cudaGetLastError());
  launches += 1;
  unsigned long long nCand64 = 0;
  CU(cudaMemcpyAsync(&nCand64, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaMemcpyAsync(out.unitCand.data(), ctx->d_candOut.p, kb.units.size() * 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const uint32_t nCand = (uint32_t)nCand64;
  out.nCand = nCand;
  ctx->stats.n_candidates = nCand;

  // ---- K3: hit filter -> survivors --------------------------------------------------------------------
  ENSURE(ctx->d_pass, (size_t)nCand * 4);
  ENSURE(ctx->d_passScan, (size_t)nCand * 8);
  ENSURE(ctx->d_list1, (size_t)nCand * 4);
  hit_filter_kernel<<<(nCand + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_cand.as<Cand>(), nCand, cfg.hitThr, ctx->d_pass.as<uint32_t>());
  launches += 1;
  if ((r = scan_u32(ctx, ctx->d_pass.as<uint32_t>(), nCand, ctx->d_passScan.as<unsigned long long>(), d_tot, &launches))) return r;
  compact_index_kernel<<<(nCand + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_pass.as<uint32_t>(), ctx->d_passScan.as<unsigned long long>(), nCand, ctx->d_list1.as<uint32_t>());
  CU(cudaGetLastError());
  launches += 1;
  unsigned long long n1_64 = 0;
  CU(cudaMemcpyAsync(&n1_64, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(cudaStreamSynchronize(ctx->stream));
  const uint32_t n1 = (uint32_t)n1_64;

  std::vector<uint32_t> list1(n1), pass2(n1), counts((size_t)n1 * n);
  std::vector<double> kurt(n1), corr;
  std::vector<Cand> cand1(n1);
  if (n1) {
    // ---- per-sample counts, kurtosis ------------------------------------------------------------------
    ENSURE(ctx->d_counts, (size_t)n1 * n * 4);
    ENSURE(ctx->d_offsum, (size_t)n1 * n * 4);
    ENSURE(ctx->d_kurt, (size_t)n1 * 8);
    ENSURE(ctx->d_pass2, (size_t)n1 * 4);
    CountArgs ca;
    ca.T = ka.T; ca.units = ka.units; ca.nTilesIdx = ka.nTilesIdx; ca.cand = ctx->d_cand.as<Cand>(); ca.list = ctx->d_list1.as<uint32_t>();
    ca.nList = n1; ca.nondir = cfg.nondir ? 1 : 0; ca.counts = ctx->d_counts.as<uint32_t>(); ca.offsum = ctx->d_offsum.as<uint32_t>();
    const uint64_t cthreads = (uint64_t)n1 * n;
    region_counts_kernel<<<(unsigned)((cthreads + 255) / 256), 256, 0, ctx->stream>>>(ca);
    KurtArgs kg;
    kg.T = ka.T; kg.units = ka.units; kg.nTilesIdx = ka.nTilesIdx; kg.cand = ctx->d_cand.as<Cand>(); kg.list = ctx->d_list1.as<uint32_t>();
    kg.nList = n1; kg.nondir = cfg.nondir ? 1 : 0; kg.counts = ctx->d_counts.as<uint32_t>(); kg.offsum = ctx->d_offsum.as<uint32_t>();
    kg.kurtThr = cfg.kurtThr; kg.kurt = ctx->d_kurt.as<double>(); kg.pass = ctx->d_pass2.as<uint32_t>();
    region_kurtosis_kernel<<<(n1 + KURT_WARPS - 1) / KURT_WARPS, KURT_WARPS * 32, 0, ctx->stream>>>(kg);
    CU(cudaGetLastError());
    launches += 2;

    // ---- strand correlation for the regions that got this far (misc/peakcall.cpp:41 evaluates it last) ----
    std::vector<uint32_t> sel;                                       // indices into list1 space
    if (cfg.needCorr) {
      ENSURE(ctx->d_pass2Scan, (size_t)n1 * 8);
      ENSURE(ctx->d_list2, (size_t)n1 * 4);
      if ((r = scan_u32(ctx, ctx->d_pass2.as<uint32_t>(), n1, ctx->d_pass2Scan.as<unsigned long long>(), d_tot, &launches))) return r;
      compact_index_kernel<<<(n1 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_pass2.as<uint32_t>(), ctx->d_pass2Scan.as<unsigned long long>(), n1, ctx->d_list2.as<uint32_t>());
      launches += 1;
      unsigned long long n2_64 = 0;
      CU(cudaMemcpyAsync(&n2_64, d_tot, 8, cudaMemcpyDeviceToHost, ctx->stream));
      CU(cudaStreamSynchronize(ctx->stream));
      const uint32_t n2 = (uint32_t)n2_64;
      sel.resize(n2);
      corr.assign(n1, std::nan(""));
      if (n2) {
        CU(cudaMemcpyAsync(sel.data(), ctx->d_list2.p, (size_t)n2 * 4, cudaMemcpyDeviceToHost, ctx->stream));
        // list2 holds indices into list1; the dump descriptors need candidate indices
        ENSURE(ctx->d_dUnit, (size_t)n2 * 4); ENSURE(ctx->d_dLeft, (size_t)n2 * 4); ENSURE(ctx->d_dLen, (size_t)n2 * 4); ENSURE(ctx->d_dChunks, (size_t)n2 * 4);
        ENSURE(ctx->d_bestShift, (size_t)n2 * 4);
        gather_u32_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_list1.as<uint32_t>(), ctx->d_list2.as<uint32_t>(), n2, ctx->d_bestShift.as<uint32_t>());
        dump_desc_kernel<<<(n2 + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_cand.as<Cand>(), ctx->d_bestShift.as<uint32_t>(), n2,
            ctx->d_dUnit.as<uint32_t>(), ctx->d_dLeft.as<uint32_t>(), ctx->d_dLen.as<uint32_t>(), ctx->d_dChunks.as<uint32_t>());
        CU(cudaGetLastError());
        launches += 2;
        std::vector<double> c2;
        if ((r = dump_and_corr(ctx, ka, cfg.nondir, n2, 1, &launches, &c2))) return r;
        for (uint32_t k = 0; k < n2; ++k) corr[sel[k]] = c2[k];
      }
    }

    // ---- bring the survivors back and assemble rows ---------------------------------------------------
    ENSURE(ctx->d_ordRuns, (size_t)n1 * sizeof(Cand));               // reuse as gather target
    CU(cudaMemcpyAsync(list1.data(), ctx->d_list1.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(pass2.data(), ctx->d_pass2.p, (size_t)n1 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(kurt.data(), ctx->d_kurt.p, (size_t)n1 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(counts.data(), ctx->d_counts.p, (size_t)n1 * n * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    // candidates of the survivors: copy the whole candidate array only if small, else gather row by row on host
    std::vector<Cand> allCand;
    allCand.resize(nCand);
    CU(cudaMemcpyAsync(allCand.data(), ctx->d_cand.p, (size_t)nCand * sizeof(Cand), cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < n1; ++i) cand1[i] = allCand[list1[i]];
  }
  CU(cudaEventRecord(ctx->ev[3], ctx->stream));
  CU(cudaEventSynchronize(ctx->ev[3]));
  { float ms = 0; cudaEventElapsedTime(&ms, ctx->ev[0], ctx->ev[3]); ctx->stats.ms_total = ms; }
  ctx->stats.launches = launches;

  // first candidate ordinal of each unit
  std::vector<uint64_t> unitFirst(kb.units.size() + 1, 0);
  for (size_t u = 0; u < kb.units.size(); ++u) unitFirst[u + 1] = unitFirst[u] + out.unitCand[u];
  for (uint32_t i = 0; i < n1; ++i) {
    if (!pass2[i]) continue;
    double cr = std::nan("");
    if (cfg.needCorr) {
      cr = corr[i];
      if (cfg.corrThr > -1 && !(cr >= cfg.corrThr)) continue;        // misc/peakcall.cpp:41 third clause
    }
    HostRegion hr;
    const Cand& c = cand1[i];
    hr.unit = c.unit; hr.buffer = kb.units[c.unit].buffer; hr.contig = kb.units[c.unit].contig;
    hr.left = c.left; hr.len = c.len; hr.peak = c.peak; hr.wsum = c.wsum; hr.score = c.score;
    hr.kurt = kurt[i]; hr.corr = cr;
    hr.candIndex = (uint64_t)list1[i] - unitFirst[c.unit];
    hr.counts.assign(counts.begin() + (size_t)i * n, counts.begin() + (size_t)(i + 1) * n);
    out.regions.push_back(std::move(hr));
  }
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// reference emission order, labels and dropped tail (src/regions.cpp:311-391)
// ------------------------------------------------------------------------------------------------
template <class T> static T* alloc_arr (size_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }

extern "C" void upb_free_region_table (upb_region_table* t) {
  if (!t) return;
  free(t->buffer); free(t->forward_label); free(t->dropped); free(t->contig); free(t->left); free(t->len); free(t->peak);
  free(t->cand_index); free(t->peak_score); free(t->kurtosis); free(t->corr); free(t->counts); free(t->tags_in_regions);
  free(t->unit_candidates); free(t->unit_last_tag);
  memset(t, 0, sizeof *t);
}

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

extern "C" int upb_reference_order (uint32_t C, uint32_t n, int directional, uint32_t bw, const uint32_t* firstPos, const uint32_t* lastPos,
                                    upb_region_table* t, uint64_t* perm) {
  if (!t || !perm || !firstPos || !lastPos) return UPB_ERR_BAD_ARG;
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
  for (uint64_t i = 0; i < R; ++i) { t->forward_label[i] = t->buffer[i] == 0 ? 1 : 0; t->dropped[i] = 0; }

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

  // Q5: the first Region object of reverseBuffer is constructed forward (misc/peakcall.cpp:103); if the
  // first candidate of the reverse buffer is accepted it is printed left-right
  if (directional && t->unit_candidates) {
    for (uint32_t c = 0; c < C; ++c) {
      if (t->unit_candidates[(size_t)C + c] == 0) continue;
      for (uint64_t i = 0; i < R; ++i)
        if (t->buffer[i] == 1 && t->contig[i] == c && t->cand_index[i] == 0) t->forward_label[i] = 1;
      break;
    }
  }

  // canonical rows are sorted by (buffer, contig, left)
  std::vector<uint64_t> order;
  order.reserve(R);
  if (!directional) {
    for (uint64_t i = 0; i < R; ++i) order.push_back(i);
    // only what the last contig's flush finalises is lost (Q6)
    for (uint64_t i = 0; i < R; ++i) if (t->contig[i] == C - 1 && tail(i)) t->dropped[i] = 1;
  } else if (!anyEarly) {
    for (uint64_t i = 0; i < R; ++i) order.push_back(i);            // all '+' regions, then all '-' regions
    for (uint64_t i = 0; i < R; ++i) if (t->buffer[i] == 1 && t->contig[i] == C - 1 && tail(i)) t->dropped[i] = 1;
  } else {
    // single sample whose '-' track is read in pass 1: per contig iteration the '+' tags come first, then
    // the '-' tags; regions closed by add() are written at once, the two flushes at the end of the
    // iteration hand over the '+' tail and then the '-' tail; pass 2 then writes what is pending (nothing is lost)
    std::vector<std::vector<uint64_t>> byUnit(2 * (size_t)C);
    for (uint64_t i = 0; i < R; ++i) byUnit[(size_t)t->buffer[i] * C + t->contig[i]].push_back(i);
    for (uint32_t c = 0; c < C; ++c) {
      for (int b = 0; b < 2; ++b) for (uint64_t i : byUnit[(size_t)b * C + c]) if (!tail(i)) order.push_back(i);
      for (int b = 0; b < 2; ++b) for (uint64_t i : byUnit[(size_t)b * C + c]) if (tail(i)) order.push_back(i);
    }
  }
  // dropped rows keep their slot at the end so that callers can still inspect them
  uint64_t k = 0;
  for (uint64_t i : order) if (!t->dropped[i]) perm[k++] = i;
  for (uint64_t i : order) if (t->dropped[i]) perm[k++] = i;
  return UPB_OK;
}

static int fill_table (upb_ctx* ctx, const PipeOut& po, upb_region_table* t) {
  const uint32_t n = ctx->nSamples, C = (uint32_t)ctx->sizes.size();
  const size_t R = po.regions.size();
  memset(t, 0, sizeof *t);
  t->n_rows = R; t->n_samples = n;
  t->buffer = alloc_arr<uint8_t>(R); t->forward_label = alloc_arr<uint8_t>(R); t->dropped = alloc_arr<uint8_t>(R);
  t->contig = alloc_arr<uint32_t>(R); t->left = alloc_arr<uint32_t>(R); t->len = alloc_arr<uint32_t>(R); t->peak = alloc_arr<uint32_t>(R);
  t->cand_index = alloc_arr<uint64_t>(R); t->peak_score = alloc_arr<double>(R); t->kurtosis = alloc_arr<double>(R); t->corr = alloc_arr<double>(R);
  t->counts = alloc_arr<uint32_t>(R * n); t->tags_in_regions = alloc_arr<uint64_t>(n);
  t->unit_candidates = alloc_arr<uint64_t>(2 * (size_t)C); t->unit_last_tag = alloc_arr<uint32_t>(2 * (size_t)C);
  if (!t->buffer || !t->forward_label || !t->dropped || !t->contig || !t->left || !t->len || !t->peak || !t->cand_index || !t->peak_score ||
      !t->kurtosis || !t->corr || !t->counts || !t->tags_in_regions || !t->unit_candidates || !t->unit_last_tag) {
    upb_free_region_table(t);
    return fail(ctx, UPB_ERR_NOMEM, "out of host memory for the region table");
  }
  for (size_t i = 0; i < R; ++i) {
    const HostRegion& h = po.regions[i];
    t->buffer[i] = (uint8_t)h.buffer; t->forward_label[i] = h.buffer == 0; t->contig[i] = h.contig; t->left[i] = h.left; t->len[i] = h.len;
    t->peak[i] = h.peak; t->cand_index[i] = h.candIndex; t->peak_score[i] = h.score; t->kurtosis[i] = h.kurt; t->corr[i] = h.corr;
    for (uint32_t s = 0; s < n; ++s) { t->counts[i * n + s] = h.counts[s]; t->tags_in_regions[s] += h.counts[s]; }   // misc/peakcall.cpp:44
  }
  for (size_t u = 0; u < po.units.size(); ++u) t->unit_candidates[(size_t)po.units[u].buffer * C + po.units[u].contig] = po.unitCand[u];
  t->n_passed = R;
  t->n_rejected = po.nCand - R;
  return UPB_OK;
}

extern "C" int upb_call_regions (upb_ctx* ctx, const upb_region_params* p, upb_region_table* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !out) return fail(ctx, UPB_ERR_BAD_ARG, "upb_call_regions: NULL argument");
  memset(out, 0, sizeof *out);
  if (p->n_coeffs != 0) return fail(ctx, UPB_ERR_UNSUPPORTED, "scaling coefficients (-z) are not implemented yet");
  PipeCfg cfg;
  cfg.bw = p->bandwidth; cfg.thr = p->region_threshold; cfg.kurtThr = p->kurtosis_threshold; cfg.corrThr = p->corr_threshold;
  cfg.hitThr = p->hit_threshold; cfg.background = p->background; cfg.nondir = !p->directional; cfg.control = p->control;
  cfg.needCorr = cfg.nondir;        // -D: the filter (if corr_threshold > -1) and the optional correlation column
  PipeOut po;
  int r = run_pipeline(ctx, cfg, po);
  if (r) return r;
  // canonical order: (buffer, contig, left) — units are already in that order and candidates are position-ordered
  if ((r = fill_table(ctx, po, out))) return r;
  const uint32_t C = (uint32_t)ctx->sizes.size(), n = ctx->nSamples;
  for (uint32_t s = 0; s < n; ++s) for (uint32_t d = 0; d < 2; ++d) for (uint32_t c = 0; c < C; ++c) {
    const uint32_t lp = ctx->lastPos[((size_t)s * 2 + d) * C + c];
    const uint32_t b = p->directional ? d : 0;
    if (lp > out->unit_last_tag[(size_t)b * C + c]) out->unit_last_tag[(size_t)b * C + c] = lp;
  }
  if (p->reference_order && out->n_rows) {
    std::vector<uint64_t> perm(out->n_rows);
    r = upb_reference_order(C, n, p->directional, p->bandwidth, ctx->firstPos.data(), ctx->lastPos.data(), out, perm.data());
    if (r) { upb_free_region_table(out); return fail(ctx, r, "the reference's merge loop would abort on this input (several directional samples whose '-' track is read during pass 1; SURVEY Q7)"); }
    // apply the permutation to every column
    const uint64_t R = out->n_rows;
    auto permute = [&] (auto* col, size_t width) {
      typedef typename std::remove_pointer<decltype(col)>::type T;
      std::vector<T> tmp(col, col + R * width);
      for (uint64_t k = 0; k < R; ++k) for (size_t w = 0; w < width; ++w) col[k * width + w] = tmp[perm[k] * width + w];
    };
    permute(out->buffer, 1); permute(out->forward_label, 1); permute(out->dropped, 1); permute(out->contig, 1); permute(out->left, 1);
    permute(out->len, 1); permute(out->peak, 1); permute(out->cand_index, 1); permute(out->peak_score, 1); permute(out->kurtosis, 1);
    permute(out->corr, 1); permute(out->counts, n);
  } else if (p->reference_order && p->directional) {
    // still report the abort condition even when nothing was called
    std::vector<uint64_t> none(1);
    r = upb_reference_order(C, n, p->directional, p->bandwidth, ctx->firstPos.data(), ctx->lastPos.data(), out, none.data());
    if (r) { upb_free_region_table(out); return fail(ctx, r, "the reference's merge loop would abort on this input (SURVEY Q7)"); }
  }
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

struct SumIdx { uint32_t sum; uint32_t idx; };
static bool compareRegionCounts (const SumIdx& i, const SumIdx& j) { return i.sum > j.sum; }   // src/strand_shift.cpp:26-29

extern "C" int upb_strand_shift (upb_ctx* ctx, const upb_shift_params* p, upb_shift_report* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !out) return fail(ctx, UPB_ERR_BAD_ARG, "upb_strand_shift: NULL argument");
  memset(out, 0, sizeof *out);
  PipeCfg cfg;
  cfg.bw = p->bandwidth; cfg.thr = p->region_threshold; cfg.kurtThr = p->kurtosis_threshold; cfg.corrThr = -1;   // :142
  cfg.hitThr = (double)p->hit_threshold; cfg.background = p->background; cfg.nondir = true; cfg.control = nullptr; cfg.needCorr = false;
  PipeOut po;
  int r = run_pipeline(ctx, cfg, po);
  if (r) return r;
  const size_t R = po.regions.size();
  out->n_regions_found = R;
  out->n_freq = p->max_shift + 1;
  out->shift_freq = alloc_arr<uint64_t>(out->n_freq);
  const uint16_t toTest = (uint16_t)p->n_regions_to_test;            // UShort (Q12)
  const size_t cap = toTest ? toTest : 1;
  out->tested_contig = alloc_arr<uint32_t>(cap); out->tested_left = alloc_arr<uint32_t>(cap); out->tested_len = alloc_arr<uint32_t>(cap);
  out->tested_best_shift = alloc_arr<uint32_t>(cap); out->tested_best_corr = alloc_arr<double>(cap);

  // :198 std::sort by Region::sum() descending — same comparator outcomes, same libstdc++ algorithm, hence
  // the same permutation as the reference's sort of Region pointers
  std::vector<SumIdx> order(R);
  for (size_t i = 0; i < R; ++i) { order[i].sum = po.regions[i].wsum; order[i].idx = (uint32_t)i; }   // no controls: sum() == pooled weight
  std::sort(order.begin(), order.end(), compareRegionCounts);

  // :201-226 — regions are examined in sorted order until enough qualify; the device sweeps a batch at a time
  uint16_t tested = 0;
  size_t cursor = 0;
  uint32_t launches = ctx->stats.launches;
  KdeBase kb;
  if (R && toTest) { uint32_t l2 = 0; if ((r = prepare_kde(ctx, cfg, kb, &l2))) return r; launches += l2; }
  const uint32_t nShifts = p->max_shift + 1;
  while (tested < toTest && cursor < R) {
    std::vector<uint32_t> batch;                                     // indices into po.regions
    const size_t want = (size_t)(toTest - tested) + (size_t)(toTest - tested) / 4 + 8;
    while (cursor < R && batch.size() < want) {
      const HostRegion& h = po.regions[order[cursor].idx];
      if (h.len > 2 * p->max_shift + 3) batch.push_back(order[cursor].idx);   // :206 large enough to try all shifts
      ++cursor;
    }
    if (batch.empty()) break;
    const uint32_t nb = (uint32_t)batch.size();
    std::vector<uint32_t> dUnit(nb), dLeft(nb), dLen(nb), dChunks(nb);
    for (uint32_t k = 0; k < nb; ++k) {
      const HostRegion& h = po.regions[batch[k]];
      dUnit[k] = h.unit; dLeft[k] = h.left; dLen[k] = h.len; dChunks[k] = (h.len + TILE - 1) / TILE;
    }
    ENSURE(ctx->d_dUnit, (size_t)nb * 4); ENSURE(ctx->d_dLeft, (size_t)nb * 4); ENSURE(ctx->d_dLen, (size_t)nb * 4); ENSURE(ctx->d_dChunks, (size_t)nb * 4);
    CU(cudaMemcpyAsync(ctx->d_dUnit.p, dUnit.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_dLeft.p, dLeft.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_dLen.p, dLen.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_dChunks.p, dChunks.data(), (size_t)nb * 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((r = dump_and_corr(ctx, kb.ka, true, nb, nShifts, &launches, nullptr))) return r;
    ENSURE(ctx->d_bestShift, (size_t)nb * 4);
    ENSURE(ctx->d_bestCorr, (size_t)nb * 8);
    best_shift_kernel<<<(nb + 255) / 256, 256, 0, ctx->stream>>>(ctx->d_corr.as<double>(), nb, nShifts, ctx->d_bestShift.as<uint32_t>(), ctx->d_bestCorr.as<double>());
    CU(cudaGetLastError());
    launches += 1;
    std::vector<uint32_t> bs(nb);
    std::vector<double> bc(nb);
    CU(cudaMemcpyAsync(bs.data(), ctx->d_bestShift.p, (size_t)nb * 4, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(bc.data(), ctx->d_bestCorr.p, (size_t)nb * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t k = 0; k < nb && tested < toTest; ++k) {
      if (bc[k] >= p->corr_threshold) {                              // :220 disqualify regions under threshold
        const HostRegion& h = po.regions[batch[k]];
        ++out->shift_freq[bs[k]];
        out->n_tags_in_regions += h.wsum;
        out->tested_contig[tested] = h.contig; out->tested_left[tested] = h.left; out->tested_len[tested] = h.len;
        out->tested_best_shift[tested] = bs[k]; out->tested_best_corr[tested] = bc[k];
        ++tested;
      }
    }
  }
  ctx->stats.launches = launches;
  out->n_regions_tested = tested;
  if (tested == 0) return fail(ctx, UPB_ERR_NO_REGIONS, "no regions qualified with given settings");   // :228-230

  // :239-258 smooth the histogram of maxima with Kernel(5) and take the mode (151 numbers: host)
  double modeKernel[11];
  upb_kernel_taps(5, 1, modeKernel);
  const size_t m = out->n_freq;
  std::vector<double> density(m, 0.0);
  for (size_t i = 0; i < m; ++i)
    for (int j = 0; j < 11; ++j) {
      const long k = (long)i - 5 + j;
      if (k >= 0 && k < (long)m) { volatile double term = (double)out->shift_freq[i] * modeKernel[j]; density[k] = density[k] + term; }
    }
  double bestDensity = 0;
  uint32_t best = 0;
  if (m >= 5) for (size_t i = p->min_shift; i < m - 5; ++i) if (density[i] > bestDensity) { best = (uint32_t)i; bestDensity = density[i]; }
  out->best_shift = best;
  return UPB_OK;
}

// ------------------------------------------------------------------------------------------------
// not yet implemented entry points (declared in the header so the ABI is stable)
// ------------------------------------------------------------------------------------------------
extern "C" void upb_free_profile (upb_profile* p) { if (!p) return; free(p->pos); free(p->score); memset(p, 0, sizeof *p); }

extern "C" int upb_profile_contig (upb_ctx* ctx, const upb_region_params* p, int buffer, uint32_t contig, upb_profile* out) {
  if (!ctx) return UPB_ERR_BAD_ARG;
  if (!p || !out) return fail(ctx, UPB_ERR_BAD_ARG, "upb_profile_contig: NULL argument");
  memset(out, 0, sizeof *out);
  const uint32_t C = (uint32_t)ctx->sizes.size();
  if (contig >= C || buffer < 0 || buffer > 1 || (!p->directional && buffer != 0)) return fail(ctx, UPB_ERR_BAD_ARG, "upb_profile_contig: bad buffer/contig");
  if (p->n_coeffs != 0) return fail(ctx, UPB_ERR_UNSUPPORTED, "scaling coefficients (-z) are not implemented yet");
  CU(cudaSetDevice(ctx->device));
  PipeCfg cfg;
  cfg.bw = p->bandwidth; cfg.thr = p->region_threshold; cfg.kurtThr = p->kurtosis_threshold; cfg.corrThr = p->corr_threshold;
  cfg.hitThr = p->hit_threshold; cfg.background = p->background; cfg.nondir = !p->directional; cfg.control = p->control; cfg.needCorr = false;
  KdeBase kb;
  uint32_t launches = 0;
  int r = prepare_kde(ctx, cfg, kb, &launches);
  if (r) return r;
  int unit = -1;
  for (size_t u = 0; u < kb.units.size(); ++u) if (kb.units[u].contig == contig && kb.units[u].buffer == (uint32_t)buffer) unit = (int)u;
  if (unit < 0) return UPB_OK;      // no non-control tags: the reference writes nothing for it
  // evaluate the contig in slabs of 8M positions; the reference evaluates firstTag-bw .. lastTag+bw (Q20)
  const uint64_t last = (uint64_t)ctx->sizes[contig] + cfg.bw;
  const uint32_t SLAB = 8u << 20;
  std::vector<uint32_t> posOut;
  std::vector<double> scoreOut;
  std::vector<double> hF, hR;
  for (uint64_t start = 1; start <= last; start += SLAB) {
    const uint32_t len = (uint32_t)std::min<uint64_t>(SLAB, last - start + 1);
    const uint32_t dUnit = (uint32_t)unit, dLeft = (uint32_t)start, dLen = len, dChunks = (len + TILE - 1) / TILE;
    ENSURE(ctx->d_dUnit, 4); ENSURE(ctx->d_dLeft, 4); ENSURE(ctx->d_dLen, 4); ENSURE(ctx->d_dChunks, 4);
    CU(cudaMemcpyAsync(ctx->d_dUnit.p, &dUnit, 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_dLeft.p, &dLeft, 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_dLen.p, &dLen, 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(cudaMemcpyAsync(ctx->d_dChunks.p, &dChunks, 4, cudaMemcpyHostToDevice, ctx->stream));
    if ((r = dump_and_corr(ctx, kb.ka, cfg.nondir, 1, 0, &launches, nullptr))) return r;
    hF.resize(len); hR.resize(len);
    CU(cudaMemcpyAsync(hF.data(), ctx->d_F.p, (size_t)len * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaMemcpyAsync(hR.data(), ctx->d_R.p, (size_t)len * 8, cudaMemcpyDeviceToHost, ctx->stream));
    CU(cudaStreamSynchronize(ctx->stream));
    for (uint32_t i = 0; i < len; ++i) {
      volatile double scv = hF[i] + hR[i];                           // misc/peakcall.cpp:57
      const double sc = scv;
      if (sc != 0) { posOut.push_back((uint32_t)(start + i)); scoreOut.push_back(sc); }
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
  (void)directional; (void)extension; (void)n_regions; (void)regions; (void)counts_out;
  if (!ctx) return UPB_ERR_BAD_ARG;
  return fail(ctx, UPB_ERR_UNSUPPORTED, "upb_count_tags_in_regions is not implemented yet");
}
