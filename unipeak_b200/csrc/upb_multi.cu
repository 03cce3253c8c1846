// upb_multi.cu — several GPUs of one box behind one handle (include/unipeak_b200.h "upb_multi_*").
//
// The reference is one thread on one core (README:60-61): its merge loop walks the contigs one after the other
// (src/regions.cpp:311-391, src/strand_shift.cpp:143-190) and its tags_in_regions walks the samples one after the other
// (src/tags_in_regions.cpp:131-199).  Both loops shard without a data-path exchange: whole contigs per GPU for region
// calling and the shift calibration, whole samples per GPU for tags_in_regions.  One worker thread per GPU owns one upb_ctx; the
// handle only orchestrates:
//   regions        every rank calls upb_call_regions_shared on ONE pinned host table (upb_host_alloc memory is mapped into every
//                  GPU): the per-(buffer, contig) row counts are exchanged on the devices through that table and every GPU writes
//                  its rows to their place in table order over its own PCIe link; rank 0 assembles totals and emission order.
//   strand_shift   the protocol of unipeak_b200/multi.py::strand_shift_sharded (upb_shift_regions / upb_shift_examine_order /
//                  upb_shift_sweep / upb_shift_mode) with the ranks' lists concatenated in host memory instead of all-gathered.
//   tags_in_regions  every rank counts its own samples; the columns are interleaved on the host.
// Ranks are threads of one process, so what they exchange is either in the shared mapped table (device side) or plain host
// memory: there is no collective to call here.  Ranks that are PROCESSES (torchrun, bench.py --gpus N) use the same C entry
// points with torch.distributed / NCCL for the tag totals and the list all-gathers (unipeak_b200/multi.py).
#include <algorithm>
#include <atomic>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

#include "../../include/unipeak_b200.h"

namespace {

// one persistent thread per GPU: a CUDA context is cheapest to drive from the thread that always drives it, and a call then costs
// two condition-variable hand-offs instead of a thread creation per rank
class Worker {
 public:
  Worker () : th_([this] { loop(); }) {}
  ~Worker () { { std::lock_guard<std::mutex> l(m_); quit_ = true; } cv_.notify_all(); th_.join(); }
  void post (std::function<void()> f) { { std::lock_guard<std::mutex> l(m_); job_ = std::move(f); busy_ = true; } cv_.notify_all(); }
  void wait () { std::unique_lock<std::mutex> l(m_); done_.wait(l, [this] { return !busy_; }); }
 private:
  void loop () {
    for (;;) {
      std::function<void()> f;
      { std::unique_lock<std::mutex> l(m_); cv_.wait(l, [this] { return quit_ || (busy_ && job_); }); if (quit_) return; f = std::move(job_); job_ = nullptr; }
      f();
      { std::lock_guard<std::mutex> l(m_); busy_ = false; }
      done_.notify_all();
    }
  }
  std::mutex m_;
  std::condition_variable cv_, done_;
  std::function<void()> job_;
  bool busy_ = false, quit_ = false;
  std::thread th_;
};

}  // namespace

struct upb_multi {
  int n = 0;
  std::vector<int> dev;
  std::vector<upb_ctx*> ctx;
  std::vector<Worker*> worker;
  std::vector<uint32_t> sizes, owner;         // owner[c]: rank that loads contig c
  std::vector<std::vector<uint8_t>> keep;     // [rank][contig]
  uint32_t nSamples = 0;
  bool loaded = false;
  // samples-sharded load (tags_in_regions)
  std::vector<uint32_t> sampleLo;             // [n + 1] rank r holds samples [sampleLo[r], sampleLo[r + 1])
  bool bySample = false;
  // shared table
  void* seg = nullptr;
  uint64_t segBytes = 0, rowCap = 0, step = 0;
  uint32_t segSamples = 0;
  std::string err;
};

static int mfail (upb_multi* m, int code, const std::string& msg) { if (m) m->err = msg; return code; }

// run f(rank) on every worker; returns the first non-zero status (and that rank's message)
static int run_all (upb_multi* m, const std::function<int(int)>& f) {
  std::vector<int> rc(m->n, 0);
  for (int r = 0; r < m->n; ++r) m->worker[r]->post([&, r] { rc[r] = f(r); });
  for (int r = 0; r < m->n; ++r) m->worker[r]->wait();
  // the rank that failed first-hand, not the one that reports that another rank failed
  int pick = -1;
  for (int r = 0; r < m->n; ++r) if (rc[r]) {
    const char* e = upb_last_error(m->ctx[r]);
    const bool secondHand = e && strstr(e, "a rank reported a failed step");
    if (pick < 0 || (!secondHand && strstr(upb_last_error(m->ctx[pick]), "a rank reported a failed step"))) pick = r;
  }
  if (pick < 0) return UPB_OK;
  const char* e = upb_last_error(m->ctx[pick]);
  return mfail(m, rc[pick], std::string("rank ") + std::to_string(pick) + " (GPU " + std::to_string(m->dev[pick]) + "): " + (e ? e : ""));
}

extern "C" const char* upb_multi_last_error (const upb_multi* m) { return m ? m->err.c_str() : "no multi-GPU handle"; }
extern "C" int upb_multi_gpus (const upb_multi* m) { return m ? m->n : 0; }
extern "C" upb_ctx* upb_multi_context (upb_multi* m, int rank) { return (m && rank >= 0 && rank < m->n) ? m->ctx[rank] : nullptr; }

extern "C" void upb_multi_destroy (upb_multi* m) {
  if (!m) return;
  for (int r = 0; r < (int)m->ctx.size(); ++r) if (m->ctx[r]) { upb_ctx* c = m->ctx[r]; m->worker[r]->post([c] { upb_destroy(c); }); m->worker[r]->wait(); }
  for (Worker* w : m->worker) delete w;
  if (m->seg) upb_host_free(m->seg);
  delete m;
}

extern "C" int upb_multi_create (int n_gpus, const int* devices, upb_multi** out) {
  if (!out || n_gpus < 1) return UPB_ERR_BAD_ARG;
  *out = nullptr;
  const int have = upb_device_count();
  if (have < 1) return UPB_ERR_NO_DEVICE;
  upb_multi* m = new upb_multi;
  m->n = n_gpus;
  m->dev.resize(n_gpus); m->ctx.assign(n_gpus, nullptr);
  for (int r = 0; r < n_gpus; ++r) {
    m->dev[r] = devices ? devices[r] : r;
    if (m->dev[r] < 0 || m->dev[r] >= have) { delete m; return UPB_ERR_NO_DEVICE; }
  }
  for (int r = 0; r < n_gpus; ++r) m->worker.push_back(new Worker);
  const int rc = run_all(m, [m] (int r) { return upb_create(m->dev[r], &m->ctx[r]); });
  if (rc) { upb_multi_destroy(m); return rc; }
  *out = m;
  return UPB_OK;
}

extern "C" int upb_multi_set_contigs (upb_multi* m, uint32_t n_contigs, const uint32_t* sizes) {
  if (!m) return UPB_ERR_BAD_ARG;
  if (!n_contigs || !sizes) return mfail(m, UPB_ERR_BAD_ARG, "upb_multi_set_contigs: no contigs");
  m->sizes.assign(sizes, sizes + n_contigs);
  // whole contigs to ranks, longest first to the least loaded rank (tag density is roughly uniform along the genome)
  std::vector<uint32_t> order(n_contigs);
  for (uint32_t c = 0; c < n_contigs; ++c) order[c] = c;
  std::stable_sort(order.begin(), order.end(), [&] (uint32_t a, uint32_t b) { return sizes[a] > sizes[b]; });
  std::vector<uint64_t> load(m->n, 0);
  m->owner.assign(n_contigs, 0);
  for (uint32_t c : order) {
    int best = 0;
    for (int r = 1; r < m->n; ++r) if (load[r] < load[best]) best = r;
    m->owner[c] = (uint32_t)best; load[best] += sizes[c];
  }
  m->keep.assign(m->n, std::vector<uint8_t>(n_contigs, 0));
  for (uint32_t c = 0; c < n_contigs; ++c) m->keep[m->owner[c]][c] = 1;
  m->loaded = false;
  return run_all(m, [m, n_contigs, sizes] (int r) { return upb_set_contigs(m->ctx[r], n_contigs, sizes); });
}

extern "C" int upb_multi_owner (const upb_multi* m, uint32_t* owner_out) {
  if (!m || !owner_out || m->owner.empty()) return UPB_ERR_BAD_ARG;
  memcpy(owner_out, m->owner.data(), m->owner.size() * 4);
  return UPB_OK;
}

extern "C" int upb_multi_load_tags (upb_multi* m, const upb_tag_source* src) {
  if (!m) return UPB_ERR_BAD_ARG;
  if (!src || m->sizes.empty()) return mfail(m, UPB_ERR_BAD_ARG, "upb_multi_load_tags: upb_multi_set_contigs first");
  m->nSamples = src->n_samples; m->bySample = false;
  const int rc = run_all(m, [m, src] (int r) {
    upb_tag_source s = *src;
    s.keep_contig = m->keep[r].data();                                // the arrays describe every contig; only the owned ones travel to this GPU
    return upb_load_tags_async(m->ctx[r], &s);
  });
  m->loaded = rc == UPB_OK;
  return rc;
}

extern "C" int upb_multi_load_wait (upb_multi* m) {
  if (!m) return UPB_ERR_BAD_ARG;
  return run_all(m, [m] (int r) { return upb_load_wait(m->ctx[r]); });
}

// samples to ranks in contiguous blocks (tags_in_regions walks its samples independently: src/tags_in_regions.cpp:175-199)
extern "C" int upb_multi_load_tags_by_sample (upb_multi* m, uint32_t n_samples, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt) {
  if (!m) return UPB_ERR_BAD_ARG;
  if (!n_samples || !off || m->sizes.empty()) return mfail(m, UPB_ERR_BAD_ARG, "upb_multi_load_tags_by_sample: bad argument");
  const uint32_t C = (uint32_t)m->sizes.size();
  m->nSamples = n_samples; m->bySample = true;
  m->sampleLo.assign(m->n + 1, 0);
  for (int r = 0; r <= m->n; ++r) m->sampleLo[r] = (uint32_t)((uint64_t)n_samples * r / m->n);
  const int rc = run_all(m, [&] (int r) -> int {
    const uint32_t s0 = m->sampleLo[r], s1 = m->sampleLo[r + 1];
    if (s1 == s0) return UPB_OK;
    // this rank's samples as arrays of their own: offsets rebased to the first entry of the block
    const uint64_t* o = off + (size_t)s0 * 2 * (C + 1);
    const uint64_t base = o[0];
    std::vector<uint64_t> local((size_t)(s1 - s0) * 2 * (C + 1));
    for (size_t k = 0; k < local.size(); ++k) local[k] = o[k] - base;
    return upb_load_tags(m->ctx[r], s1 - s0, local.data(), pos + base, cnt + base);
  });
  m->loaded = rc == UPB_OK;
  return rc;
}

static int ensure_table (upb_multi* m, uint64_t rowCap) {
  const uint32_t C = (uint32_t)m->sizes.size();
  const uint64_t need = upb_shared_table_bytes(C, m->nSamples, (uint32_t)m->n, rowCap);
  if (m->seg && m->segBytes >= need && m->segSamples == m->nSamples && m->rowCap >= rowCap) return UPB_OK;
  if (m->seg) { upb_host_free(m->seg); m->seg = nullptr; }
  const int rc = upb_host_alloc(&m->seg, need);
  if (rc) return mfail(m, rc, "upb_multi: cannot allocate the shared region table");
  m->segBytes = need; m->rowCap = rowCap; m->segSamples = m->nSamples; m->step = 0;
  upb_shared_table t;
  memset(&t, 0, sizeof t);
  t.host = m->seg; t.device = m->seg; t.bytes = need; t.rank = 0; t.world = (uint32_t)m->n; t.n_contigs = C; t.n_samples = m->nSamples; t.row_capacity = rowCap;
  return upb_shared_table_init(&t);
}

// src/regions.cpp:304-393 over the GPUs of the handle.  `out` is rank 0's view of the shared table: valid until the next call on
// this handle; upb_free_region_table on it is a no-op.
extern "C" int upb_multi_call_regions (upb_multi* m, const upb_region_params* p, upb_region_table* out) {
  if (!m) return UPB_ERR_BAD_ARG;
  if (!p || !out) return mfail(m, UPB_ERR_BAD_ARG, "upb_multi_call_regions: NULL argument");
  if (!m->loaded || m->bySample) return mfail(m, UPB_ERR_NOT_LOADED, "upb_multi_call_regions: upb_multi_load_tags first");
  memset(out, 0, sizeof *out);
  const uint32_t C = (uint32_t)m->sizes.size();
  int rc;
  if (!m->seg || m->segSamples != m->nSamples) {
    // first call: every rank calls its own contigs once, which sizes its arrays; the row counts size the table
    std::vector<uint64_t> rows(m->n, 0);
    upb_region_params q = *p;
    q.reference_order = 0;
    if ((rc = run_all(m, [&] (int r) { upb_region_table t; const int e = upb_call_regions(m->ctx[r], &q, &t); if (!e) { rows[r] = t.n_rows; upb_free_region_table(&t); } return e; }))) return rc;
    uint64_t total = 0;
    for (uint64_t v : rows) total += v;
    if ((rc = ensure_table(m, total + total / 2 + 4096))) return rc;
  }
  for (int attempt = 0; attempt < 3; ++attempt) {
    ++m->step;
    std::vector<upb_region_table> tabs(m->n);
    rc = run_all(m, [&] (int r) {
      upb_shared_table t;
      memset(&t, 0, sizeof t);
      t.host = m->seg; t.device = m->seg; t.bytes = m->segBytes; t.rank = (uint32_t)r; t.world = (uint32_t)m->n; t.n_contigs = C; t.n_samples = m->nSamples; t.row_capacity = m->rowCap;
      return upb_call_regions_shared(m->ctx[r], p, &t, m->step, &tabs[r]);
    });
    if (rc == UPB_OK) { *out = tabs[0]; return UPB_OK; }
    // Ranks that share one GPU (devices = {0, 0}: tests) can stall each other: a rank that has to grow a scratch array frees device memory,
    // which waits for every kernel on that GPU — including the other rank's row_place_kernel, which waits for this rank's counts.  The wait is
    // bounded and reported; the arrays have grown by then, so the step is simply run again.  One rank per GPU never gets here.
    if (rc == UPB_ERR_CUDA && m->err.find("never published its row counts") != std::string::npos && attempt < 2) {
      upb_shared_table t;                                             // every rank has returned: wipe the exchange area (a failed rank left its poison there)
      memset(&t, 0, sizeof t);
      t.host = m->seg; t.device = m->seg; t.bytes = m->segBytes; t.world = (uint32_t)m->n; t.n_contigs = C; t.n_samples = m->nSamples; t.row_capacity = m->rowCap;
      upb_shared_table_init(&t);
      m->step = 0;
      continue;
    }
    if (rc != UPB_ERR_NOMEM) return rc;
    // more rows than the table has room for (other settings than the sizing call's): a larger table, and once more
    uint64_t total = 0;
    for (int r = 0; r < m->n; ++r) total += tabs[r].n_rows;
    if ((rc = ensure_table(m, std::max<uint64_t>(2 * m->rowCap, total + total / 2 + 4096)))) return rc;
  }
  return mfail(m, UPB_ERR_NOMEM, "upb_multi_call_regions: the shared table stayed too small");
}

extern "C" int upb_multi_profile_contig (upb_multi* m, const upb_region_params* p, int buffer, uint32_t contig, upb_profile* out) {
  if (!m || contig >= m->owner.size()) return UPB_ERR_BAD_ARG;
  const int r = (int)m->owner[contig];
  int rc = 0;
  m->worker[r]->post([&] { rc = upb_profile_contig(m->ctx[r], p, buffer, contig, out); });
  m->worker[r]->wait();
  if (rc) { const char* e = upb_last_error(m->ctx[r]); mfail(m, rc, e ? e : ""); }
  return rc;
}

// ---- bin/strand_shift: src/strand_shift.cpp:139-259 with the contigs spread over the GPUs ---------------------------------------
template <class T> static T* alloc_arr (size_t n) { return (T*)calloc(n ? n : 1, sizeof(T)); }

extern "C" int upb_multi_strand_shift (upb_multi* m, const upb_shift_params* p, upb_shift_report* out) {
  if (!m) return UPB_ERR_BAD_ARG;
  if (!p || !out) return mfail(m, UPB_ERR_BAD_ARG, "upb_multi_strand_shift: NULL argument");
  if (!m->loaded || m->bySample) return mfail(m, UPB_ERR_NOT_LOADED, "upb_multi_strand_shift: upb_multi_load_tags first");
  memset(out, 0, sizeof *out);
  const uint32_t C = (uint32_t)m->sizes.size();
  // 1 + 2: every rank's Region vector (:143-190); concatenated by contig they are the reference's vector
  std::vector<upb_shift_region_list> lists(m->n);
  for (upb_shift_region_list& l : lists) memset(&l, 0, sizeof l);
  int rc = run_all(m, [&] (int r) { return upb_shift_regions(m->ctx[r], p, &lists[r]); });
  struct Freer { std::vector<upb_shift_region_list>& v; ~Freer () { for (upb_shift_region_list& l : v) upb_free_shift_regions(&l); } } freer{lists};
  if (rc) return rc;
  uint64_t R = 0;
  for (const upb_shift_region_list& l : lists) R += l.n;
  if (R > 0xFFFFFFFFull) return mfail(m, UPB_ERR_UNSUPPORTED, "upb_multi_strand_shift: too many regions");
  std::vector<uint32_t> gContig(R), gLeft(R), gLen(R), gSum(R);
  {
    // each rank's list is in (contig, left) order and a contig lives on one rank: walk the contigs in table order
    std::vector<uint64_t> cur(m->n, 0);
    uint64_t k = 0;
    for (uint32_t c = 0; c < C; ++c) {
      const int r = (int)m->owner[c];
      const upb_shift_region_list& l = lists[r];
      uint64_t& i = cur[r];
      while (i < l.n && l.contig[i] == c) { gContig[k] = c; gLeft[k] = l.left[i]; gLen[k] = l.len[i]; gSum[k] = l.sum[i]; ++k; ++i; }
    }
    if (k != R) return mfail(m, UPB_ERR_CUDA, "upb_multi_strand_shift: a rank returned a region outside its contigs");
  }
  out->n_regions_found = R;
  out->n_freq = p->max_shift + 1;
  out->shift_freq = alloc_arr<uint64_t>(out->n_freq);
  const uint16_t toTest = (uint16_t)p->n_regions_to_test;            // UShort
  const size_t cap = toTest ? toTest : 1;
  out->tested_contig = alloc_arr<uint32_t>(cap); out->tested_left = alloc_arr<uint32_t>(cap); out->tested_len = alloc_arr<uint32_t>(cap);
  out->tested_best_shift = alloc_arr<uint32_t>(cap); out->tested_best_corr = alloc_arr<double>(cap);
  if (!out->shift_freq || !out->tested_contig || !out->tested_left || !out->tested_len || !out->tested_best_shift || !out->tested_best_corr) {
    upb_free_shift_report(out);
    return mfail(m, UPB_ERR_NOMEM, "upb_multi_strand_shift: out of host memory");
  }
  // 3: examination order = the head of the reference's std::sort by Region::sum() (:198), grown on demand
  std::vector<uint32_t> order(R ? R : 1);
  uint64_t sortedEnd = 0;
  auto sort_more = [&] (uint64_t want) { return upb_shift_examine_order(R, gSum.data(), want, order.data(), &sortedEnd); };
  if (R && (rc = sort_more(std::min<uint64_t>(R, 2 * (uint64_t)toTest + 256)))) { upb_free_shift_report(out); return mfail(m, rc, "upb_shift_examine_order"); }
  uint16_t tested = 0;
  uint64_t cursor = 0;
  while (tested < toTest && cursor < R) {
    std::vector<uint32_t> batch;
    const size_t want = (size_t)(toTest - tested) + (size_t)(toTest - tested) / 4 + 8;
    while (cursor < R && batch.size() < want) {
      if (cursor >= sortedEnd && (rc = sort_more(std::min<uint64_t>(R, std::max<uint64_t>(2 * sortedEnd, cursor + 1024))))) { upb_free_shift_report(out); return mfail(m, rc, "upb_shift_examine_order"); }
      const uint32_t i = order[cursor];
      if (gLen[i] > 2 * p->max_shift + 3) batch.push_back(i);         // :206 large enough to try all shifts
      ++cursor;
    }
    if (batch.empty()) break;
    // 4: every rank sweeps the regions of the batch that lie on its contigs
    const size_t nb = batch.size();
    std::vector<uint32_t> bs(nb, 0);
    std::vector<double> bc(nb, 0.0);
    rc = run_all(m, [&] (int r) -> int {
      std::vector<uint32_t> slot, contig, left, len;
      for (size_t k = 0; k < nb; ++k) if ((int)m->owner[gContig[batch[k]]] == r) { slot.push_back((uint32_t)k); contig.push_back(gContig[batch[k]]); left.push_back(gLeft[batch[k]]); len.push_back(gLen[batch[k]]); }
      if (slot.empty()) return UPB_OK;
      std::vector<uint32_t> s(slot.size());
      std::vector<double> c(slot.size());
      const int e = upb_shift_sweep(m->ctx[r], p, slot.size(), contig.data(), left.data(), len.data(), s.data(), c.data());
      if (e) return e;
      for (size_t j = 0; j < slot.size(); ++j) { bs[slot[j]] = s[j]; bc[slot[j]] = c[j]; }
      return UPB_OK;
    });
    if (rc) { upb_free_shift_report(out); return rc; }
    // 5: :220-225 in examination order
    for (size_t k = 0; k < nb && tested < toTest; ++k) {
      if (bc[k] >= p->corr_threshold) {
        const uint32_t i = batch[k];
        ++out->shift_freq[bs[k]];
        out->n_tags_in_regions += gSum[i];
        out->tested_contig[tested] = gContig[i]; out->tested_left[tested] = gLeft[i]; out->tested_len[tested] = gLen[i];
        out->tested_best_shift[tested] = bs[k]; out->tested_best_corr[tested] = bc[k];
        ++tested;
      }
    }
  }
  out->n_regions_tested = tested;
  if (tested == 0) return mfail(m, UPB_ERR_NO_REGIONS, "no regions qualified with given settings");   // :228-230
  return upb_shift_mode(out->shift_freq, out->n_freq, p->min_shift, &out->best_shift);
}

// ---- bin/tags_in_regions: src/tags_in_regions.cpp:131-199 with the samples spread over the GPUs -----------------------------------
extern "C" int upb_multi_count_tags_in_regions (upb_multi* m, int directional, uint32_t extension, uint64_t n_regions, const upb_region_key* regions, uint32_t* counts_out) {
  if (!m) return UPB_ERR_BAD_ARG;
  if (!m->loaded || !m->bySample) return mfail(m, UPB_ERR_NOT_LOADED, "upb_multi_count_tags_in_regions: upb_multi_load_tags_by_sample first");
  if (n_regions && (!regions || !counts_out)) return mfail(m, UPB_ERR_BAD_ARG, "upb_multi_count_tags_in_regions: NULL argument");
  const uint32_t n = m->nSamples;
  return run_all(m, [&] (int r) -> int {
    const uint32_t s0 = m->sampleLo[r], s1 = m->sampleLo[r + 1], k = s1 - s0;
    if (!k || !n_regions) return UPB_OK;
    std::vector<uint32_t> part((size_t)n_regions * k);
    const int e = upb_count_tags_in_regions(m->ctx[r], directional, extension, n_regions, regions, part.data());
    if (e) return e;
    for (uint64_t i = 0; i < n_regions; ++i) memcpy(counts_out + i * n + s0, part.data() + i * k, (size_t)k * 4);   // disjoint columns per rank
    return UPB_OK;
  });
}
