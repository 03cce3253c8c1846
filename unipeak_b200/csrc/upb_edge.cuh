// upb_edge.cuh — the reference's left-edge behaviour, reproduced on the device.
//
// ProfileBuffer::add (misc/peakcall.cpp:161-222) keeps its pop_front/push_back INSIDE the test
// `bufferPos_ + i > bandwidth_` (:176-184): while the previous tag lies within `bandwidth` of the contig start the
// window is not shifted for the positions <= 0, so the cells that are in the window at that moment keep their content
// but are later reported at other positions — densities and hit vectors of the first tags of a contig are displaced,
// and up to `bandwidth` stale cells surface just before the next tag, however far away that is.  Everything the
// window holds after the last such add() has been pushed by correct shifts, so positions beyond X = (position of that
// add) + bandwidth are exactly what span_kernel computes.
//
// One thread per affected (ProfileBuffer, contig) unit therefore replays the reference's streaming algorithm — the
// k-way merge of the samples (src/regions.cpp:326-359), add(), processPosition(), processRegion(), Region::addPos and
// the region statistics, statement by statement — from the first tag until the first position beyond X at which no
// region is open (Xstop).  Its regions replace whatever span_kernel finds at or before Xstop.  The replay is strictly
// sequential, as the reference is; it touches a few hundred positions per affected unit in practice.
#pragma once

#include "upb_device.cuh"

namespace upb {

struct EdgeRow {
  uint32_t left, len, peak, wsum;
  uint32_t candIndex, pad;
  double score, kurt, corr;
};

struct EdgeArgs {
  Tags T;
  const Unit* units;
  const uint32_t* affected;         // [nAffected] unit indices
  uint32_t nAffected;
  const uint8_t* control;           // [nSamples]
  const double* coeffBySample;      // [nSamples] or NULL
  const double* taps;               // 2bw+1
  uint32_t bw;
  int nondir;
  double thr, kurtThr, corrThr, hitThr;
  // per-unit scratch
  unsigned long long* cursors;      // [nAffected * 2 * nSamples]
  double* ringF; double* ringR;     // [nAffected * w]
  uint32_t* ringHits;               // [nAffected * w * nSamples]
  uint8_t* ringHas;                 // [nAffected * w]
  uint32_t* regPc;                  // [nAffected * regCap] pooled hits of every position added to the open region
  uint8_t* regHas;                  // [nAffected * regCap]
  double* regF; double* regR;       // [nAffected * regCap]
  uint32_t* regSums;                // [nAffected * nSamples]
  uint32_t* callCounts;             // [nAffected * nSamples] counts vector of the current add()
  uint32_t regCap, rowCap;
  unsigned long long maxPositions;  // replay budget (positions); exceeding it is reported, not guessed around
  // outputs
  uint32_t* stop;                   // [nAffected] Xstop
  uint32_t* nCand;                  // [nAffected] regions closed by the replay (accepted + rejected)
  uint32_t* nRows;                  // [nAffected] accepted
  EdgeRow* rows;                    // [nAffected * rowCap]
  uint32_t* rowCounts;              // [nAffected * rowCap * nSamples]
  uint32_t* flags;                  // bit 0: row buffer full, bit 1: region scratch full, bit 2: replay budget exceeded
};

struct EdgeState {
  // ProfileBuffer
  uint32_t w, head, bw, bufferPos, lastPos;
  // Region under construction (misc/data.hpp:54-73)
  uint32_t left, len, peakPos;
  double peakScore;
  // bookkeeping
  uint32_t lastSkipAdd;             // position of the last add() that skipped a shift
  bool skipped;
  uint32_t nCand, nRows;
  unsigned long long processed;
  bool done;
  uint32_t stop;
};

__device__ inline void edge_process_region (const EdgeArgs& a, uint32_t ai, EdgeState& st) {
  // misc/peakcall.cpp:33-53 over misc/data.cpp:104-193
  const uint32_t n = a.T.nSamples;
  uint32_t* sums = a.regSums + (size_t)ai * n;
  const uint32_t* pc = a.regPc + (size_t)ai * a.regCap;
  const uint8_t* has = a.regHas + (size_t)ai * a.regCap;
  const double* f = a.regF + (size_t)ai * a.regCap;
  const double* r = a.regR + (size_t)ai * a.regCap;
  uint32_t nonControl = 0;
  for (uint32_t e = 0; e < n; ++e) if (!a.control[e]) nonControl += sums[e];
  // posMean / posKurtosis: offsets are UShort, count and sum HitCount
  uint32_t count = 0, sum = 0;
  for (uint32_t i = 0; i < st.len; ++i) if (has[i]) { count += pc[i]; sum += pc[i] * (i & 0xFFFFu); }
  const double xbar = (double)sum / (double)count;
  double sum2 = 0.0, sum4 = 0.0;
  for (uint32_t i = 0; i < st.len; ++i) if (has[i]) {
    const double d = __dsub_rn((double)(i & 0xFFFFu), xbar);
    const double d2 = __dmul_rn(d, d), d4 = __dmul_rn(d2, d2);
    sum2 = __dadd_rn(sum2, __dmul_rn((double)pc[i], d2));
    sum4 = __dadd_rn(sum4, __dmul_rn((double)pc[i], d4));
  }
  const double kurt = __ddiv_rn(__dmul_rn(__dsub_rn((double)count, 1.0), sum4), __dmul_rn(sum2, sum2));
  // strandCorr(0): misc/data.cpp:22-58,184-193
  double corr = __longlong_as_double(0x7ff8000000000000ll);
  if (a.nondir && st.len > 3) {
    const double m = (double)st.len;
    double s1 = 0.0, s2 = 0.0;
    for (uint32_t i = 0; i < st.len; ++i) { s1 = __dadd_rn(s1, f[i]); s2 = __dadd_rn(s2, r[i]); }
    const double mean1 = __ddiv_rn(s1, m), mean2 = __ddiv_rn(s2, m);
    double q1 = 0.0, q2 = 0.0, q12 = 0.0;
    for (uint32_t i = 0; i < st.len; ++i) {
      const double x = __dsub_rn(f[i], mean1), y = __dsub_rn(r[i], mean2);
      q1 = __dadd_rn(q1, __dmul_rn(x, x)); q2 = __dadd_rn(q2, __dmul_rn(y, y)); q12 = __dadd_rn(q12, __dmul_rn(x, y));
    }
    const double m1 = __dsub_rn(m, 1.0);
    const double sd1 = __dsqrt_rn(__ddiv_rn(q1, m1)), sd2 = __dsqrt_rn(__ddiv_rn(q2, m1));
    corr = __ddiv_rn(q12, __dmul_rn(__dmul_rn(m1, sd1), sd2));
  }
  bool pass = (double)nonControl >= a.hitThr;
  if (pass) pass = (a.kurtThr == 0.0) || (st.len > 1 && kurt <= a.kurtThr);
  if (pass) pass = (a.corrThr <= -1.0) || (corr >= a.corrThr);
  if (pass) {
    if (st.nRows < a.rowCap) {
      EdgeRow row;
      row.left = st.left; row.len = st.len; row.peak = st.peakPos; row.wsum = nonControl; row.candIndex = st.nCand; row.pad = 0;
      row.score = st.peakScore; row.kurt = kurt; row.corr = corr;
      a.rows[(size_t)ai * a.rowCap + st.nRows] = row;
      uint32_t* rc = a.rowCounts + ((size_t)ai * a.rowCap + st.nRows) * n;
      for (uint32_t e = 0; e < n; ++e) rc[e] = sums[e];
    } else atomicOr(a.flags, 1u);
    ++st.nRows;
  }
  ++st.nCand;
  // region_ = new Region(forward_, contig_, 0, nExpt_)
  st.left = 0; st.len = 0; st.peakPos = 0; st.peakScore = 0.0;
  for (uint32_t e = 0; e < n; ++e) sums[e] = 0;
}

__device__ inline void edge_add_pos (const EdgeArgs& a, uint32_t ai, EdgeState& st, bool hasHits, const uint32_t* hits, double fs, double rs) {
  // Region::addPos, misc/data.cpp:92-102
  const uint32_t n = a.T.nSamples;
  if (st.len >= a.regCap) { atomicOr(a.flags, 2u); return; }
  uint32_t pc = 0;
  if (hasHits) { uint32_t* sums = a.regSums + (size_t)ai * n; for (uint32_t e = 0; e < n; ++e) { pc += hits[e]; sums[e] += hits[e]; } }
  const size_t o = (size_t)ai * a.regCap + st.len;
  a.regPc[o] = pc; a.regHas[o] = hasHits ? 1 : 0; a.regF[o] = fs; a.regR[o] = rs;
  ++st.len;
  const double score = __dadd_rn(fs, rs);
  if (st.peakPos == 0 || score > st.peakScore) { st.peakPos = st.left + st.len - 1; st.peakScore = score; }
}

__device__ inline void edge_process_position (const EdgeArgs& a, uint32_t ai, EdgeState& st, uint32_t pos, double fs, double rs, bool hasHits, const uint32_t* hits) {
  // misc/peakcall.cpp:55-86 (the assert pos > lastPos_ cannot fire here: positions only grow)
  const double score = __dadd_rn(fs, rs);
  if (pos == st.lastPos + 1) {
    if (st.left != 0) {
      if (score >= a.thr) edge_add_pos(a, ai, st, hasHits, hits, fs, rs);
      else edge_process_region(a, ai, st);
    } else if (score >= a.thr) { st.left = pos; edge_add_pos(a, ai, st, hasHits, hits, fs, rs); }
  } else {                                                            // leap: left is NOT set here
    if (st.left != 0) edge_process_region(a, ai, st);
    if (score >= a.thr) edge_add_pos(a, ai, st, hasHits, hits, fs, rs);
  }
  st.lastPos = pos;
  ++st.processed;
  // hand over to span_kernel at the first position beyond the displaced cells with no region open
  if (st.skipped && st.bufferPos > st.bw && (unsigned long long)pos >= (unsigned long long)st.lastSkipAdd + st.bw && st.left == 0 && st.len == 0) { st.done = true; st.stop = pos; }
  if (st.processed > a.maxPositions) { atomicOr(a.flags, 4u); st.done = true; st.stop = pos; }
}

// ProfileBuffer::add, misc/peakcall.cpp:161-222.  counts == NULL stands for the empty vector flushContig passes.
__device__ inline void edge_add (const EdgeArgs& a, uint32_t ai, EdgeState& st, const uint32_t* counts, uint32_t pos, bool forward) {
  const uint32_t n = a.T.nSamples, w = st.w;
  double* F = a.ringF + (size_t)ai * w; double* R = a.ringR + (size_t)ai * w;
  uint32_t* H = a.ringHits + (size_t)ai * w * n; uint8_t* has = a.ringHas + (size_t)ai * w;
  uint32_t nStatic = w;
  if ((unsigned long long)pos <= (unsigned long long)st.bufferPos + 2ull * st.bw) nStatic = pos - st.bufferPos;
  if (st.bufferPos != 0) {
    for (uint32_t i = 0; i < nStatic && !st.done; ++i) {
      if (st.bufferPos + i > st.bw) {
        const uint32_t c = st.head;
        edge_process_position(a, ai, st, st.bufferPos + i - st.bw, F[c], R[c], has[c] != 0, H + (size_t)c * n);
        F[c] = 0.0; R[c] = 0.0; has[c] = 0;                           // pop_front + push_back(BufferPos())
        st.head = (st.head + 1) % w;
      } else { st.skipped = true; st.lastSkipAdd = pos; }             // the shift the reference forgets (Q2)
    }
  }
  if (st.done) return;
  double countSum = 0.0;
  if (counts) {
    if (!a.coeffBySample) { for (uint32_t e = 0; e < n; ++e) if (!a.control[e]) countSum = __dadd_rn(countSum, (double)counts[e]); }
    else {
      for (uint32_t e = 0; e < n; ++e) if (!a.control[e]) countSum = __dadd_rn(countSum, __dmul_rn((double)counts[e], a.coeffBySample[e]));
      for (uint32_t e = 0; e < n; ++e) if (!a.control[e]) countSum = __dadd_rn(countSum, (double)counts[e]);
    }
  }
  if (countSum != 0.0) {
    for (uint32_t j = 0; j < w; ++j) {
      const uint32_t c = (st.head + j) % w;
      const double term = __dmul_rn(a.taps[j], countSum);
      if (forward) F[c] = __dadd_rn(F[c], term); else R[c] = __dadd_rn(R[c], term);
    }
    const uint32_t c = (st.head + st.bw) % w;
    uint32_t* h = H + (size_t)c * n;
    if (has[c]) { for (uint32_t e = 0; e < n; ++e) h[e] += counts[e]; }
    else { for (uint32_t e = 0; e < n; ++e) h[e] = counts[e]; has[c] = 1; }
  }
  st.bufferPos = pos;
}

__global__ void left_edge_kernel (EdgeArgs a) {
  const uint32_t ai = blockIdx.x;
  if (ai >= a.nAffected || threadIdx.x != 0) return;
  const Unit U = a.units[a.affected[ai]];
  const uint32_t n = a.T.nSamples, C = a.T.C, w = 2 * a.bw + 1;
  EdgeState st;
  st.w = w; st.head = 0; st.bw = a.bw; st.bufferPos = 0; st.lastPos = 0;
  st.left = 0; st.len = 0; st.peakPos = 0; st.peakScore = 0.0;
  st.lastSkipAdd = 0; st.skipped = false; st.nCand = 0; st.nRows = 0; st.processed = 0; st.done = false; st.stop = 0;
  for (uint32_t j = 0; j < w; ++j) { a.ringF[(size_t)ai * w + j] = 0.0; a.ringR[(size_t)ai * w + j] = 0.0; a.ringHas[(size_t)ai * w + j] = 0; }
  uint32_t* sums = a.regSums + (size_t)ai * n;
  for (uint32_t e = 0; e < n; ++e) sums[e] = 0;
  // stream cursors of every sample (and strand) on this contig: src/regions.cpp:335-354
  unsigned long long* cur = a.cursors + (size_t)ai * 2 * n;
  const uint32_t d0 = a.nondir ? 0u : U.buffer, d1 = a.nondir ? 1u : U.buffer;
  for (uint32_t s = 0; s < n; ++s) for (uint32_t d = 0; d < 2; ++d) cur[s * 2 + d] = a.T.off[(uint64_t)(s * 2 + d) * (C + 1) + U.contig];
  uint32_t* counts = a.callCounts + (size_t)ai * n;
  while (!st.done) {
    // next position carrying a tag of any sample on the strands this buffer sees
    uint32_t pos = 0xFFFFFFFFu;
    for (uint32_t s = 0; s < n; ++s) for (uint32_t d = d0; d <= d1; ++d) {
      const unsigned long long e = cur[s * 2 + d];
      if (e < a.T.off[(uint64_t)(s * 2 + d) * (C + 1) + U.contig + 1]) pos = min(pos, a.T.pos[e]);
    }
    if (pos == 0xFFFFFFFFu) break;
    for (uint32_t d = d0; d <= d1 && !st.done; ++d) {                // forward call first, then reverse (src/regions.cpp:356-359)
      bool any = false;
      for (uint32_t s = 0; s < n; ++s) {
        counts[s] = 0;
        const unsigned long long e = cur[s * 2 + d];
        if (e < a.T.off[(uint64_t)(s * 2 + d) * (C + 1) + U.contig + 1] && a.T.pos[e] == pos) { counts[s] = a.T.cnt[e]; cur[s * 2 + d] = e + 1; any = true; }
      }
      if (any) edge_add(a, ai, st, counts, pos, d == 0);
    }
  }
  if (!st.done) {
    // flushContig (misc/peakcall.cpp:224-231): the window is drained.  Its last cell is an end tap (exactly 0), so any
    // open region is closed by the flush itself; the replay owns the whole unit
    edge_add(a, ai, st, nullptr, st.bufferPos + w, true);
    if (!st.done) st.stop = 0xFFFFFFFFu;
  }
  a.stop[ai] = st.stop; a.nCand[ai] = st.nCand; a.nRows[ai] = st.nRows;
}

}  // namespace upb
