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
//
// Chains.  A unit whose EVERY tag lies within `bandwidth` of the contig start never shifts the window far enough for the flush
// (flushContig, misc/peakcall.cpp:224-231) to drain it: bandwidth + 1 - lastTag cells survive at the front of the deque — and a
// region may stay open — when the buffer moves on to its next contig, where the stale cells are reported at the first positions
// processed.  The thread that replays such a unit therefore carries its window and its open Region into the next unit of the same
// ProfileBuffer (the next contig on which any sample, controls included, has a tag) and replays that one as well, until the
// stale cells are gone; `chainNext` links the affected units in that order and only chain heads start a thread.
//
// Two optional recordings, for the callers that need more than the region rows: the -w profile points of one unit
// (processPosition's profileOut_->write, misc/peakcall.cpp:81-84) and the forward / reverse score vectors of the accepted regions
// long enough for the shift sweep of bin/strand_shift (Region::forwardScores / reverseScores, src/strand_shift.cpp:206-219).
#pragma once

#include "upb_device.cuh"

namespace upb {

struct EdgeRow {
  uint32_t left, len, peak, wsum;
  uint32_t candIndex, vecOff;       // vecOff: 1 + offset of the row's score vectors in vecF / vecR, 0 if not recorded
  double score, kurt, corr;
};

struct EdgeArgs {
  Tags T;
  const Unit* units;
  const uint32_t* affected;         // [nAffected] unit indices
  const uint32_t* chainNext;        // [nAffected] affected index of the unit that inherits this one's window if it leaves stale cells, 0xFFFFFFFF if none
  const uint8_t* chainHead;         // [nAffected] 1: a thread starts here; 0: reached through chainNext
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
  uint32_t* flags;                  // bit 0: row buffer full, bit 1: region scratch full, bit 2: replay budget exceeded, bit 3: profile buffer full, bit 4: vector pool full
  // optional recordings
  uint32_t profAff;                 // affected index whose profile points are recorded, 0xFFFFFFFF: none
  uint32_t* profPos; double* profScore; uint32_t profCap; uint32_t* profN;
  uint32_t vecMinLen;               // accepted rows with len > vecMinLen keep their score vectors; 0xFFFFFFFF: none
  double* vecF; double* vecR; uint32_t vecCap; uint32_t* vecUsed;
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
  uint32_t ri;                      // scratch slot of the chain (the head's affected index): window, open region, per-sample sums
  bool carried;                     // this unit started on the window its predecessor left behind
};

__device__ inline void edge_process_region (const EdgeArgs& a, uint32_t ai, EdgeState& st) {
  // misc/peakcall.cpp:33-53 over misc/data.cpp:104-193
  const uint32_t n = a.T.nSamples;
  uint32_t* sums = a.regSums + (size_t)st.ri * n;
  const uint32_t* pc = a.regPc + (size_t)st.ri * a.regCap;
  const uint8_t* has = a.regHas + (size_t)st.ri * a.regCap;
  const double* f = a.regF + (size_t)st.ri * a.regCap;
  const double* r = a.regR + (size_t)st.ri * a.regCap;
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
      row.left = st.left; row.len = st.len; row.peak = st.peakPos; row.wsum = nonControl; row.candIndex = st.nCand; row.vecOff = 0;
      row.score = st.peakScore; row.kurt = kurt; row.corr = corr;
      if (st.len > a.vecMinLen) {
        const uint32_t at = atomicAdd(a.vecUsed, st.len);
        if (at <= a.vecCap && st.len <= a.vecCap - at) {
          for (uint32_t i = 0; i < st.len; ++i) { a.vecF[at + i] = f[i]; a.vecR[at + i] = r[i]; }
          row.vecOff = at + 1;
        } else atomicOr(a.flags, 16u);
      }
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
  if (hasHits) { uint32_t* sums = a.regSums + (size_t)st.ri * n; for (uint32_t e = 0; e < n; ++e) { pc += hits[e]; sums[e] += hits[e]; } }
  const size_t o = (size_t)st.ri * a.regCap + st.len;
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
  if (score != 0.0 && ai == a.profAff) {                             // profileOut_->write(PosScore(forward_, contig_, pos, score)), :81-84
    const uint32_t k = *a.profN;
    if (k < a.profCap) { a.profPos[k] = pos; a.profScore[k] = score; } else atomicOr(a.flags, 8u);
    *a.profN = k + 1;
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
  double* F = a.ringF + (size_t)st.ri * w; double* R = a.ringR + (size_t)st.ri * w;
  uint32_t* H = a.ringHits + (size_t)st.ri * w * n; uint8_t* has = a.ringHas + (size_t)st.ri * w;
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
  if (st.carried && st.bufferPos == 0) { st.skipped = true; st.lastSkipAdd = pos; }   // stale cells sit in front of this first tag's window
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
  if (blockIdx.x >= a.nAffected || threadIdx.x != 0 || !a.chainHead[blockIdx.x]) return;
  const uint32_t n = a.T.nSamples, C = a.T.C, w = 2 * a.bw + 1;
  EdgeState st;
  st.ri = blockIdx.x;
  st.w = w; st.head = 0; st.bw = a.bw;
  st.left = 0; st.len = 0; st.peakPos = 0; st.peakScore = 0.0;
  st.processed = 0;
  for (uint32_t j = 0; j < w; ++j) { a.ringF[(size_t)st.ri * w + j] = 0.0; a.ringR[(size_t)st.ri * w + j] = 0.0; a.ringHas[(size_t)st.ri * w + j] = 0; }
  uint32_t* sums = a.regSums + (size_t)st.ri * n;
  for (uint32_t e = 0; e < n; ++e) sums[e] = 0;
  unsigned long long* cur = a.cursors + (size_t)st.ri * 2 * n;
  uint32_t* counts = a.callCounts + (size_t)st.ri * n;
  bool carried = false;
  for (uint32_t ai = blockIdx.x; ai != 0xFFFFFFFFu; ai = a.chainNext[ai]) {
    const Unit U = a.units[a.affected[ai]];
    // what flushContig leaves behind for the next contig: bufferPos_ = 0, lastPos_ = 0 (misc/peakcall.cpp:226-227); window and Region stay
    st.bufferPos = 0; st.lastPos = 0;
    st.lastSkipAdd = 0; st.skipped = false; st.nCand = 0; st.nRows = 0; st.done = false; st.stop = 0; st.carried = carried;
    // stream cursors of every sample (and strand) on this contig: src/regions.cpp:335-354
    const uint32_t d0 = a.nondir ? 0u : U.buffer, d1 = a.nondir ? 1u : U.buffer;
    for (uint32_t s = 0; s < n; ++s) for (uint32_t d = 0; d < 2; ++d) cur[s * 2 + d] = a.T.off[(uint64_t)(s * 2 + d) * (C + 1) + U.contig];
    while (!st.done) {
      // next position carrying a tag of any sample on the strands this buffer sees
      uint32_t pos = 0xFFFFFFFFu;
      for (uint32_t s = 0; s < n; ++s) for (uint32_t d = d0; d <= d1; ++d) {
        const unsigned long long e = cur[s * 2 + d];
        if (e < a.T.off[(uint64_t)(s * 2 + d) * (C + 1) + U.contig + 1]) pos = min(pos, a.T.pos[e]);
      }
      if (pos == 0xFFFFFFFFu) break;
      for (uint32_t d = d0; d <= d1 && !st.done; ++d) {              // forward call first, then reverse (src/regions.cpp:356-359)
        bool any = false;
        for (uint32_t s = 0; s < n; ++s) {
          counts[s] = 0;
          const unsigned long long e = cur[s * 2 + d];
          if (e < a.T.off[(uint64_t)(s * 2 + d) * (C + 1) + U.contig + 1] && a.T.pos[e] == pos) { counts[s] = a.T.cnt[e]; cur[s * 2 + d] = e + 1; any = true; }
        }
        if (any) edge_add(a, ai, st, counts, pos, d == 0);
      }
    }
    bool leaks = false;
    if (!st.done) {
      // flushContig (misc/peakcall.cpp:224-231).  If the last tag lies beyond `bandwidth` every cell is shifted out: the last one is an end
      // tap (exactly 0), so any open region is closed by the flush itself and the window is clean.  Otherwise bandwidth + 1 - lastTag cells
      // (and possibly an open region) stay for the next contig of this buffer
      leaks = st.bufferPos != 0 && st.bufferPos <= st.bw;
      edge_add(a, ai, st, nullptr, st.bufferPos + w, true);
      if (!st.done) st.stop = 0xFFFFFFFFu;                            // the replay owns the whole unit
    }
    a.stop[ai] = st.stop; a.nCand[ai] = st.nCand; a.nRows[ai] = st.nRows;
    if (!leaks) break;                                                // prepare() links a successor exactly when the last tag lies within `bandwidth`
    carried = true;
  }
}

}  // namespace upb
