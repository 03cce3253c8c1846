/*
 * unipeak_oracle.c — CPU restatement (plain C99) of UniPeak's region-calling hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see unipeak_oracle.h).  It follows the reference's streaming algorithm
 * step by step — including the behaviours SURVEY.md Appendix A.4 lists as quirks — so that its
 * results are the reference's results.  It is written from the reference's behaviour, not from its
 * text: the sliding window is a ring of cells instead of a std::deque, regions are flat arrays, and
 * the three drivers read tag arrays instead of parsing wiggle text.
 *
 * Compile with -ffp-contract=off and no -ffast-math: every product and sum below must round
 * separately, exactly as the reference's x86-64 SSE2 build does.
 *
 * Parity: pinned against oracle/_ref (the unmodified reference) — see the header.
 */
#include "unipeak_oracle.h"

#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ------------------------------------------------------------------------------------------ */
/* small helpers                                                                               */
/* ------------------------------------------------------------------------------------------ */

const char* upo_strerror (int status) {
  switch (status) {
    case UPO_OK: return "ok";
    case UPO_ERR_ASSERT_ORDER: return "reference assert: tags not sorted within buffer (peakcall.cpp:171)";
    case UPO_ERR_ASSERT_POS: return "reference assert: position processed twice (peakcall.cpp:56)";
    case UPO_ERR_NO_REGIONS: return "no regions qualified with given settings (strand_shift.cpp:228)";
    case UPO_ERR_REVERSE_NONDIR: return "reverse regions in non-directional analysis (tags_in_regions.cpp:167)";
    case UPO_ERR_NOMEM: return "out of memory";
    case UPO_ERR_BAD_ARG: return "bad argument";
    default: return "unknown";
  }
}

static void* xrealloc (void* p, size_t n) {
  void* q = realloc(p, n ? n : 1);
  if (!q) abort();
  return q;
}

/* ------------------------------------------------------------------------------------------ */
/* Kernel — misc/kernel.cpp:12-35                                                              */
/* ------------------------------------------------------------------------------------------ */

void upo_kernel (uint32_t bandwidth, double sum_arg, double* taps) {
  /* kernel.cpp:12-14: f(x) = 3*(1 - x^2)/4 with pow(x,2) == x*x;
   * kernel.cpp:21-27: taps for i = -bw..bw, running sum in that order;
   * kernel.cpp:30-34: every tap multiplied by scale = sumArg / sum. */
  const int bw = (int)bandwidth;
  double sum = 0;
  int n = 0;
  for (int i = -bw; i <= bw; ++i) {
    const double x = (double)i / (double)bw;
    const double xx = x * x;
    const double one_minus = 1 - xx;
    const double three = 3 * one_minus;
    taps[n] = three / 4;
    sum += taps[n];
    ++n;
  }
  const double scale = sum_arg / sum;
  for (int j = 0; j < n; ++j) taps[j] *= scale;
}

/* ------------------------------------------------------------------------------------------ */
/* Region — misc/data.hpp:54-73, misc/data.cpp:79-193                                          */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  int forward;
  uint8_t buffer;    /* which ProfileBuffer finalised it (bookkeeping only; not in the reference) */
  uint32_t contig;
  uint32_t left;
  uint32_t peak_pos;
  double peak_score;
  uint32_t n_expts;
  size_t len, cap;
  uint32_t** hits;   /* per position: NULL or n_expts counts (owned) */
  double* f;
  double* r;
} region_t;

static region_t* region_new (int forward, uint32_t contig, uint32_t left, uint32_t n_expts) {
  region_t* g = (region_t*)xrealloc(NULL, sizeof *g);
  memset(g, 0, sizeof *g);
  g->forward = forward;
  g->contig = contig;
  g->left = left;
  g->n_expts = n_expts;
  return g;
}

static void region_free (region_t* g) {
  if (!g) return;
  for (size_t i = 0; i < g->len; ++i) free(g->hits[i]);
  free(g->hits);
  free(g->f);
  free(g->r);
  free(g);
}

/* data.cpp:92-102 */
static void region_add_pos (region_t* g, uint32_t* hits, double fs, double rs) {
  if (g->len == g->cap) {
    g->cap = g->cap ? g->cap * 2 : 16;
    g->hits = (uint32_t**)xrealloc(g->hits, g->cap * sizeof *g->hits);
    g->f = (double*)xrealloc(g->f, g->cap * sizeof *g->f);
    g->r = (double*)xrealloc(g->r, g->cap * sizeof *g->r);
  }
  g->hits[g->len] = hits;
  g->f[g->len] = fs;
  g->r[g->len] = rs;
  ++g->len;
  const double score = fs + rs;
  if (g->peak_pos == 0 || score > g->peak_score) {   /* first strict maximum */
    g->peak_pos = g->left + (uint32_t)g->len - 1;
    g->peak_score = score;
  }
}

/* data.cpp:104-114 (HitCount wraps mod 2^32) */
static uint32_t region_sum (const region_t* g) {
  uint32_t result = 0;
  for (size_t i = 0; i < g->len; ++i)
    if (g->hits[i]) for (uint32_t e = 0; e < g->n_expts; ++e) result += g->hits[i][e];
  return result;
}

/* data.cpp:116-131 */
static void region_expt_sums (const region_t* g, uint32_t* out) {
  for (uint32_t e = 0; e < g->n_expts; ++e) out[e] = 0;
  for (size_t i = 0; i < g->len; ++i)
    if (g->hits[i]) for (uint32_t e = 0; e < g->n_expts; ++e) out[e] += g->hits[i][e];
}

/* data.cpp:133-148: offsets are UShort (wrap at 65536), count and sum are HitCount (wrap at 2^32) */
static double region_pos_mean (const region_t* g) {
  uint32_t count = 0, sum = 0;
  uint16_t pos = 0;
  for (size_t i = 0; i < g->len; ++i) {
    if (g->hits[i]) {
      uint32_t pos_count = 0;
      for (uint32_t e = 0; e < g->n_expts; ++e) pos_count += g->hits[i][e];
      count += pos_count;
      sum += pos_count * (uint32_t)pos;
    }
    ++pos;
  }
  return (double)sum / (double)count;
}

/* data.cpp:164-182; pow(d,2) = d*d and pow(d,4) = (d*d)*(d*d) in the reference's -ansi build (Q9) */
static double region_pos_kurtosis (const region_t* g) {
  const double x_bar = region_pos_mean(g);
  uint32_t count = 0;
  double sum2 = 0, sum4 = 0;
  uint16_t pos = 0;
  for (size_t i = 0; i < g->len; ++i) {
    if (g->hits[i]) {
      uint32_t pos_count = 0;
      for (uint32_t e = 0; e < g->n_expts; ++e) pos_count += g->hits[i][e];
      count += pos_count;
      const double d = (double)pos - x_bar;
      const double d2 = d * d;
      const double d4 = d2 * d2;
      sum2 += (double)pos_count * d2;
      sum4 += (double)pos_count * d4;
    }
    ++pos;
  }
  return ((double)count - 1) * sum4 / (sum2 * sum2);
}

/* data.cpp:22-26 */
static double vec_mean (const double* b, size_t n) {
  double sum = 0;
  for (size_t i = 0; i < n; ++i) sum += b[i];
  return sum / (double)n;
}

/* data.cpp:28-32 */
static double vec_sd (const double* b, size_t n, double x_bar) {
  double ssr = 0;
  for (size_t i = 0; i < n; ++i) { const double d = b[i] - x_bar; ssr += d * d; }
  return sqrt(ssr / ((double)n - 1));
}

/* data.cpp:38-58 */
static double vec_corr (const double* a, const double* b, size_t n) {
  const double mean1 = vec_mean(a, n);
  const double mean2 = vec_mean(b, n);
  const double sd1 = vec_sd(a, n, mean1);
  const double sd2 = vec_sd(b, n, mean2);
  double ssr = 0;
  for (size_t i = 0; i < n; ++i) ssr += (a[i] - mean1) * (b[i] - mean2);
  return ssr / (((double)n - 1) * sd1 * sd2);
}

/* data.cpp:184-193: forward scores [0, len-2s) against reverse scores [2s, len) */
static double region_strand_corr (const region_t* g, uint32_t shift) {
  if (g->len > (size_t)(2 * shift + 3)) return vec_corr(g->f, g->r + 2 * shift, g->len - 2 * shift);
  return NAN;
}

/* ------------------------------------------------------------------------------------------ */
/* ProfileBuffer — misc/peakcall.hpp:17-76, misc/peakcall.cpp:33-231                           */
/* ------------------------------------------------------------------------------------------ */

typedef struct { uint32_t* hits; double f, r; } cell_t;

typedef struct {
  region_t** v;
  size_t n, cap;
} region_vec;

static void rv_push (region_vec* rv, region_t* g) {
  if (rv->n == rv->cap) { rv->cap = rv->cap ? rv->cap * 2 : 64; rv->v = (region_t**)xrealloc(rv->v, rv->cap * sizeof *rv->v); }
  rv->v[rv->n++] = g;
}

typedef struct {
  upo_profile_point* v;
  size_t n, cap;
} profile_vec;

typedef struct {
  cell_t* ring;           /* the 2bw+1 window cells; ring[(head+j) % w] is deque element j */
  uint32_t w, head;
  uint32_t bw;
  const double* taps;
  region_t* region;
  region_vec* regions_out;
  double region_thr, kurt_thr, corr_thr, hit_thr;
  int forward;
  uint8_t buffer_id;
  uint32_t contig;
  uint32_t buffer_pos;    /* centre of the window (bufferPos_) */
  uint32_t last_pos;      /* last processed position (lastPos_) */
  uint64_t n_regions, n_rejects, n_contig_regions, n_nonzero;
  uint64_t* tags_in_regions;
  const uint8_t* control;
  const double* coeffs;
  uint32_t n_coeffs;
  uint32_t n_expt;
  profile_vec* profile;
  int status;
} pbuf_t;

static cell_t* pb_cell (pbuf_t* b, uint32_t j) { return &b->ring[(b->head + j) % b->w]; }

/* peakcall.cpp:88-135 */
static pbuf_t* pb_new (const double* taps, uint32_t bw, double region_thr, double kurt_thr, double corr_thr,
                       double hit_thr, int forward, uint8_t buffer_id, const uint8_t* control, uint32_t n_expt,
                       const double* coeffs, uint32_t n_coeffs, region_vec* out, profile_vec* profile) {
  pbuf_t* b = (pbuf_t*)xrealloc(NULL, sizeof *b);
  memset(b, 0, sizeof *b);
  b->w = 2 * bw + 1;
  b->bw = bw;
  b->ring = (cell_t*)xrealloc(NULL, b->w * sizeof *b->ring);
  memset(b->ring, 0, b->w * sizeof *b->ring);
  b->taps = taps;
  b->region = region_new(1, 0, 0, n_expt);   /* peakcall.cpp:103: the first region is ALWAYS forward (Q5) */
  b->regions_out = out;
  b->region_thr = region_thr;
  b->kurt_thr = kurt_thr;
  b->corr_thr = corr_thr;
  b->hit_thr = hit_thr;
  b->forward = forward;
  b->buffer_id = buffer_id;
  b->tags_in_regions = (uint64_t*)xrealloc(NULL, (n_expt ? n_expt : 1) * sizeof(uint64_t));
  memset(b->tags_in_regions, 0, (n_expt ? n_expt : 1) * sizeof(uint64_t));
  b->control = control;
  b->coeffs = coeffs;
  b->n_coeffs = n_coeffs;
  b->n_expt = n_expt;
  b->profile = profile;
  return b;
}

static void pb_free (pbuf_t* b) {
  if (!b) return;
  for (uint32_t j = 0; j < b->w; ++j) free(b->ring[j].hits);
  free(b->ring);
  region_free(b->region);
  free(b->tags_in_regions);
  free(b);
}

/* peakcall.cpp:33-53 */
static void pb_process_region (pbuf_t* b) {
  region_t* g = b->region;
  uint32_t* sums = (uint32_t*)xrealloc(NULL, (b->n_expt ? b->n_expt : 1) * sizeof(uint32_t));
  region_expt_sums(g, sums);
  uint32_t non_control = 0;
  for (uint32_t e = 0; e < b->n_expt; ++e) if (!b->control[e]) non_control += sums[e];
  /* short-circuit order matters only for speed: none of the three tests has side effects */
  int pass = ((double)non_control >= b->hit_thr);
  if (pass) pass = (b->kurt_thr == 0) || (g->len > 1 && region_pos_kurtosis(g) <= b->kurt_thr);
  if (pass) pass = (b->corr_thr <= -1) || (region_strand_corr(g, 0) >= b->corr_thr);
  if (pass) {
    for (uint32_t e = 0; e < b->n_expt; ++e) b->tags_in_regions[e] += sums[e];
    g->buffer = b->buffer_id;
    rv_push(b->regions_out, g);
    ++b->n_regions;
    ++b->n_contig_regions;
  } else {
    region_free(g);
    ++b->n_rejects;
  }
  free(sums);
  b->region = region_new(b->forward, b->contig, 0, b->n_expt);   /* peakcall.cpp:52 */
}

/* peakcall.cpp:55-86 */
static void pb_process_position (pbuf_t* b, uint32_t pos, double fs, double rs, uint32_t* hits) {
  if (!(pos > b->last_pos)) { b->status = UPO_ERR_ASSERT_POS; free(hits); return; }
  const double score = fs + rs;
  if (pos == b->last_pos + 1) {              /* single step */
    if (b->region->left != 0) {              /* currently in a region */
      if (score >= b->region_thr) region_add_pos(b->region, hits, fs, rs);
      else { pb_process_region(b); free(hits); }
    } else {
      if (score >= b->region_thr) { b->region->left = pos; region_add_pos(b->region, hits, fs, rs); }
      else free(hits);
    }
    ++b->last_pos;
  } else {                                   /* leap over zeroes: left is NOT set here (Q10) */
    if (b->region->left != 0) pb_process_region(b);
    if (score >= b->region_thr) region_add_pos(b->region, hits, fs, rs); else free(hits);
  }
  if (score != 0) {
    if (b->profile) {
      profile_vec* pv = b->profile;
      if (pv->n == pv->cap) { pv->cap = pv->cap ? pv->cap * 2 : 1024; pv->v = (upo_profile_point*)xrealloc(pv->v, pv->cap * sizeof *pv->v); }
      pv->v[pv->n].buffer = b->buffer_id;
      pv->v[pv->n].contig = b->contig;
      pv->v[pv->n].pos = pos;
      pv->v[pv->n].score = score;
      ++pv->n;
    }
    ++b->n_nonzero;
  }
  b->last_pos = pos;
}

static uint64_t pb_flush_contig (pbuf_t* b);

/* peakcall.cpp:161-222.  counts == NULL stands for the empty vector flushContig passes. */
static void pb_add (pbuf_t* b, const uint32_t* counts, uint32_t contig, uint32_t pos, int forward) {
  if (b->status) return;
  if (contig != b->contig) {                 /* :164-168 */
    pb_flush_contig(b);
    b->contig = contig;
    b->region->contig = contig;
  }
  if (!(pos >= b->buffer_pos)) { b->status = UPO_ERR_ASSERT_ORDER; return; }   /* :171 */

  /* :172-173 (UShort in the reference; the window never exceeds 65535 cells here) */
  uint32_t n_static = b->w;
  if (pos <= b->buffer_pos + 2 * b->bw) n_static = pos - b->buffer_pos;

  /* :176-184 — note the pop/push sit INSIDE the left-edge test (Q2) */
  if (b->buffer_pos != 0) {
    for (uint32_t i = 0; i < n_static; ++i) {
      if (b->buffer_pos + i > b->bw) {
        cell_t* front = pb_cell(b, 0);
        pb_process_position(b, b->buffer_pos + i - b->bw, front->f, front->r, front->hits);
        front->hits = NULL; front->f = 0; front->r = 0;   /* pop_front + push_back(BufferPos()) */
        b->head = (b->head + 1) % b->w;
        if (b->status) return;
      }
    }
  }

  /* :187-200 — non-control count sum; with coefficients the plain counts are added on top (Q3) */
  double count_sum = 0;
  if (counts) {
    if (b->n_coeffs == 0) {
      for (uint32_t e = 0; e < b->n_expt; ++e) if (!b->control[e]) count_sum += counts[e];
    } else {
      uint32_t k = 0;
      for (uint32_t e = 0; e < b->n_expt && k < b->n_coeffs; ++e) if (!b->control[e]) { count_sum += (double)counts[e] * b->coeffs[k]; ++k; }
      for (uint32_t e = 0; e < b->n_expt; ++e) if (!b->control[e]) count_sum += counts[e];
    }
  }

  /* :203-220 */
  if (count_sum != 0) {
    for (uint32_t j = 0; j < b->w; ++j) {
      cell_t* c = pb_cell(b, j);
      const double term = b->taps[j] * count_sum;
      if (forward) c->f += term; else c->r += term;
    }
    cell_t* centre = pb_cell(b, b->bw);
    if (centre->hits) {
      for (uint32_t e = 0; e < b->n_expt; ++e) centre->hits[e] += counts[e];
    } else {
      centre->hits = (uint32_t*)xrealloc(NULL, (b->n_expt ? b->n_expt : 1) * sizeof(uint32_t));
      memcpy(centre->hits, counts, b->n_expt * sizeof(uint32_t));
    }
  }
  b->buffer_pos = pos;
}

/* peakcall.cpp:224-231 */
static uint64_t pb_flush_contig (pbuf_t* b) {
  pb_add(b, NULL, b->contig, b->buffer_pos + b->w, 1);
  b->buffer_pos = 0;
  b->last_pos = 0;
  const uint64_t result = b->n_contig_regions;
  b->n_contig_regions = 0;
  return result;
}

/* ------------------------------------------------------------------------------------------ */
/* Tag streams — what ParseAlignStream / NondirParseAlignStream deliver (misc/format.cpp)      */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
  const upo_input* in;
  uint32_t sample;
  int nondir;
  /* directional: strand-major walk (all '+' contigs, then all '-'), format.cpp:503-535 */
  uint32_t d, c;
  uint64_t i;
  /* non-directional: two strand cursors merged by (contig, pos), forward first on ties,
   * format.cpp:832-850, :894-915 */
  uint32_t cf, cr;
  uint64_t jf, jr;
  /* current alignment (lastAlign()); count == 0 means exhausted */
  int forward;
  uint32_t contig, pos, count;
} stream_t;

static const uint64_t* st_off (const stream_t* s, uint32_t d) {
  return s->in->off + ((size_t)s->sample * 2 + d) * (s->in->n_contigs + 1);
}

static void st_next (stream_t* s) {
  const uint32_t nc = s->in->n_contigs;
  if (!s->nondir) {
    /* strand-major walk: every '+' contig in table order, then every '-' contig */
    while (s->d < 2) {
      const uint64_t* off = st_off(s, s->d);
      if (s->c == nc) { ++s->d; s->c = 0; if (s->d < 2) s->i = st_off(s, s->d)[0]; continue; }
      if (s->i < off[s->c + 1]) {
        s->forward = (s->d == 0); s->contig = s->c; s->pos = s->in->pos[s->i]; s->count = s->in->cnt[s->i];
        ++s->i;
        return;
      }
      ++s->c;   /* off[c+1] is the start of the next contig, so i is already there */
    }
    s->count = 0; s->contig = nc;   /* format.cpp:693-700: an exhausted stream reports contig == size() */
    return;
  }
  /* non-directional: merge the two strand cursors */
  const uint64_t* of = st_off(s, 0);
  const uint64_t* orv = st_off(s, 1);
  while (s->cf < nc && s->jf >= of[s->cf + 1]) ++s->cf;
  while (s->cr < nc && s->jr >= orv[s->cr + 1]) ++s->cr;
  const int have_f = s->cf < nc, have_r = s->cr < nc;
  int take_f;
  if (have_f && have_r) {
    /* format.cpp:832-850 whichLower(): same contig -> forward unless strictly further; else lower contig */
    if (s->cf == s->cr) take_f = s->in->pos[s->jf] <= s->in->pos[s->jr]; else take_f = s->cf < s->cr;
  } else if (have_f) take_f = 1;
  else if (have_r) take_f = 0;
  else { s->count = 0; s->contig = nc; return; }
  if (take_f) { s->forward = 1; s->contig = s->cf; s->pos = s->in->pos[s->jf]; s->count = s->in->cnt[s->jf]; ++s->jf; }
  else { s->forward = 0; s->contig = s->cr; s->pos = s->in->pos[s->jr]; s->count = s->in->cnt[s->jr]; ++s->jr; }
}

static void st_open (stream_t* s, const upo_input* in, uint32_t sample, int nondir) {
  memset(s, 0, sizeof *s);
  s->in = in;
  s->sample = sample;
  s->nondir = nondir;
  s->i = st_off(s, 0)[0];
  s->jf = st_off(s, 0)[0];
  s->jr = st_off(s, 1)[0];
  st_next(s);   /* the drivers prime every stream with one readAlign(), regions.cpp:199 */
}

/* ------------------------------------------------------------------------------------------ */
/* regions driver — src/regions.cpp:304-396                                                    */
/* ------------------------------------------------------------------------------------------ */

static void fill_row (const region_t* g, upo_region* row, uint32_t* counts) {
  row->forward_label = (uint8_t)g->forward;
  row->buffer = g->buffer;
  row->contig = g->contig;
  row->left = g->left;
  row->len = (uint32_t)g->len;
  row->peak_pos = g->peak_pos;
  row->peak_score = g->peak_score;
  row->kurtosis = region_pos_kurtosis(g);      /* format.cpp:1152 recomputes it at write time */
  row->corr = region_strand_corr(g, 0);        /* format.cpp:1151 */
  region_expt_sums(g, counts);
}

/* one pass of the merge loop body for (contig, pos): regions.cpp:327-359.  Returns nextPos. */
static uint32_t merge_step (stream_t* st, uint32_t n, uint32_t contig, uint32_t pos, uint32_t pos_limit,
                            uint32_t* fwd_hits, uint32_t* rev_hits, int* found_f, int* found_r) {
  uint32_t next_pos = pos_limit + 1;
  *found_f = 0; *found_r = 0;
  for (uint32_t s = 0; s < n; ++s) {
    fwd_hits[s] = 0; rev_hits[s] = 0;
    while (st[s].count != 0 && st[s].pos == pos && st[s].contig == contig) {   /* :341 — no strand test (Q7) */
      if (st[s].forward) { fwd_hits[s] += st[s].count; *found_f = 1; }
      else { rev_hits[s] += st[s].count; *found_r = 1; }
      st_next(&st[s]);
    }
    if (st[s].count != 0 && st[s].contig == contig && st[s].pos < next_pos) next_pos = st[s].pos;   /* :353 */
  }
  return next_pos;
}

int upo_regions (const upo_input* in, const upo_params* p, upo_regions_result* out) {
  memset(out, 0, sizeof *out);
  if (!in || !p || p->bandwidth == 0 || p->bandwidth > 32767) return out->status = UPO_ERR_BAD_ARG;
  const uint32_t n = in->n_samples;
  double* taps = (double*)xrealloc(NULL, (2 * (size_t)p->bandwidth + 1) * sizeof(double));
  upo_kernel(p->bandwidth, 1 / p->background, taps);   /* regions.cpp:232 */

  region_vec pending = {0, 0, 0};
  profile_vec prof = {0, 0, 0};
  uint8_t* no_control = (uint8_t*)xrealloc(NULL, n ? n : 1);
  memset(no_control, 0, n ? n : 1);
  const uint8_t* control = p->control ? p->control : no_control;
  pbuf_t* fb = pb_new(taps, p->bandwidth, p->region_threshold, p->kurtosis_threshold, p->corr_threshold, p->hit_threshold,
                      1, 0, control, n, p->coeffs, p->n_coeffs, &pending, p->want_profile ? &prof : NULL);
  pbuf_t* rb = pb_new(taps, p->bandwidth, p->region_threshold, p->kurtosis_threshold, p->corr_threshold, p->hit_threshold,
                      0, 1, control, n, p->coeffs, p->n_coeffs, &pending, p->want_profile ? &prof : NULL);

  stream_t* st = (stream_t*)xrealloc(NULL, (n ? n : 1) * sizeof *st);
  for (uint32_t s = 0; s < n; ++s) st_open(&st[s], in, s, !p->directional);
  uint32_t* fwd_hits = (uint32_t*)xrealloc(NULL, (n ? n : 1) * sizeof(uint32_t));
  uint32_t* rev_hits = (uint32_t*)xrealloc(NULL, (n ? n : 1) * sizeof(uint32_t));

  size_t rows_cap = 0;
  uint32_t contig = 0;
  int forward_pass = 1;
  int status = 0;
  while (contig < in->n_contigs && !status) {                    /* :311 */
    const uint32_t pos_limit = in->contig_sizes[contig];
    uint32_t pos = 1;
    while (pos <= pos_limit && !status) {                        /* :326 */
      int found_f, found_r;
      const uint32_t next_pos = merge_step(st, n, contig, pos, pos_limit, fwd_hits, rev_hits, &found_f, &found_r);
      if (found_f) pb_add(fb, fwd_hits, contig, pos, 1);         /* :356 */
      if (found_r) { if (p->directional) pb_add(rb, rev_hits, contig, pos, 0); else pb_add(fb, rev_hits, contig, pos, 0); }   /* :357-359 */
      status = fb->status ? fb->status : rb->status;
      /* :361-368 — write and free whatever has been finalised so far */
      if (pending.n != 0) {
        if (out->n_rows + pending.n > rows_cap) {
          rows_cap = (out->n_rows + pending.n) * 2;
          out->rows = (upo_region*)xrealloc(out->rows, rows_cap * sizeof *out->rows);
          out->counts = (uint32_t*)xrealloc(out->counts, rows_cap * (n ? n : 1) * sizeof(uint32_t));
        }
        for (size_t k = 0; k < pending.n; ++k) {
          region_t* g = pending.v[k];
          fill_row(g, &out->rows[out->n_rows], out->counts + out->n_rows * n);
          ++out->n_rows;
          region_free(g);
        }
        pending.n = 0;
      }
      pos = next_pos;                                            /* :370 */
    }
    pb_flush_contig(fb);                                         /* :373 */
    pb_flush_contig(rb);                                         /* :374 */
    status = status ? status : (fb->status ? fb->status : rb->status);
    ++contig;
    if (p->directional && contig == in->n_contigs && forward_pass) { contig = 0; forward_pass = 0; }   /* :387-390 */
  }

  /* whatever the last flushes finalised is never written (Q6) */
  out->n_dropped = pending.n;
  if (pending.n) {
    out->dropped = (upo_region*)xrealloc(NULL, pending.n * sizeof *out->dropped);
    out->dropped_counts = (uint32_t*)xrealloc(NULL, pending.n * (n ? n : 1) * sizeof(uint32_t));
    for (size_t k = 0; k < pending.n; ++k) {
      fill_row(pending.v[k], &out->dropped[k], out->dropped_counts + k * n);
      region_free(pending.v[k]);
    }
  }
  free(pending.v);

  out->n_passed = fb->n_regions + rb->n_regions;                 /* :396 */
  out->n_rejected = fb->n_rejects + rb->n_rejects;
  out->tags_in_regions = (uint64_t*)xrealloc(NULL, (n ? n : 1) * sizeof(uint64_t));
  for (uint32_t s = 0; s < n; ++s) out->tags_in_regions[s] = fb->tags_in_regions[s] + rb->tags_in_regions[s];   /* :402 */
  out->n_profile = prof.n;
  out->profile = prof.v;
  out->status = status;

  pb_free(fb); pb_free(rb);
  free(st); free(fwd_hits); free(rev_hits); free(taps); free(no_control);
  return status;
}

void upo_regions_free (upo_regions_result* out) {
  if (!out) return;
  free(out->rows); free(out->counts); free(out->dropped); free(out->dropped_counts);
  free(out->tags_in_regions); free(out->profile);
  memset(out, 0, sizeof *out);
}

/* ------------------------------------------------------------------------------------------ */
/* std::sort restatement — libstdc++ (GCC 13) bits/stl_algo.h: introsort + final insertion sort */
/* The reference calls std::sort with compareRegionCounts (src/strand_shift.cpp:26-29,198); the   */
/* sort is unstable, so the order of equal sums is whatever this exact algorithm leaves (Q12).    */
/* ------------------------------------------------------------------------------------------ */

typedef struct { const uint32_t* sums; } cmp_t;
#define CMP(c, a, b) ((c)->sums[(a)] > (c)->sums[(b)])

static void is_swap (uint32_t* a, uint32_t* b) { const uint32_t t = *a; *a = *b; *b = t; }

static void is_push_heap (uint32_t* first, ptrdiff_t hole, ptrdiff_t top, uint32_t value, const cmp_t* c) {
  ptrdiff_t parent = (hole - 1) / 2;
  while (hole > top && CMP(c, first[parent], value)) {
    first[hole] = first[parent];
    hole = parent;
    parent = (hole - 1) / 2;
  }
  first[hole] = value;
}

static void is_adjust_heap (uint32_t* first, ptrdiff_t hole, ptrdiff_t len, uint32_t value, const cmp_t* c) {
  const ptrdiff_t top = hole;
  ptrdiff_t child = hole;
  while (child < (len - 1) / 2) {
    child = 2 * (child + 1);
    if (CMP(c, first[child], first[child - 1])) --child;
    first[hole] = first[child];
    hole = child;
  }
  if ((len & 1) == 0 && child == (len - 2) / 2) {
    child = 2 * (child + 1);
    first[hole] = first[child - 1];
    hole = child - 1;
  }
  is_push_heap(first, hole, top, value, c);
}

static void is_heapsort (uint32_t* first, uint32_t* last, const cmp_t* c) {
  const ptrdiff_t len = last - first;
  if (len >= 2) {
    ptrdiff_t parent = (len - 2) / 2;
    for (;;) {
      const uint32_t v = first[parent];
      is_adjust_heap(first, parent, len, v, c);
      if (parent == 0) break;
      --parent;
    }
  }
  while (last - first > 1) {
    --last;
    const uint32_t v = *last;
    *last = *first;
    is_adjust_heap(first, 0, last - first, v, c);
  }
}

static void is_median_to_first (uint32_t* result, uint32_t* a, uint32_t* b, uint32_t* cc, const cmp_t* c) {
  if (CMP(c, *a, *b)) {
    if (CMP(c, *b, *cc)) is_swap(result, b);
    else if (CMP(c, *a, *cc)) is_swap(result, cc);
    else is_swap(result, a);
  } else if (CMP(c, *a, *cc)) is_swap(result, a);
  else if (CMP(c, *b, *cc)) is_swap(result, cc);
  else is_swap(result, b);
}

static uint32_t* is_partition (uint32_t* first, uint32_t* last, uint32_t* pivot, const cmp_t* c) {
  for (;;) {
    while (CMP(c, *first, *pivot)) ++first;
    --last;
    while (CMP(c, *pivot, *last)) --last;
    if (!(first < last)) return first;
    is_swap(first, last);
    ++first;
  }
}

static void is_introsort_loop (uint32_t* first, uint32_t* last, long depth, const cmp_t* c) {
  while (last - first > 16) {
    if (depth == 0) { is_heapsort(first, last, c); return; }
    --depth;
    uint32_t* mid = first + (last - first) / 2;
    is_median_to_first(first, first + 1, mid, last - 1, c);
    uint32_t* cut = is_partition(first + 1, last, first, c);
    is_introsort_loop(cut, last, depth, c);
    last = cut;
  }
}

static void is_unguarded_linear_insert (uint32_t* last, const cmp_t* c) {
  const uint32_t val = *last;
  uint32_t* next = last - 1;
  while (CMP(c, val, *next)) { *last = *next; last = next; --next; }
  *last = val;
}

static void is_insertion_sort (uint32_t* first, uint32_t* last, const cmp_t* c) {
  if (first == last) return;
  for (uint32_t* i = first + 1; i != last; ++i) {
    if (CMP(c, *i, *first)) {
      const uint32_t val = *i;
      memmove(first + 1, first, (size_t)(i - first) * sizeof *first);
      *first = val;
    } else is_unguarded_linear_insert(i, c);
  }
}

#ifndef UPO_EXTERNAL_SORT
void upo_sort_by_sum_desc (size_t n, const uint32_t* sums, uint32_t* idx) {
  if (n == 0) return;
  cmp_t c; c.sums = sums;
  long lg = 0;
  for (size_t m = n; m > 1; m >>= 1) ++lg;
  is_introsort_loop(idx, idx + n, lg * 2, &c);
  if (n > 16) {
    is_insertion_sort(idx, idx + 16, &c);
    for (uint32_t* i = idx + 16; i != idx + n; ++i) is_unguarded_linear_insert(i, &c);
  } else is_insertion_sort(idx, idx + n, &c);
}
#endif

/* ------------------------------------------------------------------------------------------ */
/* strand_shift driver — src/strand_shift.cpp:139-259                                          */
/* ------------------------------------------------------------------------------------------ */

int upo_strand_shift (const upo_input* in, const upo_shift_params* p, upo_shift_result* out) {
  memset(out, 0, sizeof *out);
  if (!in || !p || p->bandwidth == 0 || p->bandwidth > 32767) return out->status = UPO_ERR_BAD_ARG;
  const uint32_t n = in->n_samples;
  double* taps = (double*)xrealloc(NULL, (2 * (size_t)p->bandwidth + 1) * sizeof(double));
  upo_kernel(p->bandwidth, 1 / p->background, taps);              /* :137 */
  uint8_t* control = (uint8_t*)xrealloc(NULL, n ? n : 1);
  memset(control, 0, n ? n : 1);
  region_vec regions = {0, 0, 0};
  /* :142 — one buffer, correlation filter off, unscaled integer hit threshold (Q11) */
  pbuf_t* b = pb_new(taps, p->bandwidth, p->region_threshold, p->kurtosis_threshold, -1, (double)p->hit_threshold,
                     1, 0, control, n, NULL, 0, &regions, NULL);
  stream_t* st = (stream_t*)xrealloc(NULL, (n ? n : 1) * sizeof *st);
  for (uint32_t s = 0; s < n; ++s) st_open(&st[s], in, s, 1);     /* :120 always NondirParseAlignStream */
  uint32_t* fwd_hits = (uint32_t*)xrealloc(NULL, (n ? n : 1) * sizeof(uint32_t));
  uint32_t* rev_hits = (uint32_t*)xrealloc(NULL, (n ? n : 1) * sizeof(uint32_t));

  int status = 0;
  for (uint32_t contig = 0; contig < in->n_contigs && !status; ++contig) {   /* :144-190 */
    const uint32_t pos_limit = in->contig_sizes[contig];
    uint32_t pos = 1;
    while (pos <= pos_limit && !status) {
      int found_f, found_r;
      const uint32_t next_pos = merge_step(st, n, contig, pos, pos_limit, fwd_hits, rev_hits, &found_f, &found_r);
      if (found_f) pb_add(b, fwd_hits, contig, pos, 1);
      if (found_r) pb_add(b, rev_hits, contig, pos, 0);
      status = b->status;
      pos = next_pos;
    }
    pb_flush_contig(b);
    status = status ? status : b->status;
  }
  out->n_regions_found = regions.n;

  if (!status) {
    /* :198 sort by Region::sum() descending with std::sort's tie order */
    uint32_t* sums = (uint32_t*)xrealloc(NULL, (regions.n ? regions.n : 1) * sizeof(uint32_t));
    uint32_t* idx = (uint32_t*)xrealloc(NULL, (regions.n ? regions.n : 1) * sizeof(uint32_t));
    for (size_t k = 0; k < regions.n; ++k) { sums[k] = region_sum(regions.v[k]); idx[k] = (uint32_t)k; }
    upo_sort_by_sum_desc(regions.n, sums, idx);

    /* :201-226 */
    const uint16_t to_test = (uint16_t)p->n_regions_to_test;
    uint16_t tested = 0;
    out->shift_freq = (uint64_t*)xrealloc(NULL, ((size_t)p->max_shift + 1) * sizeof(uint64_t));
    memset(out->shift_freq, 0, ((size_t)p->max_shift + 1) * sizeof(uint64_t));
    const size_t cap = to_test ? to_test : 1;
    out->tested_contig = (uint32_t*)xrealloc(NULL, cap * sizeof(uint32_t));
    out->tested_left = (uint32_t*)xrealloc(NULL, cap * sizeof(uint32_t));
    out->tested_len = (uint32_t*)xrealloc(NULL, cap * sizeof(uint32_t));
    out->tested_best_shift = (uint32_t*)xrealloc(NULL, cap * sizeof(uint32_t));
    out->tested_best_corr = (double*)xrealloc(NULL, cap * sizeof(double));
    for (size_t k = 0; tested < to_test && k < regions.n; ++k) {
      const region_t* g = regions.v[idx[k]];
      if (g->len > (size_t)(2 * p->max_shift + 3)) {
        uint32_t best_shift = 0;
        double best_corr = -1;
        for (uint32_t shift = 0; shift <= p->max_shift; ++shift) {
          const double c = region_strand_corr(g, shift);
          if (c > best_corr) { best_shift = shift; best_corr = c; }   /* lowest shift wins ties; NaN never wins */
        }
        if (best_corr >= p->corr_threshold) {
          ++out->shift_freq[best_shift];
          out->n_tags_in_regions += sums[idx[k]];
          out->tested_contig[tested] = g->contig;
          out->tested_left[tested] = g->left;
          out->tested_len[tested] = (uint32_t)g->len;
          out->tested_best_shift[tested] = best_shift;
          out->tested_best_corr[tested] = best_corr;
          ++tested;
        }
      }
    }
    out->n_regions_tested = tested;
    out->n_tested_rows = tested;
    free(sums); free(idx);

    if (tested == 0) status = UPO_ERR_NO_REGIONS;                /* :228-230 */
    else {
      /* :239-248 smooth the histogram of maxima with Kernel(5) (sum 1) */
      double mode_kernel[11];
      upo_kernel(5, 1, mode_kernel);
      const size_t m = (size_t)p->max_shift + 1;
      double* density = (double*)xrealloc(NULL, m * sizeof(double));
      for (size_t i = 0; i < m; ++i) density[i] = 0;
      for (size_t i = 0; i < m; ++i)
        for (int j = 0; j < 11; ++j) {
          const long k = (long)i - 5 + j;
          if (k >= 0 && k < (long)m) density[k] += (double)out->shift_freq[i] * mode_kernel[j];
        }
      /* :251-258 — UShort i; size()-5 is unsigned arithmetic */
      double best_density = 0;
      uint32_t best = 0;
      if (m >= 5) for (size_t i = p->min_shift; i < m - 5; ++i) if (density[i] > best_density) { best = (uint32_t)i; best_density = density[i]; }
      out->best_shift = best;
      free(density);
    }
  }

  for (size_t k = 0; k < regions.n; ++k) region_free(regions.v[k]);
  free(regions.v);
  pb_free(b);
  free(st); free(fwd_hits); free(rev_hits); free(taps); free(control);
  out->status = status;
  return status;
}

void upo_shift_free (upo_shift_result* out) {
  if (!out) return;
  free(out->shift_freq); free(out->tested_contig); free(out->tested_left); free(out->tested_len);
  free(out->tested_best_shift); free(out->tested_best_corr);
  memset(out, 0, sizeof *out);
}

/* ------------------------------------------------------------------------------------------ */
/* tags_in_regions driver — src/tags_in_regions.cpp:131-199                                    */
/* ------------------------------------------------------------------------------------------ */

int upo_tags_in_regions (const upo_input* in, int directional, uint32_t extension,
                         size_t n_regions, const upo_region_key* regions, uint32_t* counts_out) {
  const uint32_t n = in->n_samples;
  stream_t* st = (stream_t*)xrealloc(NULL, (n ? n : 1) * sizeof *st);
  for (uint32_t s = 0; s < n; ++s) st_open(&st[s], in, s, !directional);
  int status = 0;
  for (size_t k = 0; k < n_regions; ++k) {
    const uint32_t contig = regions[k].contig, start = regions[k].start, end = regions[k].end;
    const int forward = end >= start;                             /* :166 */
    if (!forward && !directional) { status = UPO_ERR_REVERSE_NONDIR; break; }
    const uint32_t lo = forward ? start : end;
    const uint32_t left = lo < extension ? 0 : lo - extension;    /* :171 */
    const uint32_t right = (forward ? end : start) + extension;   /* :172 */
    for (uint32_t s = 0; s < n; ++s) {
      stream_t* a = &st[s];
      /* :185 — skip everything before the region; a strand mismatch also skips (Q13 i) */
      while (a->count != 0 && (forward != a->forward || a->contig < contig || (a->contig == contig && a->pos < left))) st_next(a);
      uint32_t hits = 0;
      /* :187-190 — neither strand nor count is tested here (Q13 ii); an exhausted stream reports
       * contig == n_contigs (format.cpp:693-700), which ends the loop */
      while (a->contig == contig && a->pos <= right) { hits += a->count; st_next(a); }
      counts_out[k * n + s] = hits;
    }
  }
  free(st);
  return status;
}
