/*
 * unipeak_oracle.h — CPU restatement of UniPeak's region-calling hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The CUDA path (unipeak_b200/csrc) never calls into this file.
 *
 * Parity status: PINNED against the reference itself.  The reference ships no tests or golden
 * vectors, so this restatement is checked against the UNMODIFIED reference
 * compiled by oracle/Makefile into oracle/_ref (tests/test_oracle_vs_reference.py, run in the
 * build container) and against fixtures generated from those binaries and committed under
 * tests/golden/ (tests/golden/make_golden.py).
 *
 * Every function names the reference file:line it follows (paths relative to the reference root).
 */
#ifndef UNIPEAK_ORACLE_H
#define UNIPEAK_ORACLE_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Sorted tag-frequency input: for sample s, strand d (0 = '+', 1 = '-') and contig c the entries are
 * pos[off[(s*2+d)*(n_contigs+1)+c] .. off[...+c+1]) — strictly ascending 1-based positions with
 * non-zero counts, offsets (-s) already applied.  This is what misc/format.cpp:503-565 hands to the
 * drivers one tag line at a time. */
typedef struct {
  uint32_t n_contigs;
  const uint32_t* contig_sizes;
  uint32_t n_samples;
  const uint64_t* off;
  const uint32_t* pos;
  const uint32_t* cnt;
} upo_input;

typedef struct {
  uint32_t bandwidth;          /* -b */
  double region_threshold;     /* -r */
  double kurtosis_threshold;   /* -k */
  double corr_threshold;       /* -u (forced to -1 when directional, src/regions.cpp:104-109) */
  double hit_threshold;        /* ALREADY scaled by the number of non-control samples (src/regions.cpp:155) */
  int directional;             /* 1 = two strand buffers (default), 0 = -D */
  const uint8_t* control;      /* [n_samples] 1 = negative control (-e) */
  uint32_t n_coeffs;           /* 0 = unscaled; else == number of non-control samples (-z) */
  const double* coeffs;        /* already normalised as src/regions.cpp:216-229 does */
  double background;           /* tags per position (per strand if directional), src/regions.cpp:209-210 */
  int want_profile;            /* collect the -w profile */
} upo_params;

typedef struct {
  uint8_t forward_label;       /* Region::forward as printed (Q5 makes it differ from the buffer) */
  uint8_t buffer;              /* 0 = forwardBuffer, 1 = reverseBuffer */
  uint32_t contig;
  uint32_t left;
  uint32_t len;
  uint32_t peak_pos;
  double peak_score;
  double kurtosis;
  double corr;                 /* strandCorr(0) */
} upo_region;

typedef struct {
  uint8_t buffer;
  uint32_t contig;
  uint32_t pos;
  double score;
} upo_profile_point;

typedef struct {
  /* rows in the reference's emission order (src/regions.cpp:361-368) */
  size_t n_rows;
  upo_region* rows;
  uint32_t* counts;            /* [n_rows * n_samples] exptSums */
  /* regions finalised by the last flush: counted but never written */
  size_t n_dropped;
  upo_region* dropped;
  uint32_t* dropped_counts;
  uint64_t n_passed;           /* forwardBuffer.nRegions()+reverseBuffer.nRegions() */
  uint64_t n_rejected;
  uint64_t* tags_in_regions;   /* [n_samples] */
  size_t n_profile;
  upo_profile_point* profile;
  int status;                  /* 0 ok; else a reference assert would have fired (see upo_strerror) */
} upo_regions_result;

typedef struct {
  uint32_t bandwidth;
  double region_threshold;
  double kurtosis_threshold;
  double corr_threshold;
  uint32_t hit_threshold;      /* unscaled integer (src/strand_shift.cpp:46,142) */
  uint32_t min_shift;
  uint32_t max_shift;
  uint32_t n_regions_to_test;  /* UShort in the reference */
  double background;
} upo_shift_params;

typedef struct {
  uint64_t n_regions_found;
  uint32_t n_regions_tested;
  uint64_t n_tags_in_regions;
  uint64_t* shift_freq;        /* [max_shift+1] */
  uint32_t best_shift;
  /* per tested region, in test order (diagnostics for parity tests) */
  size_t n_tested_rows;
  uint32_t* tested_contig;
  uint32_t* tested_left;
  uint32_t* tested_len;
  uint32_t* tested_best_shift;
  double* tested_best_corr;
  int status;
} upo_shift_result;

typedef struct {
  uint32_t contig;
  uint32_t start;              /* as written in the table: start > end means a reverse-strand row */
  uint32_t end;
} upo_region_key;

enum {
  UPO_OK = 0,
  UPO_ERR_ASSERT_ORDER = 1,    /* misc/peakcall.cpp:171 assert(pos >= bufferPos_) */
  UPO_ERR_ASSERT_POS = 2,      /* misc/peakcall.cpp:56 assert(pos > lastPos_) */
  UPO_ERR_NO_REGIONS = 3,      /* src/strand_shift.cpp:228-230 */
  UPO_ERR_REVERSE_NONDIR = 4,  /* src/tags_in_regions.cpp:167-170 */
  UPO_ERR_NOMEM = 5,
  UPO_ERR_BAD_ARG = 6
};

const char* upo_strerror (int status);

/* misc/kernel.cpp:12-35 */
void upo_kernel (uint32_t bandwidth, double sum_arg, double* taps /* [2*bandwidth+1] */);

/* src/regions.cpp:304-396 over misc/peakcall.cpp + misc/data.cpp */
int upo_regions (const upo_input* in, const upo_params* p, upo_regions_result* out);
void upo_regions_free (upo_regions_result* out);

/* src/strand_shift.cpp:139-259 */
int upo_strand_shift (const upo_input* in, const upo_shift_params* p, upo_shift_result* out);
void upo_shift_free (upo_shift_result* out);

/* src/tags_in_regions.cpp:131-199; counts_out is [n_regions * in->n_samples] */
int upo_tags_in_regions (const upo_input* in, int directional, uint32_t extension,
                         size_t n_regions, const upo_region_key* regions, uint32_t* counts_out);

/* libstdc++ std::sort with the reference's comparator (src/strand_shift.cpp:26-29,198): sorts idx[] so
 * that sums[idx[.]] is descending, with exactly the tie order std::sort produces.  Implemented in
 * upo_sort.cpp when linked; unipeak_oracle.c carries a C restatement of GCC 13's introsort. */
void upo_sort_by_sum_desc (size_t n, const uint32_t* sums, uint32_t* idx);

#ifdef __cplusplus
}
#endif
#endif
