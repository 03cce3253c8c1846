/*
 * unipeak_b200.h — C-ABI of the B200-native UniPeak region-calling hot path.
 *
 * The reference (jwfoley/UniPeak) has no plugin or FFI layer: its hot path is the C++ class surface
 * Kernel / ProfileBuffer / Region, driven one tag position at a time from the three main()s.  A
 * streaming add() cannot be a device call, so this boundary is batch-shaped: the caller hands over
 * what the reference's wiggle front end would have streamed (sorted, unique per-strand tag positions
 * with counts) and gets back what the reference's drivers would have written.
 * Every entry point names the reference interface it replaces (paths relative to the reference root).
 *
 * Conventions: plain C types only; every function returns UPB_OK (0) or a negative error code and
 * never throws, exits or aborts; upb_last_error() gives the message; one context per host thread;
 * all host buffers are the caller's unless a upb_*_free function is named; no CPU fallback exists —
 * without a CUDA device upb_create fails.
 */
#ifndef UNIPEAK_B200_H
#define UNIPEAK_B200_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UPB_VERSION 1

enum {
  UPB_OK = 0,
  UPB_ERR_CUDA = -1,            /* a CUDA runtime call failed (message has the cudaError string) */
  UPB_ERR_BAD_ARG = -2,
  UPB_ERR_NO_DEVICE = -3,       /* no CUDA device: there is deliberately no CPU path */
  UPB_ERR_NOT_LOADED = -4,      /* contigs / tags not set before a call that needs them */
  UPB_ERR_UNSORTED = -5,        /* reference would assert: misc/peakcall.cpp:171 */
  UPB_ERR_LEFT_EDGE = -6,       /* a tag lies within `bandwidth` of a contig start and the call is one for which the
                                   reference's left-edge behaviour (misc/peakcall.cpp:176-184) is not replayed:
                                   the -w profile, a strand-shift sweep over a replayed region, a contig with nothing else */
  UPB_ERR_REFERENCE_ABORTS = -7,/* the reference's merge loop would trip its own assert on this input
                                   (several directional samples whose '-' track is read in pass 1) */
  UPB_ERR_NO_REGIONS = -8,      /* src/strand_shift.cpp:228-230 "no regions qualified" */
  UPB_ERR_REVERSE_NONDIR = -9,  /* src/tags_in_regions.cpp:167-170 */
  UPB_ERR_NOMEM = -10,
  UPB_ERR_UNSUPPORTED = -11
};

typedef struct upb_ctx upb_ctx;

/* ---- context ------------------------------------------------------------------------------- */
int upb_create (int device, upb_ctx** out);
void upb_destroy (upb_ctx* ctx);
const char* upb_last_error (const upb_ctx* ctx);   /* ctx may be NULL: last error of upb_create on this thread */
int upb_device_count (void);

/* pinned host memory for the arrays handed to upb_load_tags (optional; any host memory works) */
int upb_host_alloc (void** out, size_t bytes);
void upb_host_free (void* p);
/* page-lock and map host memory the caller owns — e.g. a POSIX shared-memory segment shared by the ranks of a multi-GPU job, so that
 * every rank's upb_scatter_rows writes its rows straight into ONE host table (the destination of the region-list gather);
 * *device_ptr is the address to hand to upb_scatter_rows */
int upb_host_register (void* p, size_t bytes, void** device_ptr);
void upb_host_unregister (void* p);

/* ---- inputs -------------------------------------------------------------------------------- */
/* Replaces ContigTable (misc/data.hpp:75-100, filled by parseContigTable, misc/format.cpp:27-57):
 * sizes in table order; contig index = row. */
int upb_set_contigs (upb_ctx* ctx, uint32_t n_contigs, const uint32_t* sizes);

/* Replaces the per-sample tag streams ParseAlignStream / NondirParseAlignStream deliver
 * (misc/format.cpp:503-565, :798-915; consumed at src/regions.cpp:335-354).  For sample s, strand d
 * (0 '+', 1 '-') and contig c the entries are pos[off[(s*2+d)*(n_contigs+1)+c] .. off[..+c+1]):
 * strictly ascending 1-based positions (the -s offset already applied, out-of-bounds tags dropped, as
 * misc/format.cpp:654-678 does) with non-zero counts.  `off` has n_samples*2*(n_contigs+1) entries.
 * Copies to the device and builds the per-tile search index; the arrays may be freed afterwards.
 * A rank of a multi-GPU job loads only the contigs it owns (empty segments elsewhere). */
int upb_load_tags (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt);

/* The same tag streams in the compact host form (3 bytes per entry instead of 8 — the load is PCIe-bound): dpos[e] = position
 * minus the previous position of the same (sample, strand, contig) segment, cnt8[e] = count; 0xFFFF / 0xFF are escapes whose
 * absolute position / full count sit in `exceptions` (sorted by entry index; pos or cnt is 0 where that field is not escaped).
 * The FIRST entry of every segment must be a position escape.  Decoded on the device (segmented scan), validated exactly as
 * upb_load_tags validates.  upb_pack_tags is the host-only encoder a front end runs while it parses (pass exc_cap = 0 to
 * count the exceptions first; returns UPB_ERR_NOMEM with *n_exc_out set when exc_cap is too small). */
typedef struct { uint64_t index; uint32_t pos; uint32_t cnt; } upb_packed_exception;
int upb_load_tags_packed (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint16_t* dpos, const uint8_t* cnt8,
                          uint64_t n_exceptions, const upb_packed_exception* exceptions);
int upb_pack_tags (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                   uint16_t* dpos_out, uint8_t* cnt8_out, upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out);
/* The 2-byte form: one u16 per entry, position delta in the low 12 bits (0xFFF = escape), count in the high 4 (0xF = escape);
 * exceptions as above.  On RNA-seq / ChIP-seq-like data well under 1 % of the entries escape, so the load moves 2.1 instead of
 * 3 bytes per entry.  Same decoder, same validation.  (Added at the end of round 1: encoder and a numpy decoder are tested on the
 * CPU, the device side has one GPU test that has not run yet — upb_load_tags_packed stays the default of bench.py until it has.) */
int upb_load_tags_packed2 (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint16_t* words,
                           uint64_t n_exceptions, const upb_packed_exception* exceptions);
int upb_pack_tags2 (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                    uint16_t* words_out, upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out);

/* ---- bin/regions --------------------------------------------------------------------------- */
/* What src/regions.cpp:104-232 derives from its flags before the hot loop. */
typedef struct {
  uint32_t bandwidth;           /* -b; Kernel(bandwidth, 1/background), misc/kernel.cpp:16-35 */
  double region_threshold;      /* -r (must be > 0) */
  double kurtosis_threshold;    /* -k (0 disables) */
  double corr_threshold;        /* -u; callers pass -1 when directional (src/regions.cpp:104-109) */
  double hit_threshold;         /* -t ALREADY multiplied by the number of non-control samples (src/regions.cpp:155) */
  int directional;              /* 1: forwardBuffer + reverseBuffer; 0: -D, one buffer sees both strands */
  const uint8_t* control;       /* [n_samples], 1 = negative control (-e); NULL = none */
  uint32_t n_coeffs;            /* 0, or the number of non-control samples (-z) */
  const double* coeffs;         /* already normalised as src/regions.cpp:216-229 */
  double background;            /* tags per mappable position, halved if directional (src/regions.cpp:209-210) */
  int reference_order;          /* 1: rows come back in the reference's emission order with its labels and its
                                   dropped tail; needs every contig loaded in this context.
                                   0: canonical order (buffer, contig, left), nothing dropped.
                                   2: as 0, but the rows stay in device memory only — for a rank of a multi-GPU job that hands them
                                      to upb_scatter_rows; n_rows, the totals and the unit_* arrays of the returned table are valid,
                                      the per-row columns are not filled. */
} upb_region_params;

/* One row per accepted region == one Region the reference pushes into regionsOut (misc/peakcall.cpp:45)
 * and FormatOutStream::write(Region) prints (misc/format.cpp:1141-1162).  Library-allocated: every array
 * points into one pinned host slab the device wrote directly; upb_free_region_table hands the slab back to
 * the context's pool, so free the table before destroying the context. */
typedef struct {
  uint64_t n_rows;
  uint32_t n_samples;
  uint8_t* buffer;              /* 0 forwardBuffer, 1 reverseBuffer */
  uint8_t* forward_label;       /* Region::forward as the reference prints it (differs from buffer only for Q5) */
  uint8_t* dropped;             /* 1: the reference finalises this region in its last flush and never writes it (Q6) */
  uint32_t* contig;
  uint32_t* left;               /* Region::left */
  uint32_t* len;                /* Region::counts->size() */
  uint32_t* peak;               /* Region::peakPos */
  uint32_t* hit_sum;            /* nonControlSum of misc/peakcall.cpp:36-40 (== Region::sum() when there are no controls) */
  uint64_t* cand_index;         /* ordinal of this region among ALL candidate regions of its (buffer, contig) */
  double* peak_score;           /* Region::peakScore */
  double* kurtosis;             /* Region::posKurtosis() */
  double* corr;                 /* Region::strandCorr(0); NaN where the reference would not have computed it */
  uint32_t* counts;             /* [n_rows * n_samples] Region::exptSums() — the regions x experiments matrix */
  uint64_t n_passed;            /* ProfileBuffer::nRegions() summed over buffers (includes dropped rows) */
  uint64_t n_rejected;          /* ProfileBuffer::nRegionRejects() */
  uint64_t* tags_in_regions;    /* [n_samples] ProfileBuffer::nTagsInRegions() */
  uint64_t* unit_candidates;    /* [2 * n_contigs] candidate regions per (buffer, contig) */
  uint32_t* unit_last_tag;      /* [2 * n_contigs] last add() position per (buffer, contig), 0 if none */
  uint64_t* unit_rows;          /* [2 * n_contigs] accepted rows per (buffer, contig) */
  void* _slab;                  /* private: the slab, its size and its owner */
  uint64_t _slab_bytes;
  void* _ctx;
} upb_region_table;

int upb_call_regions (upb_ctx* ctx, const upb_region_params* params, upb_region_table* out);
void upb_free_region_table (upb_region_table* t);

/* Host-only: put canonical rows (reference_order == 0), e.g. gathered from several ranks, into the
 * reference's emission order and set forward_label / dropped (src/regions.cpp:311-391).  `first_pos`
 * and `last_pos` are [n_samples*2*n_contigs] first/last tag position per (sample, strand, contig),
 * 0 where empty.  perm_out receives n_rows row indices. */
int upb_reference_order (uint32_t n_contigs, uint32_t n_samples, int directional, uint32_t bandwidth,
                         const uint32_t* first_pos, const uint32_t* last_pos,
                         upb_region_table* t, uint64_t* perm_out);

/* ---- several GPUs: one table out of the ranks' tables ---------------------------------------- */
/* Every rank calls upb_call_regions on the contigs it loaded (reference_order = 0), the ranks sum their unit_rows
 * (one small all-reduce), and each rank scatters its rows to their place in ONE table-wide device slab that the
 * caller zeroed (or, when every row is written by exactly one rank, simply allocated — device memory or mapped host memory from
 * upb_host_register): slab columns are laid out as upb_slab_layout says, rows in canonical order (buffer, contig, left).
 * Summing the slabs over the ranks (ncclReduce on 64-bit words; the row sets are disjoint) gives rank 0 the whole
 * table without a host round trip; upb_table_from_slab wraps a host copy of it. */
enum { UPB_COL_CAND_INDEX = 0, UPB_COL_PEAK_SCORE, UPB_COL_KURTOSIS, UPB_COL_CORR, UPB_COL_CONTIG, UPB_COL_LEFT, UPB_COL_LEN,
       UPB_COL_PEAK, UPB_COL_HIT_SUM, UPB_COL_COUNTS, UPB_COL_BUFFER, UPB_COL_END /* = bytes of the slab */, UPB_N_COLS };
int upb_slab_layout (uint64_t n_rows, uint32_t n_samples, uint64_t* offsets_out /* [UPB_N_COLS] byte offsets */);
int upb_scatter_rows (upb_ctx* ctx, uint64_t total_rows, const uint64_t* unit_rows_total /* [2 * n_contigs] summed over ranks */,
                      const uint64_t* unit_rows_local /* [2 * n_contigs] of this rank's last upb_call_regions (its table's unit_rows) */,
                      void* d_slab /* device, upb_slab_layout(total_rows) bytes, zeroed */);
/* Host-only: column pointers of `out` into a host copy of the summed slab (the caller keeps the memory alive; do not
 * pass such a table to upb_free_region_table). forward_label / dropped / unit arrays are left NULL. */
int upb_table_from_slab (void* h_slab, uint64_t n_rows, uint32_t n_samples, upb_region_table* out);

/* ---- bin/regions -w (profile) -------------------------------------------------------------- */
/* Replaces the FormatOutStream::write(PosScore) calls of ProfileBuffer::processPosition
 * (misc/peakcall.cpp:81-84): every position with score != 0 of one (buffer, contig), ascending.
 * Library-allocated arrays; free with upb_free_profile. */
typedef struct {
  uint64_t n;
  uint32_t* pos;
  double* score;
} upb_profile;
int upb_profile_contig (upb_ctx* ctx, const upb_region_params* params, int buffer, uint32_t contig, upb_profile* out);
void upb_free_profile (upb_profile* p);

/* ---- bin/strand_shift ---------------------------------------------------------------------- */
/* What src/strand_shift.cpp:52-137 derives from its flags. */
typedef struct {
  uint32_t bandwidth;
  double region_threshold;
  double kurtosis_threshold;
  double corr_threshold;
  uint32_t hit_threshold;       /* unscaled integer (src/strand_shift.cpp:46,142) */
  uint32_t min_shift;           /* -n */
  uint32_t max_shift;           /* -x */
  uint32_t n_regions_to_test;   /* -g (UShort in the reference) */
  double background;            /* total tags / mappable (src/strand_shift.cpp:133) */
} upb_shift_params;

/* What src/strand_shift.cpp:196-300 reports. */
typedef struct {
  uint64_t n_regions_found;
  uint32_t n_regions_tested;
  uint64_t n_tags_in_regions;
  uint32_t best_shift;
  uint32_t n_freq;              /* max_shift + 1 */
  uint64_t* shift_freq;         /* shiftFreq */
  /* per tested region, in test order */
  uint32_t* tested_contig;
  uint32_t* tested_left;
  uint32_t* tested_len;
  uint32_t* tested_best_shift;
  double* tested_best_corr;
} upb_shift_report;

int upb_strand_shift (upb_ctx* ctx, const upb_shift_params* params, upb_shift_report* out);
void upb_free_shift_report (upb_shift_report* r);

/* ---- bin/strand_shift over contig-sharded contexts (one per GPU) -----------------
 * The same steps as upb_strand_shift, cut where the ranks have to exchange something:
 *   1. upb_shift_regions        the Region vector of src/strand_shift.cpp:143-190 for the contigs loaded on this context,
 *                               in the order the reference builds it (contig, left)
 *   2. (caller) all-gather of the lists, 16 B per region; concatenated by contig they are the reference's whole vector
 *   3. upb_shift_examine_order  the order in which src/strand_shift.cpp:198-226 examines that vector: the head of
 *                               libstdc++'s std::sort by Region::sum() descending, ties as the reference leaves them (Q12)
 *   4. upb_shift_sweep          src/strand_shift.cpp:206-219 for the regions of a batch that lie on this context's contigs;
 *                               (caller) all-gather of the (best shift, best correlation) pairs
 *   5. (caller) src/strand_shift.cpp:220-225 in examination order, then upb_shift_mode = src/strand_shift.cpp:239-258
 * unipeak_b200/multi.py::strand_shift_sharded is this protocol over torch.distributed; with one rank it returns exactly
 * what upb_strand_shift returns. */
typedef struct {
  uint64_t n;
  uint32_t* contig;
  uint32_t* left;
  uint32_t* len;
  uint32_t* sum;                /* Region::sum() (no controls in strand_shift: the pooled hit sum) */
  uint8_t* replayed;            /* 1: produced by the left-edge replay (Q2); upb_strand_shift refuses to sweep such a region */
} upb_shift_region_list;

int upb_shift_regions (upb_ctx* ctx, const upb_shift_params* params, upb_shift_region_list* out);
void upb_free_shift_regions (upb_shift_region_list* l);
/* order_out[0 .. *sorted_end) = indices of the first *sorted_end (>= min(want, n)) elements of the reference's sorted vector;
 * the rest of order_out is a permutation of the remaining indices in no particular order.  Host only. */
int upb_shift_examine_order (uint64_t n, const uint32_t* sums, uint64_t want, uint32_t* order_out, uint64_t* sorted_end);
/* every region must be longer than 2 * max_shift + 3 (src/strand_shift.cpp:206) and lie on a contig with tags on this context */
int upb_shift_sweep (upb_ctx* ctx, const upb_shift_params* params, uint64_t n, const uint32_t* contig, const uint32_t* left,
                     const uint32_t* len, uint32_t* best_shift_out, double* best_corr_out);
int upb_shift_mode (const uint64_t* shift_freq, uint32_t n_freq, uint32_t min_shift, uint32_t* best_out);   /* host only */

/* ---- bin/tags_in_regions ------------------------------------------------------------------- */
typedef struct {
  uint32_t contig;
  uint32_t start;               /* as written in the table; start > end marks a reverse-strand row */
  uint32_t end;
} upb_region_key;

/* Replaces the per-row counting loop of src/tags_in_regions.cpp:131-199 over the loaded samples,
 * including its stream-order semantics.  counts_out is [n_regions * n_samples]. */
int upb_count_tags_in_regions (upb_ctx* ctx, int directional, uint32_t extension, uint64_t n_regions,
                               const upb_region_key* regions, uint32_t* counts_out);

/* ---- bin/convert_align --------------------------------------------------------------------- */
/* Replaces CountMap (misc/data.hpp, CountMap::add misc/data.cpp:301-314) as filled by src/convert_align.cpp:70-90: the number of
 * tags per (strand, contig, 5' position) of a batch of alignments.  key[i] = strand * n_contigs + contig (strand 0 '+', 1 '-'),
 * pos[i] the 1-based 5' position (offset applied, out-of-bounds tags already dropped, as ParseAlignStream delivers them),
 * cnt[i] the multiplicity (NULL: 1 each).  combine_strands != 0 pools both strands on strand 0 (the -D output,
 * CountMap::ConstNondirIterator).  Result: distinct positions ascending inside each (strand, contig), library-allocated. */
typedef struct {
  uint64_t n;                   /* distinct (strand, contig, position) */
  uint64_t tag_count;           /* CountMap::tagCount() */
  uint64_t* off;                /* [2 * n_contigs + 1] entries of (strand d, contig c): [off[d*C+c], off[d*C+c+1]) */
  uint32_t* pos;
  uint32_t* cnt;
} upb_tag_counts;
int upb_aggregate_tags (upb_ctx* ctx, uint64_t n, const uint32_t* key, const uint32_t* pos, const uint32_t* cnt, int combine_strands,
                        upb_tag_counts* out);
void upb_free_tag_counts (upb_tag_counts* t);

/* ---- Kernel -------------------------------------------------------------------------------- */
/* Replaces Kernel::Kernel (misc/kernel.cpp:16-35): 2*bandwidth+1 Epanechnikov taps scaled to `sum`.
 * Host arithmetic, identical expression order; the device uses exactly this table. */
int upb_kernel_taps (uint32_t bandwidth, double sum, double* taps_out);

/* ---- instrumentation ----------------------------------------------------------------------- */
/* Device-side timings (CUDA events on the context's stream) and launch counts of the last call. */
typedef struct {
  float ms_total;               /* first kernel start -> last kernel end */
  float ms_kde;                 /* the fused pool + KDE + threshold + region-statistics kernel (span_kernel) */
  uint32_t launches;            /* kernels launched by the last call */
  uint64_t kde_tiles;           /* spans (16384 positions of one buffer x contig) the kernel visited */
  uint64_t kde_hot_tiles;       /* pieces (groups of <= 16 hot 32-position blocks) evaluated position by position */
  uint64_t n_candidates;        /* candidate regions == Region objects the reference would have built */
  uint64_t n_entries;           /* tag entries resident */
  uint64_t hot_blocks;          /* 32-position blocks that could reach the threshold */
  uint64_t exact_evals;         /* positions evaluated in the reference's own summation order */
  uint64_t spans_with_tags;
  uint64_t n_stitched;          /* run pieces merged into a neighbour (runs cut by a piece / span boundary) */
  uint64_t warm_spans;          /* spans that survived the zone-map prefilter */
  float ms_call;                /* host wall clock of the last upb_call_regions, entry to return */
} upb_stats;
int upb_last_stats (const upb_ctx* ctx, upb_stats* out);

/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on, so that a caller can bracket calls with
 * its own CUDA events (bench.py times the step this way).  No reference counterpart: instrumentation only. */
void* upb_stream (const upb_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
