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
  UPB_ERR_LEFT_EDGE = -6,       /* no longer returned (kept for ABI stability): every shape of the reference's left-edge behaviour
                                   (misc/peakcall.cpp:176-184,224-231) is replayed - regions, the -w profile, the shift
                                   sweep, and contigs whose every tag lies within `bandwidth` of the start */
  UPB_ERR_REFERENCE_ABORTS = -7,/* reference_order == 1 only: several directional samples, and the reference's merge loop reads the
                                   '-' track of at least one of them during pass 1 (it starts at or after the last contig of that
                                   sample's '+' track: one-contig tables, toy inputs).  Depending on the tag positions the
                                   reference then trips its own assert (misc/peakcall.cpp:171) or sweeps that sample's '-' tags
                                   through reverseBuffer separately from the other samples'; neither is reproduced.
                                   reference_order == 0 still returns the canonical, pooled table. */
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
/* page-lock and map host memory the caller owns — e.g. a POSIX shared-memory segment shared by the ranks of a multi-GPU job
 * (upb_shared_table); *device_ptr is its device address */
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

/* The same tag streams in the compact host form (2 bytes per entry instead of 8 — the load is PCIe-bound): one u16 per entry,
 * words[e] = (position minus the previous position of the same (sample, strand, contig) segment) in the low 12 bits | count << 12;
 * 0xFFF / 0xF are escapes whose absolute position / full count sit in `exceptions` (sorted by entry index; pos or cnt is 0 where
 * that field is not escaped).  The first entry of every segment, and every 2048th entry after it, MUST be a position escape, so
 * that a tile of 2048 entries decodes on its own.  Decoded on the device, validated exactly as upb_load_tags validates (plus: every
 * escape has its exception, every tile starts with one).  On RNA-seq / ChIP-seq-like data well under 1 % of the entries escape
 * (2.1 bytes per entry); sparse tracks (mean gap > 4095) are better off plain — upb_pack_tags2 reports the exception count, so a
 * front end can choose (2 E + 16 n_exc against 8 E).  upb_pack_tags2 is the host-only encoder a front end runs while it parses (pass
 * exc_cap = 0 to count the exceptions first; returns UPB_ERR_NOMEM with *n_exc_out set when exc_cap is too small). */
typedef struct { uint64_t index; uint32_t pos; uint32_t cnt; } upb_packed_exception;
int upb_load_tags_packed2 (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint16_t* words,
                           uint64_t n_exceptions, const upb_packed_exception* exceptions);
int upb_pack_tags2 (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                    uint16_t* words_out, upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out);

/* The densest host form: a self-synchronising variable-length byte code, ~1.5 bytes per entry on RNA-seq / ChIP-seq-like data (the
 * load is PCIe-bound, so bytes are time).  Per entry, in segment order:
 *   first byte   1 c v5..v0        bit 7 marks the start of an entry; v = position minus the previous position of the same
 *                                  (sample, strand, contig) segment — the ABSOLUTE position for the first entry of every tile of 2048
 *                                  entries of a segment, so that a tile decodes on its own; c = 1: a count byte closes the entry
 *   then         0 xxxxxxx ...     0 - 4 bytes with the higher bits of v, 7 per byte, low to high
 *   then (c = 1) 0 kkkkkkk         the count, 2 - 127; 0 = the count sits in `exceptions` (index = entry index, pos = 0).  c = 0: count 1
 * Entry starts are found from bit 7 alone (a prefix count over the bytes), so a tile is decoded by a whole thread block at once.
 * byte_off has the layout of `off` (n_samples * 2 * (n_contigs + 1) entries): the bytes of segment (s, d, c) are
 * bytes[byte_off[(s*2+d)*(n_contigs+1)+c] .. byte_off[.. + c + 1]); tile_off[t] is the first byte of tile t, tiles numbered through the
 * segments in order (a segment of n entries has (n + 2047) / 2048 tiles), tile_off[n_tiles] = total bytes.  Validated on the device
 * like the other forms (plus: the byte code is well formed and every tile holds exactly its entries).
 * upb_pack_tags_vbyte is the host-only encoder (one task per (sample, strand)); call it with bytes_cap = 0 to learn the sizes
 * (*n_bytes_out, *n_tiles_out, *n_exc_out; returns UPB_ERR_NOMEM), then with buffers of those sizes. */
int upb_load_tags_vbyte (upb_ctx* ctx, uint32_t n_samples, const uint64_t* off, const uint8_t* bytes, const uint64_t* byte_off,
                         uint64_t n_tiles, const uint64_t* tile_off, uint64_t n_exceptions, const upb_packed_exception* exceptions);
int upb_pack_tags_vbyte (uint32_t n_samples, uint32_t n_contigs, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt,
                         uint8_t* bytes_out, uint64_t bytes_cap, uint64_t* n_bytes_out, uint64_t* byte_off_out,
                         uint64_t* tile_off_out, uint64_t tile_cap, uint64_t* n_tiles_out,
                         upb_packed_exception* exc_out, uint64_t exc_cap, uint64_t* n_exc_out);

/* The general, asynchronous form of the loads above.  The transfer is cut into groups — (strand, range of contigs) — and
 * pipelined: while a group's bytes cross PCIe the previous group is decoded, validated and indexed; upb_load_tags_async returns as
 * soon as everything is queued, and a upb_call_regions that follows starts calling regions on the groups that have arrived while
 * the rest is still on its way (the reference overlaps nothing: src/regions.cpp:195-199 parses, then :309-391 computes).
 * The host arrays must stay valid (and, to overlap at all, page-locked) until upb_load_wait or the next call that uses the data
 * returns; what the synchronous loads report at once (unsorted / out-of-range input, a malformed packed form) is reported there.
 *   pos/cnt     plain form, or NULL
 *   words ...   2-byte form, or NULL
 *   vbytes ...  variable-length byte form, or NULL
 *   keep_contig [n_contigs] 1 = load this contig's segments, 0 = leave them out (a rank of a multi-GPU job keeps the contigs it
 *               owns: the arrays describe all contigs, only the kept ones travel); NULL = all
 *   directional_hint  order of the groups: 1 = strand-major, the order a directional upb_call_regions works through them;
 *               0 = contig-major (both strands of a contig range together: -D, strand_shift).  Only the overlap depends on it. */
typedef struct {
  uint32_t n_samples;
  const uint64_t* off;
  const uint32_t* pos;
  const uint32_t* cnt;
  const uint16_t* words;
  uint64_t n_exceptions;
  const upb_packed_exception* exceptions;
  const uint8_t* keep_contig;
  int directional_hint;
  /* variable-length byte form (upb_load_tags_vbyte), or NULL; uses n_exceptions / exceptions for the counts that do not fit */
  const uint8_t* vbytes;
  const uint64_t* vbyte_off;
  const uint64_t* vtile_off;
  uint64_t n_vtiles;
} upb_tag_source;
int upb_load_tags_async (upb_ctx* ctx, const upb_tag_source* src);
int upb_load_wait (upb_ctx* ctx);
/* Instrumentation: when the last byte of the last (completed) load was on the device and when its index was complete, in ms after the
 * first copy was queued. */
int upb_last_load_ms (const upb_ctx* ctx, float* copied_ms, float* indexed_ms);
/* Rebuild the region-calling index (tile-major copy + zone maps: the device-side replacement of the k-way merge of
 * src/regions.cpp:335-354) from the sample-major arrays already resident — what a cold call pays on top of upb_call_regions. */
int upb_reindex (upb_ctx* ctx);

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
                                   0: canonical order (buffer, contig, left), nothing dropped. */
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

/* ---- several GPUs: one table every rank writes its rows into ----------------------------------- */
/* The reference is one thread on one core (README:60-61); its regions table is written by src/regions.cpp:361-368 as regions
 * close.  Here one rank per GPU (a thread or a process each) loads the contigs it owns and calls upb_call_regions_shared with the
 * same `step` (1, 2, 3 ...).  The table lives in ONE page-locked host segment mapped by every rank (upb_host_alloc + threads, or a
 * POSIX shared-memory segment + upb_host_register per process): each rank's GPU publishes its per-(buffer, contig) row counts
 * there, waits ON THE DEVICE for the other ranks' counts and writes its rows straight to their place in table order over its own
 * PCIe link — no collective and no host round trip inside the call.  On rank 0 the call returns the whole table (columns point
 * into the segment: valid until rank 0's next call; upb_free_region_table is a no-op for it), reference_order honoured; on the
 * other ranks only n_rows (their own rows) is set.  Rank 0 zeroes the exchange area once with upb_shared_table_init before any
 * rank calls.  A rank that owns no contig still calls. */
typedef struct {
  void* host;                   /* the segment as this process sees it */
  void* device;                 /* its device address (upb_host_register; == host for upb_host_alloc memory) */
  uint64_t bytes;               /* >= upb_shared_table_bytes(...) */
  uint32_t rank, world;
  uint32_t n_contigs, n_samples;
  uint64_t row_capacity;        /* rows the table has room for (all ranks together) */
} upb_shared_table;
uint64_t upb_shared_table_bytes (uint32_t n_contigs, uint32_t n_samples, uint32_t world, uint64_t row_capacity);
int upb_shared_table_init (const upb_shared_table* t);
int upb_call_regions_shared (upb_ctx* ctx, const upb_region_params* params, const upb_shared_table* t, uint64_t step, upb_region_table* out);

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
  uint8_t* replayed;            /* 1: produced by the left-edge replay (Q2): swept over the replay's own score vectors */
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
  float ms_host_prepare;        /* host side of the last region call: set-up before the first launch (units, taps, small tables) */
  float ms_host_enqueue;        /*   issuing the launches (excluding the waits below) */
  float ms_host_wait;           /*   blocked on the device (the mid-call read-back of a sizing / shared call, and the final one) */
  float ms_host_post;           /*   after the last read-back (counters, per-unit numbers) */
  float ms_host_assemble;       /*   upb_call_regions_shared on rank 0: waiting for the other ranks' rows + totals + emission order */
} upb_stats;
int upb_last_stats (const upb_ctx* ctx, upb_stats* out);

/* Per-stage device time of the last region-calling call (CUDA events between the stages of the pipeline, on the context's stream),
 * for the per-kernel roofline entries of bench.py.  Off by default.  names_out[k] points at static strings. */
int upb_set_profiling (upb_ctx* ctx, int on);
int upb_last_stage_times (const upb_ctx* ctx, uint32_t cap, const char** names_out, float* ms_out, uint32_t* n_out);

/* The CUDA stream (cudaStream_t) every kernel and copy of this context is issued on, so that a caller can bracket calls with
 * its own CUDA events (bench.py times the step this way).  No reference counterpart: instrumentation only. */
void* upb_stream (const upb_ctx* ctx);

/* ---- several GPUs of one box behind one handle ------------------------------------------------------------------------------
 * The three drivers' loops (src/regions.cpp:304-393, src/strand_shift.cpp:139-259, src/tags_in_regions.cpp:131-199) run on one core
 * in the reference.  A upb_multi owns one context and one worker thread per GPU and shards them the way SURVEY 8e describes: whole
 * contigs per GPU for region calling and the shift calibration (no halo crosses a contig), whole samples per GPU for
 * tags_in_regions.  Results are what the one-GPU entry points return, bit for bit.  The region table is ONE pinned host table
 * every GPU writes its rows into (upb_call_regions_shared); there is no collective on the path.  bin/regions, bin/strand_shift and
 * bin/tags_in_regions use it when UPB_GPUS=N is set. */
typedef struct upb_multi upb_multi;
int upb_multi_create (int n_gpus, const int* devices /* NULL: 0 .. n_gpus-1 */, upb_multi** out);
void upb_multi_destroy (upb_multi* m);
const char* upb_multi_last_error (const upb_multi* m);
int upb_multi_gpus (const upb_multi* m);
upb_ctx* upb_multi_context (upb_multi* m, int rank);                  /* statistics, profiling */
int upb_multi_set_contigs (upb_multi* m, uint32_t n_contigs, const uint32_t* sizes);
int upb_multi_owner (const upb_multi* m, uint32_t* owner_out);        /* [n_contigs] rank that holds each contig */
/* Contig-sharded load: the arrays describe every contig (as for upb_load_tags_async; src->keep_contig is ignored), each GPU takes
 * the segments of the contigs it owns.  Asynchronous like upb_load_tags_async; upb_multi_load_wait reports load errors. */
int upb_multi_load_tags (upb_multi* m, const upb_tag_source* src);
int upb_multi_load_wait (upb_multi* m);
/* upb_call_regions over all GPUs.  `out` points into the handle's shared table: valid until the next call on the handle,
 * upb_free_region_table on it is a no-op. */
int upb_multi_call_regions (upb_multi* m, const upb_region_params* params, upb_region_table* out);
int upb_multi_profile_contig (upb_multi* m, const upb_region_params* params, int buffer, uint32_t contig, upb_profile* out);
/* upb_strand_shift over all GPUs (after upb_multi_load_tags): the report of the one-GPU call; free with upb_free_shift_report. */
int upb_multi_strand_shift (upb_multi* m, const upb_shift_params* params, upb_shift_report* out);
/* Sample-sharded load (plain form) and upb_count_tags_in_regions over all GPUs: counts_out is [n_regions * n_samples]. */
int upb_multi_load_tags_by_sample (upb_multi* m, uint32_t n_samples, const uint64_t* off, const uint32_t* pos, const uint32_t* cnt);
int upb_multi_count_tags_in_regions (upb_multi* m, int directional, uint32_t extension, uint64_t n_regions,
                                     const upb_region_key* regions, uint32_t* counts_out);

#ifdef __cplusplus
}
#endif
#endif
