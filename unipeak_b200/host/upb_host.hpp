// upb_host.hpp — host-side front end shared by the drop-in command-line drivers (bin/regions, bin/strand_shift,
// bin/tags_in_regions): flag parsing with the reference's flag surface (the TCLAP declarations in src/*.cpp), the contig
// table (misc/format.cpp:27-57), the tag-frequency wiggle reader (misc/format.cpp:236-286,503-565,648-766) and the
// reference's number formatting (misc/filterstream.cpp:134-137, misc/format.cpp:1141-1162).  Everything here stays on
// the CPU by design (BASELINE.json north_star: parsers are host-side front ends); the hot path is behind
// include/unipeak_b200.h and only there.
#pragma once

#include <cstdint>
#include <cstdio>
#include <map>
#include <string>
#include <vector>

#include "unipeak_b200.h"

namespace upbhost {

// ---- flags (what TCLAP::CmdLine does for the three mains) ---------------------------------------------------------------
enum ArgKind { ARG_SWITCH, ARG_STRING, ARG_U16, ARG_U32, ARG_DOUBLE, ARG_I16 };
struct ArgSpec {                           // in the order the reference declares them
  char flag; const char* name; ArgKind kind; bool required; const char* desc; const char* typeDesc;
};

class CmdLine {
 public:
  CmdLine (const std::vector<ArgSpec>& specs, const char* unlabeledDesc, const char* unlabeledType, const char* unlabeledName = "");
  void parse (int argc, char** argv);      // prints TCLAP's messages and exits (1 on error, 0 for -h / --version)
  bool isSet (char flag) const { return values_.count(flag) != 0; }
  bool sw (char flag) const { return switches_.count(flag) != 0; }
  std::string str (char flag, const std::string& dflt = "") const;
  double dbl (char flag, double dflt) const;
  uint32_t u32 (char flag, uint32_t dflt) const;
  uint16_t u16 (char flag, uint16_t dflt) const;
  short i16 (char flag, short dflt) const;
  const std::vector<std::string>& unlabeled () const { return unlabeled_; }

 private:
  [[noreturn]] void failure (const std::string& argId, const std::string& what) const;
  void shortUsage (FILE* f) const;
  void longUsage (FILE* f) const;
  std::string argId (const ArgSpec& a) const;
  std::vector<ArgSpec> specs_;
  std::string unlabeledDesc_, unlabeledType_, unlabeledName_, prog_;
  std::map<char, std::string> values_;
  std::map<char, bool> switches_;
  std::vector<std::string> unlabeled_;
};

// the "-s" list shared by the three drivers (src/regions.cpp:111-127): exits with the reference's messages
std::vector<short> parse_offsets (const std::string& s, size_t nFiles);
std::vector<std::string> split_commas (const std::string& s);      // boost::tokenizer<char_separator>(","): drops empty tokens
bool lexical_short (const std::string& s, short* out);             // boost::lexical_cast<short>
bool lexical_double (const std::string& s, double* out);           // boost::lexical_cast<double>
bool lexical_u32 (const char* s, size_t n, uint32_t* out);         // boost::lexical_cast<Pos>

// ---- contig table --------------------------------------------------------------------------------------------------------
struct ContigTable {
  std::vector<std::string> names;
  std::vector<uint32_t> sizes;
  std::map<std::string, uint32_t> index;
  uint32_t size () const { return (uint32_t)names.size(); }
  uint32_t find (const std::string& name) const { auto it = index.find(name); return it == index.end() ? size() : it->second; }
  uint32_t genome_size () const { uint32_t g = 0; for (uint32_t s : sizes) g += s; return g; }   // Pos arithmetic, as ContigTable::getGenomeSize
};
ContigTable parse_contig_table (const std::string& fname);         // prints "reading <f>... N contigs" like the reference

// ---- tag-frequency wiggle ---------------------------------------------------------------------------------------------------
struct Sample {
  std::string fname, name;                 // exptName()
  uint64_t expectedTags = 0;               // "# tags=" header, else the confident count (getExpectedTags)
  uint64_t confidentTags = 0, outOfBoundsTags = 0, totalTags = 0;
  // [strand][contig] ascending unique positions with counts (the -s offset applied, out-of-bounds tags dropped)
  std::vector<std::vector<uint32_t> > pos[2], cnt[2];
  std::string error;                       // non-empty: the reference's "error: ..." text for this file
};
// useLength: -l (only moves the out-of-bounds test for wiggle input, misc/format.cpp:514,655)
void read_wig (const std::string& fname, const ContigTable& contigs, int offset, unsigned useLength, Sample* out);
// every file on its own thread; returns in input order; exits with the first file's error like the reference would
std::vector<Sample> read_samples (const std::vector<std::string>& fnames, const ContigTable& contigs,
                                  const std::vector<short>& offsets, unsigned useLength);

// ---- alignments (bin/convert_align) ------------------------------------------------------------------------------------------------
// What ParseAlignStream delivers to CountMap::add (src/convert_align.cpp:70-90): one (strand, contig, 5' position, count) per usable
// alignment, plus the counters printSummary reports (misc/format.cpp:213-232).  Formats: BED, SAM, tag-frequency wiggle.
struct AlignOpts { unsigned mismatchTolerance = 1; unsigned useLength = 0; int offset = 0; double probThreshold = 0.9; };
struct AlignBatch {
  std::string name;                        // exptName()
  int format = 0;                          // 1 BED, 4 SAM, 6 / 7 wiggle (misc/format.hpp:36-43)
  uint64_t totalTags = 0, confidentTags = 0, rejectTags = 0, outOfBoundsTags = 0;
  std::vector<uint32_t> key, pos, cnt;     // key = strand * nContigs + contig
  std::string error;
};
void read_alignments (const std::string& fname, const ContigTable& contigs, const AlignOpts& opts, AlignBatch* out);
std::string align_summary (const AlignBatch& b, double probThreshold);   // ParseAlignStream::printSummary

// flat layout of upb_load_tags (include/unipeak_b200.h)
struct Flat { std::vector<uint64_t> off; uint32_t* pos = nullptr; uint32_t* cnt = nullptr; uint64_t n = 0; bool pinned = false; ~Flat (); };
void flatten (const std::vector<Sample>& samples, uint32_t nContigs, Flat* out);

// What the three drivers do with the samples they parsed: hand them to the device in the 2-byte host form (one u16 per entry +
// exception list, include/unipeak_b200.h) — encoded here, one thread per sample, straight from the per-segment vectors the wiggle
// reader filled — or in the plain form where the 2-byte form would be larger (sparse tracks: mean gap > 4095 positions).  start()
// returns as soon as the transfers are queued (upb_load_tags_async): a upb_call_regions that follows works on the groups that have
// arrived while the rest is still crossing PCIe.  The buffers live until the loader is destroyed: keep it until the first call
// that uses the data has returned (or call wait()).
class TagLoader {
 public:
  TagLoader () {}
  ~TagLoader ();
  int start (upb_ctx* ctx, const std::vector<Sample>& samples, uint32_t nContigs, bool directionalHint);
  int start (upb_multi* multi, const std::vector<Sample>& samples, uint32_t nContigs, bool directionalHint);   // contig-sharded over the GPUs of the handle
  int wait (upb_ctx* ctx) { return upb_load_wait(ctx); }
  bool packed () const { return words_ != nullptr || vbytes_ != nullptr; }
  uint64_t entries () const { return n_; }
  uint64_t exceptions () const { return exc_.size(); }
  double packSeconds () const { return packSeconds_; }
 private:
  TagLoader (const TagLoader&);
  void encode (const std::vector<Sample>& samples, uint32_t nContigs, bool directionalHint, upb_tag_source* src);
  std::vector<uint64_t> off_;
  uint16_t* words_ = nullptr; bool wordsPinned_ = false;
  uint8_t* vbytes_ = nullptr; bool vbytesPinned_ = false;             // variable-length byte form (UPB_HOST_FORM=vbyte)
  std::vector<uint64_t> vbyteOff_, vtileOff_;
  std::vector<upb_packed_exception> exc_;
  Flat flat_;
  uint64_t n_ = 0;
  double packSeconds_ = 0;
};

// ---- formatting ----------------------------------------------------------------------------------------------------------------
std::string fmt_lexical (double x);        // OutStream << double == boost::lexical_cast<string>(double): 17 significant digits
std::string fmt_g (double x);              // std::ostream << double
std::string fmt_fixed (double x, int digits);   // boost::format("%.Nf")
std::string fname_prefix (const std::string& name);   // getFnamePrefix, misc/format.cpp:60-67
int fast_g6 (double x, char* buf);         // == snprintf(buf, .., "%g", x) for finite x, without going through printf

// OutStream (misc/filterstream.cpp:87-107): file, "stdout", gzip by suffix
class Out {
 public:
  explicit Out (const std::string& fname);
  ~Out ();
  void write (const char* p, size_t n);
  void write (const std::string& s) { write(s.data(), s.size()); }
  void close ();
 private:
  void flush ();
  FILE* f_ = nullptr; void* gz_ = nullptr; bool isStdout_ = false;
  std::vector<char> buf_;
};

[[noreturn]] void die (const std::string& msg);                    // "error: <msg>\n\n", exit(1)
[[noreturn]] void die_upb (upb_ctx* ctx, int code);

// The device side of a driver: one GPU (upb_ctx), or — with UPB_GPUS=N in the environment — N GPUs of the box behind a upb_multi
// (contigs spread over the GPUs for regions / strand_shift, samples for tags_in_regions).  Output is byte-identical either way.
// Every method dies with the library's message on failure, as the drivers did with die_upb.
class Device {
 public:
  Device ();                                   // returns at once: the CUDA context(s) are created by a background thread while the caller parses its inputs
  ~Device ();
  int gpus () const { return want_ > 1 ? want_ : 1; }
  void join ();                                // (die() calls this so that the process does not exit under a thread that is inside the driver)
  void setContigs (uint32_t nContigs, const uint32_t* sizes);
  void load (TagLoader& loader, const std::vector<Sample>& samples, uint32_t nContigs, bool directionalHint, bool wait);   // by contig
  void loadBySample (const std::vector<Sample>& samples, uint32_t nContigs, bool directionalHint);                       // tags_in_regions
  void callRegions (const upb_region_params* p, upb_region_table* out);
  void profileContig (const upb_region_params* p, int buffer, uint32_t contig, upb_profile* out);
  int strandShift (const upb_shift_params* p, upb_shift_report* out);       // returns UPB_OK or UPB_ERR_NO_REGIONS
  void countTagsInRegions (int directional, uint32_t extension, uint64_t nRegions, const upb_region_key* regions, uint32_t* countsOut);
 private:
  Device (const Device&);
  [[noreturn]] void fail (int rc);
  void ready ();                               // joins the background creation; dies with the library's message if it failed
  void* init_ = nullptr;                       // std::thread
  int initRc_ = 0, want_ = 1;
  std::string initMsg_;
  upb_ctx* ctx_ = nullptr;
  upb_multi* multi_ = nullptr;
  TagLoader sampleLoader_;
  Flat flat_;
};

}  // namespace upbhost
