// upb_host.cpp — see upb_host.hpp.  Host-only code: flags, contig table, wiggle reader, number formatting, output files.
#include "upb_host.hpp"

#include <zlib.h>

#include <algorithm>
#include <cctype>
#include <cerrno>
#include <climits>
#include <cmath>
#include <chrono>
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <thread>

namespace upbhost {

// ---------------------------------------------------------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------------------------------------------------------
void die (const std::string& msg) {
  fflush(stdout);
  fprintf(stderr, "error: %s\n\n", msg.c_str());
  exit(1);
}

void die_upb (upb_ctx* ctx, int code) {
  const char* m = upb_last_error(ctx);
  std::string msg = (m && *m) ? m : "device call failed";
  msg += " (unipeak_b200 error " + std::to_string(code) + ")";
  die(msg);
}

// ---------------------------------------------------------------------------------------------------------------------------
// flags: the observable behaviour of TCLAP::CmdLine + StdOutput (misc/tclap/CmdLine.h, StdOutput.h) for the argument kinds
// the three mains declare.  Unknown "-x" tokens are swallowed by the unlabeled multi-argument, as in the reference.
// ---------------------------------------------------------------------------------------------------------------------------
CmdLine::CmdLine (const std::vector<ArgSpec>& specs, const char* unlabeledDesc, const char* unlabeledType, const char* unlabeledName)
    : specs_(specs), unlabeledDesc_(unlabeledDesc), unlabeledType_(unlabeledType), unlabeledName_(unlabeledName) {}

std::string CmdLine::argId (const ArgSpec& a) const { return std::string("Argument: -") + a.flag + " (--" + a.name + ")"; }

// word wrap of StdOutput::spacePrint: break at ' ', ',' or '|', indent the continuation lines further
static void space_print (FILE* f, const std::string& s, int maxWidth, int indent, int secondLineOffset) {
  const int len = (int)s.size();
  if (len + indent <= maxWidth || maxWidth <= 0) { fprintf(f, "%*s%s\n", indent, "", s.c_str()); return; }
  int allowed = maxWidth - indent, start = 0;
  while (start < len) {
    int n = std::min(len - start, allowed);
    if (n == allowed) while (n >= 0 && s.c_str()[n + start] != ' ' && s.c_str()[n + start] != ',' && s.c_str()[n + start] != '|') --n;
    if (n <= 0) n = allowed;
    for (int i = 0; i < n; ++i) if (s[start + i] == '\n') n = i + 1;
    fprintf(f, "%*s", indent, "");
    if (start == 0) { indent += secondLineOffset; allowed -= secondLineOffset; }
    fprintf(f, "%s\n", s.substr(start, n).c_str());
    while (start < len && s.c_str()[n + start] == ' ') ++start;
    start += n;
  }
}

static std::string short_id (const ArgSpec& a) {
  std::string id = std::string("-") + a.flag;
  if (a.kind != ARG_SWITCH) id += std::string(" <") + a.typeDesc + ">";
  return a.required ? id : "[" + id + "]";
}

void CmdLine::shortUsage (FILE* f) const {
  std::string s = prog_ + " ";
  for (size_t i = specs_.size(); i-- > 0;) s += " " + short_id(specs_[i]);
  s += " [--] [--version] [-h] <" + unlabeledType_ + "> ...";
  int second = (int)prog_.size() + 2;
  if (second > 75 / 2) second = 75 / 2;
  space_print(f, s, 75, 3, second);
}

void CmdLine::longUsage (FILE* f) const {
  for (size_t i = specs_.size(); i-- > 0;) {
    const ArgSpec& a = specs_[i];
    std::string id = std::string("-") + a.flag, tail = std::string("--") + a.name;
    if (a.kind != ARG_SWITCH) { id += std::string(" <") + a.typeDesc + ">"; tail += std::string(" <") + a.typeDesc + ">"; }
    space_print(f, id + ",  " + tail, 75, 3, 3);
    space_print(f, std::string(a.required ? "(required)  " : "") + a.desc, 75, 5, 0);
    fputc('\n', f);
  }
  space_print(f, "--,  --ignore_rest", 75, 3, 3);
  space_print(f, "Ignores the rest of the labeled arguments following this flag.", 75, 5, 0);
  fputc('\n', f);
  space_print(f, "--version", 75, 3, 3);
  space_print(f, "Displays version information and exits.", 75, 5, 0);
  fputc('\n', f);
  space_print(f, "-h,  --help", 75, 3, 3);
  space_print(f, "Displays usage information and exits.", 75, 5, 0);
  fputc('\n', f);
  space_print(f, "<" + unlabeledType_ + ">  (accepted multiple times)", 75, 3, 3);
  space_print(f, "(required)  " + unlabeledDesc_, 75, 5, 0);
  fputc('\n', f);
  fputc('\n', f);
  space_print(f, "See README.TXT for more information", 75, 3, 0);
}

void CmdLine::failure (const std::string& id, const std::string& what) const {
  fflush(stdout);
  fprintf(stderr, "PARSE ERROR: %s\n             %s\n\n", id.c_str(), what.c_str());
  fprintf(stderr, "Brief USAGE: \n");
  shortUsage(stderr);
  fprintf(stderr, "\nFor complete USAGE and HELP type: \n   %s --help\n\n", prog_.c_str());
  exit(1);
}

// TCLAP::ValueArg<T>::_extractValue for value-like T: stream extraction, every token must convert, exactly one value
template <class T> static bool extract_ok (const std::string& val, bool* several) {
  std::istringstream is(val);
  T dest;
  int n = 0;
  while (is.good()) {
    if (is.peek() != EOF) is >> dest; else break;
    ++n;
  }
  if (is.fail()) return false;
  *several = n > 1;
  return true;
}

void CmdLine::parse (int argc, char** argv) {
  prog_ = argc > 0 ? argv[0] : "";
  bool ignoreRest = false;
  auto bySpec = [&] (const std::string& tok) -> const ArgSpec* {
    for (const ArgSpec& a : specs_) {
      if (tok.size() == 2 && tok[0] == '-' && tok[1] == a.flag) return &a;
      if (tok.size() > 2 && tok[0] == '-' && tok[1] == '-' && tok.substr(2) == a.name) return &a;
    }
    return nullptr;
  };
  for (int i = 1; i < argc; ++i) {
    const std::string tok = argv[i];
    if (ignoreRest) { unlabeled_.push_back(tok); continue; }
    if (tok == "--" || tok == "--ignore_rest") { ignoreRest = true; continue; }
    if (tok == "--version") { printf("\n%s  version: 1.0\n\n", prog_.c_str()); exit(0); }
    if (tok == "-h" || tok == "--help") {
      printf("\nUSAGE: \n\n");
      shortUsage(stdout);
      printf("\n\nWhere: \n\n");
      longUsage(stdout);
      printf("\n");
      exit(0);
    }
    const ArgSpec* a = bySpec(tok);
    if (a) {
      if (a->kind == ARG_SWITCH) {
        if (switches_.count(a->flag)) failure(argId(*a), "Argument already set!");
        switches_[a->flag] = true;
        continue;
      }
      if (values_.count(a->flag)) failure(argId(*a), "Argument already set!");
      if (i + 1 >= argc) failure(argId(*a), "Missing a value for this argument!");
      const std::string val = argv[++i];
      bool ok = true, several = false;
      switch (a->kind) {
        case ARG_U16: ok = extract_ok<unsigned short>(val, &several); break;
        case ARG_U32: ok = extract_ok<unsigned int>(val, &several); break;
        case ARG_DOUBLE: ok = extract_ok<double>(val, &several); break;
        case ARG_I16: ok = extract_ok<short>(val, &several); break;
        default: break;
      }
      if (!ok) failure(argId(*a), "Couldn't read argument value from string '" + val + "'");
      if (several) failure(argId(*a), "More than one valid value parsed from string '" + val + "'");
      values_[a->flag] = val;
      continue;
    }
    // combined switches, e.g. -qD
    if (tok.size() > 2 && tok[0] == '-' && tok[1] != '-') {
      bool all = true;
      for (size_t k = 1; k < tok.size() && all; ++k) {
        bool found = false;
        for (const ArgSpec& s : specs_) if (s.kind == ARG_SWITCH && s.flag == tok[k]) found = true;
        all = found;
      }
      if (all) { for (size_t k = 1; k < tok.size(); ++k) switches_[tok[k]] = true; continue; }
    }
    unlabeled_.push_back(tok);
  }
  // required arguments, listed in TCLAP's (reverse-declaration) order; the unlabeled argument has an empty name
  std::string missing;
  int nMissing = 0;
  for (size_t i = specs_.size(); i-- > 0;)
    if (specs_[i].required && !values_.count(specs_[i].flag)) { missing += std::string(specs_[i].name) + ", "; ++nMissing; }
  if (unlabeled_.empty()) { missing += unlabeledName_ + ", "; ++nMissing; }
  if (nMissing > 0)
    failure(" ", std::string(nMissing > 1 ? "Required arguments missing: " : "Required argument missing: ") + missing.substr(0, missing.size() - 2));
}

std::string CmdLine::str (char flag, const std::string& dflt) const {
  auto it = values_.find(flag);
  return it == values_.end() ? dflt : it->second;
}
template <class T> static T extract (const std::string& val) { std::istringstream is(val); T v = T(); is >> v; return v; }
double CmdLine::dbl (char flag, double dflt) const { auto it = values_.find(flag); return it == values_.end() ? dflt : extract<double>(it->second); }
uint32_t CmdLine::u32 (char flag, uint32_t dflt) const { auto it = values_.find(flag); return it == values_.end() ? dflt : extract<unsigned int>(it->second); }
short CmdLine::i16 (char flag, short dflt) const { auto it = values_.find(flag); return it == values_.end() ? dflt : extract<short>(it->second); }
uint16_t CmdLine::u16 (char flag, uint16_t dflt) const { auto it = values_.find(flag); return it == values_.end() ? dflt : extract<unsigned short>(it->second); }

// ---------------------------------------------------------------------------------------------------------------------------
// lexical_cast / tokenizer behaviour used by the mains
// ---------------------------------------------------------------------------------------------------------------------------
std::vector<std::string> split_commas (const std::string& s) {
  std::vector<std::string> out;
  size_t i = 0;
  while (i <= s.size()) {
    size_t j = s.find(',', i);
    if (j == std::string::npos) j = s.size();
    if (j > i) out.push_back(s.substr(i, j - i));
    i = j + 1;
  }
  return out;
}

bool lexical_short (const std::string& s, short* out) {
  if (s.empty() || isspace((unsigned char)s[0])) return false;
  errno = 0;
  char* end = nullptr;
  const long v = strtol(s.c_str(), &end, 10);
  if (errno || end == s.c_str() || *end != '\0' || v < SHRT_MIN || v > SHRT_MAX) return false;
  *out = (short)v;
  return true;
}

bool lexical_double (const std::string& s, double* out) {
  if (s.empty() || isspace((unsigned char)s[0])) return false;
  char* end = nullptr;
  const double v = strtod(s.c_str(), &end);
  if (end == s.c_str() || *end != '\0') return false;
  *out = v;
  return true;
}

bool lexical_u32 (const char* s, size_t n, uint32_t* out) {
  size_t i = 0;
  if (n && s[0] == '+') i = 1;
  if (i >= n) return false;
  uint64_t v = 0;
  for (; i < n; ++i) {
    const unsigned d = (unsigned)s[i] - '0';
    if (d > 9) return false;
    v = v * 10 + d;
    if (v > 0xFFFFFFFFull) return false;
  }
  *out = (uint32_t)v;
  return true;
}

std::vector<short> parse_offsets (const std::string& s, size_t nFiles) {
  std::vector<short> offsets;
  if (s.empty()) return offsets;
  for (const std::string& tok : split_commas(s)) {
    short v;
    if (!lexical_short(tok, &v)) { fprintf(stderr, "error: bad offset argument\n\n"); exit(1); }
    offsets.push_back(v);
  }
  if (!(offsets.size() == nFiles || offsets.size() == 1)) {
    fprintf(stderr, "error: wrong number of offset arguments\nmust have same number as alignment files or just one\n\n");
    exit(1);
  }
  return offsets;
}

// ---------------------------------------------------------------------------------------------------------------------------
// whole-file reader (InStream, misc/filterstream.cpp:30-50): plain, gzip by suffix, "stdin"
// ---------------------------------------------------------------------------------------------------------------------------
static bool ends_with (const std::string& s, const char* suf) {
  const size_t n = strlen(suf);
  return s.size() >= n && s.compare(s.size() - n, n, suf) == 0;
}

static bool slurp (const std::string& fname, std::vector<char>* buf, std::string* shownName, std::string* err) {
  buf->clear();
  *shownName = fname;
  if (fname == "stdin") {
    *shownName = "standard input stream";
    char tmp[1 << 16];
    size_t k;
    while ((k = fread(tmp, 1, sizeof tmp, stdin)) > 0) buf->insert(buf->end(), tmp, tmp + k);
    return true;
  }
  if (ends_with(fname, ".bz2")) { *err = "bzip2 input is not supported by this build (no bzlib.h): " + fname; return false; }
  if (ends_with(fname, ".gz")) {
    gzFile g = gzopen(fname.c_str(), "rb");
    if (!g) { *err = "could not read " + fname; return false; }
    gzbuffer(g, 1 << 20);
    std::vector<char> tmp(1 << 22);
    int k;
    while ((k = gzread(g, tmp.data(), (unsigned)tmp.size())) > 0) buf->insert(buf->end(), tmp.begin(), tmp.begin() + k);
    gzclose(g);
    return true;
  }
  FILE* f = fopen(fname.c_str(), "rb");
  if (!f) { *err = "could not read " + fname; return false; }
  if (fseek(f, 0, SEEK_END) == 0) {
    const long sz = ftell(f);
    if (sz > 0) buf->reserve((size_t)sz);
    fseek(f, 0, SEEK_SET);
  }
  std::vector<char> tmp(1 << 22);
  size_t k;
  while ((k = fread(tmp.data(), 1, tmp.size(), f)) > 0) buf->insert(buf->end(), tmp.begin(), tmp.begin() + k);
  fclose(f);
  return true;
}

// ---------------------------------------------------------------------------------------------------------------------------
// contig table (misc/format.cpp:27-57; CONTIG_TABLE_REGEX "^(\w+)\W+(\d+)", misc/format.hpp:22)
// ---------------------------------------------------------------------------------------------------------------------------
static inline bool is_word (char c) { return isalnum((unsigned char)c) || c == '_'; }

ContigTable parse_contig_table (const std::string& fname) {
  fprintf(stderr, "reading %s... ", fname.c_str());
  fflush(stderr);
  ContigTable t;
  std::vector<char> buf;
  std::string shown, err;
  if (!slurp(fname, &buf, &shown, &err)) { fprintf(stderr, "error: %s\n\n", err.c_str()); exit(1); }
  size_t p = 0;
  while (p < buf.size()) {
    const char* nl = (const char*)memchr(buf.data() + p, '\n', buf.size() - p);
    const size_t e = nl ? (size_t)(nl - buf.data()) : buf.size();
    const char* s = buf.data() + p;
    const size_t len = e - p;
    p = e + 1;
    if (!len || s[0] == '#') continue;
    size_t i = 0;
    while (i < len && is_word(s[i])) ++i;
    if (i == 0) continue;
    size_t j = i;
    while (j < len && !is_word(s[j])) ++j;
    if (j == i) continue;
    size_t k = j;
    while (k < len && isdigit((unsigned char)s[k])) ++k;
    if (k == j) continue;
    const std::string name(s, i);
    uint32_t size;
    if (!lexical_u32(s + j, k - j, &size)) die("bad contig size for " + name + " in " + fname);
    if (t.index.count(name)) { fprintf(stderr, "error: %s defined twice in contig table\n\n", name.c_str()); exit(1); }
    t.index[name] = t.size();
    t.names.push_back(name);
    t.sizes.push_back(size);
  }
  if (t.size() == 0) { fprintf(stderr, "error: no contigs in table\n\n"); exit(1); }
  fprintf(stderr, "%u contigs\n", t.size());
  return t;
}

std::string fname_prefix (const std::string& name) {
  std::string r = name;
  const size_t slash = r.rfind('/');
  if (slash != std::string::npos) r = r.substr(slash + 1);
  const size_t dot = r.find('.');
  if (dot != std::string::npos) r = r.substr(0, dot);
  return r;
}

// ---------------------------------------------------------------------------------------------------------------------------
// tag-frequency wiggle (misc/format.cpp:242-286 format detection, :503-565 the two wiggle cases, :654-678 offset and
// bounds, :737-766 the "# tags=" header).  The reference streams one Alignment at a time into the merge loop; here the
// whole file becomes per-(strand, contig) arrays, which is what the device path consumes.
// ---------------------------------------------------------------------------------------------------------------------------
// NAME_REGEX1 name="(.+?)"  /  NAME_REGEX2 name=(.+?)<space>   (misc/format.hpp:31-32)
static bool match_name (const char* s, size_t len, const char* open, char closer, std::string* out) {
  const size_t lo = strlen(open);
  for (size_t p = 0; p + lo < len; ++p) {
    if (memcmp(s + p, open, lo) != 0) continue;
    for (size_t q = p + lo + 1; q < len; ++q)
      if (s[q] == closer) { out->assign(s + p + lo, q - p - lo); return true; }
  }
  return false;
}

// DIRECTIONAL_WIG_NAME_REGEX "(.+) ([+-])" searched in the track name: the last " +"/" -" that has a character before it
static bool split_directional (const std::string& name, std::string* base, bool* forward) {
  for (size_t i = name.size(); i-- > 1;)
    if (name[i] == ' ' && i + 1 < name.size() && (name[i + 1] == '+' || name[i + 1] == '-')) {
      *base = name.substr(0, i);
      *forward = name[i + 1] == '+';
      return true;
    }
  return false;
}

static std::string line_error (const std::string& msg, const std::string& shown, uint64_t lineNo) {
  return msg + " in " + shown + " line " + std::to_string(lineNo);
}

namespace {

// The state machine of the wiggle reader: one instance per file, fed line by line (sequential path) or with pre-tokenised data
// lines (parallel path).  Every rule of the reference's parser that this front end reproduces lives here, once.
struct WigParser {
  Sample& S;
  const ContigTable& contigs;
  const int offset;
  const unsigned useLength;
  const std::string& fname;
  const std::string& shown;
  const uint32_t C;
  int format = 0;                   // 0 unknown, 6 directional wiggle, 7 non-directional wiggle
  bool forward = true;
  uint32_t contig = 0;              // ParseAlignStream::open resets the Alignment to contig 0 (misc/format.cpp:122-127)
  bool inHeader = true;             // getExpectedTags only looks at the '#' lines before the first other line
  long lastKey = -1;                // (strand, contig) blocks must come in stream order
  std::vector<uint32_t>* vp = nullptr;
  std::vector<uint32_t>* vc = nullptr;

  WigParser (Sample& s, const ContigTable& c, int off, unsigned len, const std::string& f, const std::string& sh)
    : S(s), contigs(c), offset(off), useLength(len), fname(f), shown(sh), C(c.size()) {}

  bool not_a_wiggle () { S.error = fname + " is not a tag-frequency wiggle file; this build reads only the output of convert_align"; return false; }

  // a data line whose two numbers are already parsed; false: S.error is set
  bool entry (uint32_t first, uint32_t count, uint64_t lineNo) {
    S.totalTags += count;
    if (count == 0) return true;
    uint32_t last = first;
    if (useLength != 0) last = forward ? first + (useLength - 1) : first - (useLength - 1);
    if (offset != 0) {
      if ((forward && (int)first > -offset) || (!forward && (int)last > offset)) {
        first += (uint32_t)(forward ? offset : -offset);
        last += (uint32_t)(forward ? offset : -offset);
      } else { S.outOfBoundsTags += count; return true; }
    }
    const uint32_t size = contigs.sizes[contig];
    if (first == 0 || first > size || last == 0 || last > size) { S.outOfBoundsTags += count; return true; }
    S.confidentTags += count;
    if (!vp) {
      const long key = (long)(forward ? 0 : 1) * (long)C + contig;
      if (key <= lastKey) { S.error = "alignments out of order in " + shown + " line " + std::to_string(lineNo) + " (strand/contig blocks must follow the contig table, '+' track first)"; return false; }
      lastKey = key;
      vp = &S.pos[forward ? 0 : 1][contig];
      vc = &S.cnt[forward ? 0 : 1][contig];
    }
    if (!vp->empty() && first <= vp->back()) {
      if (first == vp->back()) { vc->back() += count; return true; }   // the merge loop adds up equal positions (src/regions.cpp:341-350)
      S.error = line_error("alignments out of order", shown, lineNo);
      return false;
    }
    vp->push_back(first);
    vc->push_back(count);
    return true;
  }

  // a data line as text (it starts with a digit and the format is known)
  bool data_line (const char* s, size_t len, uint64_t lineNo) {
    if (contig == C) return true;                                   // contig not in the table: skipped
    size_t delim = len;
    while (delim-- > 0) if (s[delim] == '\t' || s[delim] == ' ') break;
    if (delim == (size_t)-1) { S.error = line_error("bad format", shown, lineNo); return false; }
    uint32_t first, count;
    const size_t cs = delim + ((delim + 1 < len && s[delim + 1] == '-') ? 2 : 1);
    if (!lexical_u32(s, delim, &first) || cs > len || !lexical_u32(s + cs, len - cs, &count)) { S.error = line_error("bad format", shown, lineNo); return false; }
    return entry(first, count, lineNo);
  }

  // any line; false: S.error is set and parsing stops
  bool line (const char* s, size_t len, uint64_t lineNo) {
    if (len == 0) { inHeader = false; return true; }               // an empty line ends getExpectedTags' header scan too
    if (s[0] == '#') {
      if (inHeader) {
        // TAG_COUNT_REGEX "# tags=(\d+)" anywhere in the line
        const std::string ln(s, len);
        size_t q = 0;
        while ((q = ln.find("# tags=", q)) != std::string::npos) {
          size_t k = q + 7;
          while (k < ln.size() && isdigit((unsigned char)ln[k])) ++k;
          if (k > q + 7) {
            if (S.expectedTags != 0) { S.error = line_error("multiple tag count headers", shown, lineNo); return false; }
            S.expectedTags = strtoull(ln.c_str() + q + 7, nullptr, 10);
            if (S.expectedTags == 0) { S.error = line_error("zero tag count", shown, lineNo); return false; }
            break;
          }
          ++q;
        }
      }
      return true;
    }
    inHeader = false;
    if (format == 0) {
      if (len >= 5 && memcmp(s, "track", 5) == 0) {
        std::string nm;
        const bool named = match_name(s, len, "name=\"", '"', &nm) || match_name(s, len, "name=", ' ', &nm);
        if (named) S.name = nm;
        const std::string ln(s, len);
        if (ln.find("type=wiggle_0") == std::string::npos) return not_a_wiggle();
        std::string b;
        bool f;
        if (split_directional(S.name, &b, &f)) { format = 6; S.name = b; forward = f; }
        else { format = 7; S.name = named ? nm : std::string(); forward = true; }
        return true;
      }
      return not_a_wiggle();
    }
    if (isdigit((unsigned char)s[0])) return data_line(s, len, lineNo);
    if (len > 19 && memcmp(s, "variableStep chrom=", 19) == 0) {
      contig = contigs.find(std::string(s + 19, len - 19));
      vp = vc = nullptr;
      return true;
    }
    if (len >= 5 && memcmp(s, "track", 5) == 0) {
      std::string nm;
      if (match_name(s, len, "name=\"", '"', &nm)) {
        contig = C;
        vp = vc = nullptr;
        if (format == 6) {
          std::string b;
          bool f;
          if (!split_directional(nm, &b, &f)) { S.error = line_error("strand not defined", shown, lineNo); return false; }
          S.name = b;
          forward = f;
        } else {
          S.name = nm;
        }
        return true;
      }
    }
    S.error = line_error("bad format", shown, lineNo);
    return false;
  }
};

void reset_sample (Sample& S, const std::string& fname, uint32_t C) {
  S = Sample();
  S.fname = fname;
  S.name = fname_prefix(fname);
  for (int d = 0; d < 2; ++d) { S.pos[d].assign(C, std::vector<uint32_t>()); S.cnt[d].assign(C, std::vector<uint32_t>()); }
}

void parse_sequential (const std::vector<char>& buf, WigParser& P) {
  const char* base = buf.data();
  size_t p = 0;
  const size_t N = buf.size();
  uint64_t lineNo = 0;
  // InStream::good() is still true after the last '\n', so the reference reads one more (empty) line; it changes nothing
  while (p < N) {
    const char* nl = (const char*)memchr(base + p, '\n', N - p);
    const size_t e = nl ? (size_t)(nl - base) : N;
    ++lineNo;
    if (!P.line(base + p, e - p, lineNo)) return;
    p = e + 1;
  }
}

// Parallel path (at one or two big files the single parsing thread is the longest part of the program): the file is cut
// into byte ranges, every range is tokenised by its own thread — a line of the canonical shape  digits (blank | tab) [-] digits
// becomes a (position, count) token, every other line a reference to its text — and the tokens then run through the same state
// machine in file order.  Anything that is not plainly fine (a data line of another shape, any error of the state machine) makes
// the caller parse the file again sequentially, so error texts, line numbers and every corner of the sequential path stay what
// they are.  Returns false when that second pass is needed.
struct WigToken { uint32_t first, count; };                          // first == 0xFFFFFFFF: count indexes `special`
struct WigChunk { std::vector<WigToken> tok; std::vector<std::pair<size_t, size_t> > special; bool odd = false; };

void tokenise_range (const char* base, size_t N, size_t lo, size_t hi, WigChunk* out) {
  size_t p = lo;
  if (lo != 0) {                                                      // first line that STARTS inside [lo, hi)
    const char* nl = (const char*)memchr(base + lo - 1, '\n', N - (lo - 1));
    if (!nl) return;
    p = (size_t)(nl - base) + 1;
  }
  out->tok.reserve((hi - lo) / 9 + 16);
  while (p < hi && p < N) {
    const char* s = base + p;
    const char* nl = (const char*)memchr(s, '\n', N - p);
    const size_t len = nl ? (size_t)(nl - s) : N - p;
    p += len + 1;
    size_t i = 0;
    uint64_t a = 0, b = 0;
    while (i < len && (unsigned)(s[i] - '0') <= 9 && i < 10) { a = a * 10 + (unsigned)(s[i] - '0'); ++i; }
    if (i == 0) {                                                     // not a data line
      WigToken t; t.first = 0xFFFFFFFFu; t.count = (uint32_t)out->special.size();
      out->special.push_back(std::make_pair((size_t)(s - base), len));
      out->tok.push_back(t);
      continue;
    }
    if (i >= len || (s[i] != ' ' && s[i] != '\t') || a > 0xFFFFFFFEull) { out->odd = true; return; }
    ++i;
    if (i < len && s[i] == '-') ++i;
    const size_t d0 = i;
    while (i < len && (unsigned)(s[i] - '0') <= 9 && i - d0 < 10) { b = b * 10 + (unsigned)(s[i] - '0'); ++i; }
    if (i == d0 || i != len || b > 0xFFFFFFFFull) { out->odd = true; return; }
    WigToken t; t.first = (uint32_t)a; t.count = (uint32_t)b;
    out->tok.push_back(t);
  }
}

bool parse_parallel (const std::vector<char>& buf, WigParser& P, unsigned nThreads) {
  const char* base = buf.data();
  const size_t N = buf.size();
  std::vector<WigChunk> chunks(nThreads);
  std::vector<std::thread> pool;
  const auto tB = std::chrono::steady_clock::now();
  for (unsigned t = 0; t < nThreads; ++t)
    pool.emplace_back(tokenise_range, base, N, N * t / nThreads, N * (t + 1) / nThreads, &chunks[t]);
  for (std::thread& th : pool) th.join();
  if (getenv("UPB_WIG_DEBUG")) fprintf(stderr, "[wig] tokenise %.4f s\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tB).count());
  if (getenv("UPB_WIG_DEBUG")) { size_t nt = 0, ns = 0, odd = 0; for (const WigChunk& c : chunks) { nt += c.tok.size(); ns += c.special.size(); odd += c.odd; } fprintf(stderr, "[wig] %u ranges, %zu tokens, %zu special, %zu odd\n", nThreads, nt, ns, odd); }
  for (const WigChunk& c : chunks) if (c.odd) return false;
  for (const WigChunk& c : chunks)
    for (const WigToken& t : c.tok) {
      if (t.first == 0xFFFFFFFFu) {
        const std::pair<size_t, size_t>& sp = c.special[t.count];
        if (!P.line(base + sp.first, sp.second, 0)) return false;
      } else {
        // what line() does for a line that starts with a digit, without the text
        if (P.format == 0) return false;
        P.inHeader = false;
        if (P.contig == P.C) continue;
        if (!P.entry(t.first, t.count, 0)) return false;
      }
    }
  return true;
}

}  // namespace

static unsigned g_wigBudget = 0;     // threads one file may use (0: the whole machine); read_samples divides the machine among its files

void read_wig (const std::string& fname, const ContigTable& contigs, int offset, unsigned useLength, Sample* out) {
  Sample& S = *out;
  const uint32_t C = contigs.size();
  reset_sample(S, fname, C);
  std::vector<char> buf;
  std::string shown;
  const auto tA = std::chrono::steady_clock::now();
  if (!slurp(fname, &buf, &shown, &S.error)) return;
  if (getenv("UPB_WIG_DEBUG")) fprintf(stderr, "[wig] slurp %.4f s, %zu bytes\n", std::chrono::duration<double>(std::chrono::steady_clock::now() - tA).count(), buf.size());
  unsigned nThreads;
  if (const char* forced = getenv("UPB_WIG_THREADS")) nThreads = std::max(1, atoi(forced));   // test hook: 1 = always sequential, n = n ranges whatever the size
  else {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    const unsigned budget = g_wigBudget ? g_wigBudget : hw;
    nThreads = buf.size() < (8u << 20) ? 1u : std::min<unsigned>(budget, (unsigned)(buf.size() >> 22));   // at least 4 MB per thread
  }
  if (nThreads > 1) {
    WigParser P(S, contigs, offset, useLength, fname, shown);
    if (parse_parallel(buf, P, nThreads)) {
      if (S.expectedTags == 0) S.expectedTags = S.confidentTags;   // "count tags the hard way", misc/format.cpp:754-758
      return;
    }
    reset_sample(S, fname, C);                                       // something was not plainly fine: the sequential path decides
  }
  WigParser P(S, contigs, offset, useLength, fname, shown);
  parse_sequential(buf, P);
  if (S.error.empty() && S.expectedTags == 0) S.expectedTags = S.confidentTags;
}

std::vector<Sample> read_samples (const std::vector<std::string>& fnames, const ContigTable& contigs,
                                  const std::vector<short>& offsets, unsigned useLength) {
  std::vector<Sample> samples(fnames.size());
  unsigned nThreads = std::thread::hardware_concurrency();
  if (nThreads == 0) nThreads = 4;
  nThreads = std::min<unsigned>(nThreads, (unsigned)fnames.size());
  int nStdin = 0;
  for (const std::string& f : fnames) nStdin += f == "stdin";
  if (nStdin) nThreads = 1;
  g_wigBudget = std::max(1u, std::max(1u, std::thread::hardware_concurrency()) / std::max(1u, nThreads));   // what is left of the machine per file
  std::vector<std::thread> pool;
  for (unsigned t = 0; t < nThreads; ++t)
    pool.emplace_back([&, t] () {
      for (size_t i = t; i < fnames.size(); i += nThreads) {
        const int off = offsets.empty() ? 0 : (offsets.size() > 1 ? offsets[i] : offsets[0]);
        read_wig(fnames[i], contigs, off, useLength, &samples[i]);
      }
    });
  for (std::thread& th : pool) th.join();
  return samples;
}

// ---------------------------------------------------------------------------------------------------------------------------
// alignments: BED (misc/format.cpp:287-312), SAM (:567-610), tag-frequency wiggle; format detection :247-284; the common tail
// (read length, -s offset, contig bounds) :654-678
// ---------------------------------------------------------------------------------------------------------------------------
namespace {

struct AlignSink {
  const ContigTable& contigs;
  const AlignOpts& o;
  AlignBatch& b;
  // returns false when the tag was dropped
  bool add (bool forward, uint32_t contig, uint32_t first, uint32_t last, uint32_t count) {
    if (count == 0 || contig == contigs.size()) return false;
    if (o.useLength != 0) last = forward ? first + (o.useLength - 1) : first - (o.useLength - 1);
    if (o.offset != 0) {
      if ((forward && (int)first > -o.offset) || (!forward && (int)last > o.offset)) {
        first += (uint32_t)(forward ? o.offset : -o.offset);
        last += (uint32_t)(forward ? o.offset : -o.offset);
      } else { b.outOfBoundsTags += count; return false; }
    }
    const uint32_t size = contigs.sizes[contig];
    if (first == 0 || first > size || last == 0 || last > size) { b.outOfBoundsTags += count; return false; }
    b.confidentTags += count;
    b.key.push_back((forward ? 0u : contigs.size()) + contig);
    b.pos.push_back(first);
    b.cnt.push_back(count);
    return true;
  }
};

// boost::split(fields, line, is_any_of(seps)) without token compression: empty fields are kept
void split_any (const char* s, size_t len, const char* seps, std::vector<std::pair<const char*, size_t> >* out) {
  out->clear();
  size_t a = 0;
  for (size_t i = 0; i <= len; ++i)
    if (i == len || strchr(seps, s[i])) { out->push_back(std::make_pair(s + a, i - a)); a = i + 1; }
}

bool all_digits (const char* s, size_t n) { if (!n) return false; for (size_t i = 0; i < n; ++i) if (!isdigit((unsigned char)s[i])) return false; return true; }

// BED_REGEX  ^[^\s]+\s\d+\s\d+\s[^\s]*\s[^\s]*?\s[+-]   (misc/format.hpp:24)
bool looks_like_bed (const char* s, size_t len) {
  size_t i = 0;
  auto nonspace = [&] (bool atLeastOne) { size_t a = i; while (i < len && !isspace((unsigned char)s[i])) ++i; return !atLeastOne || i > a; };
  auto space = [&] () { if (i < len && isspace((unsigned char)s[i])) { ++i; return true; } return false; };
  auto digits = [&] () { size_t a = i; while (i < len && isdigit((unsigned char)s[i])) ++i; return i > a; };
  if (!nonspace(true) || !space() || !digits() || !space() || !digits() || !space()) return false;
  nonspace(false); if (!space()) return false;
  nonspace(false); if (!space()) return false;
  return i < len && (s[i] == '+' || s[i] == '-');
}

// SAM_HEADER_REGEX  ^@[A-Za-z][A-Za-z](\t[A-Za-z][A-Za-z0-9]:[ -~]+)+$
bool looks_like_sam_header (const char* s, size_t len) {
  if (len < 3 || s[0] != '@' || !isalpha((unsigned char)s[1]) || !isalpha((unsigned char)s[2])) return false;
  size_t i = 3;
  int groups = 0;
  while (i < len) {
    if (s[i] != '\t' || i + 4 > len || !isalpha((unsigned char)s[i + 1]) || !isalnum((unsigned char)s[i + 2]) || s[i + 3] != ':') return false;
    i += 4;
    size_t a = i;
    while (i < len && s[i] >= ' ' && s[i] <= '~') ++i;
    if (i == a) return false;
    ++groups;
  }
  return groups > 0;
}

}  // namespace

void read_alignments (const std::string& fname, const ContigTable& contigs, const AlignOpts& opts, AlignBatch* out) {
  AlignBatch& B = *out;
  B = AlignBatch();
  B.name = fname_prefix(fname);
  if (ends_with(fname, ".bam")) { B.error = fname + ": BAM input is not supported by this build (convert it with samtools view -h)"; return; }
  std::vector<char> buf;
  std::string shown;
  if (!slurp(fname, &buf, &shown, &B.error)) return;
  AlignSink sink{contigs, opts, B};
  const uint32_t C = contigs.size();
  // MAPQ threshold exactly as the reference derives it (a double, misc/format.cpp:166-168)
  const double phred = opts.probThreshold == 0 ? 0 : opts.probThreshold == 1 ? 255 : -10 * log10(1 - opts.probThreshold);
  bool forward = true;              // wiggle state
  uint32_t contig = 0;
  uint64_t lineNo = 0;
  std::vector<std::pair<const char*, size_t> > f;
  const char* base = buf.data();
  size_t p = 0;
  const size_t N = buf.size();
  auto fail = [&] (const char* msg) { B.error = line_error(msg, shown, lineNo); };
  while (p < N) {
    const char* nl = (const char*)memchr(base + p, '\n', N - p);
    const size_t e = nl ? (size_t)(nl - base) : N;
    const char* s = base + p;
    const size_t len = e - p;
    p = e + 1;
    ++lineNo;
    if (len == 0 || s[0] == '#') continue;
    if (B.format == 0) {
      if (len >= 5 && memcmp(s, "track", 5) == 0) {
        std::string nm;
        const bool named = match_name(s, len, "name=\"", '"', &nm) || match_name(s, len, "name=", ' ', &nm);
        if (named) B.name = nm;
        if (std::string(s, len).find("type=wiggle_0") != std::string::npos) {
          std::string b2; bool fw;
          if (split_directional(B.name, &b2, &fw)) { B.format = 6; B.name = b2; forward = fw; }
          else { B.format = 7; B.name = named ? nm : std::string(); forward = true; }
        } else B.format = 1;
        continue;
      } else if (looks_like_bed(s, len)) B.format = 1;
      else if (looks_like_sam_header(s, len)) B.format = 4;
      else if (s[0] == '>') { B.error = fname + ": Eland / Corona input is not supported by this build"; return; }
      else { fail("bad format"); return; }
    }
    if (B.format == 1) {                                             // BED
      if (len >= 5 && memcmp(s, "track", 5) == 0) continue;
      split_any(s, len, "\t ", &f);
      if (f.size() < 6) { fail("bad format"); return; }
      const char strand = f[5].second ? f[5].first[0] : '\0';
      if (strand != '+' && strand != '-') { fail("bad format"); return; }
      const bool fw = strand == '+';
      uint32_t a, b2;
      if (!lexical_u32(f[1].first, f[1].second, &a) || !lexical_u32(f[2].first, f[2].second, &b2)) { fail("bad format"); return; }
      ++B.totalTags;
      sink.add(fw, contigs.find(std::string(f[0].first, f[0].second)), fw ? a + 1 : b2, fw ? b2 : a + 1, 1);
    } else if (B.format == 4) {                                      // SAM
      if (s[0] == '@') continue;
      split_any(s, len, "\t", &f);
      if (f.size() < 10) { fail("bad format"); return; }
      uint32_t flag;
      if (!lexical_u32(f[1].first, f[1].second, &flag) || flag > 0xFFFF) { fail("bad format"); return; }
      if (flag & 0x0100) continue;                                   // non-primary
      ++B.totalTags;
      if (flag & (0x0004 + 0x0200)) continue;                        // unmapped or failed QC
      if (f.size() > 11) {
        long edit = 0;
        for (size_t k = 0; k < f.size(); ++k)
          if (f[k].second > 5 && memcmp(f[k].first, "NM:i:", 5) == 0 && all_digits(f[k].first + 5, f[k].second - 5)) { edit = strtol(std::string(f[k].first + 5, f[k].second - 5).c_str(), nullptr, 10); break; }
        if (edit > (long)opts.mismatchTolerance) continue;
      }
      if (phred != 0 && phred != 255) {
        uint32_t mapq;
        if (!lexical_u32(f[4].first, f[4].second, &mapq) || mapq > 0xFFFF) { fail("bad format"); return; }
        if ((double)mapq < phred) { ++B.rejectTags; continue; }
      }
      const uint32_t c = contigs.find(std::string(f[2].first, f[2].second));
      if (c == C) continue;
      const uint32_t seqLen = (uint32_t)f[9].second;
      if (opts.useLength != 0 && seqLen < opts.useLength) { fail("sequence shorter than requested length"); return; }
      const bool fw = !(flag & 0x0010);
      uint32_t pos1;
      if (!lexical_u32(f[3].first, f[3].second, &pos1)) { fail("bad format"); return; }
      const uint32_t first = pos1 + (fw ? 0u : seqLen - 1);
      const uint32_t span = (opts.useLength != 0 ? opts.useLength : seqLen) - 1;
      sink.add(fw, c, first, fw ? first + span : first - span, 1);
    } else {                                                          // tag-frequency wiggle (misc/format.cpp:503-565)
      if (isdigit((unsigned char)s[0])) {
        if (contig == C) continue;
        size_t delim = len;
        while (delim-- > 0) if (s[delim] == '\t' || s[delim] == ' ') break;
        if (delim == (size_t)-1) { fail("bad format"); return; }
        uint32_t first, count;
        const size_t cs = delim + ((delim + 1 < len && s[delim + 1] == '-') ? 2 : 1);
        if (!lexical_u32(s, delim, &first) || cs > len || !lexical_u32(s + cs, len - cs, &count)) { fail("bad format"); return; }
        B.totalTags += count;
        uint32_t last = first;
        if (opts.useLength != 0) last = (B.format == 6 && !forward) ? first - (opts.useLength - 1) : first + (opts.useLength - 1);
        sink.add(forward, contig, first, last, count);
      } else if (len > 19 && memcmp(s, "variableStep chrom=", 19) == 0) {
        contig = contigs.find(std::string(s + 19, len - 19));
      } else if (len >= 5 && memcmp(s, "track", 5) == 0) {
        std::string nm;
        if (!match_name(s, len, "name=\"", '"', &nm)) { fail("bad format"); return; }
        contig = C;
        if (B.format == 6) {
          std::string b2; bool fw;
          if (!split_directional(nm, &b2, &fw)) { fail("strand not defined"); return; }
          B.name = b2; forward = fw;
        } else B.name = nm;
      } else { fail("bad format"); return; }
    }
  }
}

std::string align_summary (const AlignBatch& b, double prob) {
  std::string s;
  if (!b.name.empty()) s += b.name + ": ";
  s += std::to_string(b.totalTags) + " tags";
  if (b.format == 4) {
    s += "\n  " + std::to_string(b.confidentTags) + (prob != 0 ? " confidently mapped" : " unique best") + " hits (" +
         fmt_fixed(100 * (double)b.confidentTags / (double)b.totalTags, 1) + "%)";
    if (prob != 0) s += "\n  " + std::to_string(b.rejectTags) + " unique best hits rejected by filter (" + fmt_fixed(100 * (double)b.rejectTags / (double)b.totalTags, 1) + "%)";
  }
  if (b.outOfBoundsTags > 0) s += "\n" + std::to_string(b.outOfBoundsTags) + " out of contig bounds";
  s += "\n";
  return s;
}

Flat::~Flat () {
  if (pinned) { upb_host_free(pos); upb_host_free(cnt); } else { free(pos); free(cnt); }
}

void flatten (const std::vector<Sample>& samples, uint32_t C, Flat* out) {
  const size_t n = samples.size();
  out->off.assign(n * 2 * ((size_t)C + 1), 0);
  uint64_t run = 0;
  for (size_t s = 0; s < n; ++s)
    for (int d = 0; d < 2; ++d) {
      uint64_t* o = &out->off[(s * 2 + d) * ((size_t)C + 1)];
      for (uint32_t c = 0; c < C; ++c) { o[c] = run; run += samples[s].pos[d][c].size(); }
      o[C] = run;
    }
  out->n = run;
  const size_t bytes = (size_t)(run ? run : 1) * 4;
  void *pp = nullptr, *pc = nullptr;
  if (upb_host_alloc(&pp, bytes) == UPB_OK && upb_host_alloc(&pc, bytes) == UPB_OK) out->pinned = true;
  else { if (pp) upb_host_free(pp); pp = malloc(bytes); pc = malloc(bytes); out->pinned = false; }
  if (!pp || !pc) die("out of host memory");
  out->pos = (uint32_t*)pp;
  out->cnt = (uint32_t*)pc;
  for (size_t s = 0; s < n; ++s)
    for (int d = 0; d < 2; ++d) {
      const uint64_t* o = &out->off[(s * 2 + d) * ((size_t)C + 1)];
      for (uint32_t c = 0; c < C; ++c) {
        const std::vector<uint32_t>& vp = samples[s].pos[d][c];
        if (vp.empty()) continue;
        memcpy(out->pos + o[c], vp.data(), vp.size() * 4);
        memcpy(out->cnt + o[c], samples[s].cnt[d][c].data(), vp.size() * 4);
      }
    }
}

// the 2-byte form of one (sample, strand, contig) segment: include/unipeak_b200.h (upb_pack_tags2 is the same encoder over flat arrays)
static void pack_segment (const uint32_t* pos, const uint32_t* cnt, size_t n, uint64_t first, uint16_t* words, std::vector<upb_packed_exception>* exc) {
  uint32_t prev = 0;
  for (size_t i = 0; i < n; ++i) {
    const uint32_t p = pos[i], k = cnt[i];
    const bool escP = (i & 2047) == 0 || p < prev || p - prev >= 0xFFFu;
    const bool escC = k >= 0xFu;
    if (escP || escC) { upb_packed_exception x; x.index = first + i; x.pos = escP ? p : 0u; x.cnt = escC ? k : 0u; exc->push_back(x); }
    words[first + i] = (uint16_t)((escP ? 0xFFFu : p - prev) | ((escC ? 0xFu : k) << 12));
    prev = p;
  }
}

TagLoader::~TagLoader () {
  if (words_) { if (wordsPinned_) upb_host_free(words_); else free(words_); }
  if (vbytes_) { if (vbytesPinned_) upb_host_free(vbytes_); else free(vbytes_); }
}

// which host form the drivers load: the variable-length byte form (1 - 6 bytes per entry, ~1.5 on real tracks: never more than the plain
// form's 8); UPB_HOST_FORM = packed2 (the 2-byte form where it pays, else plain) | plain | vbyte
static int host_form () {
  const char* e = getenv("UPB_HOST_FORM");
  if (e && !strcmp(e, "packed2")) return 1;
  if (e && !strcmp(e, "plain")) return 0;
  return 2;
}

void TagLoader::encode (const std::vector<Sample>& samples, uint32_t C, bool directionalHint, upb_tag_source* srcOut) {
  const size_t n = samples.size();
  off_.assign(n * 2 * ((size_t)C + 1), 0);
  uint64_t run = 0;
  for (size_t s = 0; s < n; ++s)
    for (int d = 0; d < 2; ++d) {
      uint64_t* o = &off_[(s * 2 + d) * ((size_t)C + 1)];
      for (uint32_t c = 0; c < C; ++c) { o[c] = run; run += samples[s].pos[d][c].size(); }
      o[C] = run;
    }
  n_ = run;
  // escapes without encoding anything: a first pass over the gaps decides the form (16 bytes per exception against 6 saved per entry)
  const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  std::vector<uint64_t> nEsc(n, 0);
  {
    std::vector<std::thread> th;
    for (size_t s = 0; s < n; ++s) th.emplace_back([&, s] {
      uint64_t e = 0;
      for (int d = 0; d < 2; ++d) for (uint32_t c = 0; c < C; ++c) {
        const std::vector<uint32_t>& vp = samples[s].pos[d][c]; const std::vector<uint32_t>& vc = samples[s].cnt[d][c];
        uint32_t prev = 0;
        for (size_t i = 0; i < vp.size(); ++i) { e += ((i & 2047) == 0 || vp[i] < prev || vp[i] - prev >= 0xFFFu || vc[i] >= 0xFu) ? 1 : 0; prev = vp[i]; }
      }
      nEsc[s] = e;
    });
    for (std::thread& t : th) t.join();
  }
  uint64_t escTotal = 0;
  for (uint64_t e : nEsc) escTotal += e;
  upb_tag_source src;
  memset(&src, 0, sizeof src);
  src.n_samples = (uint32_t)n; src.off = off_.data(); src.directional_hint = directionalHint ? 1 : 0;
  const int form = host_form();
  if (form == 2) {
    // the variable-length byte form, by the library's encoder over the flat arrays (one task per (sample, strand))
    flatten(samples, C, &flat_);
    uint64_t nb = 0, nt = 0, nx = 0;
    upb_pack_tags_vbyte((uint32_t)n, C, off_.data(), flat_.pos, flat_.cnt, nullptr, 0, &nb, nullptr, nullptr, 0, &nt, nullptr, 0, &nx);
    void* pb = nullptr;
    const size_t bytes = (size_t)(nb ? nb : 1);
    if (upb_host_alloc(&pb, bytes) == UPB_OK) vbytesPinned_ = true; else { pb = malloc(bytes); vbytesPinned_ = false; }
    if (!pb) die("out of host memory");
    vbytes_ = (uint8_t*)pb;
    vbyteOff_.assign(off_.size(), 0); vtileOff_.assign((size_t)nt + 1, 0); exc_.resize((size_t)nx);
    if (upb_pack_tags_vbyte((uint32_t)n, C, off_.data(), flat_.pos, flat_.cnt, vbytes_, bytes, &nb, vbyteOff_.data(), vtileOff_.data(), nt + 1, &nt,
                            exc_.empty() ? nullptr : exc_.data(), nx, &nx) != UPB_OK) die("encoding the tag arrays failed");
    src.vbytes = vbytes_; src.vbyte_off = vbyteOff_.data(); src.vtile_off = vtileOff_.data(); src.n_vtiles = nt;
    src.n_exceptions = exc_.size(); src.exceptions = exc_.empty() ? nullptr : exc_.data();
  } else if (form == 1 && 2 * run + 16 * escTotal < 6 * run) {
    void* pw = nullptr;
    const size_t bytes = (size_t)(run ? run : 1) * 2;
    if (upb_host_alloc(&pw, bytes) == UPB_OK) wordsPinned_ = true; else { pw = malloc(bytes); wordsPinned_ = false; }
    if (!pw) die("out of host memory");
    words_ = (uint16_t*)pw;
    std::vector<std::vector<upb_packed_exception> > per(n);
    std::vector<std::thread> th;
    for (size_t s = 0; s < n; ++s) th.emplace_back([&, s] {
      per[s].reserve((size_t)nEsc[s]);
      for (int d = 0; d < 2; ++d) {
        const uint64_t* o = &off_[(s * 2 + d) * ((size_t)C + 1)];
        for (uint32_t c = 0; c < C; ++c) {
          const std::vector<uint32_t>& vp = samples[s].pos[d][c];
          if (!vp.empty()) pack_segment(vp.data(), samples[s].cnt[d][c].data(), vp.size(), o[c], words_, &per[s]);
        }
      }
    });
    for (std::thread& t : th) t.join();
    exc_.reserve((size_t)escTotal);
    for (size_t s = 0; s < n; ++s) exc_.insert(exc_.end(), per[s].begin(), per[s].end());   // samples are laid out one after the other: sorted by index
    src.words = words_; src.n_exceptions = exc_.size(); src.exceptions = exc_.empty() ? nullptr : exc_.data();
  } else {
    flatten(samples, C, &flat_);
    src.pos = flat_.pos; src.cnt = flat_.cnt;
  }
  packSeconds_ = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  *srcOut = src;
}

int TagLoader::start (upb_ctx* ctx, const std::vector<Sample>& samples, uint32_t C, bool directionalHint) {
  upb_tag_source src;
  encode(samples, C, directionalHint, &src);
  return upb_load_tags_async(ctx, &src);
}

int TagLoader::start (upb_multi* multi, const std::vector<Sample>& samples, uint32_t C, bool directionalHint) {
  upb_tag_source src;
  encode(samples, C, directionalHint, &src);
  return upb_multi_load_tags(multi, &src);
}

// ---------------------------------------------------------------------------------------------------------------------------
// Device: one GPU, or UPB_GPUS=N of them
// ---------------------------------------------------------------------------------------------------------------------------
static Device* g_pendingDevice = nullptr;      // a device whose creation may still be running
static void join_pending_device () { if (g_pendingDevice) g_pendingDevice->join(); }   // exit() (die, usage errors) must not tear the process down under a thread inside the driver

Device::Device () {
  if (const char* e = getenv("UPB_GPUS")) want_ = atoi(e);
  if (want_ < 1) want_ = 1;
  // Driver initialisation + context creation take 0.3 - 2 s (more than the reference's whole run on a small input): they run here, on a thread
  // of their own, while the caller parses contig table and wiggle files — nothing in the reference's front end waits for a device
  static const bool registered = (atexit(join_pending_device), true);
  (void)registered;
  g_pendingDevice = this;
  init_ = new std::thread([this] {
    if (want_ > 1) {
      const int have = upb_device_count();
      if (want_ > have) { initRc_ = UPB_ERR_NO_DEVICE; initMsg_ = "UPB_GPUS=" + std::to_string(want_) + " but only " + std::to_string(have) + " CUDA device(s) are visible"; return; }
      initRc_ = upb_multi_create(want_, nullptr, &multi_);
    } else initRc_ = upb_create(0, &ctx_);
    if (initRc_) { const char* m = upb_last_error(nullptr); initMsg_ = (m && *m) ? m : "device call failed"; }
  });
  if (getenv("UPB_SYNC_INIT")) join();                               // (A/B of the overlap: create the context before parsing, as r01 did in effect)
}
void Device::join () {
  if (!init_) return;
  std::thread* t = (std::thread*)init_;
  init_ = nullptr;
  t->join();
  delete t;
  if (g_pendingDevice == this) g_pendingDevice = nullptr;
}
void Device::ready () {
  join();
  if (initRc_) { const int rc = initRc_; initRc_ = 0; die(initMsg_ + " (unipeak_b200 error " + std::to_string(rc) + ")"); }
}
Device::~Device () { join(); if (multi_) upb_multi_destroy(multi_); if (ctx_) upb_destroy(ctx_); }
void Device::fail (int rc) {
  if (!multi_) die_upb(ctx_, rc);
  const char* m = upb_multi_last_error(multi_);
  die(std::string((m && *m) ? m : "device call failed") + " (unipeak_b200 error " + std::to_string(rc) + ")");
}
void Device::setContigs (uint32_t C, const uint32_t* sizes) {
  ready();
  const int rc = multi_ ? upb_multi_set_contigs(multi_, C, sizes) : upb_set_contigs(ctx_, C, sizes);
  if (rc) fail(rc);
}
void Device::load (TagLoader& loader, const std::vector<Sample>& samples, uint32_t C, bool directionalHint, bool wait) {
  ready();
  int rc = multi_ ? loader.start(multi_, samples, C, directionalHint) : loader.start(ctx_, samples, C, directionalHint);
  if (!rc && wait) rc = multi_ ? upb_multi_load_wait(multi_) : upb_load_wait(ctx_);
  if (rc) fail(rc);
}
void Device::loadBySample (const std::vector<Sample>& samples, uint32_t C, bool directionalHint) {
  ready();
  int rc;
  if (multi_) {
    flatten(samples, C, &flat_);
    rc = upb_multi_load_tags_by_sample(multi_, (uint32_t)samples.size(), flat_.off.data(), flat_.pos, flat_.cnt);
  } else {
    rc = sampleLoader_.start(ctx_, samples, C, directionalHint);
    if (!rc) rc = upb_load_wait(ctx_);
  }
  if (rc) fail(rc);
}
void Device::callRegions (const upb_region_params* p, upb_region_table* out) {
  ready();
  const int rc = multi_ ? upb_multi_call_regions(multi_, p, out) : upb_call_regions(ctx_, p, out);
  if (rc) fail(rc);
}
void Device::profileContig (const upb_region_params* p, int buffer, uint32_t contig, upb_profile* out) {
  ready();
  const int rc = multi_ ? upb_multi_profile_contig(multi_, p, buffer, contig, out) : upb_profile_contig(ctx_, p, buffer, contig, out);
  if (rc) fail(rc);
}
int Device::strandShift (const upb_shift_params* p, upb_shift_report* out) {
  ready();
  const int rc = multi_ ? upb_multi_strand_shift(multi_, p, out) : upb_strand_shift(ctx_, p, out);
  if (rc && rc != UPB_ERR_NO_REGIONS) { fprintf(stderr, "\n"); fail(rc); }
  return rc;
}
void Device::countTagsInRegions (int directional, uint32_t extension, uint64_t nRegions, const upb_region_key* regions, uint32_t* countsOut) {
  ready();
  const int rc = multi_ ? upb_multi_count_tags_in_regions(multi_, directional, extension, nRegions, regions, countsOut)
                        : upb_count_tags_in_regions(ctx_, directional, extension, nRegions, regions, countsOut);
  if (rc) fail(rc);
}

// ---------------------------------------------------------------------------------------------------------------------------
// number formatting
// ---------------------------------------------------------------------------------------------------------------------------
std::string fmt_lexical (double x) { char b[64]; snprintf(b, sizeof b, "%.17g", x); return b; }
std::string fmt_g (double x) { char b[64]; snprintf(b, sizeof b, "%g", x); return b; }
std::string fmt_fixed (double x, int digits) { char b[400]; snprintf(b, sizeof b, "%.*f", digits, x); return b; }

// %g with the default precision 6 for a finite, non-negative double — the profile wiggle writes one per non-zero position
// (misc/format.cpp:1112-1129), so it is worth not going through printf: scale to a 6-digit integer, round half to even on
// the exact value (via long double / fma-free remainder check), strip trailing zeros.  Falls back to snprintf in the rare
// cases where the fast scaling cannot guarantee printf's correctly rounded result.
int fast_g6 (double x, char* buf) {
  if (!(x > 0) || !std::isfinite(x) || x < 1e-4 || x >= 1e15) return snprintf(buf, 32, "%g", x);
  // decimal exponent e with 10^e <= x < 10^(e+1)
  int e = (int)std::floor(std::log10(x));
  static const double P10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
  auto scaled = [&] (int ee) -> long double {                      // x * 10^(5-ee) in extended precision (64-bit mantissa)
    const int k = 5 - ee;
    return k >= 0 ? (long double)x * (long double)P10[k] : (long double)x / (long double)P10[-k];
  };
  long double v = scaled(e);
  if (v < 100000.0L) { --e; v = scaled(e); } else if (v >= 1000000.0L) { ++e; v = scaled(e); }
  const long double fl = std::floor(v);
  const long double frac = v - fl;
  // the extended product carries ~1e-19 relative error; anything that close to a tie goes to printf
  if (std::fabs((double)(frac - 0.5L)) < 1e-9) return snprintf(buf, 32, "%g", x);
  long m = (long)fl + (frac > 0.5L ? 1 : 0);
  if (m >= 1000000) { m /= 10; ++e; }
  if (e < -4 || e >= 6) {
    // scientific: d.ddddde+XX with trailing zeros removed
    char digits[8];
    for (int i = 5; i >= 0; --i) { digits[i] = (char)('0' + m % 10); m /= 10; }
    int nd = 6;
    while (nd > 1 && digits[nd - 1] == '0') --nd;
    char* o = buf;
    *o++ = digits[0];
    if (nd > 1) { *o++ = '.'; for (int i = 1; i < nd; ++i) *o++ = digits[i]; }
    *o++ = 'e';
    *o++ = e < 0 ? '-' : '+';
    const int ae = e < 0 ? -e : e;
    if (ae >= 100) { *o++ = (char)('0' + ae / 100); }
    *o++ = (char)('0' + (ae / 10) % 10);
    *o++ = (char)('0' + ae % 10);
    *o = 0;
    return (int)(o - buf);
  }
  char digits[8];
  for (int i = 5; i >= 0; --i) { digits[i] = (char)('0' + m % 10); m /= 10; }
  char* o = buf;
  if (e >= 0) {
    for (int i = 0; i <= e; ++i) *o++ = digits[i];
    int nd = 6;
    while (nd > e + 1 && digits[nd - 1] == '0') --nd;
    if (nd > e + 1) { *o++ = '.'; for (int i = e + 1; i < nd; ++i) *o++ = digits[i]; }
  } else {
    *o++ = '0'; *o++ = '.';
    for (int i = -1; i > e; --i) *o++ = '0';
    int nd = 6;
    while (nd > 1 && digits[nd - 1] == '0') --nd;
    for (int i = 0; i < nd; ++i) *o++ = digits[i];
  }
  *o = 0;
  return (int)(o - buf);
}

// ---------------------------------------------------------------------------------------------------------------------------
// output file
// ---------------------------------------------------------------------------------------------------------------------------
Out::Out (const std::string& fname) {
  buf_.reserve(1 << 20);
  if (fname == "stdout") { f_ = stdout; isStdout_ = true; return; }
  if (ends_with(fname, ".bz2")) die("bzip2 output is not supported by this build (no bzlib.h): " + fname);
  if (ends_with(fname, ".gz")) {
    gz_ = gzopen(fname.c_str(), "wb");
    if (!gz_) { fprintf(stderr, "error: could not write %s\n\n", fname.c_str()); exit(1); }
    return;
  }
  f_ = fopen(fname.c_str(), "wb");
  if (!f_) { fprintf(stderr, "error: could not write %s\n\n", fname.c_str()); exit(1); }
}

void Out::flush () {
  if (buf_.empty()) return;
  if (gz_) gzwrite((gzFile)gz_, buf_.data(), (unsigned)buf_.size());
  else if (f_) fwrite(buf_.data(), 1, buf_.size(), f_);
  buf_.clear();
}

void Out::write (const char* p, size_t n) {
  if (buf_.size() + n > (1u << 20)) flush();
  if (n > (1u << 20)) { if (gz_) gzwrite((gzFile)gz_, p, (unsigned)n); else if (f_) fwrite(p, 1, n, f_); return; }
  buf_.insert(buf_.end(), p, p + n);
}

void Out::close () {
  flush();
  if (gz_) { gzclose((gzFile)gz_); gz_ = nullptr; }
  if (f_) { if (isStdout_) fflush(f_); else fclose(f_); f_ = nullptr; }
}

Out::~Out () { close(); }

}  // namespace upbhost
