// bin/tags_in_regions — drop-in for the reference's counter of extra samples' tags in an existing regions table
// (src/tags_in_regions.cpp).  Same flags, stderr messages and table text; the per-row, per-sample counting (including the
// reference's stream-order behaviour) runs on the device through upb_count_tags_in_regions.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "upb_host.hpp"

using namespace upbhost;

namespace {

// InStream::readLine over a file read whole (misc/filterstream.cpp:62-72): good() stays true until a read hits the end
struct Lines {
  std::vector<char> buf;
  size_t p = 0;
  uint64_t lineNo = 0;
  bool eof = false;
  bool good () const { return !eof; }
  std::string next () {
    ++lineNo;
    if (p >= buf.size()) { eof = true; return std::string(); }
    const char* nl = (const char*)memchr(buf.data() + p, '\n', buf.size() - p);
    const size_t e = nl ? (size_t)(nl - buf.data()) : buf.size();
    std::string s(buf.data() + p, e - p);
    if (!nl) eof = true;
    p = e + 1;
    return s;
  }
};

bool read_all (const std::string& fname, std::vector<char>* buf) {
  FILE* f = fname == "stdin" ? stdin : fopen(fname.c_str(), "rb");
  if (!f) return false;
  char tmp[1 << 16];
  size_t k;
  while ((k = fread(tmp, 1, sizeof tmp, f)) > 0) buf->insert(buf->end(), tmp, tmp + k);
  if (f != stdin) fclose(f);
  return true;
}

}  // namespace

int main (int argc, char* argv[]) {
  // the TCLAP declarations of src/tags_in_regions.cpp:36-46, in the same order
  const std::vector<ArgSpec> specs = {
    {'D', "non-directional", ARG_SWITCH, false, "don't separate strands (data aren't from a directional protocol), default false", ""},
    {'i', "mismatches", ARG_U16, false, "mismatch tolerance, default 2 (Eland extended only)", "non-negative integer"},
    {'l', "length", ARG_U16, false, "sequence length to use, default entire sequence (Eland extended only)", "positive integer"},
    {'s', "shift", ARG_STRING, false, "number of bases downstream to shift alignments (barcode length and/or strand shift), default 0", "comma-separated integers"},
    {'p', "prob", ARG_DOUBLE, false, "posterior probability threshold for alignments", "probability"},
    {'e', "extend", ARG_U32, false, "extend region boundaries by a fixed amount", "positive integer"},
    {'o', "out", ARG_STRING, true, "output regions filename", "filename"},
    {'f', "in", ARG_STRING, true, "input regions filename", "filename"},
    {'c', "contig", ARG_STRING, true, "contig table filename", "filename"},
  };
  CmdLine cmd(specs, ".wig files of alignments", "alignment filenames");
  cmd.parse(argc, argv);
  const std::string regionsInFname = cmd.str('f'), regionsOutFname = cmd.str('o'), contigFname = cmd.str('c'), offsetString = cmd.str('s');
  const uint32_t extension = cmd.u32('e', 0);
  const uint16_t useLength = cmd.u16('l', 0);
  const std::vector<std::string>& alignFnames = cmd.unlabeled();
  const bool directional = !cmd.sw('D');

  const std::vector<short> offsets = parse_offsets(offsetString, alignFnames.size());
  Device dev;                                                        // one GPU, or UPB_GPUS=N of them; its CUDA context is created in the background while the inputs are parsed
  const ContigTable contigs = parse_contig_table(contigFname);
  const uint32_t C = contigs.size();

  fprintf(stderr, "reading alignment files...\n");
  std::vector<Sample> samples = read_samples(alignFnames, contigs, offsets, useLength);
  for (const Sample& s : samples) {
    if (!s.error.empty()) die(s.error);
    fprintf(stderr, "  %s: %llu tags\n", s.name.c_str(), (unsigned long long)s.expectedTags);
  }
  const size_t n = samples.size();

  // header (:102-129)
  Lines in;
  if (!read_all(regionsInFname, &in.buf)) { fprintf(stderr, "error: could not read %s\n\n", regionsInFname.c_str()); exit(1); }
  Out out(regionsOutFname);
  std::string line = in.next();
  while (line.empty() || line[0] == '#') {
    if (!in.good()) die("unexpected end of " + regionsInFname + " (no column-name line)");   // the reference's assert(good()) fires here
    out.write(line + "\n");
    line = in.next();
  }
  if (line[0] != '\t') { fprintf(stderr, "error: bad format in %s line %llu\n\n", regionsInFname.c_str(), (unsigned long long)in.lineNo); exit(1); }
  {
    std::string s;
    if (extension != 0) s += "# region_extension=" + fmt_lexical(extension) + "\n";
    for (const std::string& f : alignFnames) s += "# extra_align_file=" + f + "\n";
    if (!offsets.empty()) {
      if (offsets.size() == 1) s += "# shift=" + fmt_lexical(offsets[0]) + "\n";
      else {
        s += "# shifts=";
        for (size_t i = 0; i + 1 < offsets.size(); ++i) s += fmt_lexical(offsets[i]) + ",";
        s += fmt_lexical(offsets.back()) + "\n";
      }
    }
    s += line;
    for (const Sample& smp : samples) s += "\t" + smp.name;
    s += "\n";
    out.write(s);
  }

  // rows (:136-172): parsed up to the first bad line; the rows before it are still counted and written, as the reference's
  // streaming loop would have done before exiting
  fprintf(stderr, "processing regions... ");
  fflush(stderr);
  std::vector<upb_region_key> keys;
  std::vector<std::string> rows;
  std::string pendingError;
  while (in.good()) {
    const std::string row = in.next();
    if (row.empty()) continue;
    const size_t colon = row.find(':'), dash = row.find('-'), tab = row.find('\t');
    const std::string bad = "error: bad format in " + regionsInFname + " line " + std::to_string(in.lineNo) + "\n\n";
    if (tab == std::string::npos || colon == std::string::npos || dash == std::string::npos || dash > tab || colon > dash) { pendingError = bad; break; }
    const uint32_t contig = contigs.find(row.substr(0, colon));
    if (contig == C) { pendingError = "error: contig not in table in " + regionsInFname + " line " + std::to_string(in.lineNo) + "\n\n"; break; }
    upb_region_key k;
    k.contig = contig;
    if (!lexical_u32(row.data() + colon + 1, dash - colon - 1, &k.start) || !lexical_u32(row.data() + dash + 1, tab - dash - 1, &k.end)) { pendingError = bad; break; }
    if (k.end < k.start && !directional) { pendingError = "error: reverse regions in non-directional analysis\n\n"; break; }
    keys.push_back(k);
    rows.push_back(row);
  }

  std::vector<uint32_t> counts(keys.size() * n + 1, 0);
  if (!keys.empty()) {
    dev.setContigs(C, contigs.sizes.data());
    dev.loadBySample(samples, C, directional);
    dev.countTagsInRegions(directional ? 1 : 0, extension, keys.size(), keys.data(), counts.data());
  }
  std::vector<uint64_t> tagsInRegions(n, 0);
  {
    std::string s;
    for (size_t r = 0; r < rows.size(); ++r) {
      s = rows[r];
      for (size_t e = 0; e < n; ++e) { s += "\t" + fmt_lexical(counts[r * n + e]); tagsInRegions[e] += counts[r * n + e]; }
      s += "\n";
      out.write(s);
    }
  }
  out.close();
  if (!pendingError.empty()) { fputs(pendingError.c_str(), stderr); exit(1); }

  // summary (:203-212)
  fprintf(stderr, "%llu in %s\n", (unsigned long long)rows.size(), regionsInFname.c_str());
  fprintf(stderr, "tags in regions:\n");
  for (size_t i = 0; i < n; ++i)
    fprintf(stderr, "  %s: %llu (%s%%)\n", samples[i].name.c_str(), (unsigned long long)tagsInRegions[i],
            fmt_fixed(100 * (double)tagsInRegions[i] / (double)samples[i].expectedTags, 1).c_str());
  fprintf(stderr, "\nDone!\n\n");
  return 0;
}
