// bin/convert_align — drop-in for the reference's alignment-to-wiggle converter (src/convert_align.cpp).  Same flags, stderr
// messages and tag-frequency wiggle text.  The parsers stay on the host (BED, SAM and wiggle input; the reference's Eland,
// Corona and BAM readers are not rebuilt); the aggregation the reference does with one std::map per strand and contig
// (CountMap::add, misc/data.cpp:301-314) runs on the device through upb_aggregate_tags.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "upb_host.hpp"

using namespace upbhost;

int main (int argc, char* argv[]) {
  // the TCLAP declarations of src/convert_align.cpp:38-49, in the same order
  const std::vector<ArgSpec> specs = {
    {'q', "quiet", ARG_SWITCH, false, "don't give contig-by-contig updates", ""},
    {'D', "non-directional", ARG_SWITCH, false, "combine strands (data aren't from a directional protocol), default false", ""},
    {'i', "mismatches", ARG_U16, false, "mismatch tolerance, default 2", "non-negative integer"},
    {'l', "length", ARG_U16, false, "sequence length to use, default entire sequence", "positive integer"},
    {'s', "shift", ARG_I16, false, "number of bases downstream to shift alignments (barcode length + strand shift), default 0", "integer"},
    {'p', "prob", ARG_DOUBLE, false, "posterior probability threshold for alignments, default 0.9", "probability"},
    {'a', "assembly", ARG_STRING, false, "reference assembly name, e.g. hg18", "name"},
    {'n', "name", ARG_STRING, false, "track name", "name"},
    {'c', "contigs", ARG_STRING, true, "contig table filename", "filename"},
    {'o', "output", ARG_STRING, true, "output file", "output filename"},
  };
  CmdLine cmd(specs, "alignment files", "align filenames", "input");
  cmd.parse(argc, argv);
  std::string trackName = cmd.str('n');
  const std::string outFname = cmd.str('o'), contigFname = cmd.str('c'), assembly = cmd.str('a');
  AlignOpts opts;
  opts.mismatchTolerance = cmd.u16('i', 1);
  opts.useLength = cmd.u16('l', 0);
  opts.offset = cmd.i16('s', 0);
  opts.probThreshold = cmd.dbl('p', 0.9);
  const std::vector<std::string>& alignFnames = cmd.unlabeled();
  const bool directional = !cmd.sw('D');
  const bool quiet = cmd.sw('q');

  const ContigTable contigs = parse_contig_table(contigFname);
  const uint32_t C = contigs.size();

  // read alignments (:70-90)
  std::vector<uint32_t> key, pos, cnt;
  for (const std::string& f : alignFnames) {
    fprintf(stderr, "reading %s...\n", f.c_str());
    if (!quiet) fprintf(stderr, "0 tags read");
    AlignBatch b;
    read_alignments(f, contigs, opts, &b);
    if (!b.error.empty()) die(b.error);
    if (!quiet) fprintf(stderr, "\r");
    fputs(align_summary(b, opts.probThreshold).c_str(), stderr);
    key.insert(key.end(), b.key.begin(), b.key.end());
    pos.insert(pos.end(), b.pos.begin(), b.pos.end());
    cnt.insert(cnt.end(), b.cnt.begin(), b.cnt.end());
  }

  // aggregate on the device
  upb_ctx* ctx = nullptr;
  int rc = upb_create(0, &ctx);
  if (rc) die_upb(nullptr, rc);
  if ((rc = upb_set_contigs(ctx, C, contigs.sizes.data()))) die_upb(ctx, rc);
  upb_tag_counts tc;
  if ((rc = upb_aggregate_tags(ctx, key.size(), key.data(), pos.data(), cnt.data(), directional ? 0 : 1, &tc))) die_upb(ctx, rc);
  fprintf(stderr, "%llu usable tags\n", (unsigned long long)tc.tag_count);
  if (tc.tag_count == 0) { fprintf(stderr, "error: nothing to do\n\n"); exit(1); }

  if (trackName.empty()) trackName = fname_prefix(outFname);
  fprintf(stderr, "writing %s in wiggle format with %s... ", outFname.c_str(), directional ? "separate strands" : "strands combined");
  fflush(stderr);
  Out out(outFname);
  std::string s;
  for (const std::string& f : alignFnames) s += "# original_file=" + f + "\n";
  if (opts.probThreshold != 0) s += "# prob_threshold=" + fmt_lexical(opts.probThreshold) + "\n";
  s += "# tags=" + std::to_string(tc.tag_count) + "\n";
  out.write(s);

  // FormatOutStream::write(Alignment) + trackHeader (misc/format.cpp:1043-1075, :1164-1219), ALIGN_PRIORITY = 3
  auto trackHeader = [&] (bool forward) {
    std::string h = "track name=\"" + trackName;
    if (directional) h += forward ? " +" : " -";
    h += "\"";
    if (directional) h += std::string(" description=\"") + (forward ? trackName : std::string(" ")) + "\"";
    h += " priority=3 visibility=";
    if (directional) h += forward ? "full type=wiggle_0 alwaysZero=on color=0,0,255" : "full type=wiggle_0 alwaysZero=on color=255,0,0 altColor=255,0,0";
    else h += "full type=wiggle_0 alwaysZero=on color=191,0,191";
    if (!assembly.empty()) h += " db=" + assembly;
    h += "\n";
    out.write(h);
  };
  bool first = true;
  char line[64];
  for (uint32_t d = 0; d < (directional ? 2u : 1u); ++d) {
    bool headerDone = false;
    for (uint32_t c = 0; c < C; ++c) {
      const uint64_t a = tc.off[(size_t)d * C + c], b = tc.off[(size_t)d * C + c + 1];
      if (a == b) continue;
      if (directional ? !headerDone : first) { trackHeader(d == 0); headerDone = true; first = false; }
      out.write("variableStep chrom=" + contigs.names[c] + "\n");
      for (uint64_t e = a; e < b; ++e) {
        const int k = snprintf(line, sizeof line, (directional && d == 1) ? "%u -%u\n" : "%u %u\n", tc.pos[e], tc.cnt[e]);
        out.write(line, (size_t)k);
      }
    }
  }
  out.close();
  upb_free_tag_counts(&tc);
  upb_destroy(ctx);
  fprintf(stderr, "done!\n\n");
  return 0;
}
