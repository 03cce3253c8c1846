// bin/strand_shift — drop-in for the reference's strand-shift calibration (src/strand_shift.cpp).  Same flags, stderr
// messages and shift report; region calling, the per-region shift sweep and the correlations run on the device through
// upb_strand_shift (include/unipeak_b200.h).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "upb_host.hpp"

using namespace upbhost;

int main (int argc, char* argv[]) {
  // the TCLAP declarations of src/strand_shift.cpp:52-68, in the same order
  const std::vector<ArgSpec> specs = {
    {'m', "mappable", ARG_U32, false, "number of mappable positions, default sum of contigs", "positive integer"},
    {'t', "hitThreshold", ARG_U32, false, "minimum average hits per experiment in a region, default 10", "non-negative integer"},
    {'u', "corrThreshold", ARG_DOUBLE, false, "minimum correlation between strands' densities (non-directional only), default 0.3", "number"},
    {'k', "kurtosisThreshold", ARG_DOUBLE, false, "maximum kurtosis threshold for regions, default 50", "positive number"},
    {'r', "regionThreshold", ARG_DOUBLE, false, "minimum fold enrichment threshold for regions, default 25", "positive number"},
    {'b', "bandwidth", ARG_U16, false, "kernel bandwidth, default 50", "positive integer"},
    {'g', "testRegions", ARG_U16, false, "number of regions to use for calibration, default 1000", "non-negative integer"},
    {'x', "maxShift", ARG_U16, false, "maximum shift to try, default 150", "positive integer"},
    {'n', "minShift", ARG_U16, false, "minimum shift to try, default 25", "positive integer"},
    {'i', "mismatches", ARG_U16, false, "mismatch tolerance, default 2", "non-negative integer"},
    {'l', "length", ARG_U16, false, "sequence length to use, default entire sequence", "positive integer"},
    {'s', "shift", ARG_STRING, false, "number of bases downstream to shift alignments (barcode length and/or strand shift), default 0", "comma-separated integers"},
    {'p', "prob", ARG_DOUBLE, false, "posterior probability threshold for alignments, default 0.9", "probability"},
    {'o', "out", ARG_STRING, false, "output filename", "filename"},
    {'c', "contigs", ARG_STRING, true, "contig table filename", "filename"},
  };
  CmdLine cmd(specs, ".wig files of alignments", "alignment filenames");
  cmd.parse(argc, argv);
  const std::string contigFname = cmd.str('c'), outFname = cmd.str('o'), offsetString = cmd.str('s');
  const uint16_t useLength = cmd.u16('l', 0), bandwidth = cmd.u16('b', 50);
  uint32_t mappable = cmd.u32('m', 0);
  const double regionThreshold = cmd.dbl('r', 25), kurtosisThreshold = cmd.dbl('k', 50), corrThreshold = cmd.dbl('u', 0.3);
  const uint32_t hitThreshold = cmd.u32('t', 10);
  const uint16_t nRegionsToTest = cmd.u16('g', 1000), minShift = cmd.u16('n', 25), maxShift = cmd.u16('x', 150);
  const std::vector<std::string>& alignFnames = cmd.unlabeled();

  const std::vector<short> offsets = parse_offsets(offsetString, alignFnames.size());
  Device dev;                                                        // one GPU, or UPB_GPUS=N of them; its CUDA context is created in the background while the inputs are parsed
  const ContigTable contigs = parse_contig_table(contigFname);
  const uint32_t C = contigs.size();

  // input (:111-128): every file is read through NondirParseAlignStream, i.e. both strands feed one buffer
  fprintf(stderr, "reading alignment files...\n");
  std::vector<Sample> samples = read_samples(alignFnames, contigs, offsets, useLength);
  uint64_t nTotalTags = 0;
  for (const Sample& s : samples) {
    if (!s.error.empty()) die(s.error);
    nTotalTags += s.expectedTags;
    fprintf(stderr, "  %s: %llu tags\n", s.name.c_str(), (unsigned long long)s.expectedTags);
  }

  // background (:130-134)
  if (mappable == 0) mappable = contigs.genome_size();
  fprintf(stderr, "%llu usable tags at %u mappable positions\n", (unsigned long long)nTotalTags, mappable);
  const double background = (double)nTotalTags / (double)mappable;
  fprintf(stderr, "using background = %s tags/position\n", fmt_g(background).c_str());

  fprintf(stderr, "calling enriched regions for shift calibration... ");
  fflush(stderr);
  if (!(background > 0)) {                                             // no tag at all: the reference's loop finds no region (:143-233) and says so
    fprintf(stderr, "0 found\n");
    fprintf(stderr, "error: no regions qualified with given settings\n\n");
    exit(1);
  }
  dev.setContigs(C, contigs.sizes.data());
  TagLoader loader;
  dev.load(loader, samples, C, false, true);
  upb_shift_params prm;
  memset(&prm, 0, sizeof prm);
  prm.bandwidth = bandwidth;
  prm.region_threshold = regionThreshold;
  prm.kurtosis_threshold = kurtosisThreshold;
  prm.corr_threshold = corrThreshold;
  prm.hit_threshold = hitThreshold;
  prm.min_shift = minShift;
  prm.max_shift = maxShift;
  prm.n_regions_to_test = nRegionsToTest;
  prm.background = background;
  upb_shift_report rep;
  const int rc = dev.strandShift(&prm, &rep);
  fprintf(stderr, "%llu found\n", (unsigned long long)rep.n_regions_found);
  if (rc == UPB_ERR_NO_REGIONS || rep.n_regions_tested == 0) {      // :231-233
    fprintf(stderr, "error: no regions qualified with given settings\n\n");
    exit(1);
  } else if (rep.n_regions_tested < nRegionsToTest) {
    fprintf(stderr, "warning: too few regions qualified with given settings\n");
  }
  fprintf(stderr, "the top %u qualified regions contained %llu tags (%s%%)\n", rep.n_regions_tested, (unsigned long long)rep.n_tags_in_regions,
          fmt_fixed(100 * (double)rep.n_tags_in_regions / (double)nTotalTags, 1).c_str());
  fprintf(stderr, "estimated shift = %u\n", rep.best_shift);

  // report (:262-300); every number goes through OutStream << double, i.e. 17 significant digits
  if (!outFname.empty()) {
    fprintf(stderr, "writing output to %s... ", outFname.c_str());
    fflush(stderr);
    Out out(outFname);
    std::string s;
    for (const std::string& f : alignFnames) s += "# align_file=" + f + "\n";
    if (!offsets.empty()) {
      if (offsets.size() == 1) s += "# shift=" + fmt_lexical(offsets[0]) + "\n";
      else {
        s += "# shifts=";
        for (size_t i = 0; i + 1 < offsets.size(); ++i) s += fmt_lexical(offsets[i]) + ",";
        s += fmt_lexical(offsets.back()) + "\n";
      }
    }
    s += "# contig_table=" + contigFname + "\n";
    s += "# bandwidth=" + fmt_lexical(bandwidth) + "\n";
    s += "# tags=" + fmt_lexical((double)nTotalTags) + "\n";
    s += "# background=" + fmt_lexical(background) + "\n";
    s += "# region_threshold=" + fmt_lexical(regionThreshold) + "\n";
    s += "# kurtosis_threshold=" + fmt_lexical(kurtosisThreshold) + "\n";
    s += "# corr_threshold=" + fmt_lexical(corrThreshold) + "\n";
    s += "# hit_threshold=" + fmt_lexical(hitThreshold) + "\n";
    s += "# regions_tested=" + fmt_lexical(rep.n_regions_tested) + "\n";
    s += "# tags_in_regions=" + fmt_lexical((double)rep.n_tags_in_regions) + "\n";
    s += "# min_shift=" + fmt_lexical(minShift) + "\n";
    s += "# best_shift=" + fmt_lexical(rep.best_shift) + "\n\n";
    s += "shift\tregions\n";
    for (uint32_t i = 0; i < rep.n_freq; ++i)
      if (rep.shift_freq[i] != 0) s += fmt_lexical(i) + "\t" + fmt_lexical((double)rep.shift_freq[i]) + "\n";
    out.write(s);
    out.close();
  }
  upb_free_shift_report(&rep);
  fprintf(stderr, "done!\n\n");
  return 0;
}
