// bin/regions — drop-in for the reference's region caller (src/regions.cpp).  Same flags, stderr messages, regions-table
// and profile-wiggle text; the work between "calling enriched regions..." and the row writes goes through the C ABI
// (include/unipeak_b200.h) to the sm_100a kernels.  There is no CPU path: without a CUDA device this program exits 1.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "upb_host.hpp"

using namespace upbhost;

namespace {

// FormatOutStream::write(PosScore) + trackHeader for the profile (misc/format.cpp:1091-1139, :1164-1219)
struct ProfileWriter {
  Out out;
  const ContigTable& contigs;
  std::string exptName, assembly, contig;
  bool directional, forward = true;
  ProfileWriter (const std::string& fname, const ContigTable& c, const std::string& name, bool dir, const std::string& assm)
      : out(fname), contigs(c), exptName(name), assembly(assm), directional(dir) {}
  void trackHeader () {
    std::string s = "track name=\"" + exptName;
    if (directional) s += forward ? " +" : " -";
    s += "\"";
    if (directional) s += std::string(" description=\"") + (forward ? exptName : std::string(" ")) + "\"";
    s += " priority=2 visibility=";                                  // PROFILE_PRIORITY, misc/defaults.hpp:42
    if (directional) s += forward ? "full type=wiggle_0 alwaysZero=on color=0,0,255" : "full type=wiggle_0 alwaysZero=on color=255,0,0 altColor=255,0,0";
    else s += "full type=wiggle_0 alwaysZero=on color=191,0,191";
    if (!assembly.empty()) s += " db=" + assembly;
    s += "\n";
    out.write(s);
  }
  void points (bool fwd, uint32_t c, const uint32_t* pos, const double* score, uint64_t n) {
    if (n == 0) return;
    if (directional) {
      if (contig.empty() || forward != fwd) { forward = fwd; contig.clear(); trackHeader(); }
    } else if (contig.empty()) {
      trackHeader();
    }
    if (contig.empty() || contig != contigs.names[c]) { contig = contigs.names[c]; out.write("variableStep chrom=" + contig + "\n"); }
    char line[96];
    for (uint64_t i = 0; i < n; ++i) {
      int k = snprintf(line, sizeof line, "%u ", pos[i]);
      if (directional && !fwd) line[k++] = '-';
      k += fast_g6(score[i], line + k);
      line[k++] = '\n';
      out.write(line, (size_t)k);
    }
  }
};

}  // namespace

int main (int argc, char* argv[]) {
  // the TCLAP declarations of src/regions.cpp:53-75, in the same order
  const std::vector<ArgSpec> specs = {
    {'q', "quiet", ARG_SWITCH, false, "don't give contig-by-contig updates", ""},
    {'D', "non-directional", ARG_SWITCH, false, "don't separate strands (data aren't from a directional protocol), default false", ""},
    {'a', "assembly", ARG_STRING, false, "reference assembly name, e.g. hg18", "name"},
    {'n', "name", ARG_STRING, false, "wiggle plot track name", "name"},
    {'w', "wig", ARG_STRING, false, "output wiggle plot filename", "filename"},
    {'m', "mappable", ARG_U32, false, "number of mappable positions", "positive integer"},
    {'t', "hitThreshold", ARG_DOUBLE, false, "minimum average hits per experiment in a region, default 10", "non-negative number"},
    {'y', "corr", ARG_SWITCH, false, "output strand correlations in addition to kurtosis boundaries (non-directional only), default false", ""},
    {'u', "corrThreshold", ARG_DOUBLE, false, "minimum correlation between strands' densities (non-directional only), default 0.3", "number"},
    {'k', "kurtosisThreshold", ARG_DOUBLE, false, "maximum kurtosis threshold for regions, default 50", "non-negative number"},
    {'r', "regionThreshold", ARG_DOUBLE, false, "minimum fold enrichment threshold for regions, default 25", "positive number"},
    {'z', "coeff", ARG_STRING, false, "coefficients to scale samples, or 'p' to scale inversely to confident reads", "comma-separated numbers"},
    {'b', "bandwidth", ARG_U16, false, "kernel bandwidth, default 50", "positive integer"},
    {'i', "mismatches", ARG_U16, false, "mismatch tolerance, default 2 (Eland extended only)", "non-negative integer"},
    {'l', "length", ARG_U16, false, "sequence length to use, default entire sequence (Eland extended only)", "positive integer"},
    {'s', "shift", ARG_STRING, false, "number of bases downstream to shift alignments (barcode length and/or strand shift), default 0", "comma-separated integers"},
    {'p', "prob", ARG_DOUBLE, false, "posterior probability threshold for alignments", "probability"},
    {'e', "exclude", ARG_STRING, false, "samples to exclude from peak-calling (e.g. negative controls)", "comma-separated positive integers"},
    {'f', "peaks", ARG_SWITCH, false, "output peak positions in addition to region boundaries, default false", ""},
    {'o', "out", ARG_STRING, true, "output regions filename", "filename"},
    {'c', "contig", ARG_STRING, true, "contig table filename", "filename"},
  };
  CmdLine cmd(specs, ".wig files of alignments", "alignment filenames");
  cmd.parse(argc, argv);
  const std::string regionsFname = cmd.str('o'), contigFname = cmd.str('c');
  const bool outputPeaks = cmd.sw('f');
  const std::string controlString = cmd.str('e'), offsetString = cmd.str('s'), assembly = cmd.str('a');
  const uint16_t useLength = cmd.u16('l', 0);
  const uint16_t bandwidth = cmd.u16('b', 50);
  std::string trackName = cmd.str('n');
  const std::string profileFname = cmd.str('w');
  uint32_t mappable = cmd.u32('m', 0);
  const std::string coeffString = cmd.str('z');
  const double regionThreshold = cmd.dbl('r', 25), kurtosisThreshold = cmd.dbl('k', 50);
  bool outputCorrs = cmd.sw('y');
  double corrThreshold = cmd.dbl('u', 0.3);
  double hitThreshold = cmd.dbl('t', 10);
  const std::vector<std::string>& alignFnames = cmd.unlabeled();
  const bool directional = !cmd.sw('D');
  const bool quiet = cmd.sw('q');

  if (!profileFname.empty() && trackName.empty()) trackName = fname_prefix(profileFname);
  if (directional) {                                                // src/regions.cpp:104-109
    if (corrThreshold != 0.3) fprintf(stderr, "warning: correlation threshold is not used on strand-specific analysis\n");
    corrThreshold = -1;
    if (outputCorrs) fprintf(stderr, "warning: strand correlations are not calculated for strand-specific analysis\n");
    outputCorrs = false;
  }

  const std::vector<short> offsets = parse_offsets(offsetString, alignFnames.size());

  // controls (src/regions.cpp:129-150)
  std::vector<uint8_t> control(alignFnames.size(), 0);
  if (!controlString.empty()) {
    for (const std::string& tok : split_commas(controlString)) {
      short v;
      if (!lexical_short(tok, &v)) { fprintf(stderr, "error: bad control index\n\n"); exit(1); }
      const uint16_t index = (uint16_t)v;
      if (index == 0) { fprintf(stderr, "error: bad control index (first sample is 1)\n\n"); exit(1); }
      if (index > control.size()) { fprintf(stderr, "error: bad control index (greater than number of samples)\n\n"); exit(1); }
      control[index - 1] = 1;
    }
  }
  uint16_t nControl = 0;
  for (uint8_t c : control) if (c) ++nControl;
  hitThreshold *= (double)(alignFnames.size() - nControl);          // :155 (size_t converted to double)

  // coefficients (:157-180)
  std::vector<double> coeffs;
  bool proportional = false;
  if (!coeffString.empty()) {
    if (coeffString == "p") proportional = true;
    else {
      for (const std::string& tok : split_commas(coeffString)) {
        double v;
        if (!lexical_double(tok, &v)) { fprintf(stderr, "error: bad coeff argument\n\n"); exit(1); }
        coeffs.push_back(v);
      }
      if (coeffs.size() != alignFnames.size() - nControl) {
        fprintf(stderr, "error: wrong number of coeff arguments\nmust have same number as non-control alignment files\n\n");
        exit(1);
      }
    }
  }

  Device dev;                                                        // one GPU, or UPB_GPUS=N of them; its CUDA context is created in the background while the inputs are parsed
  const ContigTable contigs = parse_contig_table(contigFname);
  const uint32_t C = contigs.size();

  // input (:184-202); the files are parsed whole (one thread per file) instead of streamed
  fprintf(stderr, "reading alignment files...\n");
  std::vector<Sample> samples = read_samples(alignFnames, contigs, offsets, useLength);
  uint64_t nNonControlTags = 0, nTotalTags = 0;
  for (size_t i = 0; i < samples.size(); ++i) {
    if (!samples[i].error.empty()) die(samples[i].error);
    if (!control[i]) nNonControlTags += samples[i].expectedTags;
    nTotalTags += samples[i].expectedTags;
    fprintf(stderr, "  %s: %llu tags\n", samples[i].name.c_str(), (unsigned long long)samples[i].expectedTags);
  }
  const size_t n = samples.size();

  // background (:204-213)
  if (mappable == 0) mappable = contigs.genome_size();
  fprintf(stderr, "%llu usable tags", (unsigned long long)nTotalTags);
  if (nNonControlTags != nTotalTags) fprintf(stderr, ", %llu not from negative controls,", (unsigned long long)nNonControlTags);
  fprintf(stderr, " at %u mappable positions\n", mappable);
  double background = (double)nNonControlTags / (double)mappable;
  if (directional) background /= 2;
  fprintf(stderr, "using background = %s tags/position", fmt_g(background).c_str());
  if (directional) fprintf(stderr, " on each strand");
  fprintf(stderr, "\n");

  // scaling coefficients (:215-229)
  if (proportional) {
    for (size_t i = 0; i < n; ++i)
      if (!control[i]) coeffs.push_back((double)nNonControlTags / ((double)samples[i].expectedTags * (double)(n - nControl)));
  } else if (!coeffs.empty()) {
    double scaledSum = 0;
    // the reference's loop advances its coefficient iterator on every sample, control or not (:227)
    for (size_t i = 0; i < coeffs.size() && i < n; ++i) if (!control[i]) scaledSum += coeffs[i] * (double)samples[i].expectedTags;
    for (double& c : coeffs) c *= (double)nNonControlTags / scaledSum;
  }

  // header (:234-274)
  std::string commonHeader;
  for (const std::string& f : alignFnames) commonHeader += "# align_file=" + f + "\n";
  if (!offsets.empty()) {
    if (offsets.size() == 1) commonHeader += "# shift=" + std::to_string(offsets[0]) + "\n";
    else {
      commonHeader += "# shifts=";
      for (size_t i = 0; i + 1 < offsets.size(); ++i) commonHeader += std::to_string(offsets[i]) + ",";
      commonHeader += std::to_string(offsets.back()) + "\n";
    }
  }
  commonHeader += "# contig_table=" + contigFname + "\n";
  commonHeader += "# bandwidth=" + std::to_string(bandwidth) + "\n";
  if (n > 1) {
    commonHeader += "# tags=" + std::to_string(samples[0].expectedTags);
    for (size_t i = 1; i < n; ++i) commonHeader += "," + std::to_string(samples[i].expectedTags);
    commonHeader += "\n";
  }
  if (nControl > 0) {
    commonHeader += "# control=";
    bool first = true;
    for (size_t i = 0; i < n; ++i) if (control[i]) { if (!first) commonHeader += ","; first = false; commonHeader += std::to_string(i + 1); }
    commonHeader += "\n";
  }
  if (!coeffs.empty()) {
    commonHeader += "# coeffs=" + fmt_g(coeffs[0]);
    for (size_t i = 1; i < coeffs.size(); ++i) commonHeader += "," + fmt_g(coeffs[i]);
    commonHeader += "\n";
  }
  commonHeader += "# total_tags=" + std::to_string(nNonControlTags) + "\n";
  commonHeader += "# background=" + fmt_g(background) + "\n";

  ProfileWriter* profile = nullptr;
  if (!profileFname.empty()) {
    profile = new ProfileWriter(profileFname, contigs, trackName, directional, assembly);
    profile->out.write(commonHeader);
  }
  Out regionsOut(regionsFname);
  regionsOut.write(commonHeader);
  regionsOut.write("# region_threshold=" + fmt_lexical(regionThreshold) + "\n# kurtosis_threshold=" + fmt_lexical(kurtosisThreshold) +
                   "\n# corr_threshold=" + fmt_lexical(corrThreshold) + "\n# hit_threshold=" + fmt_lexical(hitThreshold) + "\n");
  {
    std::string cols;
    if (outputPeaks) cols += "\tpeak";
    if (outputCorrs) cols += "\tcorrelation";
    cols += "\tkurtosis";
    for (const Sample& s : samples) cols += "\t" + s.name;
    cols += "\n";
    regionsOut.write(cols);
  }

  // the hot path (:304-391), on the device
  fprintf(stderr, "calling enriched regions...\n");
  dev.setContigs(C, contigs.sizes.data());
  // which strand/contig units does the merge loop read in pass 1?  (single sample whose '-' track starts at or after its
  // last '+' contig: src/regions.cpp:341 consumes it there) — needed for the progress lines and the profile order
  bool early = false;
  if (directional) {
    for (const Sample& s : samples) {
      int lastPlus = -1, firstMinus = -1;
      for (uint32_t c = 0; c < C; ++c) if (!s.pos[0][c].empty()) lastPlus = (int)c;
      for (uint32_t c = 0; c < C; ++c) if (!s.pos[1][c].empty()) { firstMinus = (int)c; break; }
      if (firstMinus >= 0 && firstMinus >= lastPlus) early = true;
    }
  }
  // No usable tag at all (empty tracks, or tags in control samples only): the reference runs through — its Kernel(bandwidth, 1 / 0) is all
  // inf / NaN, no score compares >= the threshold — and writes headers, zero counts and "-nan%".  Nothing to ask a device for.
  const bool noTags = !(background > 0);
  std::vector<uint64_t> zero64(std::max<size_t>(2 * (size_t)C, n) + 1, 0);
  std::vector<uint32_t> zero32(2 * (size_t)C + 1, 0);
  TagLoader loader;                                                  // compact host form, queued asynchronously: the call below starts on what has arrived
  if (!noTags) dev.load(loader, samples, C, directional, false);
  upb_region_params prm;
  memset(&prm, 0, sizeof prm);
  prm.bandwidth = bandwidth;
  prm.region_threshold = regionThreshold;
  prm.kurtosis_threshold = kurtosisThreshold;
  prm.corr_threshold = corrThreshold;
  prm.hit_threshold = hitThreshold;
  prm.directional = directional ? 1 : 0;
  prm.control = nControl ? control.data() : nullptr;
  prm.n_coeffs = (uint32_t)coeffs.size();
  prm.coeffs = coeffs.empty() ? nullptr : coeffs.data();
  prm.background = background;
  prm.reference_order = 1;
  upb_region_table tab;
  memset(&tab, 0, sizeof tab);
  if (noTags) { tab.tags_in_regions = zero64.data(); tab.unit_rows = zero64.data(); tab.unit_last_tag = zero32.data(); }
  else dev.callRegions(&prm, &tab);

  // rows (FormatOutStream::write(Region), misc/format.cpp:1141-1162); the rows the reference's last flush loses stay lost
  {
    std::string row;
    row.reserve(256 + 12 * n);
    char num[64];
    for (uint64_t i = 0; i < tab.n_rows; ++i) {
      if (tab.dropped[i]) continue;
      row.clear();
      row += contigs.names[tab.contig[i]];
      row += ':';
      const unsigned long long left = tab.left[i], right = (unsigned long long)tab.left[i] + tab.len[i] - 1;
      if (tab.forward_label[i]) snprintf(num, sizeof num, "%llu-%llu", left, right); else snprintf(num, sizeof num, "%llu-%llu", right, left);
      row += num;
      if (outputPeaks) { snprintf(num, sizeof num, "\t%u", tab.peak[i]); row += num; }
      if (outputCorrs) { row += '\t'; row += fmt_g(tab.corr[i]); }
      row += '\t';
      row += fmt_fixed(tab.kurtosis[i], 2);
      const uint32_t* cnt = tab.counts + i * n;
      for (size_t e = 0; e < n; ++e) { snprintf(num, sizeof num, "\t%u", cnt[e]); row += num; }
      row += '\n';
      regionsOut.write(row);
    }
  }

  // contig-by-contig progress (:312-384): the counts flushContig() returns per iteration
  if (!quiet) {
    for (int pass = 0; pass < (directional ? 2 : 1); ++pass)
      for (uint32_t c = 0; c < C; ++c) {
        const unsigned long long fwd = pass == 0 ? (unsigned long long)tab.unit_rows[c] : 0ull;
        const unsigned long long rev = !directional ? 0ull : (early ? (pass == 0 ? tab.unit_rows[C + c] : 0ull) : (pass == 1 ? tab.unit_rows[C + c] : 0ull));
        fprintf(stderr, "  %s%s... ", contigs.names[c].c_str(), directional ? (pass == 0 ? "+" : "-") : "");
        if (pass == 0) {
          fprintf(stderr, "%llu\n", fwd);
          if (directional && rev > 0) fprintf(stderr, "  %s-... %llu\n", contigs.names[c].c_str(), rev);
        } else {
          fprintf(stderr, "%llu\n", rev);
        }
      }
  }

  // profile (-w): positions with score != 0 in the order ProfileBuffer::processPosition writes them
  if (profile && noTags) { profile->out.close(); delete profile; profile = nullptr; }
  if (profile) {
    auto unit = [&] (int buffer, uint32_t c, int part /* 0 all, 1 before the flush, 2 the flush */) {
      upb_profile pf;
      dev.profileContig(&prm, buffer, c, &pf);
      uint64_t lo = 0, hi = pf.n;
      if (part) {
        // add() has processed every position < lastTag - bw when flushContig() takes over (misc/peakcall.cpp:172-184)
        const uint32_t lastTag = tab.unit_last_tag[(size_t)buffer * C + c];
        const uint64_t cut = lastTag > bandwidth ? (uint64_t)lastTag - bandwidth : 0;
        uint64_t k = 0;
        while (k < pf.n && pf.pos[k] < cut) ++k;
        if (part == 1) hi = k; else lo = k;
      }
      profile->points(buffer == 0, c, pf.pos + lo, pf.score + lo, hi - lo);
      upb_free_profile(&pf);
    };
    if (!directional) { for (uint32_t c = 0; c < C; ++c) unit(0, c, 0); }
    else if (!early) { for (int b = 0; b < 2; ++b) for (uint32_t c = 0; c < C; ++c) unit(b, c, 0); }
    else for (uint32_t c = 0; c < C; ++c) { unit(0, c, 1); unit(1, c, 1); unit(0, c, 2); unit(1, c, 2); }
    profile->out.close();
    delete profile;
  }
  regionsOut.close();

  // summary (:395-407)
  fprintf(stderr, "%llu regions passed filters, %llu rejected\n", (unsigned long long)tab.n_passed, (unsigned long long)tab.n_rejected);
  fprintf(stderr, "tags in regions:\n");
  for (size_t i = 0; i < n; ++i) {
    const Sample& s = samples[i];
    if (s.confidentTags + s.outOfBoundsTags != s.expectedTags)
      fprintf(stderr, "  warning: expected %llu tags in %s but found %llu; results inaccurate\n", (unsigned long long)s.expectedTags, s.name.c_str(), (unsigned long long)s.confidentTags);
    const unsigned long long t = tab.tags_in_regions[i];
    fprintf(stderr, "  %s: %llu (%s%%)\n", s.name.c_str(), t, fmt_fixed(100 * (double)t / (double)s.confidentTags, 1).c_str());
  }
  if (!noTags && dev.gpus() == 1) upb_free_region_table(&tab);       // (several GPUs: the table is the handle's)
  fprintf(stderr, "\nAll done!\n\n");
  return 0;
}
