// ref_replay.cpp — hot-path-only replay harness around the UNMODIFIED reference objects.
//
// TEST INFRASTRUCTURE ONLY.  Built by oracle/Makefile into oracle/_ref/ref_replay, linked against
// oracle/_ref/libunipeak_ref.a (the reference's own Kernel / ProfileBuffer / Region, compiled from
// /root/reference/misc where the sources lie).  It feeds those classes from in-memory tag arrays the
// way src/regions.cpp:304-391 and src/strand_shift.cpp:139-259 feed them from wiggle text, so that
//   (a) raw doubles (peak score, kurtosis, correlation) can be compared bit-for-bit, which the
//       reference's own text output (6 / 2 decimals) does not allow, and
//   (b) the reference's hot path can be timed without its text parser (bench.py cpu_baseline,
//       kind "reference").
// The merge loop below is ours (the reference's lives inside main()); every object it drives is the
// reference's.  Nothing from the reference is copied into this repository.
//
// usage: ref_replay regions|shift <input.upbi> <output.bin> key=value ...
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <chrono>
#include <map>

#include "defaults.hpp"
#include "data.hpp"
#include "kernel.hpp"
#include "peakcall.hpp"

using namespace unipeak;

struct Input {
  uint32_t nContigs, nSamples;
  std::vector<uint32_t> sizes;
  std::vector<uint64_t> off;
  std::vector<uint32_t> pos, cnt;
};

static void die (const char* msg) { std::fprintf(stderr, "ref_replay: %s\n", msg); std::exit(2); }

static Input readInput (const char* path) {
  FILE* f = std::fopen(path, "rb");
  if (!f) die("cannot open input");
  char magic[8];
  Input in;
  uint64_t nEntries = 0;
  if (std::fread(magic, 1, 8, f) != 8 || std::memcmp(magic, "UPBI0001", 8) != 0) die("bad magic");
  if (std::fread(&in.nContigs, 4, 1, f) != 1 || std::fread(&in.nSamples, 4, 1, f) != 1) die("short header");
  in.sizes.resize(in.nContigs);
  if (std::fread(in.sizes.data(), 4, in.nContigs, f) != in.nContigs) die("short sizes");
  const size_t nOff = (size_t)in.nSamples * 2 * (in.nContigs + 1);
  in.off.resize(nOff);
  if (std::fread(in.off.data(), 8, nOff, f) != nOff) die("short offsets");
  if (std::fread(&nEntries, 8, 1, f) != 1) die("short count");
  in.pos.resize(nEntries);
  in.cnt.resize(nEntries);
  if (nEntries && std::fread(in.pos.data(), 4, nEntries, f) != nEntries) die("short pos");
  if (nEntries && std::fread(in.cnt.data(), 4, nEntries, f) != nEntries) die("short cnt");
  std::fclose(f);
  return in;
}

// Array-backed stand-in for the reference's tag streams: delivers exactly the (forward, contig,
// firstPos, count) sequence ParseAlignStream (strand-major) or NondirParseAlignStream (merged,
// forward first on ties) would deliver for a canonical wiggle file holding these arrays.
struct Stream {
  const Input* in; uint32_t sample; bool nondir;
  uint32_t d, c; uint64_t i;
  uint32_t cf, cr; uint64_t jf, jr;
  bool forward; uint32_t contig, pos, count;
  const uint64_t* offs (uint32_t strand) const { return in->off.data() + ((size_t)sample * 2 + strand) * (in->nContigs + 1); }
  void open (const Input* input, uint32_t s, bool nd) {
    in = input; sample = s; nondir = nd; d = 0; c = 0; cf = 0; cr = 0;
    i = offs(0)[0]; jf = offs(0)[0]; jr = offs(1)[0];
    forward = true; contig = 0; pos = 0; count = 0;
    next();
  }
  void next () {
    const uint32_t nc = in->nContigs;
    if (!nondir) {
      while (d < 2) {
        const uint64_t* o = offs(d);
        if (c == nc) { ++d; c = 0; if (d < 2) i = offs(d)[0]; continue; }
        if (i < o[c + 1]) { forward = (d == 0); contig = c; pos = in->pos[i]; count = in->cnt[i]; ++i; return; }
        ++c;
      }
      count = 0; contig = nc; return;
    }
    const uint64_t* of = offs(0); const uint64_t* orv = offs(1);
    while (cf < nc && jf >= of[cf + 1]) ++cf;
    while (cr < nc && jr >= orv[cr + 1]) ++cr;
    const bool hf = cf < nc, hr = cr < nc;
    bool takeF;
    if (hf && hr) takeF = (cf == cr) ? (in->pos[jf] <= in->pos[jr]) : (cf < cr);
    else if (hf) takeF = true; else if (hr) takeF = false;
    else { count = 0; contig = nc; return; }
    if (takeF) { forward = true; contig = cf; pos = in->pos[jf]; count = in->cnt[jf]; ++jf; }
    else { forward = false; contig = cr; pos = in->pos[jr]; count = in->cnt[jr]; ++jr; }
  }
};

static std::map<std::string, std::string> parseArgs (int argc, char** argv, int first) {
  std::map<std::string, std::string> m;
  for (int a = first; a < argc; ++a) {
    const char* eq = std::strchr(argv[a], '=');
    if (!eq) die("arguments must be key=value");
    m[std::string(argv[a], eq - argv[a])] = std::string(eq + 1);
  }
  return m;
}
static double argD (const std::map<std::string, std::string>& m, const char* k, double dflt) {
  std::map<std::string, std::string>::const_iterator it = m.find(k);
  return it == m.end() ? dflt : std::strtod(it->second.c_str(), 0);   // accepts hex floats, so values round-trip exactly
}
static std::vector<double> argList (const std::map<std::string, std::string>& m, const char* k) {
  std::vector<double> v;
  std::map<std::string, std::string>::const_iterator it = m.find(k);
  if (it == m.end() || it->second.empty()) return v;
  const char* p = it->second.c_str();
  while (*p) { char* e; v.push_back(std::strtod(p, &e)); p = (*e == ',') ? e + 1 : e; }
  return v;
}

template <class T> static void put (FILE* f, const T& v) { std::fwrite(&v, sizeof v, 1, f); }

static void putRegion (FILE* f, const Region& r) {
  put<uint8_t>(f, r.forward ? 1 : 0);
  put<uint32_t>(f, r.contig); put<uint32_t>(f, r.left); put<uint32_t>(f, (uint32_t)r.counts->size());
  put<uint32_t>(f, r.peakPos); put<double>(f, r.peakScore);
  put<double>(f, r.posKurtosis()); put<double>(f, r.strandCorr());
  const HitCountVec sums = r.exptSums();
  for (size_t e = 0; e < sums.size(); ++e) put<uint32_t>(f, sums[e]);
}

static bool compareRegionCounts (const Region* const i, const Region* const j) { return i->sum() > j->sum(); }   // strand_shift.cpp:26-29

int main (int argc, char** argv) {
  if (argc < 4) die("usage: ref_replay regions|shift <input.upbi> <output.bin> key=value ...");
  const std::string mode = argv[1];
  const Input in = readInput(argv[2]);
  const std::map<std::string, std::string> args = parseArgs(argc, argv, 4);
  const UShort bandwidth = (UShort)argD(args, "bw", DEFAULT_BANDWIDTH);
  const double regionThr = argD(args, "r", DEFAULT_REGION_THRESHOLD);
  const double kurtThr = argD(args, "k", DEFAULT_KURTOSIS_THRESHOLD);
  const double corrThr = argD(args, "u", DEFAULT_CORR_THRESHOLD);
  const double hitThr = argD(args, "t", DEFAULT_HIT_THRESHOLD);
  const double background = argD(args, "background", 0);
  const bool directional = argD(args, "directional", 1) != 0;
  const int reps = (int)argD(args, "reps", 1);
  if (background <= 0) die("background= is required");
  const uint32_t n = in.nSamples;
  std::vector<bool> control(n, false);
  { const std::vector<double> c = argList(args, "control"); for (size_t k = 0; k < c.size(); ++k) control[(size_t)c[k]] = true; }
  const std::vector<double> coeffs = argList(args, "coeffs");

  ContigTable contigs;
  for (uint32_t c = 0; c < in.nContigs; ++c) { char name[32]; std::snprintf(name, sizeof name, "c%u", c); contigs.add(name, in.sizes[c]); }
  const Kernel kernel(bandwidth, 1 / background);

  FILE* out = std::fopen(argv[3], "wb");
  if (!out) die("cannot open output");
  double bestSeconds = 1e300;

  if (mode == "regions") {
    for (int rep = 0; rep < reps; ++rep) {
      const bool last = rep == reps - 1;
      std::vector<Region*> regions;
      std::vector<Region*> written;
      const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
      ProfileBuffer forwardBuffer(&kernel, regionThr, kurtThr, directional ? -1 : corrThr, hitThr, true, control, coeffs, &contigs, &regions);
      ProfileBuffer reverseBuffer(&kernel, regionThr, kurtThr, directional ? -1 : corrThr, hitThr, false, control, coeffs, &contigs, &regions);
      std::vector<Stream> st(n);
      for (uint32_t s = 0; s < n; ++s) st[s].open(&in, s, !directional);
      ContigNo contig = 0;
      bool forwardPass = true;
      while (contig < in.nContigs) {
        const Pos posLimit = in.sizes[contig];
        Pos pos = 1;
        while (pos <= posLimit) {
          bool ff = false, fr = false;
          Pos nextPos = posLimit + 1;
          HitCountVec fh(n), rh(n);
          for (uint32_t s = 0; s < n; ++s) {
            fh[s] = 0; rh[s] = 0;
            while (st[s].count != 0 && st[s].pos == pos && st[s].contig == contig) {
              if (st[s].forward) { fh[s] += st[s].count; ff = true; } else { rh[s] += st[s].count; fr = true; }
              st[s].next();
            }
            if (st[s].count != 0 && st[s].contig == contig && st[s].pos < nextPos) nextPos = st[s].pos;
          }
          if (ff) forwardBuffer.add(fh, contig, pos, true);
          if (fr) { if (directional) reverseBuffer.add(rh, contig, pos, false); else forwardBuffer.add(rh, contig, pos, false); }
          if (!regions.empty()) {
            for (size_t k = 0; k < regions.size(); ++k) { if (last) written.push_back(regions[k]); else delete regions[k]; }
            regions.clear();
          }
          pos = nextPos;
        }
        forwardBuffer.flushContig();
        reverseBuffer.flushContig();
        ++contig;
        if (directional && contig == in.nContigs && forwardPass) { contig = 0; forwardPass = false; }
      }
      const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
      if (secs < bestSeconds) bestSeconds = secs;
      if (last) {
        put<uint64_t>(out, written.size());
        put<uint64_t>(out, regions.size());                 // dropped by the reference (never written)
        put<uint64_t>(out, forwardBuffer.nRegions() + reverseBuffer.nRegions());
        put<uint64_t>(out, forwardBuffer.nRegionRejects() + reverseBuffer.nRegionRejects());
        for (uint32_t s = 0; s < n; ++s) put<uint64_t>(out, forwardBuffer.nTagsInRegions()[s] + reverseBuffer.nTagsInRegions()[s]);
        for (size_t k = 0; k < written.size(); ++k) { putRegion(out, *written[k]); delete written[k]; }
        for (size_t k = 0; k < regions.size(); ++k) { putRegion(out, *regions[k]); delete regions[k]; }
      } else {
        for (size_t k = 0; k < regions.size(); ++k) delete regions[k];
      }
    }
  } else if (mode == "shift") {
    const UShort minShift = (UShort)argD(args, "min_shift", DEFAULT_MIN_SHIFT);
    const UShort maxShift = (UShort)argD(args, "max_shift", DEFAULT_MAX_SHIFT);
    const UShort nRegionsToTest = (UShort)argD(args, "g", DEFAULT_TEST_REGIONS);
    const HitCount hitThreshold = (HitCount)hitThr;
    const std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
    std::vector<Region*> regions;
    ProfileBuffer buffer(&kernel, regionThr, kurtThr, -1, hitThreshold, true, std::vector<bool>(n, false), std::vector<double>(), &contigs, &regions);
    std::vector<Stream> st(n);
    for (uint32_t s = 0; s < n; ++s) st[s].open(&in, s, true);
    for (ContigNo contig = 0; contig < in.nContigs; ++contig) {
      const Pos posLimit = in.sizes[contig];
      Pos pos = 1;
      while (pos <= posLimit) {
        bool ff = false, fr = false;
        Pos nextPos = posLimit + 1;
        HitCountVec fh(n), rh(n);
        for (uint32_t s = 0; s < n; ++s) {
          fh[s] = 0; rh[s] = 0;
          while (st[s].count != 0 && st[s].pos == pos && st[s].contig == contig) {
            if (st[s].forward) { fh[s] += st[s].count; ff = true; } else { rh[s] += st[s].count; fr = true; }
            st[s].next();
          }
          if (st[s].count != 0 && st[s].contig == contig && st[s].pos < nextPos) nextPos = st[s].pos;
        }
        if (ff) buffer.add(fh, contig, pos, true);
        if (fr) buffer.add(rh, contig, pos, false);
        pos = nextPos;
      }
      buffer.flushContig();
    }
    const uint64_t nFound = regions.size();
    std::sort(regions.begin(), regions.end(), compareRegionCounts);
    UShort nTested = 0;
    Counter nTagsInRegions = 0;
    std::vector<Counter> shiftFreq(maxShift + 1, 0);
    std::vector<uint32_t> tContig, tLeft, tLen, tBest; std::vector<double> tCorr;
    for (std::vector<Region*>::const_iterator i = regions.begin(); nTested < nRegionsToTest && i != regions.end(); ++i) {
      if ((*i)->counts->size() > (unsigned)(2 * maxShift + 3)) {
        UShort bestShift = 0; double bestCorr = -1;
        for (UShort shift = 0; shift <= maxShift; ++shift) {
          const double c = (*i)->strandCorr(shift);
          if (c > bestCorr) { bestShift = shift; bestCorr = c; }
        }
        if (bestCorr >= corrThr) {
          ++shiftFreq[bestShift]; nTagsInRegions += (*i)->sum(); ++nTested;
          tContig.push_back((*i)->contig); tLeft.push_back((*i)->left); tLen.push_back((uint32_t)(*i)->counts->size());
          tBest.push_back(bestShift); tCorr.push_back(bestCorr);
        }
      }
    }
    uint32_t best = 0;
    if (nTested != 0) {
      const Kernel modeKernel(MAXIMA_SMOOTHENING_BANDWIDTH);
      std::vector<double> modeDensity(shiftFreq.size(), 0);
      for (UShort i = 0; i < shiftFreq.size(); ++i)
        for (int j = 0; j < (int)modeKernel.size(); ++j) {
          const int k = (int)i - (int)MAXIMA_SMOOTHENING_BANDWIDTH + j;
          if (k >= 0 && k < (int)modeDensity.size()) modeDensity[k] += (double)shiftFreq[i] * modeKernel[j];
        }
      double bestDensity = 0;
      for (UShort i = minShift; i < modeDensity.size() - MAXIMA_SMOOTHENING_BANDWIDTH; ++i)
        if (modeDensity[i] > bestDensity) { best = i; bestDensity = modeDensity[i]; }
    }
    bestSeconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    put<uint64_t>(out, nFound); put<uint32_t>(out, nTested); put<uint64_t>(out, nTagsInRegions); put<uint32_t>(out, best);
    put<uint32_t>(out, (uint32_t)shiftFreq.size());
    for (size_t k = 0; k < shiftFreq.size(); ++k) put<uint64_t>(out, shiftFreq[k]);
    for (size_t k = 0; k < tContig.size(); ++k) { put<uint32_t>(out, tContig[k]); put<uint32_t>(out, tLeft[k]); put<uint32_t>(out, tLen[k]); put<uint32_t>(out, tBest[k]); put<double>(out, tCorr[k]); }
    for (size_t k = 0; k < regions.size(); ++k) delete regions[k];
  } else die("unknown mode");

  std::fclose(out);
  std::printf("seconds=%.6f\n", bestSeconds);
  return 0;
}
