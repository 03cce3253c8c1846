// bin/host_selftest — CPU-only checks of the host front end (no device call): the %g formatter against printf, and a dump
// of what the wiggle reader hands to upb_load_tags, for tests/test_host_cli.py.
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <string>

#include "upb_host.hpp"
#include "../csrc/upb_introsort.hpp"

#include <algorithm>
#include <vector>

using namespace upbhost;

int main (int argc, char** argv) {
  if (argc >= 2 && !strcmp(argv[1], "g6")) {
    const long n = argc > 2 ? atol(argv[2]) : 1000000;
    std::mt19937_64 rng(12345);
    long bad = 0;
    char a[64], b[64];
    for (long i = 0; i < n; ++i) {
      double x;
      switch (i % 4) {
        case 0: x = std::ldexp((double)(rng() >> 11), (int)(rng() % 100) - 80); break;      // wide range of magnitudes
        case 1: x = (double)(rng() % 2000000) / 1000.0; break;                              // short decimals (ties after scaling)
        case 2: x = (double)(rng() % 100000000) * 1e-7; break;
        default: x = std::exp(((double)(rng() % 2000000) / 1e6 - 1.0) * 30.0); break;
      }
      snprintf(a, sizeof a, "%g", x);
      fast_g6(x, b);
      if (strcmp(a, b)) { if (++bad < 10) fprintf(stderr, "g6 mismatch: %.17g printf=%s fast=%s\n", x, a, b); }
    }
    printf("g6 checked=%ld bad=%ld\n", n, bad);
    return bad ? 1 : 0;
  }
  if (argc >= 2 && !strcmp(argv[1], "sort")) {
    // the prefix sort against std::sort itself: unstable sort, heavy ties, the element order must be identical
    struct SI { uint32_t sum, idx; };
    auto cmp = [] (const SI& a, const SI& b) { return a.sum > b.sum; };
    std::mt19937_64 rng(99);
    long bad = 0, cases = 0;
    for (int rep = 0; rep < 60; ++rep) {
      const size_t n = rep < 6 ? (size_t)rep * 5 : 1000 + rng() % 120000;
      const uint32_t range = rep % 3 == 0 ? 7 : (rep % 3 == 1 ? 500 : 1000000);
      std::vector<SI> a(n);
      for (size_t i = 0; i < n; ++i) { a[i].sum = (uint32_t)(rng() % range); a[i].idx = (uint32_t)i; }
      if (rep % 5 == 4) std::sort(a.begin(), a.end(), [] (const SI& x, const SI& y) { return x.sum < y.sum; });   // adversarial: ascending input
      std::vector<SI> full = a;
      std::sort(full.begin(), full.end(), cmp);
      for (size_t bound : {(size_t)0, (size_t)1, (size_t)17, (size_t)1000, (size_t)2256, n / 2, n}) {
        std::vector<SI> b = a;
        const size_t end = upb::sort_prefix_like_std_sort(b.begin(), b.end(), bound, cmp);
        ++cases;
        if (end < std::min(bound, n)) { ++bad; continue; }
        for (size_t i = 0; i < end; ++i) if (b[i].idx != full[i].idx) { ++bad; break; }
      }
    }
    printf("sort cases=%ld bad=%ld\n", cases, bad);
    return bad ? 1 : 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "wig")) {
    const ContigTable contigs = parse_contig_table(argv[2]);
    Sample s;
    read_wig(argv[3], contigs, argc > 4 ? atoi(argv[4]) : 0, argc > 5 ? (unsigned)atoi(argv[5]) : 0, &s);
    if (!s.error.empty()) { printf("error\t%s\n", s.error.c_str()); return 0; }
    printf("name\t%s\nexpected\t%llu\nconfident\t%llu\noob\t%llu\ntotal\t%llu\n", s.name.c_str(), (unsigned long long)s.expectedTags,
           (unsigned long long)s.confidentTags, (unsigned long long)s.outOfBoundsTags, (unsigned long long)s.totalTags);
    for (int d = 0; d < 2; ++d)
      for (uint32_t c = 0; c < contigs.size(); ++c)
        for (size_t i = 0; i < s.pos[d][c].size(); ++i) printf("%d\t%u\t%u\t%u\n", d, c, s.pos[d][c][i], s.cnt[d][c][i]);
    return 0;
  }
  if (argc >= 4 && !strcmp(argv[1], "wigtime")) {                  // parse only: seconds, entries, a checksum of what was read
    const ContigTable contigs = parse_contig_table(argv[2]);
    Sample s;
    const auto t0 = std::chrono::steady_clock::now();
    read_wig(argv[3], contigs, argc > 4 ? atoi(argv[4]) : 0, argc > 5 ? (unsigned)atoi(argv[5]) : 0, &s);
    const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (!s.error.empty()) { printf("error\t%s\n", s.error.c_str()); return 0; }
    unsigned long long n = 0, h = 1469598103934665603ull;
    for (int d = 0; d < 2; ++d)
      for (uint32_t c = 0; c < contigs.size(); ++c)
        for (size_t i = 0; i < s.pos[d][c].size(); ++i) { ++n; h = (h ^ s.pos[d][c][i]) * 1099511628211ull; h = (h ^ s.cnt[d][c][i]) * 1099511628211ull; h = (h ^ (unsigned)(d * 131 + c)) * 1099511628211ull; }
    printf("seconds\t%.4f\nentries\t%llu\nchecksum\t%016llx\ntotal\t%llu\nconfident\t%llu\noob\t%llu\n", dt, n, h, (unsigned long long)s.totalTags,
           (unsigned long long)s.confidentTags, (unsigned long long)s.outOfBoundsTags);
    return 0;
  }
  fprintf(stderr, "usage: host_selftest g6 [n] | sort | wig <contig table> <wig> [offset] [length] | wigtime <contig table> <wig> [offset] [length]\n");
  return 2;
}
