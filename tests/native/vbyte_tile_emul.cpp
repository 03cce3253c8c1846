// Host emulation of the device decoder's tile stage (unipeak_b200/csrc/upb_pack.cuh, load_tile_kernel<LOAD_VBYTE>): the same
// upb_vbyte.hpp functions the kernel calls, run for 256 "threads" one after the other on random tiles — entries of every length,
// entries that straddle the threads' byte ranges, short last tiles — against the values that were encoded.
// g++ -std=c++17 -O1 -o vbyte_tile_emul vbyte_tile_emul.cpp && ./vbyte_tile_emul <seed> <tiles>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../../unipeak_b200/csrc/upb_vbyte.hpp"

using namespace upb_vb;

int main (int argc, char** argv) {
  const unsigned seed = argc > 1 ? (unsigned)atoi(argv[1]) : 1u;
  const int nTiles = argc > 2 ? atoi(argv[2]) : 200;
  std::mt19937_64 rng(seed);
  const uint32_t T = 256;
  long checked = 0;
  for (int t = 0; t < nTiles; ++t) {
    const uint32_t nHere = (t % 7 == 0) ? 1 + (uint32_t)(rng() % VB_TILE) : VB_TILE;
    std::vector<uint32_t> val(nHere), cnt(nHere);
    std::vector<uint8_t> bytes(VB_MAX_TILE_BYTES + 64, 0xEE);
    uint32_t len = 0;
    const int regime = t % 5;
    for (uint32_t i = 0; i < nHere; ++i) {
      const int bits = regime == 0 ? (int)(rng() % 33) : regime == 1 ? (int)(rng() % 7) : regime == 2 ? 6 + (int)(rng() % 8) : (int)(rng() % 14);
      uint32_t v = bits == 0 ? 0u : (uint32_t)(rng() & ((bits >= 32 ? 0xFFFFFFFFull : ((1ull << bits) - 1))));
      if (bits > 0 && bits <= 32) v |= 1u << (bits - 1);
      uint32_t k = 1;
      const int kr = (int)(rng() % 10);
      if (kr == 0) k = 2 + (uint32_t)(rng() % 126); else if (kr == 1) k = 128 + (uint32_t)(rng() % 100000); else if (kr == 2 && regime == 0) k = 0;
      val[i] = v; cnt[i] = k;
      bool esc = false;
      const uint32_t want = vb_size(v, k, &esc);
      const uint32_t got = vb_put(bytes.data() + len, v, k);
      if (want != got || got > VB_MAX_ENTRY) { printf("size mismatch\n"); return 1; }
      if (esc != (k != 1 && (k < 2 || k > 127))) { printf("escape flag\n"); return 1; }
      len += got;
    }
    // the kernel's passes
    std::vector<uint32_t> starts(T), eVal(VB_TILE, 0xDEADBEEF);
    std::vector<uint8_t> eCnt(VB_TILE, 0xEE);
    uint32_t total = 0;
    for (uint32_t th = 0; th < T; ++th) { uint32_t b0, b1; vb_chunk(len, T, th, &b0, &b1); starts[th] = vb_count_starts(bytes.data(), b0, b1); total += starts[th]; }
    if (total != nHere) { printf("tile %d: %u starts for %u entries\n", t, total, nHere); return 1; }
    uint32_t j = 0, covered = 0;
    for (uint32_t th = 0; th < T; ++th) {
      uint32_t b0, b1; vb_chunk(len, T, th, &b0, &b1);
      if (b0 != covered && b0 != len) { printf("chunks do not tile the bytes\n"); return 1; }
      covered = b1;
      if (!vb_decode_chunk(bytes.data(), len, b0, b1, j, nHere, eVal.data(), eCnt.data())) { printf("tile %d: malformed?\n", t); return 1; }
      j += starts[th];
    }
    if (covered != len) { printf("chunks do not cover the bytes\n"); return 1; }
    for (uint32_t i = 0; i < nHere; ++i) {
      const uint32_t k = (cnt[i] == 1) ? 1u : (cnt[i] >= 2 && cnt[i] <= 127 ? cnt[i] : 0u);
      if (eVal[i] != val[i] || eCnt[i] != k) { printf("tile %d entry %u: got (%u, %u), encoded (%u, %u)\n", t, i, eVal[i], (unsigned)eCnt[i], val[i], cnt[i]); return 1; }
      ++checked;
    }
    // the staged form the kernel runs: the tile at an arbitrary offset inside a 16-byte aligned buffer, garbage around it, sentinels behind it
    {
      const uint32_t shift = (uint32_t)(rng() % 16);
      std::vector<uint32_t> stw((VB_MAX_TILE_BYTES + 64) / 4, 0);
      uint8_t* st = (uint8_t*)stw.data();
      for (size_t i = 0; i < stw.size() * 4; ++i) st[i] = (uint8_t)rng();
      for (uint32_t i = 0; i < len; ++i) st[shift + i] = bytes[i];
      std::vector<uint32_t> starts2(T), eVal2(VB_TILE, 0xDEADBEEF);
      std::vector<uint8_t> eCnt2(VB_TILE, 0xEE);
      uint32_t total2 = 0, gPrev = 0;
      for (uint32_t th = 0; th < T; ++th) {
        uint32_t g0, g1; vb_chunk16(shift, len, T, th, &g0, &g1);
        if (g0 != gPrev && g0 != ((shift + len + 15) >> 4)) { printf("groups do not tile the buffer\n"); return 1; }
        gPrev = g1;
        uint32_t n = 0;
        for (uint32_t g = g0; g < g1; ++g) n += (uint32_t)__builtin_popcount(vb_starts16(stw.data() + 4 * g, g, shift, len));
        starts2[th] = n; total2 += n;
      }
      if (gPrev != ((shift + len + 15) >> 4)) { printf("groups do not cover the buffer\n"); return 1; }
      if (total2 != nHere) { printf("tile %d (staged, shift %u): %u starts for %u entries\n", t, shift, total2, nHere); return 1; }
      uint32_t j2 = 0;
      std::vector<uint16_t> startAt(VB_TILE + 8, 0xFFFF);
      for (uint32_t th = 0; th < T; ++th) {                                // pass 2a: every thread lists where its entries start
        uint32_t g0, g1; vb_chunk16(shift, len, T, th, &g0, &g1);
        vb_list_starts(stw.data(), shift, len, g0, g1, j2, nHere, startAt.data());
        j2 += starts2[th];
      }
      for (uint32_t q = 0; q < 8; ++q) startAt[nHere + q] = (uint16_t)(shift + len);
      for (uint32_t th = 0; th < T; ++th)                                  // pass 2b: thread th decodes entries 8 th .. 8 th + 7
        for (uint32_t q = 0; q < 8; ++q) {
          const uint32_t i = 8 * th + q;
          if (i >= nHere) continue;
          uint32_t v = 0, k = 0;
          if (!vb_get_span(stw.data(), startAt[i], (uint32_t)startAt[i + 1] - startAt[i], &v, &k)) { printf("tile %d entry %u (staged): malformed?\n", t, i); return 1; }
          eVal2[i] = v; eCnt2[i] = (uint8_t)k;
        }
      for (uint32_t i = 0; i < nHere; ++i)
        if (eVal2[i] != eVal[i] || eCnt2[i] != eCnt[i]) { printf("tile %d entry %u (staged, shift %u): (%u, %u) against (%u, %u)\n", t, i, shift, eVal2[i], (unsigned)eCnt2[i], eVal[i], (unsigned)eCnt[i]); return 1; }
    }
  }
  // malformed codes are refused
  { uint32_t v, k;
    const uint8_t a1[] = {0x80 | 0x40};                      if (vb_get(a1, 0, &v, &k)) { printf("count flag without count byte accepted\n"); return 1; }
    const uint8_t a2[] = {0x80, 1, 1, 1, 1, 1};              if (vb_get(a2, 5, &v, &k)) { printf("five value bytes accepted\n"); return 1; }
    const uint8_t a3[] = {0x80, 5, 0};                       if (vb_get(a3, 2, &v, &k)) { printf("non-canonical value accepted\n"); return 1; }
    const uint8_t a4[] = {0x80 | 0x40, 1};                   if (vb_get(a4, 1, &v, &k)) { printf("count byte 1 accepted\n"); return 1; }
    const uint8_t a5[] = {0x80, 127, 127, 127, 127};         if (vb_get(a5, 4, &v, &k)) { printf("value beyond 32 bits accepted\n"); return 1; }
    const uint8_t a6[] = {0x80 | 63, 127, 127, 127, 31};     if (!vb_get(a6, 4, &v, &k) || v != 0xFFFFFFFFu || k != 1) { printf("largest value refused\n"); return 1; }
    // the word-wise reader agrees on all of them (entry at byte 5 of a buffer)
    const uint8_t* cases[] = {a1, a2, a3, a4, a5, a6};
    const uint32_t lens[] = {1, 6, 3, 2, 5, 5};
    const bool good[] = {false, false, false, false, false, true};
    for (int q = 0; q < 6; ++q) {
      uint32_t bufw[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      uint8_t* buf = (uint8_t*)bufw;
      for (uint32_t i = 0; i < lens[q]; ++i) buf[5 + i] = cases[q][i];
      if (vb_get_span(bufw, 5, lens[q], &v, &k) != good[q]) { printf("word-wise reader disagrees on malformed case %d\n", q); return 1; }
    }
    { uint32_t bufw[8] = {0x80808080u, 0x80808080u, 0, 0, 0, 0, 0, 0};
      if (vb_get_span(bufw, 1, 0, &v, &k) || vb_get_span(bufw, 1, 7, &v, &k) || vb_get_span(bufw, 1, 0xFFFFFFF0u, &v, &k)) { printf("impossible span accepted\n"); return 1; } } }
  printf("ok %ld entries\n", checked);
  return 0;
}
