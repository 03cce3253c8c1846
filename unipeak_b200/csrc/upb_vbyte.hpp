// upb_vbyte.hpp — the variable-length byte form of the tag arrays (include/unipeak_b200.h: upb_load_tags_vbyte): the per-entry
// code, shared by the library's host encoder (upb_pack_tags_vbyte), the drivers' TagLoader and the device decoder (upb_pack.cuh).
//
//   first byte   1 c v5..v0     bit 7: an entry starts here; c: a count byte closes the entry; v: low 6 bits of the value
//   0 - 4 bytes  0 xxxxxxx      higher bits of the value, 7 per byte, low to high (6 + 4 * 7 = 34 bits hold any u32)
//   (c) 1 byte   0 kkkkkkk      count 2 .. 127; 0: the count is in the exception list
// value = position - previous position in the segment; the absolute position for the first entry of a tile of VB_TILE entries.
#pragma once

#include <stdint.h>

#if defined(__CUDACC__)
#define UPB_VB_HD __host__ __device__ __forceinline__
#else
#define UPB_VB_HD inline
#endif

namespace upb_vb {

constexpr uint32_t VB_TILE = 2048;                  // == PK_TILE (upb_pack.cuh)
constexpr uint32_t VB_MAX_ENTRY = 6;                // bytes: 1 + 4 + 1
constexpr uint32_t VB_MAX_TILE_BYTES = VB_TILE * VB_MAX_ENTRY;

// bytes entry (v, count) takes; *escaped = its count has to go to the exception list
UPB_VB_HD uint32_t vb_size (uint32_t v, uint32_t count, bool* escaped) {
  uint32_t n = 1;
  for (uint32_t r = v >> 6; r; r >>= 7) ++n;
  *escaped = count != 1 && (count < 2 || count > 127);
  return n + (count != 1 ? 1u : 0u);
}

UPB_VB_HD uint32_t vb_put (uint8_t* out, uint32_t v, uint32_t count) {
  const bool c = count != 1;
  out[0] = (uint8_t)(0x80u | (c ? 0x40u : 0u) | (v & 63u));
  uint32_t n = 1;
  for (uint32_t r = v >> 6; r; r >>= 7) out[n++] = (uint8_t)(r & 127u);
  if (c) out[n++] = (uint8_t)((count >= 2 && count <= 127) ? count : 0u);
  return n;
}

// decode the entry whose first byte is b[0] (bit 7 set) and which is followed by nCont continuation bytes (bit 7 clear);
// returns false when the entry is malformed.  count 0 = look it up in the exception list.
UPB_VB_HD bool vb_get (const uint8_t* b, uint32_t nCont, uint32_t* v, uint32_t* count) {
  const uint32_t first = b[0];
  const bool c = (first & 0x40u) != 0;
  if (nCont > 4u + (c ? 1u : 0u) || (c && nCont == 0)) return false;
  const uint32_t nv = nCont - (c ? 1u : 0u);
  uint64_t val = first & 63u;
  for (uint32_t k = 0; k < nv; ++k) val |= (uint64_t)b[1 + k] << (6 + 7 * k);
  if (val > 0xFFFFFFFFull) return false;
  if (nv > 0 && b[nv] == 0) return false;           // canonical: no leading zero bytes (one value, one code)
  *v = (uint32_t)val;
  *count = c ? (uint32_t)b[nCont] : 1u;
  if (c && *count == 1) return false;               // count 1 has no count byte
  return true;
}

// ---- a tile decoded by T cooperating threads (the device decoder's two passes; host code runs them one thread after the other) ----
// thread t owns bytes [b0, b1) of the tile's `len` bytes
UPB_VB_HD void vb_chunk (uint32_t len, uint32_t nThreads, uint32_t t, uint32_t* b0, uint32_t* b1) {
  const uint32_t per = (len + nThreads - 1) / nThreads;
  *b0 = t * per < len ? t * per : len;
  *b1 = *b0 + per < len ? *b0 + per : len;
}
// pass 1: entries that START in the chunk
UPB_VB_HD uint32_t vb_count_starts (const uint8_t* tb, uint32_t b0, uint32_t b1) {
  uint32_t n = 0;
  for (uint32_t b = b0; b < b1; ++b) n += tb[b] >> 7;
  return n;
}
// pass 2: decode them; j = number of entries that start before the chunk.  An entry may run on into the next chunks.  Entries beyond
// `cap` are dropped (the caller reports a tile that holds more than its entries).  Returns false if an entry is malformed.
template <class CountT>
UPB_VB_HD bool vb_decode_chunk (const uint8_t* tb, uint32_t len, uint32_t b0, uint32_t b1, uint32_t j, uint32_t cap, uint32_t* eVal, CountT* eCnt) {
  bool ok = true;
  for (uint32_t b = b0; b < b1; ++b) {
    if (!(tb[b] & 0x80u)) continue;
    uint32_t nCont = 0;
    while (nCont <= VB_MAX_ENTRY && b + 1 + nCont < len && !(tb[b + 1 + nCont] & 0x80u)) ++nCont;
    uint32_t v = 0, k = 1;
    if (!vb_get(tb + b, nCont, &v, &k)) { ok = false; v = 0; k = 1; }
    if (j < cap) { eVal[j] = v; eCnt[j] = (CountT)k; }
    ++j;
  }
  return ok;
}

// ---- the same two passes over a STAGED tile, 16 bytes at a time (what the device decoder runs; the byte-wise functions above are
// the plain statement of it).  The tile's `len` bytes sit at sb[shift .. shift + len) of a 16-byte aligned buffer (shift < 16: the
// device copies whole aligned 16-byte words) and readable up to sb[shift + len + 11].
UPB_VB_HD uint32_t vb_ctz64 (uint64_t x) {
#if defined(__CUDA_ARCH__)
  return (uint32_t)__ffsll((long long)x) - 1u;
#else
  return (uint32_t)__builtin_ctzll(x);
#endif
}
// thread t owns the 16-byte groups [g0, g1)
UPB_VB_HD void vb_chunk16 (uint32_t shift, uint32_t len, uint32_t nThreads, uint32_t t, uint32_t* g0, uint32_t* g1) {
  const uint32_t nG = len ? (shift + len + 15) >> 4 : 0u;
  const uint32_t per = (nG + nThreads - 1) / nThreads;
  *g0 = t * per < nG ? t * per : nG;
  *g1 = *g0 + per < nG ? *g0 + per : nG;
}
// bit i = byte i of group g (words w[0..3]) starts an entry of the tile
UPB_VB_HD uint32_t vb_starts16 (const uint32_t* w, uint32_t g, uint32_t shift, uint32_t len) {
  uint32_t m = 0;
  for (int q = 0; q < 4; ++q) { const uint32_t x = (w[q] >> 7) & 0x01010101u; m |= ((x | (x >> 7) | (x >> 14) | (x >> 21)) & 0xFu) << (4 * q); }
  const uint32_t first = 16 * g, end = shift + len;
  const uint32_t lo = first < shift ? shift - first : 0u;               // bytes of the group before the tile
  const uint32_t hi = end - first >= 16 ? 16u : end - first;            // ... and from its end on
  return m & (0xFFFFu << lo) & ((1u << hi) - 1u);
}
// pass 2a: where the entries of the groups [g0, g1) start (byte index into sb); j = entries that start before them
UPB_VB_HD void vb_list_starts (const uint32_t* sbw, uint32_t shift, uint32_t len, uint32_t g0, uint32_t g1, uint32_t j, uint32_t cap, uint16_t* startAt) {
  for (uint32_t g = g0; g < g1; ++g) {
    uint32_t m = vb_starts16(sbw + 4 * g, g, shift, len);
    while (m) {
      const uint32_t b = vb_ctz64(m);
      m &= m - 1;
      if (j < cap) startAt[j] = (uint16_t)(16 * g + b);
      ++j;
    }
  }
}
// pass 2b: the entry that occupies the n bytes sb[at .. at + n) (n = distance to the next entry's first byte, or to the end of the
// tile).  Returns false when it is malformed.
UPB_VB_HD bool vb_get_span (const uint32_t* sbw, uint32_t at, uint32_t n, uint32_t* v, uint32_t* count) {
  const uint32_t wi = at >> 2, sh = (at & 3u) * 8u;
  const uint32_t a = sbw[wi], b = sbw[wi + 1], c3 = sbw[wi + 2];
#if defined(__CUDA_ARCH__)
  const uint32_t lo = __funnelshift_r(a, b, sh), hi = __funnelshift_r(b, c3, sh);
#else
  const uint32_t lo = sh ? (a >> sh) | (b << (32u - sh)) : a, hi = sh ? (b >> sh) | (c3 << (32u - sh)) : b;
#endif
  const uint32_t c = (lo >> 6) & 1u;
  const uint32_t nv = n - 1u - c;                                       // value bytes behind the first (wraps when n is too small)
  if (n < 1u + c || nv > 4u) return false;
  // e1 .. e4, seven bits each, closed up into one 28-bit number; the value takes the lowest 7 nv bits of it
  const uint32_t pay = ((lo >> 8) & 0x7Fu) | ((lo >> 9) & 0x3F80u) | ((lo >> 10) & 0x1FC000u) | ((hi & 0x7Fu) << 21);
  const uint32_t used = pay & ((1u << (7u * nv)) - 1u);                  // (nv = 0: nothing; nv = 4: all 28 bits)
  if (nv == 4u && (used >> 26) != 0) return false;                      // beyond 32 bits
  if (nv > 0 && (used >> (7u * nv - 7u)) == 0) return false;            // canonical: the highest value byte is not 0
  const uint32_t k = (uint32_t)((((uint64_t)hi << 32) | lo) >> (8u * n - 8u)) & 127u;   // the byte at n - 1
  *v = (lo & 63u) | (used << 6);
  *count = c ? k : 1u;
  return !(c && k == 1u);
}

}  // namespace upb_vb
