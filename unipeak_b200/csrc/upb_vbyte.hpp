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

}  // namespace upb_vb
