// One BGZF payload: raw DEFLATE of up to 65 280 bytes (first part: bgzfCompress,
// bio/core/bgzf/compress.d:43-103, which hands the chunk to zlib's deflate(Z_FINISH) with winbits -15).
//
// The reference's own tests only ask that the bytes come back (bgzf/outputstream.d:225-247, test/unittests.d:286-305):
// the compressed bytes need not equal zlib's.  This encoder is a single final block — greedy LZ77 over a 4096-entry
// hash of 3-byte prefixes, fixed Huffman codes (RFC 1951 §3.2.6) — or a stored block when that is smaller (level 0
// always stores), written for one CUDA thread per BGZF block (csrc/deflate.cu) and compiled for the host as well, where
// the CPU tests inflate its output with zlib (biodb_debug_deflate_block).
#pragma once
#include <stdint.h>

#include "md_walk.h"   // BIODB_HD

namespace biodb {

constexpr uint32_t DEFL_HASH_BITS = 12;
constexpr uint32_t DEFL_HASH_SIZE = 1u << DEFL_HASH_BITS;
constexpr uint32_t DEFL_MAX_IN = 65535;          // one stored block holds at most this (a BGZF chunk is <= 0xFF00)

struct DeflBits {
  uint8_t* p;
  uint32_t n, cap;
  uint64_t acc;
  uint32_t bits;
  BIODB_HD void put(uint32_t v, uint32_t nb) {    // nb <= 16, LSB first
    acc |= (uint64_t)v << bits;
    bits += nb;
    while (bits >= 8) {
      if (n < cap) p[n] = (uint8_t)acc;
      ++n;
      acc >>= 8;
      bits -= 8;
    }
  }
  BIODB_HD void flush() {
    if (bits) {
      if (n < cap) p[n] = (uint8_t)acc;
      ++n;
      acc = 0;
      bits = 0;
    }
  }
};

BIODB_HD uint32_t defl_rev(uint32_t v, uint32_t nb) {      // Huffman codes go out most significant bit first
  uint32_t r = 0;
  for (uint32_t k = 0; k < nb; ++k) r |= ((v >> k) & 1u) << (nb - 1 - k);
  return r;
}
BIODB_HD uint32_t defl_log2(uint32_t v) {                   // floor(log2(v)), v >= 1
  uint32_t r = 0;
  while (v >>= 1) ++r;
  return r;
}
// literal / length symbol with the fixed code of RFC 1951 §3.2.6
BIODB_HD void defl_put_sym(DeflBits& b, uint32_t s) {
  if (s < 144) b.put(defl_rev(0x30 + s, 8), 8);
  else if (s < 256) b.put(defl_rev(0x190 + (s - 144), 9), 9);
  else if (s < 280) b.put(defl_rev(s - 256, 7), 7);
  else b.put(defl_rev(0xC0 + (s - 280), 8), 8);
}
// a match of `len` (3..258) bytes `dist` (1..32768) back
BIODB_HD void defl_put_match(DeflBits& b, uint32_t len, uint32_t dist) {
  if (len == 258) {
    defl_put_sym(b, 285);
  } else {
    const uint32_t l = len - 3;
    if (l < 8) {
      defl_put_sym(b, 257 + l);
    } else {
      const uint32_t e = defl_log2(l) - 2;                  // extra bits
      defl_put_sym(b, 257 + (e << 2) + ((l >> e) & 3) + 4);
      b.put(l & ((1u << e) - 1), e);
    }
  }
  const uint32_t d = dist - 1;
  if (d < 4) {
    b.put(defl_rev(d, 5), 5);
  } else {
    const uint32_t nb = defl_log2(d), e = nb - 1;
    b.put(defl_rev(2 * nb + ((d >> e) & 1), 5), 5);
    b.put(d & ((1u << e) - 1), e);
  }
}

BIODB_HD uint32_t defl_hash(const uint8_t* p) {
  const uint32_t v = (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16);
  return (v * 2654435761u) >> (32 - DEFL_HASH_BITS);
}

// Stored block(s): 5 bytes of header per block (n <= 65535: one block).  Returns the size, 0 if it does not fit.
BIODB_HD uint32_t deflate_stored(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap) {
  if (n > DEFL_MAX_IN || n + 5 > cap) return 0;
  out[0] = 1;                                               // BFINAL = 1, BTYPE = 00
  out[1] = (uint8_t)n;
  out[2] = (uint8_t)(n >> 8);
  out[3] = (uint8_t)~n;
  out[4] = (uint8_t)(~n >> 8);
  for (uint32_t i = 0; i < n; ++i) out[5 + i] = in[i];
  return n + 5;
}

// Raw DEFLATE of in[0, n) into out[0, cap).  htab: DEFL_HASH_SIZE entries of scratch.  level 0 stores.
// Returns the number of bytes written, 0 if cap is too small (cap >= n + 5 always suffices).
BIODB_HD uint32_t deflate_block(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, uint16_t* htab, int level) {
  if (n > DEFL_MAX_IN) return 0;
  if (level == 0 || n < 8) return deflate_stored(in, n, out, cap);
  for (uint32_t k = 0; k < DEFL_HASH_SIZE; ++k) htab[k] = 0xFFFF;
  // the compressed form is kept only if it beats the stored one
  const uint32_t limit = (n + 4 < cap) ? n + 4 : cap;
  DeflBits b{out, 0, limit, 0, 0};
  b.put(1, 1);                                              // BFINAL
  b.put(1, 2);                                              // BTYPE = 01 (fixed Huffman)
  uint32_t i = 0;
  while (i < n) {
    uint32_t best = 0, dist = 0;
    if (i + 3 <= n) {
      const uint32_t h = defl_hash(in + i);
      const uint32_t c = htab[h];
      htab[h] = (uint16_t)i;
      if (c != 0xFFFF && i - c <= 32768) {
        const uint32_t maxl = (n - i < 258) ? n - i : 258;
        uint32_t l = 0;
        while (l < maxl && in[c + l] == in[i + l]) ++l;
        if (l >= 4 || (l == 3 && i - c < 4096)) { best = l; dist = i - c; }
      }
    }
    if (best) {
      defl_put_match(b, best, dist);
      // index the positions the match covers (sparsely for long matches)
      const uint32_t step = best > 32 ? 8 : 1;
      for (uint32_t k = i + 1; k < i + best && k + 3 <= n; k += step) htab[defl_hash(in + k)] = (uint16_t)k;
      i += best;
    } else {
      defl_put_sym(b, in[i]);
      ++i;
    }
    if (b.n > limit) break;                                 // already larger than the stored form
  }
  defl_put_sym(b, 256);                                     // end of block
  b.flush();
  if (b.n <= limit && b.n < n + 5) return b.n;
  return deflate_stored(in, n, out, cap);
}

}  // namespace biodb
