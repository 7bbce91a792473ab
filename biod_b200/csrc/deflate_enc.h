// One BGZF payload: raw DEFLATE of up to 65 280 bytes (first part: bgzfCompress,
// bio/core/bgzf/compress.d:43-103, which hands the chunk to zlib's deflate(Z_FINISH) with winbits -15).
//
// The reference's own tests only ask that the bytes come back (bgzf/outputstream.d:225-247, test/unittests.d:286-305):
// the compressed bytes need not equal zlib's.  This encoder is a single final block — greedy LZ77 over a 4096-entry
// hash of 3-byte prefixes, coded with dynamic Huffman codes (RFC 1951 §3.2.7) or the fixed ones (§3.2.6), whichever is
// shorter — or a stored block when that is smaller still (level 0 always stores), written for one CUDA thread per BGZF
// block (csrc/deflate.cu) and compiled for the host as well, where
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

// ---- greedy LZ77 parse, shared by the counting pass and the emitting pass (same input, same table: same parse) ----
template <typename Sink>
BIODB_HD void defl_parse(const uint8_t* in, uint32_t n, uint16_t* htab, Sink& sink) {
  for (uint32_t k = 0; k < DEFL_HASH_SIZE; ++k) htab[k] = 0xFFFF;
  uint32_t i = 0;
  while (i < n && !sink.stop()) {
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
      sink.match(best, dist);
      // index the positions the match covers (sparsely for long matches)
      const uint32_t step = best > 32 ? 8 : 1;
      for (uint32_t k = i + 1; k < i + best && k + 3 <= n; k += step) htab[defl_hash(in + k)] = (uint16_t)k;
      i += best;
    } else {
      sink.literal(in[i]);
      ++i;
    }
  }
}

// length / distance -> code index and extra bits (RFC 1951 §3.2.5)
BIODB_HD void defl_len_code(uint32_t len, uint32_t* idx, uint32_t* ebits, uint32_t* eval) {
  if (len == 258) { *idx = 28; *ebits = 0; *eval = 0; return; }
  const uint32_t l = len - 3;
  if (l < 8) { *idx = l; *ebits = 0; *eval = 0; return; }
  const uint32_t e = defl_log2(l) - 2;
  *idx = (e << 2) + ((l >> e) & 3) + 4;
  *ebits = e;
  *eval = l & ((1u << e) - 1);
}
BIODB_HD void defl_dist_code(uint32_t dist, uint32_t* idx, uint32_t* ebits, uint32_t* eval) {
  const uint32_t d = dist - 1;
  if (d < 4) { *idx = d; *ebits = 0; *eval = 0; return; }
  const uint32_t nb = defl_log2(d), e = nb - 1;
  *idx = 2 * nb + ((d >> e) & 1);
  *ebits = e;
  *eval = d & ((1u << e) - 1);
}

struct DeflCount {                    // pass 1: symbol frequencies and the extra bits they drag along
  uint16_t ll[286];
  uint16_t dd[30];
  uint32_t extra;
  BIODB_HD void init() {
    for (int k = 0; k < 286; ++k) ll[k] = 0;
    for (int k = 0; k < 30; ++k) dd[k] = 0;
    extra = 0;
    ll[256] = 1;                      // end of block
  }
  BIODB_HD bool stop() const { return false; }
  BIODB_HD void literal(uint8_t c) { ++ll[c]; }
  BIODB_HD void match(uint32_t len, uint32_t dist) {
    uint32_t i, eb, ev;
    defl_len_code(len, &i, &eb, &ev);
    ++ll[257 + i];
    extra += eb;
    defl_dist_code(dist, &i, &eb, &ev);
    ++dd[i];
    extra += eb;
  }
};

struct DeflWork {                     // scratch of one block (about 6 KB): kept out of the thread's stack on the device
  DeflCount cnt;
  uint8_t ll_len[286], dd_len[30], cl_len[19];
  uint16_t ll_code[286], dd_code[30], cl_code[19], cl_freq[19];
  uint8_t cl_sym[320], cl_ext[320];
  uint32_t weight[2 * 286];
  uint16_t parent[2 * 286];
  uint16_t order[286];
};

// Code lengths (at most max_len bits) of a Huffman code for freq[0, n), 2 <= n <= 286: symbols never used get 0 — except
// that fewer than two used symbols are topped up to two codes of one bit each (as zlib does: every decoder accepts a
// complete code, not every one an empty or a one-code set).  The tree comes from repeatedly joining the two lightest nodes; lengths beyond max_len are folded
// back by moving codes down the length histogram until Kraft's sum is 1 again, and the lengths are then dealt out in
// order of frequency.
BIODB_HD void defl_code_lengths(const uint16_t* freq, uint32_t n, uint32_t max_len, uint8_t* len, DeflWork* w) {
  uint32_t* weight = w->weight;
  uint16_t* parent = w->parent;
  uint16_t* order = w->order;
  uint32_t used = 0;
  for (uint32_t k = 0; k < n; ++k) {
    len[k] = 0;
    if (freq[k]) order[used++] = (uint16_t)k;
  }
  if (used < 2) {
    const uint32_t first = used ? order[0] : 0;
    len[first] = 1;
    len[first == 0 ? 1 : 0] = 1;
    return;
  }
  // leaves in order of rising weight (insertion sort: at most 286 symbols), internal nodes after them; the nodes a
  // join creates come out in rising weight too, so the two lightest roots are always at the front of one of the two
  // queues (no search)
  for (uint32_t i = 1; i < used; ++i) {
    const uint16_t v = order[i];
    uint32_t j = i;
    while (j > 0 && freq[order[j - 1]] > freq[v]) { order[j] = order[j - 1]; --j; }
    order[j] = v;
  }
  for (uint32_t k = 0; k < used; ++k) { weight[k] = freq[order[k]]; parent[k] = 0xFFFF; }
  uint32_t li = 0, qi = used, qn = used;
  for (uint32_t m = 0; m + 1 < used; ++m) {
    uint32_t pick[2];
    for (int t = 0; t < 2; ++t) {
      if (li < used && (qi >= qn || weight[li] <= weight[qi])) pick[t] = li++;
      else pick[t] = qi++;
    }
    weight[qn] = weight[pick[0]] + weight[pick[1]];
    parent[qn] = 0xFFFF;
    parent[pick[0]] = parent[pick[1]] = (uint16_t)qn;
    ++qn;
  }
  uint32_t hist[33];
  for (int k = 0; k < 33; ++k) hist[k] = 0;
  for (uint32_t k = 0; k < used; ++k) {
    uint32_t d = 0;
    for (uint32_t x = k; parent[x] != 0xFFFF; x = parent[x]) ++d;
    if (d > 32) d = 32;
    ++hist[d];
  }
  // enforce the limit on the histogram of lengths
  for (uint32_t l = max_len + 1; l <= 32; ++l) { hist[max_len] += hist[l]; hist[l] = 0; }
  uint32_t total = 0;
  for (uint32_t l = max_len; l >= 1; --l) total += hist[l] << (max_len - l);
  while (total > (1u << max_len)) {
    --hist[max_len];
    for (uint32_t l = max_len - 1; l >= 1; --l)
      if (hist[l]) { --hist[l]; hist[l + 1] += 2; break; }
    --total;
  }
  // the most frequent symbols (the end of `order`) get the shortest lengths
  uint32_t k = used;
  for (uint32_t l = 1; l <= max_len; ++l)
    for (uint32_t c = 0; c < hist[l]; ++c) len[order[--k]] = (uint8_t)l;
}

// canonical codes (RFC 1951 §3.2.2), already bit-reversed for the LSB-first writer
BIODB_HD void defl_make_codes(const uint8_t* len, uint32_t n, uint32_t max_len, uint16_t* code) {
  uint32_t count[16], next[16];
  for (int k = 0; k < 16; ++k) count[k] = 0;
  for (uint32_t k = 0; k < n; ++k) ++count[len[k]];
  count[0] = 0;
  uint32_t c = 0;
  next[0] = 0;
  for (uint32_t l = 1; l <= max_len; ++l) { c = (c + count[l - 1]) << 1; next[l] = c; }
  for (uint32_t k = 0; k < n; ++k) code[k] = len[k] ? (uint16_t)defl_rev(next[len[k]]++, len[k]) : 0;
}

struct DeflEmit {                     // pass 2: the symbols in the chosen code
  DeflBits* b;
  const uint8_t* ll_len;  const uint16_t* ll_code;      // dynamic code, or nullptr for the fixed one
  const uint8_t* dd_len;  const uint16_t* dd_code;
  uint32_t limit;
  BIODB_HD bool stop() const { return b->n > limit; }   // already larger than the stored form
  BIODB_HD void sym(uint32_t s) {
    if (ll_len) b->put(ll_code[s], ll_len[s]); else defl_put_sym(*b, s);
  }
  BIODB_HD void literal(uint8_t c) { sym(c); }
  BIODB_HD void match(uint32_t len, uint32_t dist) {
    if (!ll_len) { defl_put_match(*b, len, dist); return; }
    uint32_t i, eb, ev;
    defl_len_code(len, &i, &eb, &ev);
    sym(257 + i);
    b->put(ev, eb);
    defl_dist_code(dist, &i, &eb, &ev);
    b->put(dd_code[i], dd_len[i]);
    b->put(ev, eb);
  }
};

// Raw DEFLATE of in[0, n) into out[0, cap).  htab: DEFL_HASH_SIZE entries of scratch.  level 0 stores; every other
// level: one final block, greedy LZ77, with dynamic Huffman codes (RFC 1951 §3.2.7) or the fixed ones, whichever is
// shorter — or stored after all if that beats both.
// Returns the number of bytes written, 0 if cap is too small (cap >= n + 5 always suffices).
BIODB_HD uint32_t deflate_block(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, uint16_t* htab, int level,
                                DeflWork* w) {
  if (n > DEFL_MAX_IN) return 0;
  if (level == 0 || n < 8) return deflate_stored(in, n, out, cap);
  // pass 1: what the parse will emit
  DeflCount& cnt = w->cnt;
  cnt.init();
  defl_parse(in, n, htab, cnt);
  uint8_t* ll_len = w->ll_len;
  uint8_t* dd_len = w->dd_len;
  uint16_t* ll_code = w->ll_code;
  uint16_t* dd_code = w->dd_code;
  defl_code_lengths(cnt.ll, 286, 15, ll_len, w);
  defl_code_lengths(cnt.dd, 30, 15, dd_len, w);
  uint32_t n_ll = 286, n_dd = 30;
  while (n_ll > 257 && ll_len[n_ll - 1] == 0) --n_ll;
  while (n_dd > 1 && dd_len[n_dd - 1] == 0) --n_dd;
  // the code lengths themselves, run-length coded with the symbols 16 / 17 / 18 (RFC 1951 §3.2.7)
  uint8_t* cl_sym = w->cl_sym;
  uint8_t* cl_ext = w->cl_ext;
  uint16_t* cl_freq = w->cl_freq;
  for (int k = 0; k < 19; ++k) cl_freq[k] = 0;
  uint32_t n_cl = 0;
  {
    const uint32_t tot = n_ll + n_dd;
    uint32_t i = 0;
    while (i < tot) {
      const uint8_t v = i < n_ll ? ll_len[i] : dd_len[i - n_ll];
      uint32_t run = 1;
      while (i + run < tot && (i + run < n_ll ? ll_len[i + run] : dd_len[i + run - n_ll]) == v) ++run;
      if (v == 0 && run >= 3) {
        uint32_t r = run;
        while (r >= 3) {
          const uint32_t t = r > 138 ? 138 : r;
          if (t >= 11) { cl_sym[n_cl] = 18; cl_ext[n_cl] = (uint8_t)(t - 11); }
          else { cl_sym[n_cl] = 17; cl_ext[n_cl] = (uint8_t)(t - 3); }
          ++cl_freq[cl_sym[n_cl]];
          ++n_cl;
          r -= t;
        }
        for (; r; --r) { cl_sym[n_cl] = 0; cl_ext[n_cl] = 0; ++cl_freq[0]; ++n_cl; }
      } else if (v != 0 && run >= 4) {
        cl_sym[n_cl] = v; cl_ext[n_cl] = 0; ++cl_freq[v]; ++n_cl;          // the length itself, then repeats of it
        uint32_t r = run - 1;
        while (r >= 3) {
          const uint32_t t = r > 6 ? 6 : r;
          cl_sym[n_cl] = 16; cl_ext[n_cl] = (uint8_t)(t - 3); ++cl_freq[16]; ++n_cl;
          r -= t;
        }
        for (; r; --r) { cl_sym[n_cl] = v; cl_ext[n_cl] = 0; ++cl_freq[v]; ++n_cl; }
      } else {
        for (uint32_t r = 0; r < run; ++r) { cl_sym[n_cl] = v; cl_ext[n_cl] = 0; ++cl_freq[v]; ++n_cl; }
      }
      i += run;
    }
  }
  uint8_t* cl_len = w->cl_len;
  uint16_t* cl_code = w->cl_code;
  defl_code_lengths(cl_freq, 19, 7, cl_len, w);
  const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  uint32_t n_clc = 19;
  while (n_clc > 4 && cl_len[cl_order[n_clc - 1]] == 0) --n_clc;
  // sizes in bits of the two forms
  uint64_t dyn_bits = 3 + 5 + 5 + 4 + 3ull * n_clc + cnt.extra, fix_bits = 3 + cnt.extra;
  for (uint32_t k = 0; k < n_cl; ++k)
    dyn_bits += cl_len[cl_sym[k]] + (cl_sym[k] == 16 ? 2 : cl_sym[k] == 17 ? 3 : cl_sym[k] == 18 ? 7 : 0);
  for (uint32_t k = 0; k < 286; ++k) {
    dyn_bits += (uint64_t)cnt.ll[k] * ll_len[k];
    fix_bits += (uint64_t)cnt.ll[k] * (k < 144 ? 8 : k < 256 ? 9 : k < 280 ? 7 : 8);
  }
  for (uint32_t k = 0; k < 30; ++k) {
    dyn_bits += (uint64_t)cnt.dd[k] * dd_len[k];
    fix_bits += (uint64_t)cnt.dd[k] * 5;
  }
  const bool dynamic = dyn_bits < fix_bits;
  // pass 2: the compressed form is kept only if it beats the stored one
  const uint32_t limit = (n + 4 < cap) ? n + 4 : cap;
  DeflBits b{out, 0, limit, 0, 0};
  b.put(1, 1);                                              // BFINAL
  if (dynamic) {
    defl_make_codes(ll_len, 286, 15, ll_code);
    defl_make_codes(dd_len, 30, 15, dd_code);
    defl_make_codes(cl_len, 19, 7, cl_code);
    b.put(2, 2);                                            // BTYPE = 10 (dynamic Huffman)
    b.put(n_ll - 257, 5);
    b.put(n_dd - 1, 5);
    b.put(n_clc - 4, 4);
    for (uint32_t k = 0; k < n_clc; ++k) b.put(cl_len[cl_order[k]], 3);
    for (uint32_t k = 0; k < n_cl; ++k) {
      b.put(cl_code[cl_sym[k]], cl_len[cl_sym[k]]);
      if (cl_sym[k] == 16) b.put(cl_ext[k], 2);
      else if (cl_sym[k] == 17) b.put(cl_ext[k], 3);
      else if (cl_sym[k] == 18) b.put(cl_ext[k], 7);
    }
  } else {
    b.put(1, 2);                                            // BTYPE = 01 (fixed Huffman)
  }
  DeflEmit em{&b, dynamic ? ll_len : nullptr, ll_code, dd_len, dd_code, limit};
  defl_parse(in, n, htab, em);
  em.sym(256);                                              // end of block
  b.flush();
  if (b.n <= limit && b.n < n + 5) return b.n;
  return deflate_stored(in, n, out, cap);
}

}  // namespace biodb
