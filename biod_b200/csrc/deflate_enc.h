// One BGZF payload: raw DEFLATE of up to 65 280 bytes (first part: bgzfCompress,
// bio/core/bgzf/compress.d:43-103, which hands the chunk to zlib's deflate(Z_FINISH) with winbits -15).
//
// The reference's own tests only ask that the bytes come back (bgzf/outputstream.d:225-247, test/unittests.d:286-305):
// the compressed bytes need not equal zlib's.  This encoder writes a single final block:
//   * LZ77 over windows of 32 consecutive positions, written so that a warp can look at the 32 positions at once
//     (csrc/deflate.cu: deflate_warp_kernel) and a plain loop over "lanes" gives the same tokens on the host
//     (deflate_block_host below — the CPU tests inflate its output with zlib, the GPU tests compare the two byte for
//     byte).  Per position: a hash of 4 bytes into a table of 4096 buckets x 2 ways, up to three candidates (the
//     nearest earlier position of the same window with that hash, the bucket's two entries), the longest match of at
//     least 4 bytes, one step of lazy evaluation (a position whose successor has a longer match stays a literal), then
//     the greedy chain through the window;
//   * dynamic Huffman codes (RFC 1951 §3.2.7) from the token statistics, or the fixed ones (§3.2.6), whichever is
//     shorter — or a stored block when that is smaller still (level 0 always stores).
#pragma once
#include <stdint.h>

#include "md_walk.h"   // BIODB_HD

namespace biodb {

constexpr uint32_t DEFL_HASH_BITS = 12;
constexpr uint32_t DEFL_HASH_SIZE = 1u << DEFL_HASH_BITS;   // buckets
constexpr uint32_t DEFL_WAYS = 2;                           // entries per bucket: the two latest positions with that hash
constexpr uint32_t DEFL_MIN_MATCH = 4;
constexpr uint32_t DEFL_WIN = 32;                           // positions looked at together
constexpr uint32_t DEFL_WCAP = 40;                          // match lengths are first measured up to this (>= DEFL_WIN:
                                                            // a match that reaches it ends the window, and only that
                                                            // one is then followed to its real end)
constexpr uint32_t DEFL_MAX_IN = 65535;          // one stored block holds at most this (a BGZF chunk is <= 0xFF00)
constexpr uint32_t DEFL_NONE = 0xFFFF;           // empty table entry (positions are below 65532)
// tokens: a literal is its byte; a match is DEFL_TOK_MATCH | (length - 3) << 16 | (distance - 1); DEFL_TOK_EOB ends the list
constexpr uint32_t DEFL_TOK_MATCH = 0x80000000u;
constexpr uint32_t DEFL_TOK_EOB = 0x40000000u;

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
#ifdef __CUDA_ARCH__
  return 31u - (uint32_t)__clz((int)v);
#else
  return 31u - (uint32_t)__builtin_clz(v);
#endif
}

BIODB_HD uint32_t defl_hash(uint32_t first4) {              // of the 4 bytes at a position, little endian
  return (first4 * 2654435761u) >> (32 - DEFL_HASH_BITS);
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

struct DeflCount {                    // symbol frequencies of the token list and the extra bits the tokens drag along
  uint16_t ll[286];
  uint16_t dd[30];
  uint32_t extra;
  BIODB_HD void init() {
    for (int k = 0; k < 286; ++k) ll[k] = 0;
    for (int k = 0; k < 30; ++k) dd[k] = 0;
    extra = 0;
    ll[256] = 1;                      // end of block
  }
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

struct DeflWork {                     // scratch of one block (about 6 KB; shared memory on the device)
  DeflCount cnt;
  uint8_t ll_len[288], dd_len[30], cl_len[19];      // (288: the fixed code counts two symbols that are never used)
  uint16_t ll_code[288], dd_code[30], cl_code[19], cl_freq[19];
  uint8_t cl_sym[320], cl_ext[320];
  uint32_t n_ll, n_dd, n_cl, n_clc;
  uint32_t weight[2 * 286];
  uint16_t parent[2 * 286];
  uint16_t order[286];
};

// The used symbols of freq[0, n) in order of rising (frequency, symbol) into order[]; returns how many.  (The device
// ranks them with the whole warp, deflate.cu: the order is total, so any correct sort gives this result.)
BIODB_HD uint32_t defl_sort_symbols(const uint16_t* freq, uint32_t n, uint16_t* order) {
  uint32_t used = 0;
  for (uint32_t k = 0; k < n; ++k)
    if (freq[k]) order[used++] = (uint16_t)k;
  for (uint32_t i = 1; i < used; ++i) {
    const uint16_t v = order[i];
    uint32_t j = i;
    while (j > 0 && freq[order[j - 1]] > freq[v]) { order[j] = order[j - 1]; --j; }
    order[j] = v;
  }
  return used;
}

// Code lengths (at most max_len bits) of a Huffman code for freq[0, n), 2 <= n <= 286, whose `used` symbols of non-zero
// frequency stand sorted in w->order (defl_sort_symbols): symbols never used get 0 — except that fewer than two used
// symbols are topped up to two codes of one bit each (as zlib does: every decoder accepts a complete code, not every
// one an empty or a one-code set).  The tree comes from repeatedly joining the two lightest nodes; lengths beyond
// max_len are folded back by moving codes down the length histogram until Kraft's sum is 1 again, and the lengths are
// then dealt out in order of frequency.
BIODB_HD void defl_code_lengths_sorted(const uint16_t* freq, uint32_t n, uint32_t max_len, uint8_t* len, DeflWork* w, uint32_t used) {
  uint32_t* weight = w->weight;
  uint16_t* parent = w->parent;
  const uint16_t* order = w->order;
  for (uint32_t k = 0; k < n; ++k) len[k] = 0;
  if (used < 2) {
    const uint32_t first = used ? order[0] : 0;
    len[first] = 1;
    len[first == 0 ? 1 : 0] = 1;
    return;
  }
  // leaves in order of rising weight, internal nodes after them; the nodes a join creates come out in rising weight
  // too, so the two lightest roots are always at the front of one of the two queues (no search)
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
  // depth of every node from the root down (a parent stands behind its children): weight[] becomes the depth
  uint32_t hist[33];
  for (int k = 0; k < 33; ++k) hist[k] = 0;
  weight[qn - 1] = 0;
  for (uint32_t x = qn - 1; x-- > 0;) {
    const uint32_t d = weight[parent[x]] + 1;
    weight[x] = d;
    if (x < used) ++hist[d > 32 ? 32 : d];
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

BIODB_HD void defl_code_lengths(const uint16_t* freq, uint32_t n, uint32_t max_len, uint8_t* len, DeflWork* w) {
  defl_code_lengths_sorted(freq, n, max_len, len, w, defl_sort_symbols(freq, n, w->order));
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

// After ll_len / dd_len are known: the code lengths themselves run-length coded with the symbols 16 / 17 / 18 (RFC 1951
// §3.2.7), their code, and the size in bits of the dynamic and of the fixed form of the block.  Returns true if the
// dynamic form is the shorter one; *bits = the size of the chosen form.
BIODB_HD bool defl_plan(DeflWork* w, uint64_t* bits) {
  const DeflCount& cnt = w->cnt;
  uint8_t* ll_len = w->ll_len;
  uint8_t* dd_len = w->dd_len;
  uint32_t n_ll = 286, n_dd = 30;
  while (n_ll > 257 && ll_len[n_ll - 1] == 0) --n_ll;
  while (n_dd > 1 && dd_len[n_dd - 1] == 0) --n_dd;
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
  defl_code_lengths(cl_freq, 19, 7, cl_len, w);
  const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
  uint32_t n_clc = 19;
  while (n_clc > 4 && cl_len[cl_order[n_clc - 1]] == 0) --n_clc;
  w->n_ll = n_ll; w->n_dd = n_dd; w->n_cl = n_cl; w->n_clc = n_clc;
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
  *bits = dynamic ? dyn_bits : fix_bits;
  return dynamic;
}

// The block header (BFINAL, BTYPE and, for the dynamic form, the code tables) and the codes the tokens are written
// with: w->ll_code / ll_len / dd_code / dd_len hold the chosen code afterwards (the fixed one of §3.2.6 if !dynamic).
BIODB_HD void defl_write_header(DeflBits& b, DeflWork* w, bool dynamic) {
  b.put(1, 1);                                              // BFINAL
  if (dynamic) {
    const uint8_t cl_order[19] = {16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15};
    defl_make_codes(w->cl_len, 19, 7, w->cl_code);
    b.put(2, 2);                                            // BTYPE = 10 (dynamic Huffman)
    b.put(w->n_ll - 257, 5);
    b.put(w->n_dd - 1, 5);
    b.put(w->n_clc - 4, 4);
    for (uint32_t k = 0; k < w->n_clc; ++k) b.put(w->cl_len[cl_order[k]], 3);
    for (uint32_t k = 0; k < w->n_cl; ++k) {
      const uint32_t sy = w->cl_sym[k];
      b.put(w->cl_code[sy], w->cl_len[sy]);
      if (sy == 16) b.put(w->cl_ext[k], 2);
      else if (sy == 17) b.put(w->cl_ext[k], 3);
      else if (sy == 18) b.put(w->cl_ext[k], 7);
    }
    w->ll_len[286] = w->ll_len[287] = 0;
  } else {
    b.put(1, 2);                                            // BTYPE = 01 (fixed Huffman)
    for (uint32_t k = 0; k < 288; ++k) w->ll_len[k] = (uint8_t)(k < 144 ? 8 : k < 256 ? 9 : k < 280 ? 7 : 8);
    for (uint32_t k = 0; k < 30; ++k) w->dd_len[k] = 5;
  }
  // (the fixed code of §3.2.6 is the canonical code of its lengths over 288 / 30 symbols)
  defl_make_codes(w->ll_len, 288, 15, w->ll_code);
  defl_make_codes(w->dd_len, 30, 15, w->dd_code);
}

// One token as bits: the value (LSB first, at most 48 bits) and its length
BIODB_HD uint64_t defl_token_bits(uint32_t t, const uint16_t* ll_code, const uint8_t* ll_len, const uint16_t* dd_code,
                                  const uint8_t* dd_len, uint32_t* nbits) {
  if (!(t & (DEFL_TOK_MATCH | DEFL_TOK_EOB))) { *nbits = ll_len[t]; return ll_code[t]; }
  if (t & DEFL_TOK_EOB) { *nbits = ll_len[256]; return ll_code[256]; }
  uint32_t i, eb, ev;
  defl_len_code(((t >> 16) & 0xff) + 3, &i, &eb, &ev);
  uint64_t v = ll_code[257 + i];
  uint32_t nb = ll_len[257 + i];
  v |= (uint64_t)ev << nb;
  nb += eb;
  defl_dist_code((t & 0x7fff) + 1, &i, &eb, &ev);
  v |= (uint64_t)dd_code[i] << nb;
  nb += dd_len[i];
  v |= (uint64_t)ev << nb;
  nb += eb;
  *nbits = nb;
  return v;
}

// ---- the host's restatement of the warp's LZ77 parse: same windows, same candidates, same tokens -----------------------
// tok: room for n + 1 tokens; htab: DEFL_HASH_SIZE * DEFL_WAYS entries.  Returns the number of tokens (the last one
// is DEFL_TOK_EOB); cnt receives the statistics.
inline uint32_t defl_parse_host(const uint8_t* in, uint32_t n, uint16_t* htab, uint32_t* tok, DeflCount& cnt) {
  for (uint32_t k = 0; k < DEFL_HASH_SIZE * DEFL_WAYS; ++k) htab[k] = (uint16_t)DEFL_NONE;
  cnt.init();
  auto first4 = [&](uint32_t p) { return (uint32_t)in[p] | ((uint32_t)in[p + 1] << 8) | ((uint32_t)in[p + 2] << 16) | ((uint32_t)in[p + 3] << 24); };
  auto insert = [&](uint32_t h, uint32_t p) {               // the bucket keeps its two latest positions
    htab[h * 2 + 1] = htab[h * 2];
    htab[h * 2] = (uint16_t)p;
  };
  uint32_t nt = 0, i = 0;
  while (i < n) {
    uint32_t len[DEFL_WIN], dist[DEFL_WIN], h[DEFL_WIN], cand[DEFL_WIN][3];
    bool valid[DEFL_WIN];
    for (uint32_t l = 0; l < DEFL_WIN; ++l) {
      const uint32_t p = i + l;
      valid[l] = p + 4 <= n;
      len[l] = dist[l] = 0;
      cand[l][0] = cand[l][1] = cand[l][2] = DEFL_NONE;
      if (!valid[l]) continue;
      h[l] = defl_hash(first4(p));
      cand[l][1] = htab[h[l] * 2];                          // as the table stood before this window
      cand[l][2] = htab[h[l] * 2 + 1];
      for (uint32_t k = l; k-- > 0;)
        if (valid[k] && h[k] == h[l]) { cand[l][0] = i + k; break; }
    }
    for (uint32_t l = 0; l < DEFL_WIN; ++l)
      if (valid[l]) insert(h[l], i + l);
    for (uint32_t l = 0; l < DEFL_WIN; ++l) {
      if (!valid[l]) continue;
      const uint32_t p = i + l, maxl = n - p < DEFL_WCAP ? n - p : DEFL_WCAP;
      for (int c = 0; c < 3; ++c) {
        const uint32_t q = cand[l][c];
        if (q == DEFL_NONE || p - q > 32768) continue;      // (the table may remember positions beyond DEFLATE's reach)
        uint32_t m = 0;
        while (m < maxl && in[q + m] == in[p + m]) ++m;
        if (m > len[l]) { len[l] = m; dist[l] = p - q; }
      }
      if (len[l] < DEFL_MIN_MATCH) len[l] = 0;
    }
    uint32_t l = 0;
    while (l < DEFL_WIN && i + l < n) {
      const bool lazy = l + 1 < DEFL_WIN && len[l + 1] > len[l];
      if (len[l] && !lazy) {
        uint32_t L = len[l];
        const uint32_t p = i + l, q = p - dist[l], maxl = n - p < 258 ? n - p : 258;
        if (L == DEFL_WCAP)
          while (L < maxl && in[q + L] == in[p + L]) ++L;
        tok[nt++] = DEFL_TOK_MATCH | ((L - 3) << 16) | (dist[l] - 1);
        cnt.match(L, dist[l]);
        l += L;
      } else {
        tok[nt++] = in[i + l];
        cnt.literal(in[i + l]);
        l += 1;
      }
    }
    const uint32_t exitp = i + l;
    for (uint32_t p = i + DEFL_WIN; p < exitp && p + 4 <= n; ++p) insert(defl_hash(first4(p)), p);   // what a long match skipped
    i = exitp;
  }
  tok[nt++] = DEFL_TOK_EOB;
  return nt;
}

// Raw DEFLATE of in[0, n) into out[0, cap) on the host — byte for byte what deflate_warp_kernel writes.  level 0 stores;
// every other level: one final block with dynamic or fixed Huffman codes, whichever is shorter, or stored after all if
// that beats both.  Returns the number of bytes written, 0 if cap is too small (cap >= n + 5 always suffices).
inline uint32_t deflate_block_host(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, int level) {
  if (n > DEFL_MAX_IN) return 0;
  if (level == 0 || n < 8) return deflate_stored(in, n, out, cap);
  uint16_t* htab = new uint16_t[DEFL_HASH_SIZE * DEFL_WAYS];
  uint32_t* tok = new uint32_t[(size_t)n + 1];
  DeflWork* w = new DeflWork;
  const uint32_t nt = defl_parse_host(in, n, htab, tok, w->cnt);
  defl_code_lengths(w->cnt.ll, 286, 15, w->ll_len, w);
  defl_code_lengths(w->cnt.dd, 30, 15, w->dd_len, w);
  uint64_t bits = 0;
  const bool dynamic = defl_plan(w, &bits);
  const uint32_t size = (uint32_t)((bits + 7) / 8);
  uint32_t ret;
  if (size >= n + 5 || size > cap) {
    ret = deflate_stored(in, n, out, cap);
  } else {
    DeflBits b{out, 0, cap, 0, 0};
    defl_write_header(b, w, dynamic);
    for (uint32_t k = 0; k < nt; ++k) {
      uint32_t nb;
      const uint64_t v = defl_token_bits(tok[k], w->ll_code, w->ll_len, w->dd_code, w->dd_len, &nb);
      b.put((uint32_t)(v & 0xffff), nb < 16 ? nb : 16);
      if (nb > 16) b.put((uint32_t)((v >> 16) & 0xffff), nb - 16 < 16 ? nb - 16 : 16);
      if (nb > 32) b.put((uint32_t)(v >> 32), nb - 32);
    }
    b.flush();
    ret = b.n;                                               // == size
  }
  delete w;
  delete[] tok;
  delete[] htab;
  return ret;
}

}  // namespace biodb
