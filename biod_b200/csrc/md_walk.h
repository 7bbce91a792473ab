// dna(read): the reference bases a read spans, rebuilt from its SEQ, CIGAR and MD tag.
//
// Reference behaviour restated here (none of it copied): BamRead["MD"] tag lookup (bam/read.d:1070-1087,1219-1230,
// tagvalue.d:106-119), mdOperations (bam/md/parse.d:13-143, a bidirectional range whose LAST operation is parsed from
// the back of the string at construction, with zero-length matches filtered from both ends), dna() (bam/md/
// reconstruct.d:38-214) and Base16 normalisation of MD characters (bio/core/base.d:39-85).
//
// DnaWalk is a generator with O(1) state that works directly on the raw record bytes — no strings, no allocation —
// so the same code runs in a CUDA thread (csrc/mdtag.cu) and on the host (biodb_debug_md_dna, used by the CPU tests to
// check it against the test suite's CPU restatement before any GPU is involved).
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BIODB_HD __host__ __device__ __forceinline__
#else
#define BIODB_HD inline
#endif

namespace biodb {

// Base16(char) -> char (base.d:39-62 _char2code, :85 _code2char): IUPAC letters in either case, '=' and the digits
// '0'..'3' (A C G T) keep a meaning; everything else is 'N'.
BIODB_HD uint8_t md_norm16(uint8_t ch) {
  if (ch >= 128) return 'N';
  if (ch < 64) {
    switch (ch) {
      case '=': return '=';
      case '0': return 'A';
      case '1': return 'C';
      case '2': return 'G';
      case '3': return 'T';
      default: return 'N';
    }
  }
  switch (ch | 0x20) {
    case 'a': return 'A'; case 'c': return 'C'; case 'm': return 'M'; case 'g': return 'G';
    case 'r': return 'R'; case 's': return 'S'; case 'v': return 'V'; case 't': return 'T';
    case 'w': return 'W'; case 'y': return 'Y'; case 'h': return 'H'; case 'k': return 'K';
    case 'd': return 'D'; case 'b': return 'B';
    default: return 'N';
  }
}

BIODB_HD uint32_t md_le32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

struct MdOpV {           // md/operation.d: 0 Match, 1 Mismatch, 2 Deletion
  int type;
  uint32_t match;
  uint8_t mismatch;
  const uint8_t* del;
  uint32_t del_len;
};

// mdOperations over the bytes [lo, hi) of the tag value
struct MdOpRange {
  const uint8_t* lo;
  const uint8_t* hi;
  MdOpV f, b;
  uint32_t rem;          // 255: front and back are distinct cached operations; 1: one operation left; 0: empty

  static BIODB_HD bool up(uint8_t c) { return c >= 'A' && c <= 'Z'; }
  static BIODB_HD bool dig(uint8_t c) { return c >= '0' && c <= '9'; }
  static BIODB_HD uint32_t to_uint(const uint8_t* p, uint32_t n) {     // saturating (restatement-defined)
    uint64_t v = 0;
    for (uint32_t i = 0; i < n; ++i) {
      v = v * 10 + (uint64_t)(p[i] - '0');
      if (v > 0xffffffffull) v = 0xffffffffull;
    }
    return (uint32_t)v;
  }
  static BIODB_HD bool zero_match(const MdOpV& o) { return o.type == 0 && o.match == 0; }

  BIODB_HD bool cache_front() {
    if (lo >= hi) return false;
    const uint8_t c = *lo;
    if (c == '^') {
      ++lo;
      uint32_t len = 0;
      while (lo + len < hi && up(lo[len])) ++len;
      f = MdOpV{2, 0, 0, lo, len};
      lo += len;
    } else if (dig(c)) {
      uint32_t len = 0;
      while (lo + len < hi && dig(lo[len])) ++len;
      f = MdOpV{0, to_uint(lo, len), 0, nullptr, 0};
      lo += len;
    } else {
      f = MdOpV{1, 0, c, nullptr, 0};
      ++lo;
    }
    return true;
  }
  BIODB_HD bool cache_back() {
    if (lo >= hi) return false;
    const uint32_t n = (uint32_t)(hi - lo);
    if (dig(hi[-1])) {
      uint32_t len = 0;
      while (len < n && dig(hi[-1 - (int32_t)len])) ++len;
      b = MdOpV{0, to_uint(hi - len, len), 0, nullptr, 0};
      hi -= len;
    } else if (n == 1 || dig(hi[-2])) {
      b = MdOpV{1, 0, hi[-1], nullptr, 0};
      hi -= 1;
    } else {
      uint32_t len = 0;                          // back to the '^' (no '^': the whole rest is the deletion)
      while (len < n && hi[-1 - (int32_t)len] != '^') ++len;
      b = MdOpV{2, 0, 0, hi - len, len};
      hi = lo + (len < n ? n - len - 1 : 0);
    }
    return true;
  }
  BIODB_HD bool empty() const { return rem == 0; }
  BIODB_HD void pop_front() {
    if (lo >= hi) { if (rem == 255) { f = b; rem = 1; } else rem = 0; }
    else if (!cache_front()) rem = 0;
  }
  BIODB_HD void pop_back() {
    if (lo >= hi) { if (rem == 255) { b = f; rem = 1; } else rem = 0; }
    else if (!cache_back()) rem = 0;
  }
  BIODB_HD void init(const uint8_t* a, const uint8_t* z) {
    lo = a;
    hi = z;
    rem = 255;
    f = MdOpV{0, 0, 0, nullptr, 0};
    b = f;
    if (!cache_front()) rem = 0;
    else if (!cache_back()) { b = f; rem = 1; }
    while (!empty() && zero_match(f)) pop_front();         // filterBidirectional, at construction
    while (!empty() && zero_match(b)) pop_back();
  }
  // next operation of the forward iteration
  BIODB_HD bool take(MdOpV* out) {
    if (empty()) return false;
    *out = f;
    do { pop_front(); } while (!empty() && zero_match(f));
    return true;
  }
};

// The MD tag's value inside the tag area [t, t + n): false when there is none (or the area is malformed before it).
BIODB_HD bool md_find_tag(const uint8_t* t, uint64_t n, const uint8_t** vb, const uint8_t** ve) {
  if (n < 4) return false;
  uint64_t o = 0;
  while (o + 1 < n) {
    const bool hit = t[o] == 'M' && t[o + 1] == 'D';
    o += 2;
    if (o >= n) return false;
    const uint8_t type = t[o++];
    int es;
    switch (type) {
      case 'A': case 'c': case 'C': es = 1; break;
      case 's': case 'S': es = 2; break;
      case 'i': case 'I': case 'f': es = 4; break;
      default: es = -1; break;
    }
    if (type == 'Z' || type == 'H') {
      const uint64_t b = o;
      while (o < n && t[o] != 0) ++o;
      if (o >= n) return false;
      if (hit) {
        if (type != 'Z') return false;
        *vb = t + b;
        *ve = t + o;
        return true;
      }
      ++o;
    } else if (type == 'B') {
      if (o + 5 > n) return false;
      int bs;
      switch (t[o]) {
        case 'A': case 'c': case 'C': bs = 1; break;
        case 's': case 'S': bs = 2; break;
        case 'i': case 'I': case 'f': bs = 4; break;
        default: bs = -1; break;
      }
      if (bs < 0) return false;
      const uint64_t cnt = md_le32(t + o + 1);
      o += 5 + (uint64_t)bs * cnt;
      if (hit || o > n) return false;
    } else {
      if (es < 0 || hit) return false;
      o += (uint64_t)es;
    }
  }
  return false;
}

// Generator of dna(read).  `body` points at the record's refID field (the byte after block_size).
struct DnaWalk {
  // query bases of the M / = / X operations, capped by l_seq
  const uint8_t* cg;
  const uint8_t* seq;
  uint32_t nc, k;
  int64_t lseq, qoff, qi, qend;
  // MD operations
  MdOpRange ops;
  MdOpV cur;
  uint32_t di;
  bool done;

  BIODB_HD void init(const uint8_t* body, int64_t block_size) {
    k = 0;
    nc = 0;
    lseq = qoff = qi = qend = 0;
    di = 0;
    done = true;
    cg = seq = nullptr;
    ops.lo = ops.hi = nullptr;
    ops.rem = 0;
    ops.f = ops.b = cur = MdOpV{0, 0, 0, nullptr, 0};
    if (block_size < 32) return;
    const uint32_t lname = body[8];
    nc = (uint32_t)body[12] | ((uint32_t)body[13] << 8);
    lseq = (int32_t)md_le32(body + 16);
    cg = body + 32 + lname;
    seq = cg + 4ull * nc;
    const uint64_t ls = lseq > 0 ? (uint64_t)lseq : 0;
    const uint64_t off = 32ull + lname + 4ull * nc + (ls + 1) / 2 + ls;
    const uint8_t *vb = nullptr, *ve = nullptr;
    if (off > (uint64_t)block_size) return;
    if (!md_find_tag(body + off, (uint64_t)block_size - off, &vb, &ve)) return;
    ops.init(vb, ve);
    done = !ops.take(&cur);
  }
  // next base of the joined M/=/X query chunks, or -1
  BIODB_HD int next_q() {
    while (qi >= qend) {
      if (k >= nc) return -1;
      const uint32_t raw = md_le32(cg + 4ull * k);
      ++k;
      const uint32_t t = (0x3C1A7u >> ((raw & 0xF) * 2)) & 3;       // bit 0: consumes query, bit 1: reference (cigar.d:116)
      if (!(t & 1)) continue;
      const int64_t len = raw >> 4;
      if (t & 2) {
        qi = qoff;
        qend = qoff + len < lseq ? qoff + len : lseq;
      }
      qoff += len;
    }
    const uint8_t byte = seq[qi >> 1];
    const uint32_t code = (qi & 1) ? (byte & 0xF) : (byte >> 4);
    ++qi;
    return "=ACMGRSVTWYHKDBN"[code];
  }
  BIODB_HD void advance() {
    di = 0;
    if (!ops.take(&cur)) done = true;
  }
  // next reference base, or -1 at the end
  BIODB_HD int next() {
    while (!done) {
      if (cur.type == 2) {
        if (di >= cur.del_len) { advance(); continue; }               // (an empty deletion)
        const uint8_t ch = md_norm16(cur.del[di]);
        if (++di >= cur.del_len) advance();
        return ch;
      }
      const int q = next_q();
      if (q < 0) { done = true; return -1; }                          // the query ran out: the sequence ends
      const int ch = cur.type == 0 ? q : (int)md_norm16(cur.mismatch);
      if (cur.type == 1 || --cur.match == 0) advance();
      return ch;
    }
    return -1;
  }
};

}  // namespace biodb
