// Deterministic synthetic coordinate-sorted BAM generator (host C++, multi-threaded, libz).
//
// Input generator for bench.py and the large parity tests — not part of the decode path.  It
// writes files the way BioD's own writer does: "BAM\1" + l_text + text + n_ref + refs
// (bam/writer.d:203-236), records never straddling a BGZF block when they fit
// (writer.d:259-267), <= 0xFF00 payload bytes per block (bgzf/constants.d:61), 18-byte header
// BLOCK_HEADER_START + BSIZE (constants.d:28-35), CRC32 + ISIZE footer (bgzf/compress.d:100-101),
// raw deflate with deflateInit2(level, -15, 8) (compress.d:72-76) and the 28-byte EOF block
// (constants.d:38-49).  Record recipe: SURVEY.md §8(d).
//
// Reads are generated in independent units of UNIT reads so that threads can work in parallel;
// every unit ends its last block (a valid, slightly shorter block).  Everything about read i is
// a pure function of (seed, i), so the output does not depend on the thread count.
#include <zlib.h>

#include <algorithm>
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

namespace {

constexpr uint32_t UNIT = 4096;
constexpr uint32_t BLOCK_PAYLOAD = 0xFF00;
constexpr int READ_LEN = 150;

inline uint64_t splitmix(uint64_t& s) {
  uint64_t z = (s += 0x9E3779B97F4A7C15ull);
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
inline uint64_t hash2(uint64_t a, uint64_t b) {
  uint64_t s = a * 0xD6E8FEB86659FD93ull + b;
  return splitmix(s);
}
struct Rng {   // xoshiro256**
  uint64_t s[4];
  explicit Rng(uint64_t seed) { for (auto& x : s) x = splitmix(seed); }
  static inline uint64_t rotl(uint64_t x, int k) { return (x << k) | (x >> (64 - k)); }
  uint64_t next() {
    uint64_t r = rotl(s[1] * 5, 7) * 9, t = s[1] << 17;
    s[2] ^= s[0]; s[3] ^= s[1]; s[1] ^= s[2]; s[0] ^= s[3]; s[2] ^= t; s[3] = rotl(s[3], 45);
    return r;
  }
  uint32_t below(uint32_t n) { return (uint32_t)(((next() >> 32) * (uint64_t)n) >> 32); }
  double uniform() { return (next() >> 11) * (1.0 / 9007199254740992.0); }
};

struct Params {
  uint64_t n_reads;
  uint32_t n_refs;
  int mixed;
  int level;
  uint64_t seed;
  uint64_t reads_per_ref;
};

// start-gap of read i: Geometric(p=0.2) on {0,1,...} shifted so the mean gap is 5 bp -> 30x at 150 bp
inline uint32_t gap_of(const Params& P, uint64_t i) {
  uint64_t h = hash2(P.seed ^ 0x6761707300ull, i);
  double u = ((h >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  // number of failures before the first success, p = 0.2: floor(ln(u)/ln(0.8)); mean 4, +1 -> mean 5
  uint32_t g = (uint32_t)(__builtin_log(u) / -0.2231435513142097);
  return 1 + std::min<uint32_t>(g, 200);
}
inline uint32_t genome_base(const Params& P, uint32_t ref, int64_t p) {   // 0..3
  uint64_t h = hash2(P.seed ^ 0x67656E6F6D65ull ^ ((uint64_t)ref << 48), (uint64_t)(p >> 5));
  return (uint32_t)(h >> (2 * (p & 31))) & 3;
}
const char BASES[] = "ACGT";
const uint8_t CODE[] = {1, 2, 4, 8};   // A C G T in BAM 4-bit codes (bio/core/base.d:85)

inline uint16_t reg2bin(int32_t beg, int32_t end) {   // bam/bai/bin.d:82-92
  if (end == beg) end = beg + 1;
  --end;
  if (beg >> 14 == end >> 14) return (uint16_t)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint16_t)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint16_t)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint16_t)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint16_t)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

inline void put32(std::vector<uint8_t>& v, uint32_t x) {
  v.push_back((uint8_t)x); v.push_back((uint8_t)(x >> 8)); v.push_back((uint8_t)(x >> 16)); v.push_back((uint8_t)(x >> 24));
}

// one alignment record (with its block_size prefix) appended to `out`
void make_record(const Params& P, uint64_t i, uint32_t ref, int32_t pos, std::vector<uint8_t>& out) {
  Rng rng(hash2(P.seed, i));
  // CIGAR plan: [S] M [I|D|N M] [S]
  uint32_t ops[5];
  int nops = 0;
  int sl = 0, sr = 0;          // soft clips
  int ev = 0, evlen = 0, evoff = 0;   // 0 none, 1 I, 2 D, 3 N
  if (P.mixed) {
    if (rng.uniform() < 0.05) {
      ev = 1 + (int)rng.below(3);
      evlen = ev == 3 ? 100 + (int)rng.below(901) : 1 + (int)rng.below(10);
    }
    if (rng.uniform() < 0.10) {
      int s = 1 + (int)rng.below(20);
      if (rng.below(2)) sl = s; else sr = s;
    }
  }
  const int aligned_q = READ_LEN - sl - sr - (ev == 1 ? evlen : 0);   // query bases in M ops
  if (ev) evoff = 10 + (int)rng.below((uint32_t)(aligned_q - 20));   // first M length
  auto op = [](uint32_t len, uint32_t code) { return (len << 4) | code; };
  if (sl) ops[nops++] = op(sl, 4);
  if (!ev) ops[nops++] = op(aligned_q, 0);
  else {
    ops[nops++] = op(evoff, 0);
    ops[nops++] = op(evlen, ev == 1 ? 1 : ev == 2 ? 2 : 3);
    ops[nops++] = op(aligned_q - evoff, 0);
  }
  if (sr) ops[nops++] = op(sr, 4);
  const int ref_span = aligned_q + (ev >= 2 ? evlen : 0);
  // sequence, MD
  uint8_t seq[READ_LEN];
  std::string md;
  int q = 0, run = 0;
  int64_t rp = pos;
  auto put_match = [&](int n) {
    for (int k = 0; k < n; ++k, ++q, ++rp) {
      uint32_t rb = genome_base(P, ref, rp), b = rb;
      if (rng.below(200) == 0) b = (rb + 1 + rng.below(3)) & 3;
      seq[q] = (uint8_t)b;
      if (b == rb) ++run;
      else { md += std::to_string(run); md += BASES[rb]; run = 0; }
    }
  };
  for (int k = 0; k < sl; ++k) seq[q++] = (uint8_t)rng.below(4);
  if (!ev) put_match(aligned_q);
  else {
    put_match(evoff);
    if (ev == 1) { for (int k = 0; k < evlen; ++k) seq[q++] = (uint8_t)rng.below(4); }
    else if (ev == 2) {
      md += std::to_string(run); run = 0; md += '^';
      for (int k = 0; k < evlen; ++k, ++rp) md += BASES[genome_base(P, ref, rp)];
    } else rp += evlen;
    put_match(aligned_q - evoff);
  }
  for (int k = 0; k < sr; ++k) seq[q++] = (uint8_t)rng.below(4);
  md += std::to_string(run);
  char name[16];
  int lname = snprintf(name, sizeof name, "r%09llu", (unsigned long long)i) + 1;
  const uint32_t flag = rng.below(2) ? 16 : 0;
  const uint32_t body = 32 + lname + 4 * nops + (READ_LEN + 1) / 2 + READ_LEN + 3 + (uint32_t)md.size() + 1;
  put32(out, body);
  put32(out, ref);
  put32(out, (uint32_t)pos);
  put32(out, ((uint32_t)reg2bin(pos, pos + ref_span) << 16) | (60u << 8) | (uint32_t)lname);
  put32(out, (flag << 16) | (uint32_t)nops);
  put32(out, READ_LEN);
  put32(out, 0xFFFFFFFFu);
  put32(out, 0xFFFFFFFFu);
  put32(out, 0);
  out.insert(out.end(), name, name + lname);
  for (int k = 0; k < nops; ++k) put32(out, ops[k]);
  for (int k = 0; k < READ_LEN; k += 2) out.push_back((uint8_t)((CODE[seq[k]] << 4) | (k + 1 < READ_LEN ? CODE[seq[k + 1]] : 0)));
  for (int k = 0; k < READ_LEN; ++k) out.push_back((uint8_t)(2 + rng.below(40)));
  out.push_back('M'); out.push_back('D'); out.push_back('Z');
  out.insert(out.end(), md.begin(), md.end());
  out.push_back(0);
}

void bgzf_block(const uint8_t* payload, uint32_t n, int level, std::vector<uint8_t>& out) {
  static const uint8_t HDR[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};
  size_t at = out.size();
  out.resize(at + 18 + compressBound(n) + 64 + 8);
  memcpy(&out[at], HDR, 16);
  z_stream zs;
  memset(&zs, 0, sizeof zs);
  deflateInit2(&zs, level, Z_DEFLATED, -15, 8, Z_DEFAULT_STRATEGY);
  zs.next_in = const_cast<Bytef*>(payload);
  zs.avail_in = n;
  zs.next_out = &out[at + 18];
  zs.avail_out = (uInt)(out.size() - at - 18 - 8);
  deflate(&zs, Z_FINISH);
  uint32_t c = (uint32_t)zs.total_out;
  deflateEnd(&zs);
  uint32_t bsize = c + 25;
  out[at + 16] = (uint8_t)bsize;
  out[at + 17] = (uint8_t)(bsize >> 8);
  uint32_t crc = (uint32_t)crc32(crc32(0, nullptr, 0), payload, n);
  uint8_t* f = &out[at + 18 + c];
  for (int k = 0; k < 4; ++k) { f[k] = (uint8_t)(crc >> (8 * k)); f[4 + k] = (uint8_t)(n >> (8 * k)); }
  out.resize(at + 18 + c + 8);
}

void pack_blocks(const std::vector<uint8_t>& raw, const std::vector<uint32_t>& rec_end, int level, std::vector<uint8_t>& out,
                 bool straddle = false) {
  // records never straddle a block when they fit (writer.d:259-267) — unless the htsjdk-style layout is asked for,
  // which fills every block to 0xFF00 bytes wherever that cuts a record
  size_t start = 0, last = 0;
  for (size_t k = 0; k < rec_end.size() && !straddle; ++k) {
    if (rec_end[k] - start > BLOCK_PAYLOAD && last > start) {
      bgzf_block(raw.data() + start, (uint32_t)(last - start), level, out);
      start = last;
    }
    last = rec_end[k];
  }
  while (raw.size() - start > BLOCK_PAYLOAD) {
    bgzf_block(raw.data() + start, BLOCK_PAYLOAD, level, out);
    start += BLOCK_PAYLOAD;
  }
  if (raw.size() > start) bgzf_block(raw.data() + start, (uint32_t)(raw.size() - start), level, out);
}

}  // namespace

extern "C" {

// Upper bound of the output size, to size the caller's buffer.
uint64_t bamgen_bound(uint64_t n_reads, int mixed) {
  return (uint64_t)((double)n_reads * (mixed ? 330.0 : 300.0) * 0.62) + (64u << 20);
}

// Writes a complete BAM into out[0..cap).  Returns the size, or 0 if cap was too small.
// ref_len_out (may be null) receives the common contig length.
uint64_t bamgen_generate(uint64_t n_reads, uint32_t n_refs, int mixed, int level, uint64_t seed, int threads, uint8_t* out,
                         uint64_t cap, uint64_t* ref_len_out) {
  Params P{n_reads, n_refs ? n_refs : 1, mixed & 1, level, seed, 0};
  P.reads_per_ref = (n_reads + P.n_refs - 1) / P.n_refs;
  // units never cross a reference
  struct Unit { uint64_t first; uint32_t n; uint32_t ref; int64_t pos0; };
  std::vector<Unit> units;
  for (uint32_t r = 0; r < P.n_refs; ++r) {
    uint64_t a = (uint64_t)r * P.reads_per_ref, b = std::min<uint64_t>(n_reads, a + P.reads_per_ref);
    for (uint64_t f = a; f < b; f += UNIT) units.push_back(Unit{f, (uint32_t)std::min<uint64_t>(UNIT, b - f), r, 0});
  }
  if (threads < 1) threads = 1;
  // pass 1: sum of start gaps per unit -> first position of every unit
  std::vector<uint64_t> gsum(units.size());
  {
    std::atomic<size_t> nxt{0};
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
      th.emplace_back([&] {
        for (size_t u; (u = nxt.fetch_add(1)) < units.size();) {
          uint64_t s = 0;
          for (uint32_t k = 0; k < units[u].n; ++k) s += gap_of(P, units[u].first + k);
          gsum[u] = s;
        }
      });
    for (auto& x : th) x.join();
  }
  int64_t max_end = 0;
  {
    int64_t p = 0;
    uint32_t cur = 0;
    for (size_t u = 0; u < units.size(); ++u) {
      if (units[u].ref != cur) { cur = units[u].ref; p = 0; }
      units[u].pos0 = p;
      p += (int64_t)gsum[u];
      max_end = std::max(max_end, p + 2000);
    }
  }
  const uint64_t ref_len = (uint64_t)max_end;
  if (ref_len_out) *ref_len_out = ref_len;
  // header
  std::vector<uint8_t> hdr;
  {
    std::string text = "@HD\tVN:1.6\tSO:coordinate\n";
    for (uint32_t r = 0; r < P.n_refs; ++r) text += "@SQ\tSN:chr" + std::to_string(r + 1) + "\tLN:" + std::to_string(ref_len) + "\n";
    std::vector<uint8_t> raw = {'B', 'A', 'M', 1};
    put32(raw, (uint32_t)text.size());
    raw.insert(raw.end(), text.begin(), text.end());
    put32(raw, P.n_refs);
    for (uint32_t r = 0; r < P.n_refs; ++r) {
      std::string nm = "chr" + std::to_string(r + 1);
      put32(raw, (uint32_t)nm.size() + 1);
      raw.insert(raw.end(), nm.begin(), nm.end());
      raw.push_back(0);
      put32(raw, (uint32_t)ref_len);
    }
    std::vector<uint32_t> ends = {(uint32_t)raw.size()};
    pack_blocks(raw, ends, level, hdr);
  }
  if (hdr.size() > cap) return 0;
  memcpy(out, hdr.data(), hdr.size());
  // pass 2: units in waves of `threads * 4`, written in order
  uint64_t w = hdr.size();
  const size_t wave = (size_t)threads * 8;
  std::vector<std::vector<uint8_t>> comp(wave);
  for (size_t base = 0; base < units.size(); base += wave) {
    const size_t cnt = std::min(wave, units.size() - base);
    std::atomic<size_t> nxt{0};
    std::vector<std::thread> th;
    for (int t = 0; t < threads; ++t)
      th.emplace_back([&] {
        std::vector<uint8_t> raw;
        std::vector<uint32_t> ends;
        for (size_t k; (k = nxt.fetch_add(1)) < cnt;) {
          const Unit& un = units[base + k];
          raw.clear();
          ends.clear();
          int64_t p = un.pos0;
          for (uint32_t j = 0; j < un.n; ++j) {
            p += gap_of(P, un.first + j);
            make_record(P, un.first + j, un.ref, (int32_t)p, raw);
            ends.push_back((uint32_t)raw.size());
          }
          comp[k].clear();
          pack_blocks(raw, ends, level, comp[k], (mixed & 2) != 0);
        }
      });
    for (auto& x : th) x.join();
    for (size_t k = 0; k < cnt; ++k) {
      if (w + comp[k].size() + 28 > cap) return 0;
      memcpy(out + w, comp[k].data(), comp[k].size());
      w += comp[k].size();
    }
  }
  static const uint8_t EOFB[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (w + 28 > cap) return 0;
  memcpy(out + w, EOFB, 28);
  return w + 28;
}

}  // extern "C"
