// Shared pieces of the BGZF inflate kernels (inflate.cu: warp-serial; inflate_tok.cu: lane-parallel decode + resolve):
// shared-memory / mbarrier / TMA primitives, canonical-Huffman table construction (RFC 1951 §3.2.2) and the
// fused record-chain walker.  Nothing here is derived from zlib's source.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "kernels.h"

namespace biodb {
namespace {

#ifndef BIODB_LIT_BITS
#define BIODB_LIT_BITS 10
#endif
constexpr int LIT_BITS = BIODB_LIT_BITS;
constexpr int DIST_BITS = 8;
constexpr int CL_BITS = 7;

// 16-bit LUT entries: [0,8) literal byte / length symbol / distance symbol / code-length symbol,
// [8,10) kind, [12,16) code length.
constexpr uint32_t K_LIT = 0, K_LEN = 1, K_EOB = 2, K_SPECIAL = 3;
constexpr uint32_t ENT_SLOW = K_SPECIAL << 8;           // code longer than the LUT index
constexpr uint32_t ENT_INVALID = (K_SPECIAL << 8) | 1;  // unused code
constexpr int Z_DATA = -3;
constexpr int Z_BUF = -5;

enum { KIND_LITLEN = 0, KIND_DIST = 1, KIND_CODELEN = 2 };

struct Code {               // slow-path side tables of one Huffman code
  uint16_t cnt[16];         // symbols per code length
  uint16_t first[16];       // first canonical code of each length
  uint16_t index[16];       // symbols with shorter codes
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// explicit shared-space accesses with a 32-bit address: keeps ptxas from rebuilding the shared-window base
// (S2R SR_CgaCtaId + LEA) inside the hot loop
__device__ __forceinline__ uint32_t lds16(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ uint32_t lds8(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts8(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u8 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ uint4 lds128(uint32_t addr) {
  uint4 v;
  asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr));
  return v;
}

__device__ __forceinline__ void mbar_init(uint32_t bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok = 0;
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion signalled on the mbarrier.
__device__ __forceinline__ void tma_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

// length / distance symbol -> (base, extra bits), RFC 1951 §3.2.5
__device__ __forceinline__ void len_base(uint32_t s /*0..28*/, uint32_t& base, uint32_t& eb) {
  if (s < 8) { base = 3 + s; eb = 0; }
  else if (s == 28) { base = 258; eb = 0; }
  else { eb = (s - 4) >> 2; base = 3 + ((4u + ((s - 4) & 3)) << eb); }
}
__device__ __forceinline__ void dist_base(uint32_t d /*0..29*/, uint32_t& base, uint32_t& eb) {
  if (d < 4) { base = 1 + d; eb = 0; }
  else { eb = (d - 2) >> 1; base = 1 + ((2u + (d & 1)) << eb); }
}
// number of extra bits of a length / distance symbol, RFC 1951 §3.2.5 (the closed forms of len_base / dist_base)
__device__ __forceinline__ uint32_t len_extra_bits(uint32_t s /*0..28*/) { return (s < 8 || s == 28) ? 0u : (s - 4) >> 2; }
__device__ __forceinline__ uint32_t dist_extra_bits(uint32_t d /*0..29*/) { return d < 4 ? 0u : (d - 2) >> 1; }
// Length and distance entries also carry the symbol's extra-bit count: bits [5,8) its low three bits, bit 10 its
// fourth (distances only) — so that a decoder can advance over a code without a second table lookup
// (ENTRY_EXTRA_BITS).  Readers of the symbol mask it with 31.
__device__ __forceinline__ uint32_t make_entry(int kind, int sym, int cl) {
  if (kind == KIND_LITLEN) {
    if (sym < 256) return ((uint32_t)cl << 12) | (uint32_t)sym;
    if (sym == 256) return ((uint32_t)cl << 12) | (K_EOB << 8);
    if (sym > 285) return ENT_INVALID;            // 286/287 exist only in the fixed code and are invalid
    return ((uint32_t)cl << 12) | (K_LEN << 8) | (len_extra_bits((uint32_t)(sym - 257)) << 5) | (uint32_t)(sym - 257);
  }
  if (kind == KIND_DIST) {
    if (sym > 29) return ENT_INVALID;
    const uint32_t eb = dist_extra_bits((uint32_t)sym);
    return ((uint32_t)cl << 12) | ((eb & 8) << 7) | ((eb & 7) << 5) | (uint32_t)sym;
  }
  return ((uint32_t)cl << 12) | (uint32_t)sym;     // code-length code
}
#define ENTRY_EXTRA_BITS(e) ((((e) >> 5) & 7u) | (((e) >> 7) & 8u))

// Warp-cooperative canonical-Huffman table build (RFC 1951 §3.2.2): lane L owns code length L.
// Returns 0 ok, 1 = empty code (LUT all-invalid), -1 = over-subscribed / incomplete set.
template <int PB>
__device__ __noinline__ int build_table(const uint8_t* lens, int n, uint16_t* lut, uint16_t* sorted, Code* code, int kind,
                                        int lane) {
  const bool owner = lane >= 1 && lane <= 15;
  int mycnt = 0;
  if (owner)
    for (int i = 0; i < n; ++i) mycnt += (lens[i] == lane);
  // Kraft sum in units of 2^-15: > 2^15 over-subscribed, < 2^15 incomplete
  int v = owner ? (mycnt << (15 - lane)) : 0;
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const uint32_t have = __ballot_sync(0xffffffffu, mycnt > 0);
  const int maxlen = have ? 31 - __clz(have) : 0;
  if (v > (1 << 15)) return -1;
  for (int i = lane; i < (1 << PB); i += 32) lut[i] = (uint16_t)ENT_INVALID;
  // first canonical code and symbol index of every length (serial recurrence over 15 lengths, warp-uniform)
  int myfirst = 0, myindex = 0;
  {
    int c = 0, idx = 0;
    for (int L = 1; L <= 15; ++L) {
      int prev = __shfl_sync(0xffffffffu, mycnt, L - 1);   // lane 0 holds 0
      c = (c + prev) << 1;
      idx += prev;
      if (lane == L) { myfirst = c; myindex = idx; }
    }
  }
  if (lane < 16) {
    code->cnt[lane] = (uint16_t)mycnt;
    code->first[lane] = (uint16_t)myfirst;
    code->index[lane] = (uint16_t)myindex;
  }
  __syncwarp();
  if (maxlen == 0) return 1;
  if (v < (1 << 15) && (kind == KIND_CODELEN || maxlen != 1)) return -1;
  if (owner && mycnt) {
    int c = myfirst, slot = myindex;
    const int l = lane;
    for (int sym = 0; sym < n; ++sym) {
      if (lens[sym] != l) continue;
      sorted[slot++] = (uint16_t)sym;
      const uint32_t rev = __brev((uint32_t)c) >> (32 - l);
      ++c;
      if (l <= PB) {
        const uint16_t e = (uint16_t)make_entry(kind, sym, l);
        for (uint32_t i = rev; i < (1u << PB); i += (1u << l)) lut[i] = e;
      } else {
        lut[rev & ((1u << PB) - 1)] = (uint16_t)ENT_SLOW;
      }
    }
  }
  __syncwarp();
  return 0;
}

// Symbol-parallel variant of build_table (same results): lane L takes symbols L, L+32, ...; the rank of a symbol among
// the symbols of equal code length comes from a warp match, so the canonical codes are assigned 32 symbols at a time
// and every lane fills the LUT slots of its own symbols.  scratch: 16 words of shared memory.
// sub (may be null): pool of sub_cap second-level entries.  Codes longer than PB bits that share their first PB bits
// get one sub-table of 2^r entries (r = longest such code - PB); the first-level slot then holds
// K_SPECIAL | r << 12 | offset/2 and the decoder indexes the sub-table with the next r bits.  Canonical codes sorted by
// (length, symbol) are also sorted by value when left-aligned, so the codes of one sub-table are neighbours in
// `sorted`.  Returns 2 when the pool is too small (the caller then gives the block to the warp-serial kernel).
template <int PB>
__device__ __noinline__ int build_table_par(const uint8_t* lens, int n, uint16_t* lut, uint16_t* sorted, Code* code,
                                            int kind, int lane, uint32_t* scratch, uint16_t* sub = nullptr,
                                            int sub_cap = 0) {
  if (lane < 16) scratch[lane] = 0;
  __syncwarp();
  for (int i = lane; i < n; i += 32) {
    const int l = lens[i];
    if (l) atomicAdd(&scratch[l], 1u);
  }
  __syncwarp();
  const bool owner = lane >= 1 && lane <= 15;
  const int mycnt = owner ? (int)scratch[lane] : 0;
  // Kraft sum in units of 2^-15: > 2^15 over-subscribed, < 2^15 incomplete
  int v = owner ? (mycnt << (15 - lane)) : 0;
#pragma unroll
  for (int d = 16; d; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
  const uint32_t have = __ballot_sync(0xffffffffu, mycnt > 0);
  const int maxlen = have ? 31 - __clz(have) : 0;
  if (v > (1 << 15)) return -1;
  int myfirst = 0, myindex = 0;
  {
    int c = 0, idx = 0;
    for (int L = 1; L <= 15; ++L) {
      int prev = __shfl_sync(0xffffffffu, mycnt, L - 1);   // lane 0 holds 0
      c = (c + prev) << 1;
      idx += prev;
      if (lane == L) { myfirst = c; myindex = idx; }
    }
  }
  __syncwarp();
  if (lane < 16) {
    code->cnt[lane] = (uint16_t)mycnt;
    code->first[lane] = (uint16_t)myfirst;
    code->index[lane] = (uint16_t)myindex;
    scratch[lane] = 0;                                     // symbols of this length placed so far
  }
  if (lane < 2) lut[lane] = (uint16_t)ENT_INVALID;        // (an incomplete code leaves slots unused: they inherit this)
  __syncwarp();
  if (maxlen == 0) {
    for (int i = lane; i < (1 << PB); i += 32) lut[i] = (uint16_t)ENT_INVALID;
    __syncwarp();
    return 1;
  }
  if (v < (1 << 15) && (kind == KIND_CODELEN || maxlen != 1)) return -1;
  // canonical order: sorted[index[l] + k] = k-th symbol of code length l
  for (int g = 0; g < n; g += 32) {
    const int sym = g + lane;
    const int l = sym < n ? (int)lens[sym] : 0;
    const uint32_t same = __match_any_sync(0xffffffffu, l);
    const uint32_t rank = (uint32_t)__popc(same & ((1u << lane) - 1));
    const uint32_t placed = l ? scratch[l] : 0;
    __syncwarp();
    if (l) {
      if (rank == 0) scratch[l] = placed + (uint32_t)__popc(same);
      sorted[code->index[l] + placed + rank] = (uint16_t)sym;
    }
    __syncwarp();
  }
  // The LUT, level by level: a code of length k sits at its bit-reversed value (< 2^k); copying [0, 2^k) onto
  // [2^k, 2^(k+1)) then repeats every code of length <= k with its period.  One store per symbol and one pass of
  // doublings (2^PB / 32 warp stores) instead of 2^(PB - k) stores per symbol.
#pragma unroll 1
  for (int k = 1; k <= PB; ++k) {
    const int c0 = code->index[k], cn = code->cnt[k];
    const uint32_t f = code->first[k];
    for (int i = lane; i < cn; i += 32)
      lut[__brev(f + (uint32_t)i) >> (32 - k)] = (uint16_t)make_entry(kind, sorted[c0 + i], k);
    __syncwarp();
    if (k < PB) {
      for (int i = lane; i < (1 << k); i += 32) lut[(1 << k) + i] = lut[i];
      __syncwarp();
    }
  }
  for (int l = PB + 1; l <= maxlen; ++l) {                 // codes longer than the index: marked, resolved elsewhere
    const int cn = code->cnt[l];
    const uint32_t f = code->first[l];
    for (int i = lane; i < cn; i += 32) lut[(__brev(f + (uint32_t)i) >> (32 - l)) & ((1u << PB) - 1)] = (uint16_t)ENT_SLOW;
  }
  __syncwarp();
  if (sub != nullptr && maxlen > PB) {
    const int k_begin = code->index[PB + 1], k_end = code->index[maxlen] + code->cnt[maxlen];
    // first PB bits (most significant first) of the k-th code in sorted order
    auto prefix_of = [&](int k, int& l, uint32_t& c) -> uint32_t {
      l = lens[sorted[k]];
      c = code->first[l] + (uint32_t)(k - code->index[l]);
      return c >> (l - PB);
    };
    uint32_t alloc = 0;
    for (int k0 = k_begin; k0 < k_end; k0 += 32) {        // pass 1: one sub-table per group of equal prefixes
      const int k = k0 + lane;
      int l = 0, l2 = 0;
      uint32_t c = 0, c2 = 0, pre = 0;
      bool is_last = false;
      if (k < k_end) {
        pre = prefix_of(k, l, c);
        is_last = k + 1 == k_end || prefix_of(k + 1, l2, c2) != pre;
      }
      const uint32_t size = is_last ? 1u << (l - PB) : 0;   // lengths grow inside a group: the last code is the longest
      const uint32_t incl = alloc + [&] {
        uint32_t v = size;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const uint32_t u = __shfl_up_sync(0xffffffffu, v, d);
          if (lane >= d) v += u;
        }
        return v;
      }();
      if (is_last && incl <= (uint32_t)sub_cap)
        lut[__brev(pre) >> (32 - PB)] = (uint16_t)((K_SPECIAL << 8) | ((uint32_t)(l - PB) << 12) | ((incl - size) >> 1));
      alloc = __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
    if (alloc > (uint32_t)sub_cap) return 2;              // pool too small: cannot happen for a complete code of <= 286 symbols
    for (int k0 = k_begin; k0 < k_end; k0 += 32) {        // pass 2: every long code fills its slots of its sub-table
      const int k = k0 + lane;
      if (k < k_end) {
        int l;
        uint32_t c;
        const uint32_t pre = prefix_of(k, l, c);
        const uint32_t e = lut[__brev(pre) >> (32 - PB)];
        if (e >> 12) {
          const uint32_t r = e >> 12, off = (e & 0xff) << 1, w = (uint32_t)(l - PB);
          const uint32_t rev = __brev(c & ((1u << w) - 1)) >> (32 - w);
          const uint16_t ent = (uint16_t)make_entry(kind, sorted[k], l);
          for (uint32_t i = rev; i < (1u << r); i += (1u << w)) sub[off + i] = ent;
        }
      }
    }
    __syncwarp();
  }
  return 0;
}

// Canonical decode of a code longer than PB bits, starting from the PB-bit prefix already known not to be a
// complete code (RFC 1951 §3.2.2 code assignment run backwards).
template <int PB>
__device__ __noinline__ uint32_t slow_decode(uint32_t bits, const Code* code, const uint16_t* sorted, int kind) {
  uint32_t c = __brev(bits) >> (32 - PB);     // first PB bits of the code, most significant first
  bits >>= PB;
  for (int len = PB + 1; len <= 15; ++len) {
    c = (c << 1) | (bits & 1);
    bits >>= 1;
    const uint32_t rel = c - code->first[len];
    if (rel < code->cnt[len]) return make_entry(kind, sorted[code->index[len] + rel], len);
  }
  return ENT_INVALID;
}

__device__ __forceinline__ void sts16(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u16 [%0], %1;" ::"r"(addr), "r"(v));
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v));
}

// Fused record-chain walk (records.cu): the inflate warp follows the block_size chain of its own block out of the
// shared-memory output ring while the bytes are still there.  RING = size of that ring (power of two); ring index of
// block-relative output byte x is (oa + x) & (RING - 1).  All members are warp-uniform.
//
// Entry point of the chain into a block: block 0 of a slice enters at a known offset.  Every other block first tries
// offset 0 (files written by BioD / htslib start every block at a record) and otherwise searches for the first offset
// where a plausible record starts (htsjdk-style files cut records anywhere).  The guess is only a speculation:
// scan_resolve_kernel checks it against the previous block's chain end and repairs it.
template <int RING>
struct Walker {
  static constexpr uint32_t OM = RING - 1;
  WalkOut wk;
  uint32_t ring;       // shared-space address of the output ring
  uint32_t oa;         // ring index of output byte 0
  uint32_t isize;
  uint64_t obase;
  uint32_t blk, sb, win0;
  bool walking, wentry;
  uint32_t wnext, wcnt, wncig, wsearch;
  int wbad;
  uint64_t win_abs;    // reported entry (absolute); stays "unknown" when no record starts in the block

  __device__ __forceinline__ void init(const WalkOut& w, uint32_t ring_, uint32_t oa_, uint32_t isize_, uint64_t obase_,
                                       uint32_t blk_) {
    wk = w; ring = ring_; oa = oa_; isize = isize_; obase = obase_; blk = blk_;
    walking = w.rel != nullptr;
    sb = blk + w.sb_offset;
    win0 = (walking && blk == 0) ? w.in0 : 0;
    wnext = win0; wcnt = 0; wncig = 0; wbad = WALK_OK;
    wentry = !walking || (blk == 0 && w.sb_offset == 0 && !w.search0);
    wsearch = 0;
    win_abs = ~0ull;
  }
  // little-endian u32 at block-relative offset x, read from the output ring (two aligned words + funnel shift)
  __device__ __forceinline__ uint32_t ring32(uint32_t x) const {
    const uint32_t r0 = (oa + x) & OM;
    const uint32_t w0 = lds32(ring + (r0 & ~3u));
    const uint32_t w1 = lds32(ring + ((r0 + 4) & OM & ~3u));
    return __funnelshift_r(w0, w1, (r0 & 3) * 8);
  }
  __device__ __forceinline__ uint32_t ring8(uint32_t x) const { return lds8(ring + ((oa + x) & OM)); }
  // is a BAM record header plausible at block-relative offset c?  1 yes, 0 no, -1 not enough bytes produced yet
  __device__ int plausible(uint32_t c, uint32_t avail) const {
    if (c + 36 > avail) return -1;
    const int32_t bs = (int32_t)ring32(c);
    if (bs < 34 || bs > (1 << 27)) return 0;
    const int32_t ref = (int32_t)ring32(c + 4), pos = (int32_t)ring32(c + 8);
    if (ref < -1 || ref >= wk.n_refs || pos < -1) return 0;
    const uint32_t bin_mq_nl = ring32(c + 12), flag_nc = ring32(c + 16);
    const int32_t l_seq = (int32_t)ring32(c + 20), nref = (int32_t)ring32(c + 24), npos = (int32_t)ring32(c + 28);
    const uint32_t lname = bin_mq_nl & 0xFF, nc = flag_nc & 0xFFFF;
    if (lname == 0 || l_seq < 0 || nref < -1 || nref >= wk.n_refs || npos < -1) return 0;
    const uint64_t need = 32ull + lname + 4ull * nc + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq;
    if (need > (uint64_t)bs) return 0;
    // read name: printable characters closed by a NUL (read.d:984-990)
    const uint32_t nm = c + 36;
    if (nm + lname > avail) return -1;
    if (ring8(nm + lname - 1) != 0) return 0;
    for (uint32_t k = 0; k + 1 < lname && k < 8; ++k) {
      const uint32_t ch = ring8(nm + k);
      if (ch < 0x21 || ch > 0x7e) return 0;
    }
    return 1;
  }
  // find the chain entry: lanes test 32 candidate offsets at a time
  __device__ void find_entry(uint32_t avail, bool final, int lane) {
    while (!wentry) {
      const uint32_t c = wsearch + lane;
      int ok = (c < isize) ? plausible(c, avail) : 0;
      if (ok == 1) {
        // a lone plausible header is not enough: the record it announces must be followed by another plausible one
        // (only when that one is close enough for the candidate itself to stay in the output ring meanwhile)
        const uint32_t nx = c + 4 + ring32(c);
        if (nx + 36 <= avail) ok = plausible(nx, avail) == 0 ? 0 : 1;
        else if (nx + 36 <= isize && nx - c <= 1024 && !final) ok = -1;
      }
      const uint32_t yes = __ballot_sync(0xffffffffu, ok == 1), wait = __ballot_sync(0xffffffffu, ok == -1);
      const uint32_t first_yes = yes ? (uint32_t)__ffs(yes) - 1 : 32, first_wait = wait ? (uint32_t)__ffs(wait) - 1 : 32;
      if (first_yes < first_wait) {
        wnext = wsearch + first_yes;
        win_abs = obase + wnext;
        wentry = true;
      } else if (first_wait < 32) {
        if (!final) { wsearch += first_wait; return; }      // come back when more bytes are there
        wsearch += first_wait + 1;                          // end of block: what cannot be checked is not an entry
      } else {
        wsearch += 32;
      }
      if (!wentry && wsearch >= isize) { wnext = isize; wentry = true; }   // no record starts in this block
    }
  }
  // follow the block_size chain (readrange.d:118-173) over the records whose 24 leading bytes are already produced
  __device__ void walk_upto(uint32_t avail, bool final, int lane) {
    if (!walking) return;
    if (!wentry) {
      find_entry(avail, final, lane);
      if (!wentry) return;
    }
    while (wbad == WALK_OK && wnext < isize) {
      if (wnext + 24 > avail) {
        if (final) wbad = WALK_INCOMPLETE;     // the header straddles the block end: the resolve kernel finishes it
        break;
      }
      const int32_t bs = (int32_t)ring32(wnext);
      if (bs < 32) { wbad = WALK_BAD_SIZE; break; }
      if (obase + wnext + 4 + (uint64_t)bs > wk.u_len) { wbad = WALK_TAIL; break; }
      const uint32_t bin_mq_nl = ring32(wnext + 12), flag_nc = ring32(wnext + 16);
      const int32_t l_seq = (int32_t)ring32(wnext + 20);
      const uint32_t lname = bin_mq_nl & 0xFF, nc = flag_nc & 0xFFFF;
      const uint64_t need = 32ull + lname + 4ull * nc + (l_seq > 0 ? ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq : 0);
      if (l_seq < 0 || need > (uint64_t)bs) { wbad = WALK_BAD_FIELDS; break; }
      // relative to the chain entry of the block, as scan_extract_kernel expects (block_uoff[0] includes in0)
      if (lane == 0 && wcnt < (uint32_t)SCAN_SLOTS) wk.rel[(size_t)sb * SCAN_SLOTS + wcnt] = (uint16_t)(wnext - win0);
      ++wcnt;
      wncig += nc;
      wnext += 4 + (uint32_t)bs;
    }
  }
  // per-block results of the walk (lane 0)
  __device__ void store(int status) const {
    if (!walking) return;
    wk.cnt[sb] = wcnt;
    wk.ncig[sb] = wncig;
    wk.in[sb] = (blk == 0 && wk.sb_offset == 0 && !wk.search0) ? obase + wk.in0 : win_abs;
    wk.out[sb] = obase + wnext;
    wk.bad[sb] = status ? WALK_TAIL : wbad;
  }
};

// status written by the lane-parallel kernels (inflate_tok.cu) for a block they could not finish (malformed or unusual stream): the warp-serial
// kernel then redoes the block and produces zlib's exact return code
constexpr int STATUS_RETRY = 0x7fffff01;

}  // namespace
}  // namespace biodb
