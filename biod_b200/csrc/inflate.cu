// BGZF block inflater for sm_100a: one warp per BGZF block.
//
// Replaces decompressBgzfBlock (bio/core/bgzf/block.d:127-216), i.e. libz's
// inflateInit2(-15) / inflate(Z_FINISH) / inflateEnd on one <=64 KiB raw-DEFLATE
// payload, including the error classes the reference surfaces as ZlibException
// (Z_DATA_ERROR / Z_BUF_ERROR).  The algorithm is RFC 1951; nothing here is
// derived from zlib's source.
//
// Design (per warp == per BGZF block, one 32-thread CTA each, ~10 KB of shared memory so that
// ~20 blocks are resident per SM):
//  * the compressed payload is staged through a 2 x IN_HALF shared-memory ring by
//    the TMA bulk-copy engine (cp.async.bulk global->shared, completion on an
//    mbarrier) so the decoder never waits on a global load;
//  * all 32 lanes run the (inherently serial) Huffman decode redundantly and
//    warp-uniformly out of 16-bit shared-memory LUTs -> no divergence, LUT reads are
//    broadcasts, every lane knows every symbol; literals run in a 13-instruction inner loop;
//  * literals/matches land in a shared-memory output ring that doubles as the
//    LZ77 window for near matches; far matches (older than the ring) read back the block's own
//    already-flushed bytes from L2 — the loads are issued at once and their stores deferred to
//    the next match/flush, so their latency hides behind the literals that follow;
//  * the ring is flushed to HBM in 512-byte, 16-byte-per-lane aligned vector
//    stores (ring index == global address mod ring size, so alignment carries).
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include "kernels.h"

namespace biodb {

namespace {

#ifndef BIODB_IN_HALF
#define BIODB_IN_HALF 1024
#endif
#ifndef BIODB_OUT_RING
#define BIODB_OUT_RING 4096
#endif
#ifndef BIODB_LIT_UNROLL2
#define BIODB_LIT_UNROLL2 0
#endif
#ifndef BIODB_LIT_BITS
#define BIODB_LIT_BITS 10
#endif
constexpr int IN_HALF = BIODB_IN_HALF;    // bytes per TMA chunk
constexpr int IN_RING = 2 * IN_HALF;
constexpr int IN_WORDS = IN_RING / 4;
constexpr int IN_HALF_WORDS = IN_HALF / 4;
constexpr int OUT_RING = BIODB_OUT_RING;
constexpr uint32_t OMASK = OUT_RING - 1;
constexpr int FLUSH = 512;
// bytes that may sit in the ring not yet flushed: < 2*FLUSH + one match (258) + one literal run (<= 32)
constexpr int MAX_PENDING = 2 * FLUSH + 258 + 64;
constexpr int RING_VALID = OUT_RING - MAX_PENDING - 64;  // any source byte this close to opos is still in the ring
static_assert(RING_VALID >= 1024, "output ring too small");
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

struct __align__(16) WarpSmem {
  uint32_t in_ring[IN_WORDS];
  uint8_t out_ring[OUT_RING];
  uint16_t lut_lit[1 << LIT_BITS];
  uint16_t lut_dist[1 << DIST_BITS];   // also hosts the 128-entry code-length LUT
  uint16_t sorted_lit[288];
  uint16_t sorted_dist[32];
  Code code_lit, code_dist;
  uint8_t lens[352];                   // [0,19) code-length code, [32,32+316) litlen+dist lengths
  unsigned long long mbar[2];
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
__device__ __forceinline__ uint32_t make_entry(int kind, int sym, int cl) {
  if (kind == KIND_LITLEN) {
    if (sym < 256) return ((uint32_t)cl << 12) | (uint32_t)sym;
    if (sym == 256) return ((uint32_t)cl << 12) | (K_EOB << 8);
    if (sym > 285) return ENT_INVALID;            // 286/287 exist only in the fixed code and are invalid
    return ((uint32_t)cl << 12) | (K_LEN << 8) | (uint32_t)(sym - 257);
  }
  if (kind == KIND_DIST) return sym > 29 ? ENT_INVALID : (((uint32_t)cl << 12) | (uint32_t)sym);
  return ((uint32_t)cl << 12) | (uint32_t)sym;     // code-length code
}

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

}  // namespace

__global__ void __launch_bounds__(32) inflate_kernel(InflateArgs a) {
  __shared__ WarpSmem sm;
  WarpSmem* s = &sm;
  const int lane = threadIdx.x;
  const uint32_t blk = blockIdx.x;
  if (blk >= a.n_blocks) return;

  const uint64_t poff = a.payload_off[blk];
  const uint32_t csize = a.cdata_size[blk];
  const uint32_t isize = a.isize[blk];
  const uint64_t obase = a.out_off[blk];
  uint8_t* gout = a.out + obase;
  const uint32_t oa = (uint32_t)(((uintptr_t)gout) & OMASK);   // ring index of output byte 0

  // shared-window addresses, made opaque so that they live in registers instead of being rebuilt
  uint32_t sbase = smem_u32(s);
  asm volatile("mov.u32 %0, %0;" : "+r"(sbase));
  const uint32_t in_ring = sbase + (uint32_t)offsetof(WarpSmem, in_ring);
  const uint32_t ring = sbase + (uint32_t)offsetof(WarpSmem, out_ring);
  const uint32_t lutl = sbase + (uint32_t)offsetof(WarpSmem, lut_lit);
  const uint32_t lutd = sbase + (uint32_t)offsetof(WarpSmem, lut_dist);
  const uint32_t mbar = sbase + (uint32_t)offsetof(WarpSmem, mbar);

  if (lane == 0) {
    mbar_init(mbar, 1);
    mbar_init(mbar + 8, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const uint8_t* pay = a.comp + poff;
  const uint32_t skip = (uint32_t)(((uintptr_t)pay) & 15);
  const uint8_t* src = pay - skip;                         // 16-byte aligned start of the staged stream
  const uint32_t staged = skip + csize;                    // bytes from src that matter
  uint32_t n_chunks = (staged + IN_HALF - 1) / IN_HALF;
  if (n_chunks == 0) n_chunks = 1;
  uint32_t last_bytes = (staged - (n_chunks - 1) * IN_HALF + 15) & ~15u;
  if (last_bytes == 0) last_bytes = 16;
  uint32_t issued = 0, waited = 0;

  auto issue = [&](uint32_t k) {
    __syncwarp();        // every lane has finished reading the half being overwritten (calls are warp-uniform)
    if (lane == 0) {
      const uint32_t bytes = (k + 1 == n_chunks) ? last_bytes : IN_HALF;
      const uint32_t bar = mbar + 8 * (k & 1);
      mbar_expect_tx(bar, bytes);
      tma_load(in_ring + (k & 1) * IN_HALF, src + (size_t)k * IN_HALF, bytes, bar);
    }
    issued = k + 1;
  };
  auto wait_chunk = [&](uint32_t k) {
    mbar_wait(mbar + 8 * (k & 1), (k >> 1) & 1);
    waited = k + 1;
  };
  issue(0);
  if (n_chunks > 1) issue(1);
  wait_chunk(0);

  uint32_t w = skip >> 2;   // next 32-bit word to pull from the staging ring
  uint64_t bitbuf = 0;
  int bitcnt = 0;
  const uint64_t total_bits = (uint64_t)csize * 8;
  const uint32_t skip_bits = skip * 8;
  int status = 0;
  uint32_t o = oa;          // oa + bytes produced: ring index is (o & OMASK)
  uint32_t flushed = 0;     // bytes already stored to HBM
  // deferred far match: bytes already requested from L2, to be stored into the ring later
  uint32_t pend_len = 0, pend_o = 0, pv0 = 0, pv1 = 0;
  // fused record-chain walk (records.cu): next record start inside this block, records / cigar words found so far
  const bool walking = a.walk.rel != nullptr;
  const uint32_t sb = blk + a.walk.sb_offset;
  const uint32_t win0 = (walking && blk == 0) ? a.walk.in0 : 0;
  uint32_t wnext = win0, wcnt = 0, wncig = 0;
  int wbad = WALK_OK;
  // Entry point of the chain into this block.  Block 0 of a slice enters at a known offset.  Every other block
  // first tries offset 0 (files written by BioD / htslib start every block at a record) and otherwise searches
  // for the first offset where a plausible record starts (htsjdk-style files cut records anywhere).  The guess is
  // only a speculation: scan_resolve_kernel checks it against the previous block's chain end and repairs it.
  bool wentry = !walking || (blk == 0 && a.walk.sb_offset == 0);
  uint32_t wsearch = 0;        // candidates below this offset are ruled out
  uint64_t win_abs = ~0ull;    // reported entry (absolute); stays "unknown" when no record starts in the block
#define OPOS() (o - oa)

  // pull one 32-bit word from the staging ring (warp-uniform)
  auto pull = [&]() {
    const uint32_t wd = lds32(in_ring + ((w & (IN_WORDS - 1)) << 2));
    bitbuf |= (uint64_t)wd << bitcnt;
    bitcnt += 32;
    ++w;
    if ((w & (IN_HALF_WORDS - 1)) == 0) {
      const uint32_t k = w / IN_HALF_WORDS;                // chunk about to be read
      if (k < n_chunks) {
        if (k + 1 < n_chunks && issued < k + 2) issue(k + 1);   // refill the half just drained
        if (waited < k + 1) wait_chunk(k);
      }
    }
  };
#define REFILL() do { if (bitcnt <= 32) pull(); } while (0)
#define DROP(n) do { bitbuf >>= (n); bitcnt -= (n); } while (0)
#define CONSUMED() ((uint64_t)w * 32 - (uint64_t)bitcnt - skip_bits)

  // first word may start mid-word
  pull();
  if (skip & 3) { int sh = (skip & 3) * 8; DROP(sh); }

  auto complete_pending = [&]() {
    if (pend_len) {
      if ((uint32_t)lane < pend_len) sts8(ring + ((pend_o + lane) & OMASK), pv0);
      if ((uint32_t)lane + 32 < pend_len) sts8(ring + ((pend_o + 32 + lane) & OMASK), pv1);
      pend_len = 0;
      __syncwarp();
    }
  };
  // little-endian u32 at block-relative offset x, read from the output ring
  auto ring32 = [&](uint32_t x) -> uint32_t {
    const uint32_t r0 = oa + x;
    return lds8(ring + (r0 & OMASK)) | (lds8(ring + ((r0 + 1) & OMASK)) << 8) | (lds8(ring + ((r0 + 2) & OMASK)) << 16) |
           (lds8(ring + ((r0 + 3) & OMASK)) << 24);
  };
  // is a BAM record header plausible at block-relative offset c?  1 yes, 0 no, -1 not enough bytes produced yet
  auto plausible = [&](uint32_t c, uint32_t avail) -> int {
    if (c + 36 > avail) return -1;
    const int32_t bs = (int32_t)ring32(c);
    if (bs < 34 || bs > (1 << 27)) return 0;
    const int32_t ref = (int32_t)ring32(c + 4), pos = (int32_t)ring32(c + 8);
    if (ref < -1 || ref >= a.walk.n_refs || pos < -1) return 0;
    const uint32_t bin_mq_nl = ring32(c + 12), flag_nc = ring32(c + 16);
    const int32_t l_seq = (int32_t)ring32(c + 20), nref = (int32_t)ring32(c + 24), npos = (int32_t)ring32(c + 28);
    const uint32_t lname = bin_mq_nl & 0xFF, nc = flag_nc & 0xFFFF;
    if (lname == 0 || l_seq < 0 || nref < -1 || nref >= a.walk.n_refs || npos < -1) return 0;
    const uint64_t need = 32ull + lname + 4ull * nc + ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq;
    if (need > (uint64_t)bs) return 0;
    // read name: printable characters closed by a NUL (read.d:984-990)
    const uint32_t nm = c + 36;
    if (nm + lname > avail) return -1;
    if (lds8(ring + ((oa + nm + lname - 1) & OMASK)) != 0) return 0;
    for (uint32_t k = 0; k + 1 < lname && k < 8; ++k) {
      const uint32_t ch = lds8(ring + ((oa + nm + k) & OMASK));
      if (ch < 0x21 || ch > 0x7e) return 0;
    }
    return 1;
  };
  // find the chain entry: lanes test 32 candidate offsets at a time
  auto find_entry = [&](uint32_t avail, bool final) {
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
  };
  // follow the block_size chain (readrange.d:118-173) over the records whose 24 leading bytes are already produced
  auto walk_upto = [&](uint32_t avail, bool final) {
    if (!wentry) {
      find_entry(avail, final);
      if (!wentry) return;
    }
    while (wbad == WALK_OK && wnext < isize) {
      if (wnext + 24 > avail) {
        if (final) wbad = WALK_INCOMPLETE;     // the header straddles the block end: the resolve kernel finishes it
        break;
      }
      const int32_t bs = (int32_t)ring32(wnext);
      if (bs < 32) { wbad = WALK_BAD_SIZE; break; }
      if (obase + wnext + 4 + (uint64_t)bs > a.walk.u_len) { wbad = WALK_TAIL; break; }
      const uint32_t bin_mq_nl = ring32(wnext + 12), flag_nc = ring32(wnext + 16);
      const int32_t l_seq = (int32_t)ring32(wnext + 20);
      const uint32_t lname = bin_mq_nl & 0xFF, nc = flag_nc & 0xFFFF;
      const uint64_t need = 32ull + lname + 4ull * nc + (l_seq > 0 ? ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq : 0);
      if (l_seq < 0 || need > (uint64_t)bs) { wbad = WALK_BAD_FIELDS; break; }
      // relative to the chain entry of the block, as scan_extract_kernel expects (block_uoff[0] includes in0)
      if (lane == 0 && wcnt < (uint32_t)SCAN_SLOTS) a.walk.rel[(size_t)sb * SCAN_SLOTS + wcnt] = (uint16_t)(wnext - win0);
      ++wcnt;
      wncig += nc;
      wnext += 4 + (uint32_t)bs;
    }
  };
  auto flush_to = [&](uint32_t fe) {
    // copy ring bytes [flushed, fe) to HBM; 16-byte vector stores where the global address allows
    uint32_t f = flushed;
    uint32_t head = (16 - ((oa + f) & 15)) & 15;
    if (head > fe - f) head = fe - f;
    if (head) {
      if ((uint32_t)lane < head) gout[f + lane] = (uint8_t)lds8(ring + ((oa + f + lane) & OMASK));
      f += head;
    }
    const uint32_t n16 = (fe - f) >> 4;
    for (uint32_t i = lane; i < n16; i += 32) {
      uint4 v = lds128(ring + ((oa + f + 16 * i) & OMASK));
      __stcs(reinterpret_cast<uint4*>(gout + f + 16 * i), v);
    }
    f += n16 << 4;
    const uint32_t tail = fe - f;
    if ((uint32_t)lane < tail) gout[f + lane] = (uint8_t)lds8(ring + ((oa + f + lane) & OMASK));
    flushed = fe;
    __syncwarp();
  };
  // flush whole 512-byte granules once two are pending; an overrun of ISIZE ends the block (Z_BUF_ERROR)
  auto maybe_flush = [&]() {
    if (OPOS() - flushed >= 2 * FLUSH) {
      if (OPOS() > isize) { status = Z_BUF; return; }
      complete_pending();
      __syncwarp();
      if (walking) walk_upto(OPOS(), false);
      const uint32_t fe = OPOS() - (o & (FLUSH - 1));
      if (fe > flushed) flush_to(fe);
    }
  };

  bool last = false;
  while (!last && status == 0) {
    REFILL();
    last = bitbuf & 1;
    const int btype = (int)((bitbuf >> 1) & 3);
    DROP(3);
    if (btype == 3) { status = Z_DATA; break; }

    if (btype == 0) {
      // ---- stored block -------------------------------------------------------------
      const int pad = bitcnt & 7;
      DROP(pad);
      REFILL();
      const uint32_t lw = (uint32_t)bitbuf;
      const uint32_t len = lw & 0xffff, nlen = lw >> 16;
      if (CONSUMED() + 32 > total_bits) { status = Z_BUF; break; }
      DROP(32);
      if ((len ^ 0xffff) != nlen) { status = Z_DATA; break; }
      if ((uint64_t)len * 8 + CONSUMED() > total_bits) { status = Z_BUF; break; }   // input runs out first ...
      if (OPOS() + len > isize) { status = Z_BUF; break; }                          // ... or the output does
      uint32_t left = len;
      while (left && status == 0) {
        REFILL();
        const uint32_t take = left < 4 ? left : 4;
        const uint32_t v = (uint32_t)bitbuf;
        if ((uint32_t)lane < take) sts8(ring + ((o + lane) & OMASK), v >> (8 * lane));
        DROP((int)take * 8);
        o += take;
        left -= take;
        maybe_flush();
      }
      __syncwarp();
      continue;
    }

    if (btype == 1) {
      // ---- fixed Huffman code (RFC 1951 §3.2.6) ---------------------------------------
      for (int i = lane; i < 288; i += 32) s->lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
      __syncwarp();
      build_table<LIT_BITS>(s->lens, 288, s->lut_lit, s->sorted_lit, &s->code_lit, KIND_LITLEN, lane);
      __syncwarp();
      s->lens[lane] = 5;
      __syncwarp();
      build_table<DIST_BITS>(s->lens, 32, s->lut_dist, s->sorted_dist, &s->code_dist, KIND_DIST, lane);
    } else {
      // ---- dynamic Huffman code (RFC 1951 §3.2.7) --------------------------------------
      REFILL();
      const int hlit = (int)(bitbuf & 31) + 257;
      const int hdist = (int)((bitbuf >> 5) & 31) + 1;
      const int hclen = (int)((bitbuf >> 10) & 15) + 4;
      DROP(14);
      if (hlit > 286 || hdist > 30) { status = (CONSUMED() > total_bits) ? Z_BUF : Z_DATA; break; }
      if (lane < 19) s->lens[lane] = 0;
      __syncwarp();
      for (int i = 0; i < hclen; ++i) {
        REFILL();
        // order of code-length code lengths, RFC 1951 §3.2.7, packed 5 bits each
        const uint64_t ord_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 |
                                6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
        const uint64_t ord_hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
        const int sym = i < 12 ? (int)((ord_lo >> (5 * i)) & 31) : (int)((ord_hi >> (5 * (i - 12))) & 31);
        s->lens[sym] = (uint8_t)(bitbuf & 7);
        DROP(3);
      }
      if (CONSUMED() > total_bits) { status = Z_BUF; break; }
      __syncwarp();
      uint16_t* cl_lut = s->lut_dist;
      int r = build_table<CL_BITS>(s->lens, 19, cl_lut, s->sorted_dist, &s->code_dist, KIND_CODELEN, lane);
      if (r < 0) { status = Z_DATA; break; }
      const int total = hlit + hdist;
      if (r == 1) {
        // no code-length codes at all: zlib reads every length as 0 (1 bit each) and then fails on the
        // missing end-of-block code — unless the input runs out first.
        status = (CONSUMED() + (uint64_t)total > total_bits) ? Z_BUF : Z_DATA;
        break;
      }
      __syncwarp();
      int idx = 0;
      int prev = 0;
      while (idx < total) {
        REFILL();
        const uint32_t e = lds16(lutd + (((uint32_t)bitbuf & ((1u << CL_BITS) - 1)) << 1));
        const int cl = e >> 12;
        if (cl == 0) { status = Z_DATA; break; }         // unused code of an (impossible here) incomplete set
        const int sym = e & 31;
        DROP(cl);
        if (sym < 16) {
          s->lens[32 + idx] = (uint8_t)sym;
          prev = sym;
          ++idx;
        } else {
          int rep, val;
          if (sym == 16) {
            if (idx == 0) { status = Z_DATA; break; }
            rep = 3 + (int)(bitbuf & 3);
            DROP(2);
            val = prev;
          } else if (sym == 17) {
            rep = 3 + (int)(bitbuf & 7);
            DROP(3);
            val = 0;
          } else {
            rep = 11 + (int)(bitbuf & 127);
            DROP(7);
            val = 0;
          }
          if (idx + rep > total) { status = Z_DATA; break; }
          for (int k = lane; k < rep; k += 32) s->lens[32 + idx + k] = (uint8_t)val;
          prev = val;
          idx += rep;
        }
      }
      if (CONSUMED() > total_bits) { status = Z_BUF; break; }
      if (status) break;
      __syncwarp();
      if (s->lens[32 + 256] == 0) { status = Z_DATA; break; }   // no end-of-block code
      __syncwarp();
      r = build_table<LIT_BITS>(s->lens + 32, hlit, s->lut_lit, s->sorted_lit, &s->code_lit, KIND_LITLEN, lane);
      if (r < 0) { status = Z_DATA; break; }
      r = build_table<DIST_BITS>(s->lens + 32 + hlit, hdist, s->lut_dist, s->sorted_dist, &s->code_dist, KIND_DIST, lane);
      if (r < 0) { status = Z_DATA; break; }
    }
    __syncwarp();

    // ---- symbol loop ---------------------------------------------------------------------
    while (true) {
      REFILL();                       // >= 33 valid bits
      maybe_flush();
      if (status) break;
      uint32_t e;
      // literal fast loop on the low 32 bits of the bit buffer: every lane stores the same byte to the same ring
      // slot (one shared-memory wavefront); it runs while at least LIT_BITS of the 32 bits are unread
      {
        uint32_t lo = (uint32_t)bitbuf, used = 0;
#if BIODB_LIT_UNROLL2
        // two literals per trip: the second lookup is issued before the first literal is known to be one
        while (true) {
          e = lds16(lutl + ((lo << 1) & (((1u << LIT_BITS) - 1) << 1)));
          const uint32_t cl = e >> 12;
          const uint32_t lo1 = lo >> cl;
          const uint32_t e1 = lds16(lutl + ((lo1 << 1) & (((1u << LIT_BITS) - 1) << 1)));
          if (e & (3u << 8)) break;     // not a literal
          sts8(ring + (o & OMASK), e);
          ++o;
          lo = lo1;
          used += cl;
          if (used > 32 - LIT_BITS) break;
          e = e1;
          if (e & (3u << 8)) break;
          sts8(ring + (o & OMASK), e);
          ++o;
          const uint32_t cl1 = e >> 12;
          lo >>= cl1;
          used += cl1;
          if (used > 32 - LIT_BITS) break;
        }
#else
        while (true) {
          e = lds16(lutl + ((lo << 1) & (((1u << LIT_BITS) - 1) << 1)));
          if (e & (3u << 8)) break;     // not a literal
          sts8(ring + (o & OMASK), e);
          ++o;
          const uint32_t cl = e >> 12;
          lo >>= cl;
          used += cl;
          if (used > 32 - LIT_BITS) break;
        }
#endif
        DROP(used);
      }
      if ((e & (3u << 8)) == 0) continue;   // ran low on bits after a literal
      if (((e >> 8) & 3) == K_SPECIAL) {
        REFILL();
        if (e == ENT_SLOW) e = slow_decode<LIT_BITS>((uint32_t)bitbuf, &s->code_lit, s->sorted_lit, KIND_LITLEN);
        if (((e >> 8) & 3) == K_SPECIAL) {   // invalid code
          status = (CONSUMED() + 1 > total_bits) ? Z_BUF : Z_DATA;
          break;
        }
        if (((e >> 8) & 3) == K_LIT) {       // a literal with a long code
          sts8(ring + (o & OMASK), e);
          ++o;
          DROP(e >> 12);
          continue;
        }
      }
      DROP(e >> 12);
      if (((e >> 8) & 3) == K_EOB) break;
      // ---- length / distance pair ---------------------------------------------------------
      REFILL();                       // the literal run may have left fewer bits than the extra bits need
      uint32_t lbase, eb;
      len_base(e & 31, lbase, eb);
      const uint32_t len = lbase + ((uint32_t)bitbuf & ((1u << eb) - 1));
      DROP((int)eb);
      REFILL();
      uint32_t e2 = lds16(lutd + (((uint32_t)bitbuf & ((1u << DIST_BITS) - 1)) << 1));
      if (((e2 >> 8) & 3) == K_SPECIAL) {
        if (e2 == ENT_SLOW) e2 = slow_decode<DIST_BITS>((uint32_t)bitbuf, &s->code_dist, s->sorted_dist, KIND_DIST);
        if (((e2 >> 8) & 3) == K_SPECIAL) {
          status = (CONSUMED() + 1 > total_bits) ? Z_BUF : Z_DATA;
          break;
        }
      }
      DROP(e2 >> 12);
      uint32_t dbase, eb2;
      dist_base(e2 & 31, dbase, eb2);
      const uint32_t dist = dbase + ((uint32_t)bitbuf & ((1u << eb2) - 1));
      DROP((int)eb2);
      const uint32_t opos = OPOS();
      if (opos > isize) { status = Z_BUF; break; }             // an earlier literal overran the output
      if (CONSUMED() > total_bits) { status = Z_BUF; break; }
      if (dist > opos) { status = Z_DATA; break; }             // distance too far back
      if (opos + len > isize) { status = Z_BUF; break; }       // output space exhausted mid-match
      // ---- LZ77 copy, warp-cooperative ---------------------------------------------------
      complete_pending();
      if (dist <= (uint32_t)RING_VALID) {
        const uint32_t sp = o - dist;                          // ring-relative source start
        if (dist >= len) {
          for (uint32_t i = lane; i < len; i += 32) sts8(ring + ((o + i) & OMASK), lds8(ring + ((sp + i) & OMASK)));
        } else {
          uint32_t m = (uint32_t)lane % dist;
          const uint32_t k = 32u % dist;
          for (uint32_t i = lane; i < len; i += 32) {
            sts8(ring + ((o + i) & OMASK), lds8(ring + ((sp + m) & OMASK)));
            m += k;
            if (m >= dist) m -= dist;
          }
        }
        __syncwarp();
      } else {
        // far match: the source is older than the ring and therefore already flushed (dist > len here)
        const uint8_t* g = gout + (opos - dist);
        if (len <= 64) {
          // request the bytes now, store them into the ring at the next match / flush: the L2 round trip
          // overlaps with the literals that follow
          if ((uint32_t)lane < len) pv0 = __ldcg(g + lane);
          if ((uint32_t)lane + 32 < len) pv1 = __ldcg(g + 32 + lane);
          pend_len = len;
          pend_o = o;
        } else {
          for (uint32_t i = lane; i < len; i += 32) sts8(ring + ((o + i) & OMASK), __ldcg(g + i));
          __syncwarp();
        }
      }
      o += len;
    }
    if (status == 0 && CONSUMED() > total_bits) status = Z_BUF;
    if (status == 0 && OPOS() > isize) status = Z_BUF;
  }

  // drain any TMA chunk still in flight before the CTA (and its shared memory) retires
  while (waited < issued) wait_chunk(waited);

  if (status != 0 && status != Z_BUF && OPOS() > isize) status = Z_BUF;   // the output overran before the fault
  if (status == Z_DATA && CONSUMED() > total_bits) status = Z_BUF;        // zlib would have run out of input first
  // stream ended short of ISIZE: -release BioD would hand out garbage (block.d:175); reported as a data error
  if (status == 0 && OPOS() != isize) status = Z_DATA;
  if (status == 0) {
    complete_pending();
    __syncwarp();
    if (walking) walk_upto(isize, true);
    if (OPOS() > flushed) flush_to(OPOS());
  }
  if (lane == 0) {
    a.status[blk] = status;
    if (walking) {
      a.walk.cnt[sb] = wcnt;
      a.walk.ncig[sb] = wncig;
      a.walk.in[sb] = (blk == 0 && a.walk.sb_offset == 0) ? obase + a.walk.in0 : win_abs;
      a.walk.out[sb] = obase + wnext;
      a.walk.bad[sb] = status ? WALK_TAIL : wbad;
    }
  }
#undef REFILL
#undef DROP
#undef CONSUMED
#undef OPOS
}

unsigned long long g_kernel_launches = 0;

size_t inflate_smem_bytes() { return sizeof(WarpSmem); }

cudaError_t launch_inflate(const InflateArgs& a, cudaStream_t st) {
  if (a.n_blocks == 0) return cudaSuccess;
  inflate_kernel<<<a.n_blocks, 32, 0, st>>>(a);
  ++g_kernel_launches;
  return cudaGetLastError();
}

}  // namespace biodb
