// BGZF block inflater for sm_100a: one warp per BGZF block.
//
// Replaces decompressBgzfBlock (bio/core/bgzf/block.d:127-216), i.e. libz's
// inflateInit2(-15) / inflate(Z_FINISH) / inflateEnd on one <=64 KiB raw-DEFLATE
// payload, including the error classes the reference surfaces as ZlibException
// (Z_DATA_ERROR / Z_BUF_ERROR).  The algorithm is RFC 1951; nothing here is
// derived from zlib's source.
//
// Design (per warp == per BGZF block, one 32-thread CTA each):
//  * the compressed payload is staged through a 2 x IN_HALF shared-memory ring by
//    the TMA bulk-copy engine (cp.async.bulk global->shared, completion on an
//    mbarrier) so the decoder never waits on a global load;
//  * all 32 lanes run the (inherently serial) Huffman decode redundantly and
//    warp-uniformly out of shared-memory LUTs -> no divergence, LUT reads are
//    broadcasts, every lane knows every symbol;
//  * literals/matches land in a shared-memory output ring that doubles as the
//    LZ77 window for near matches; far matches (> ring) read back the block's own
//    already-flushed bytes from L2;
//  * the ring is flushed to HBM in 512-byte, 16-byte-per-lane aligned vector
//    stores (ring index == global address mod ring size, so alignment carries).
#include <cuda_runtime.h>
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
constexpr int IN_HALF = BIODB_IN_HALF;    // bytes per TMA chunk
constexpr int IN_RING = 2 * IN_HALF;
constexpr int IN_WORDS = IN_RING / 4;
constexpr int IN_HALF_WORDS = IN_HALF / 4;
constexpr int OUT_RING = BIODB_OUT_RING;
constexpr uint32_t OMASK = OUT_RING - 1;
constexpr int FLUSH = 512;
// bytes that may sit in the ring not yet flushed: < 2*FLUSH + one match (258) + one literal run (<= 54)
constexpr int MAX_PENDING = 2 * FLUSH + 258 + 64;
constexpr int RING_VALID = OUT_RING - MAX_PENDING - 64;  // any source byte this close to opos is still in the ring
static_assert(RING_VALID >= 1024, "output ring too small");
constexpr int LIT_BITS = 10;
constexpr int DIST_BITS = 8;
constexpr int CL_BITS = 7;

// LUT entry layout (u32).  litlen: [0,8) literal byte | [8,10) kind | [10,19) length base | [19,22) extra bits |
// [28,32) code length.  dist: [0,2) kind | [8,23) distance base | [24,28) extra bits | [28,32) code length.
// code-length code: [8,13) symbol | [28,32) code length.
constexpr uint32_t K_LIT = 0, K_LEN = 1, K_EOB = 2, K_SPECIAL = 3;
constexpr uint32_t ENT_SLOW = K_SPECIAL << 8;                 // code longer than the LUT index: canonical slow path
constexpr uint32_t ENT_INVALID = (K_SPECIAL << 8) | (1u << 10);
constexpr uint32_t DENT_SLOW = K_SPECIAL;
constexpr uint32_t DENT_INVALID = K_SPECIAL | (1u << 8);
constexpr int Z_DATA = -3;
constexpr int Z_BUF = -5;

enum { KIND_LITLEN = 0, KIND_DIST = 1, KIND_CODELEN = 2 };

struct __align__(16) WarpSmem {
  uint32_t in_ring[IN_WORDS];
  uint8_t out_ring[OUT_RING];
  uint32_t lut_lit[1 << LIT_BITS];     // 4096
  uint32_t lut_dist[1 << DIST_BITS];   // 1024 (also hosts the 128-entry code-length LUT)
  uint16_t sorted_lit[288];
  uint16_t sorted_dist[32];
  uint16_t cnt_lit[16];
  uint16_t cnt_dist[16];
  uint8_t lens[352];                   // [0,19) code-length code, [32,32+316) litlen+dist lengths
  unsigned long long mbar[2];
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// explicit shared-space accesses with a 32-bit address: keeps ptxas from rebuilding the shared-window base
// (S2R SR_CgaCtaId + LEA) inside the hot loop
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

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok = 0;
  const uint32_t addr = smem_u32(bar);
  do {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
  } while (!ok);
}
// TMA bulk copy global -> shared (SASS: UBLKCP), completion signalled on the mbarrier.
__device__ __forceinline__ void tma_load(void* dst, const void* src, uint32_t bytes, unsigned long long* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                   smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

__device__ __forceinline__ uint32_t len_entry(int s /*0..28*/, int cl) {
  uint32_t base, eb;
  if (s < 8) { base = 3 + s; eb = 0; }
  else if (s == 28) { base = 258; eb = 0; }
  else { eb = (uint32_t)(s - 4) >> 2; base = 3 + ((4u + ((s - 4) & 3)) << eb); }
  return ((uint32_t)cl << 28) | (K_LEN << 8) | (base << 10) | (eb << 19);
}
__device__ __forceinline__ uint32_t dist_entry(int d /*0..29*/, int cl) {
  uint32_t base, eb;
  if (d < 4) { base = 1 + d; eb = 0; }
  else { eb = (uint32_t)(d - 2) >> 1; base = 1 + ((2u + (d & 1)) << eb); }
  return ((uint32_t)cl << 28) | (base << 8) | (eb << 24);
}
__device__ __forceinline__ uint32_t make_entry(int kind, int sym, int cl) {
  if (kind == KIND_LITLEN) {
    if (sym < 256) return ((uint32_t)cl << 28) | (uint32_t)sym;
    if (sym == 256) return ((uint32_t)cl << 28) | (K_EOB << 8);
    if (sym > 285) return ENT_INVALID;            // 286/287 exist only in the fixed code and are invalid
    return len_entry(sym - 257, cl);
  }
  if (kind == KIND_DIST) return sym > 29 ? DENT_INVALID : dist_entry(sym, cl);
  return ((uint32_t)cl << 28) | ((uint32_t)sym << 8);     // code-length code
}
__device__ __forceinline__ uint32_t slow_marker(int kind) { return kind == KIND_DIST ? DENT_SLOW : ENT_SLOW; }
__device__ __forceinline__ uint32_t invalid_marker(int kind) { return kind == KIND_DIST ? DENT_INVALID : ENT_INVALID; }

// Warp-cooperative canonical-Huffman table build (RFC 1951 §3.2.2).
// Returns 0 ok, 1 = empty code (LUT all-invalid), -1 = over-subscribed / incomplete set.
template <int PB>
__device__ int build_table(const uint8_t* lens, int n, uint32_t* lut, uint16_t* sorted, uint16_t* cnt_out, int kind,
                           int lane) {
  int cnt[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) cnt[i] = 0;
  for (int base = 0; base < n; base += 32) {
    int l = (base + lane < n) ? lens[base + lane] : 0;
#pragma unroll
    for (int len = 1; len <= 15; ++len) cnt[len] += __popc(__ballot_sync(0xffffffffu, l == len));
  }
  int left = 1, maxlen = 0;
  bool over = false;
#pragma unroll
  for (int len = 1; len <= 15; ++len) {
    left = (left << 1) - cnt[len];
    if (left < 0) { over = true; left = 0; }
    if (cnt[len]) maxlen = len;
  }
  if (over) return -1;
  for (int i = lane; i < (1 << PB); i += 32) lut[i] = invalid_marker(kind);
  if (lane < 16) cnt_out[lane] = 0;
  __syncwarp();
  if (maxlen == 0) return 1;
  if (left > 0 && (kind == KIND_CODELEN || maxlen != 1)) return -1;
  int next[16], offs[16];
  {
    int code = 0, o = 0;
    next[0] = 0;
    offs[0] = 0;
#pragma unroll
    for (int len = 1; len <= 15; ++len) {
      code = (code + (len > 1 ? cnt[len - 1] : 0)) << 1;
      next[len] = code;
      offs[len] = o;
      o += cnt[len];
    }
  }
  if (lane >= 1 && lane < 16) {
#pragma unroll
    for (int len = 1; len <= 15; ++len)
      if (lane == len) cnt_out[len] = (uint16_t)cnt[len];
  }
  int run[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) run[i] = 0;
  const uint32_t lt = (1u << lane) - 1;
  for (int base = 0; base < n; base += 32) {
    int sym = base + lane;
    int l = (sym < n) ? lens[sym] : 0;
    int code = 0, slot = 0;
#pragma unroll
    for (int len = 1; len <= 15; ++len) {
      uint32_t b = __ballot_sync(0xffffffffu, l == len);
      if (l == len) {
        int rank = run[len] + __popc(b & lt);
        code = next[len] + rank;
        slot = offs[len] + rank;
      }
      run[len] += __popc(b);
    }
    if (l > 0) {
      sorted[slot] = (uint16_t)sym;
      uint32_t rev = __brev((uint32_t)code) >> (32 - l);
      if (l <= PB) {
        uint32_t e = make_entry(kind, sym, l);
        for (uint32_t i = rev; i < (1u << PB); i += (1u << l)) lut[i] = e;
      } else {
        lut[rev & ((1u << PB) - 1)] = slow_marker(kind);
      }
    }
  }
  __syncwarp();
  return 0;
}

// Canonical bit-by-bit decode for codes longer than the LUT index (rare symbols).
__device__ __forceinline__ uint32_t slow_decode(uint64_t bitbuf, const uint16_t* cnt, const uint16_t* sorted, int kind) {
  int code = 0, first = 0, index = 0;
  uint32_t bits = (uint32_t)bitbuf;
  for (int len = 1; len <= 15; ++len) {
    code |= (int)(bits & 1);
    bits >>= 1;
    int c = cnt[len];
    if (code - c < first) return make_entry(kind, sorted[index + (code - first)], len);
    index += c;
    first += c;
    first <<= 1;
    code <<= 1;
  }
  return invalid_marker(kind);
}

struct Decoder {
  WarpSmem* s;
  const uint8_t* src;     // 16-byte aligned start of the staged byte stream
  uint32_t n_chunks;      // TMA chunks that cover the payload
  uint32_t last_bytes;    // size of the final chunk (multiple of 16)
  uint32_t issued, waited;
  uint32_t w;             // next 32-bit word to pull from the ring
  uint64_t bitbuf;
  int bitcnt;
  int lane;

  __device__ __forceinline__ void issue(uint32_t k) {
    // all lanes have finished reading the half being overwritten (calls are warp-uniform)
    __syncwarp();
    if (lane == 0) {
      uint32_t bytes = (k + 1 == n_chunks) ? last_bytes : IN_HALF;
      unsigned long long* bar = &s->mbar[k & 1];
      mbar_expect_tx(bar, bytes);
      tma_load(&s->in_ring[(k & 1) * IN_HALF_WORDS], src + (size_t)k * IN_HALF, bytes, bar);
    }
    issued = k + 1;
  }
  __device__ __forceinline__ void wait_chunk(uint32_t k, uint32_t phase_base) {
    mbar_wait(&s->mbar[k & 1], (phase_base + (k >> 1)) & 1);
    waited = k + 1;
  }
};

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

  if (lane == 0) {
    mbar_init(&s->mbar[0], 1);
    mbar_init(&s->mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncwarp();

  const uint8_t* pay = a.comp + poff;
  const uint32_t skip = (uint32_t)(((uintptr_t)pay) & 15);
  Decoder d;
  d.s = s;
  d.src = pay - skip;
  d.lane = lane;
  const uint32_t staged = skip + csize;                    // bytes from src that matter
  d.n_chunks = (staged + IN_HALF - 1) / IN_HALF;
  if (d.n_chunks == 0) d.n_chunks = 1;
  {
    uint32_t rem = staged - (d.n_chunks - 1) * IN_HALF;
    d.last_bytes = (rem + 15) & ~15u;
    if (d.last_bytes == 0) d.last_bytes = 16;
  }
  d.issued = d.waited = 0;
  d.issue(0);
  if (d.n_chunks > 1) d.issue(1);
  d.wait_chunk(0, 0);
  d.w = skip >> 2;
  d.bitbuf = 0;
  d.bitcnt = 0;

  const uint64_t total_bits = (uint64_t)csize * 8;
  const uint32_t skip_bits = skip * 8;
  int status = 0;
  uint32_t o = oa;          // oa + bytes produced: ring index is (o & OMASK)
  uint32_t flushed = 0;     // bytes already stored to HBM
  const uint32_t ring = smem_u32(s->out_ring);
  const uint32_t lutl = smem_u32(s->lut_lit);
  const uint32_t lutd = smem_u32(s->lut_dist);
#define OPOS() (o - oa)

  // pull one 32-bit word from the staging ring (warp-uniform)
  auto pull = [&]() {
    uint32_t wd = s->in_ring[d.w & (IN_WORDS - 1)];
    d.bitbuf |= (uint64_t)wd << d.bitcnt;
    d.bitcnt += 32;
    ++d.w;
    if ((d.w & (IN_HALF_WORDS - 1)) == 0) {
      uint32_t k = d.w / IN_HALF_WORDS;                  // chunk about to be read
      if (k < d.n_chunks) {
        if (k + 1 < d.n_chunks && d.issued < k + 2) d.issue(k + 1);   // refill the half just drained
        if (d.waited < k + 1) d.wait_chunk(k, 0);
      }
    }
  };
#define REFILL() do { if (d.bitcnt <= 32) pull(); } while (0)
#define DROP(n) do { d.bitbuf >>= (n); d.bitcnt -= (n); } while (0)
#define CONSUMED() ((uint64_t)d.w * 32 - (uint64_t)d.bitcnt - skip_bits)

  // first word may start mid-word
  pull();
  if (skip & 3) { int sh = (skip & 3) * 8; DROP(sh); }

  auto flush_to = [&](uint32_t fe) {
    // copy ring bytes [flushed, fe) to HBM; 16-byte vector stores where the global address allows
    uint32_t f = flushed;
    uint32_t head = (16 - ((oa + f) & 15)) & 15;
    if (head > fe - f) head = fe - f;
    if (head) {
      if ((uint32_t)lane < head) gout[f + lane] = (uint8_t)lds8(ring + ((oa + f + lane) & OMASK));
      f += head;
    }
    uint32_t n16 = (fe - f) >> 4;
    for (uint32_t i = lane; i < n16; i += 32) {
      uint4 v = lds128(ring + ((oa + f + 16 * i) & OMASK));
      __stcs(reinterpret_cast<uint4*>(gout + f + 16 * i), v);
    }
    f += n16 << 4;
    uint32_t tail = fe - f;
    if ((uint32_t)lane < tail) gout[f + lane] = (uint8_t)lds8(ring + ((oa + f + lane) & OMASK));
    flushed = fe;
    __syncwarp();
  };
  // flush whole 512-byte granules once two are pending; an overrun of ISIZE ends the block (Z_BUF_ERROR)
#define MAYBE_FLUSH()                                                  \
  do {                                                                 \
    if (OPOS() - flushed >= 2 * FLUSH) {                               \
      if (OPOS() > isize) { status = Z_BUF; break; }                   \
      __syncwarp();                                                    \
      uint32_t fe_ = OPOS() - (o & (FLUSH - 1));                       \
      if (fe_ > flushed) flush_to(fe_);                                \
    }                                                                  \
  } while (0)

  bool last = false;
  while (!last && status == 0) {
    REFILL();
    last = d.bitbuf & 1;
    int btype = (int)((d.bitbuf >> 1) & 3);
    DROP(3);
    if (btype == 3) { status = Z_DATA; break; }

    if (btype == 0) {
      // ---- stored block -------------------------------------------------------------
      int pad = d.bitcnt & 7;
      DROP(pad);
      REFILL();
      uint32_t lw = (uint32_t)d.bitbuf;
      uint32_t len = lw & 0xffff, nlen = lw >> 16;
      if (CONSUMED() + 32 > total_bits) { status = Z_BUF; break; }
      DROP(32);
      if ((len ^ 0xffff) != nlen) { status = Z_DATA; break; }
      if ((uint64_t)len * 8 + CONSUMED() > total_bits) { status = Z_BUF; break; }   // input runs out first ...
      if (OPOS() + len > isize) { status = Z_BUF; break; }                          // ... or the output does
      uint32_t left = len;
      while (left && status == 0) {
        REFILL();
        uint32_t take = left < 4 ? left : 4;
        uint32_t v = (uint32_t)d.bitbuf;
        if ((uint32_t)lane < take) sts8(ring + ((o + lane) & OMASK), v >> (8 * lane));
        DROP((int)take * 8);
        o += take;
        left -= take;
        MAYBE_FLUSH();
      }
      __syncwarp();
      continue;
    }

    if (btype == 1) {
      // ---- fixed Huffman code (RFC 1951 §3.2.6) ---------------------------------------
      for (int i = lane; i < 288; i += 32) s->lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
      __syncwarp();
      build_table<LIT_BITS>(s->lens, 288, s->lut_lit, s->sorted_lit, s->cnt_lit, KIND_LITLEN, lane);
      __syncwarp();
      s->lens[lane] = 5;
      __syncwarp();
      build_table<DIST_BITS>(s->lens, 32, s->lut_dist, s->sorted_dist, s->cnt_dist, KIND_DIST, lane);
    } else {
      // ---- dynamic Huffman code (RFC 1951 §3.2.7) --------------------------------------
      REFILL();
      int hlit = (int)(d.bitbuf & 31) + 257;
      int hdist = (int)((d.bitbuf >> 5) & 31) + 1;
      int hclen = (int)((d.bitbuf >> 10) & 15) + 4;
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
        int sym = i < 12 ? (int)((ord_lo >> (5 * i)) & 31) : (int)((ord_hi >> (5 * (i - 12))) & 31);
        s->lens[sym] = (uint8_t)(d.bitbuf & 7);
        DROP(3);
      }
      if (CONSUMED() > total_bits) { status = Z_BUF; break; }
      __syncwarp();
      uint32_t* cl_lut = s->lut_dist;
      int r = build_table<CL_BITS>(s->lens, 19, cl_lut, s->sorted_dist, s->cnt_dist, KIND_CODELEN, lane);
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
        uint32_t e = cl_lut[d.bitbuf & ((1u << CL_BITS) - 1)];
        int cl = e >> 28;
        if (cl == 0) { status = Z_DATA; break; }         // unused code of an (impossible here) incomplete set
        int sym = (e >> 8) & 31;
        DROP(cl);
        if (sym < 16) {
          s->lens[32 + idx] = (uint8_t)sym;
          prev = sym;
          ++idx;
        } else {
          int rep, val;
          if (sym == 16) {
            if (idx == 0) { status = Z_DATA; break; }
            rep = 3 + (int)(d.bitbuf & 3);
            DROP(2);
            val = prev;
          } else if (sym == 17) {
            rep = 3 + (int)(d.bitbuf & 7);
            DROP(3);
            val = 0;
          } else {
            rep = 11 + (int)(d.bitbuf & 127);
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
      r = build_table<LIT_BITS>(s->lens + 32, hlit, s->lut_lit, s->sorted_lit, s->cnt_lit, KIND_LITLEN, lane);
      if (r < 0) { status = Z_DATA; break; }
      r = build_table<DIST_BITS>(s->lens + 32 + hlit, hdist, s->lut_dist, s->sorted_dist, s->cnt_dist, KIND_DIST, lane);
      if (r < 0) { status = Z_DATA; break; }
    }
    __syncwarp();

    // ---- symbol loop ---------------------------------------------------------------------
    while (true) {
      REFILL();                       // >= 33 valid bits
      MAYBE_FLUSH();
      if (status) break;
      uint32_t e;
      // literal fast loop on the low 32 bits of the bit buffer: every lane stores the same byte to the same ring
      // slot (one shared-memory wavefront); it runs while at least LIT_BITS of the 32 bits are unread
      {
        uint32_t lo = (uint32_t)d.bitbuf, used = 0;
        while (true) {
          e = lds32(lutl + ((lo << 2) & (((1u << LIT_BITS) - 1) << 2)));
          if (e & (3u << 8)) break;     // not a literal
          sts8(ring + (o & OMASK), e);
          ++o;
          const uint32_t cl = e >> 28;
          lo >>= cl;
          used += cl;
          if (used > 32 - LIT_BITS) break;
        }
        DROP(used);
      }
      if ((e & (3u << 8)) == 0) continue;   // ran low on bits after a literal
      if (((e >> 8) & 3) == K_SPECIAL) {
        REFILL();
        if (e == ENT_SLOW) e = slow_decode(d.bitbuf, s->cnt_lit, s->sorted_lit, KIND_LITLEN);
        if (((e >> 8) & 3) == K_SPECIAL) {   // invalid code
          status = (CONSUMED() + 1 > total_bits) ? Z_BUF : Z_DATA;
          break;
        }
        if (((e >> 8) & 3) == K_LIT) {       // a literal with a long code
          sts8(ring + (o & OMASK), e);
          ++o;
          DROP(e >> 28);
          continue;
        }
      }
      DROP(e >> 28);
      const uint32_t kind = (e >> 8) & 3;
      if (kind == K_EOB) break;
      // ---- length / distance pair ---------------------------------------------------------
      REFILL();                       // the literal run may have left fewer bits than the extra bits need
      const uint32_t eb = (e >> 19) & 7;
      const uint32_t len = ((e >> 10) & 0x1ff) + ((uint32_t)d.bitbuf & ((1u << eb) - 1));
      DROP((int)eb);
      REFILL();
      uint32_t e2 = lds32(lutd + (((uint32_t)d.bitbuf << 2) & (((1u << DIST_BITS) - 1) << 2)));
      if ((e2 & 3) == K_SPECIAL) {
        if (e2 == DENT_SLOW) e2 = slow_decode(d.bitbuf, s->cnt_dist, s->sorted_dist, KIND_DIST);
        if ((e2 & 3) == K_SPECIAL) {
          status = (CONSUMED() + 1 > total_bits) ? Z_BUF : Z_DATA;
          break;
        }
      }
      DROP(e2 >> 28);
      const uint32_t eb2 = (e2 >> 24) & 15;
      const uint32_t dist = ((e2 >> 8) & 0x7fff) + ((uint32_t)d.bitbuf & ((1u << eb2) - 1));
      DROP((int)eb2);
      const uint32_t opos = OPOS();
      if (opos > isize) { status = Z_BUF; break; }             // an earlier literal overran the output
      if (CONSUMED() > total_bits) { status = Z_BUF; break; }
      if (dist > opos) { status = Z_DATA; break; }             // distance too far back
      if (opos + len > isize) { status = Z_BUF; break; }       // output space exhausted mid-match
      // ---- LZ77 copy, warp-cooperative ---------------------------------------------------
      const uint32_t sp = o - dist;                            // ring-relative source start
      if (dist <= (uint32_t)RING_VALID) {
        if (dist >= len) {
          for (uint32_t i = lane; i < len; i += 32) sts8(ring + ((o + i) & OMASK), lds8(ring + ((sp + i) & OMASK)));
        } else {
          uint32_t m = (uint32_t)lane % dist, k = 32u % dist;
          for (uint32_t i = lane; i < len; i += 32) {
            sts8(ring + ((o + i) & OMASK), lds8(ring + ((sp + m) & OMASK)));
            m += k;
            if (m >= dist) m -= dist;
          }
        }
      } else {
        // far match: the source is older than the ring and therefore already flushed (dist > len here)
        const uint8_t* g = gout + (opos - dist);
        for (uint32_t i = lane; i < len; i += 32) sts8(ring + ((o + i) & OMASK), __ldcg(g + i));
      }
      o += len;
      __syncwarp();
    }
    if (status == 0 && CONSUMED() > total_bits) status = Z_BUF;
    if (status == 0 && OPOS() > isize) status = Z_BUF;
  }

  // drain any TMA chunk still in flight before the CTA (and its shared memory) retires
  while (d.waited < d.issued) d.wait_chunk(d.waited, 0);

  if (status != 0 && status != Z_BUF && OPOS() > isize) status = Z_BUF;   // the output overran before the fault
  if (status == Z_DATA && CONSUMED() > total_bits) status = Z_BUF;   // zlib would have run out of input first
  if (status == 0 && OPOS() != isize) status = Z_DATA;   // stream ended short of ISIZE: -release BioD would hand out garbage (block.d:175); reported as a data error
  if (status == 0) {
    __syncwarp();
    if (OPOS() > flushed) flush_to(OPOS());
  }
  if (lane == 0) a.status[blk] = status;
#undef REFILL
#undef DROP
#undef CONSUMED
#undef OPOS
#undef MAYBE_FLUSH
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
