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
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "inflate_common.cuh"

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

}  // namespace

// retry_only != 0: redo only the blocks that the decode / resolve kernels gave up on (status == STATUS_RETRY)
__global__ void __launch_bounds__(32) inflate_kernel(InflateArgs a, int retry_only) {
  __shared__ WarpSmem sm;
  WarpSmem* s = &sm;
  const int lane = threadIdx.x;
  const uint32_t blk = blockIdx.x;
  if (blk >= a.n_blocks) return;
  if (retry_only && a.status[blk] != STATUS_RETRY) return;

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
  Walker<OUT_RING> wk;
  wk.init(a.walk, ring, oa, isize, obase, blk);
  const bool walking = wk.walking;
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
      if (walking) wk.walk_upto(OPOS(), false, lane);
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
    if (walking) wk.walk_upto(isize, true, lane);
    if (OPOS() > flushed) flush_to(OPOS());
  }
  if (lane == 0) {
    a.status[blk] = status;
    wk.store(status);
  }
#undef REFILL
#undef DROP
#undef CONSUMED
#undef OPOS
}

unsigned long long g_kernel_launches = 0;

size_t inflate_smem_bytes() { return sizeof(WarpSmem); }

// Which kernels inflate: the decode + resolve pair (inflate_tok.cu) followed by this file's warp-serial kernel on the
// blocks they gave up on, unless BIODB_INFLATE=serial selects the warp-serial kernel alone (A/B measurements).
enum { MODE_TOK = 0, MODE_SERIAL = 1 };
static int inflate_mode() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("BIODB_INFLATE");
    v = (e && e[0] == 's') ? MODE_SERIAL : MODE_TOK;
  }
  return v;
}

int inflate_resident_blocks(int device) { return inflate_tok_resident_blocks(device); }

size_t inflate_token_bytes(uint32_t n_blocks) { return inflate_mode() == MODE_TOK ? inflate_tok_token_bytes(n_blocks) : 0; }

cudaError_t inflate_counters(unsigned long long* out8, int reset) { return inflate_tok_counters(out8, reset); }

static InflateArgs slice_args(const InflateArgs& a, uint32_t b0, uint32_t n) {
  InflateArgs s = a;
  s.payload_off += b0; s.cdata_size += b0; s.out_off += b0; s.isize += b0; s.status += b0;
  s.n_blocks = n;
  return s;
}

cudaError_t launch_inflate(const InflateArgs& a, cudaStream_t st) {
  if (a.n_blocks == 0) return cudaSuccess;
  if (inflate_mode() == MODE_SERIAL) {
    inflate_kernel<<<a.n_blocks, 32, 0, st>>>(a, 0);
    ++g_kernel_launches;
    return cudaGetLastError();
  }
  cudaError_t e = cudaSuccess;
  if (a.tok) {
    e = launch_inflate_tok(a, st);
    g_kernel_launches += 2;
  } else {
    // callers without a token area of their own (the device-resident stage API, no fused record walk): one is allocated
    // for the call — stream-ordered, from a pool of this library's own that keeps its memory between calls — and the
    // blocks go through it in slabs of three waves of the decode kernel
    static cudaMemPool_t pools[64] = {};
    static std::mutex pool_mu;
    int dev = 0;
    if ((e = cudaGetDevice(&dev)) != cudaSuccess) return e;
    cudaMemPool_t pool = nullptr;
    {
      std::lock_guard<std::mutex> lk(pool_mu);
      if (dev >= 0 && dev < 64 && !pools[dev]) {
        cudaMemPoolProps props = {};
        props.allocType = cudaMemAllocationTypePinned;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        if (cudaMemPoolCreate(&pools[dev], &props) == cudaSuccess) {
          uint64_t keep = ~0ull;
          cudaMemPoolSetAttribute(pools[dev], cudaMemPoolAttrReleaseThreshold, &keep);
        } else {
          pools[dev] = nullptr;
          cudaGetLastError();
        }
      }
      if (dev >= 0 && dev < 64) pool = pools[dev];
    }
    const uint32_t slab = std::min<uint32_t>(a.n_blocks, std::max<uint32_t>(1u, 3u * (uint32_t)std::max(0, inflate_resident_blocks(dev))));
    void* tmp = nullptr;
    e = pool ? cudaMallocFromPoolAsync(&tmp, inflate_tok_token_bytes(slab), pool, st)
             : cudaMallocAsync(&tmp, inflate_tok_token_bytes(slab), st);
    if (e != cudaSuccess) return e;
    for (uint32_t b0 = 0; b0 < a.n_blocks && e == cudaSuccess; b0 += slab) {
      InflateArgs s = slice_args(a, b0, std::min<uint32_t>(slab, a.n_blocks - b0));
      s.tok = (uint16_t*)tmp;
      e = launch_inflate_tok(s, st);
      g_kernel_launches += 2;
    }
    cudaFreeAsync(tmp, st);
  }
  if (e != cudaSuccess) return e;
  inflate_kernel<<<a.n_blocks, 32, 0, st>>>(a, 1);      // redo what they gave up on (status STATUS_RETRY)
  g_kernel_launches += 1;
  return cudaGetLastError();
}

}  // namespace biodb
