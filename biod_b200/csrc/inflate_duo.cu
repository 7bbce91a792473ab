// BGZF block inflater for sm_100a, two warps per BGZF block: a DECODER warp that turns the DEFLATE bit stream into
// tokens and a RESOLVER warp that turns tokens into bytes, working on consecutive pieces of the stream at the same time.
//
// Replaces decompressBgzfBlock (bio/core/bgzf/block.d:127-216), i.e. libz's inflateInit2(-15) / inflate(Z_FINISH) /
// inflateEnd on one <=64 KiB raw-DEFLATE payload.  The algorithm is RFC 1951; nothing here is derived from zlib.
//
// inflate_par.cu (one warp per block) showed where the time goes: every code was decoded 3.24 times — a speculative
// round that finds where the lanes' sub-sequences really start, 1.24 repair rounds, and one more full Huffman decode
// whose only job was to write the bytes now that the output offsets were known — and decoding and LZ77 copying
// alternated in the one warp.  Here:
//   * DECODER (warp 0): TMA staging of the payload, block headers, Huffman tables, and the lane-parallel decode of
//     "super-chunks" of 32 sub-sequences of SUB_BITS bits (see inflate_par.cu for the idea).  Round 1 only looks for the
//     synchronisation points (code lengths alone: the LUT entry carries the extra-bit count, no value is computed, nothing
//     is counted).  From round 2 on a lane decodes from where its predecessor ended and RECORDS what it decodes as 16-bit
//     tokens — literal / length / distance, one per loop trip — in a per-block scratch area in global memory (8 KB per
//     block, L2-resident, written 64 contiguous bytes per warp store).  When the chain of lanes is consistent the
//     per-lane byte / match / token counts go to shared memory and the resolver is signalled (named barrier); the decoder
//     goes straight on to the next super-chunk.
//   * RESOLVER (warp 1): scans the counts into output offsets, REPLAYS the tokens (literals into the shared-memory
//     output ring, matches into a list: ~10 instructions per token instead of a Huffman decode), releases the token
//     area, copies the LZ77 matches (in-ring sources in stream order, older sources read back from L2 all at once),
//     lets the record-chain walker (records.cu) look at the new bytes and flushes whole 128-byte lines to HBM.
// Anything unusual makes the block a STATUS_RETRY, redone by the warp-serial kernel (inflate.cu), which also produces
// zlib's exact error code — as in inflate_par.cu.
#include "inflate_common.cuh"

namespace biodb {

namespace {

#ifndef BIODB_DUO_SUB_BITS
#define BIODB_DUO_SUB_BITS 224
#endif
#ifndef BIODB_DUO_MIN_CTAS
#define BIODB_DUO_MIN_CTAS 16
#endif
#ifndef BIODB_DUO_MLIST
#define BIODB_DUO_MLIST 128
#endif
#ifndef BIODB_DUO_MAX_ROUNDS
#define BIODB_DUO_MAX_ROUNDS 5
#endif
constexpr int SUB_BITS = BIODB_DUO_SUB_BITS;   // bits of one lane's sub-sequence
constexpr int SUPER_BYTES = SUB_BITS * 4;      // compressed bytes of one nominal super-chunk
constexpr int NCH = 8;                         // chunks in the staging ring
constexpr int PIN_RING = 2048;
constexpr int CH = PIN_RING / NCH;             // bytes per TMA chunk
constexpr int PIN_WORDS = PIN_RING / 4;
constexpr int POUT = 4096;
constexpr uint32_t POM = POUT - 1;
// Output bytes one super-chunk may produce.  The ring must keep, besides them, the unflushed tail (< FLUSH_ALIGN), the
// longest match (258) for the "older than the ring => already flushed" rule, and ~1.1 KB of history for the walker.
constexpr int OUT_BUDGET = POUT - 1536;
constexpr int LANE_CAP = 512;                  // a lane stops taking codes once it has produced this many bytes ...
constexpr int MLIST = BIODB_DUO_MLIST;         // matches one super-chunk may hold
constexpr int LANE_MCAP = DUO_TOK_TRIPS / 2;   // (a lane stops after DUO_TOK_TRIPS - 1 tokens: at most this many matches)
constexpr int FLUSH_ALIGN = 128;
constexpr int SUB_CAP = LIT_BITS >= 10 ? 320 : 352;
constexpr int STORE_PIECE = 1024;              // stored blocks are copied in pieces of this many bytes
constexpr int HDR_BYTES = 640;                 // >= longest dynamic block header
constexpr int MAX_ROUNDS = BIODB_DUO_MAX_ROUNDS;   // decode rounds per super-chunk before the consistent prefix is committed as it is
static_assert(LANE_CAP > SUB_BITS, "literals alone never reach the cap, so it is checked between codes only");
static_assert(LANE_CAP + SUB_BITS + 257 <= OUT_BUDGET, "one lane must always fit");
static_assert(LANE_MCAP <= MLIST, "one lane must always fit");
static_assert(DUO_TOK_TRIPS >= 64 && DUO_TOK_TRIPS <= 255, "token count of a lane travels in 8 bits");   // ... or DUO_TOK_TRIPS - 1 tokens
static_assert((NCH - 1) * CH >= HDR_BYTES + 16 && (NCH - 1) * CH >= STORE_PIECE + 16 &&
                  (NCH - 1) * CH >= SUPER_BYTES + 32 && CH % 16 == 0,
              "staging ring too small");
static_assert(STORE_PIECE <= OUT_BUDGET, "");

// tokens (16 bits, one per decode trip of a lane)
constexpr uint32_t TOK_LEN = K_LEN << 13, TOK_EOB = K_EOB << 13, TOK_DIST = 0x8000;   // literal: the byte itself
// lane stop reasons
constexpr uint32_t F_EOB = 1, F_ERR = 2, F_INEND = 3;
// decoder -> resolver messages
enum { MSG_CHUNK = 0, MSG_STORED = 1, MSG_DONE = 2 };

struct __align__(16) DuoSmem {
  // ---- decoder warp ----
  uint32_t in_ring[PIN_WORDS + 4];     // + guard word (copy of word 0) so that a 64-bit window never wraps
  uint16_t lut_lit[1 << LIT_BITS];
  uint16_t lut_dist[1 << DIST_BITS];   // also hosts the 128-entry code-length LUT
  uint16_t sorted_dist[32];
  Code code_lit, code_dist;
  uint16_t sorted_lit[288];            // literal/length symbols in canonical order (build_table_par only)
  uint8_t lens[352];                   // [0,19) code-length code, [32,32+316) litlen+dist lengths
  uint32_t auxtab[64];                 // [0,32) length symbol, [32,64) distance symbol -> base value
  uint32_t scratch[16];                // build_table_par
  uint16_t sub_lit[SUB_CAP];           // second-level tables of the literal/length codes longer than LIT_BITS
  unsigned long long mbar[NCH];        // TMA completion, one per staging chunk
  // ---- decoder -> resolver ----
  uint32_t msg_kind, msg_a, msg_b, r_bad;
  uint32_t msg_lane[32];               // MSG_CHUNK: bytes | matches << 16 | tokens << 24 of every lane
  // ---- resolver warp ----
  uint8_t out_ring[POUT];
  uint32_t m_ld[MLIST];                // matches of the super-chunk: (length-3) | (distance-1) << 8
  uint16_t m_pos[MLIST];               // and their block-relative output offset
};

struct DCtx {             // shared-space addresses and limits every decoder lane needs
  uint32_t in_ring, lutl, lutd, auxtab, subl;
  uint32_t total_bits;
  const Code* code_dist;
  const uint16_t* sorted_dist;
  uint16_t* tok;          // this block's token area + lane
};

// Decoder <-> resolver hand-over on two named barriers (bar.arrive by the 32 lanes of the signalling warp + bar.sync by
// the 32 lanes of the waiting warp = 64 arrivals): a warp blocked in bar.sync is descheduled by the hardware and costs no
// issue slots.  (A first version polled an mbarrier with try_wait: the idle resolver warps then spent 18 % of the
// kernel's instructions spinning — ncu, profiles/ — and the kernel was no faster than its one-warp predecessor.)
constexpr int BAR_FULL = 1, BAR_FREE = 2;
__device__ __forceinline__ void bar_signal(int id) {
  __threadfence_block();                        // tokens (global) and message (shared) before the arrival
  asm volatile("bar.arrive %0, 64;" ::"r"(id) : "memory");
}
__device__ __forceinline__ void bar_await(int id) {
  asm volatile("bar.sync %0, 64;" ::"r"(id) : "memory");
}

// 32 bits of the staged stream starting at bit `pos`
__device__ __forceinline__ uint32_t fetch32(uint32_t in_ring, uint32_t pos) {
  const uint32_t a = in_ring + ((pos >> 3) & (uint32_t)(PIN_RING - 4));
  const uint32_t lo = lds32(a);
  const uint32_t hi = lds32(a + 4);         // the word after the last one of the ring is a copy of word 0 (guard)
  return __funnelshift_r(lo, hi, pos);
}

// Every lane with `active` decodes the codes that start in [t, limit) of its own sub-sequence; all 32 lanes of the
// decoder warp must call this together.  One loop trip decodes ONE Huffman code per lane, whichever kind the lane needs
// next — a literal/length code or the distance code of the length it met in the previous trip — as straight-line,
// select-based code, so that the lanes stay converged.
// RECORD = false: only the bit position moves (round 1: where do the sub-sequences synchronise?).
// RECORD = true: the lane also counts the bytes and matches it produces and writes one token per trip.
template <bool RECORD>
__device__ __forceinline__ void lane_decode(const DCtx& c, bool active, uint32_t t, uint32_t limit, uint32_t& end,
                                            uint32_t& out, uint32_t& nm, uint32_t& nt, uint32_t& flag) {
  const uint32_t lim = limit < c.total_bits ? limit : c.total_bits;
  uint32_t pos = t, o = 0, m = 0, fl = 0, trips = 0;
  uint32_t st = 0;          // 0: the next code is a literal/length code, 1: a distance code
  uint32_t len = 0;
  uint32_t lut = c.lutl, msk = ((1u << LIT_BITS) - 1) << 1;
  uint32_t run = (active && pos < lim) ? 1u : 0u;
  uint16_t* tp = c.tok;
  // The body is written with selects on 0/1 flags, not with if/else: whatever nvcc turns into branches here runs
  // divergently (a lane with a literal, a lane with a length, a lane with a distance) and costs every path's instructions.
  // A stopped lane (run == 0) keeps computing on its last position; nothing it computes is kept.
  while (__any_sync(0xffffffffu, run)) {
    const uint32_t bits = fetch32(c.in_ring, pos);
    uint32_t e = lds16(lut + ((bits << 1) & msk));
    if (run && (e & (3u << 8)) == (K_SPECIAL << 8)) {      // rare: code longer than the LUT index, or invalid
      if (e >> 12)                                         // second-level table of the literal/length code
        e = lds16(c.subl + ((((e & 0xff) << 1) + ((bits >> LIT_BITS) & ~(0xffffffffu << (e >> 12)))) << 1));
      else if (e == ENT_SLOW && st)                        // (every long literal/length code has a second-level table)
        e = slow_decode<DIST_BITS>(bits, c.code_dist, c.sorted_dist, KIND_DIST);
      if ((e & (3u << 8)) == (K_SPECIAL << 8)) {
        fl = F_ERR;
        run = 0;
      }
    }
    const uint32_t cl = e >> 12;
    const uint32_t kraw = (e >> 8) & 3;                            // kind of a literal/length entry (distance entries: 0)
    const uint32_t is_len = st | (kraw == K_LEN ? 1u : 0u);        // a distance code is handled like a length code
    const uint32_t eb = is_len ? ENTRY_EXTRA_BITS(e) : 0u;         // cl + eb <= 28 bits of the 32
    const uint32_t adv = cl + eb;
    pos += run ? adv : 0u;
    if (RECORD) {
      const uint32_t val = lds32(c.auxtab + (((st << 5) | (e & 31)) << 2)) + ((bits >> cl) & ~(0xffffffffu << eb));
      const uint32_t tk = (st ? TOK_DIST : (kraw << 13)) | (is_len ? val - (st ? 1u : 3u) : (e & 0xffu));
      if (run) *tp = (uint16_t)tk;
      tp += 32;
      trips += run;
      const uint32_t lit = (is_len | kraw) ? 0u : run;             // a literal: one byte
      const uint32_t mat = st & run;                               // a distance: the match is complete
      o += lit + (mat ? len : 0u);
      m += mat;
      len = val;
    }
    const uint32_t eob = (st ? 0u : (kraw == K_EOB ? 1u : 0u)) & run;
    fl = eob ? F_EOB : fl;
    st = (st ^ 1u) & is_len;                         // length -> distance next; anything else -> literal/length next
    lut = st ? c.lutd : c.lutl;
    msk = st ? ((1u << DIST_BITS) - 1) << 1 : ((1u << LIT_BITS) - 1) << 1;
    // a lane stops only between codes of the literal/length alphabet: at its boundary, or (RECORD) when it has produced
    // LANE_CAP bytes or DUO_TOK_TRIPS - 1 tokens (hence at most LANE_MCAP matches) — the next lane continues from there
    uint32_t stop = pos >= lim ? 1u : 0u;
    if (RECORD) stop |= (o >= (uint32_t)LANE_CAP ? 1u : 0u) | (trips >= (uint32_t)DUO_TOK_TRIPS - 1 ? 1u : 0u);
    run &= ~(eob | (st ? 0u : stop));
  }
  if (fl == 0 && active && pos >= c.total_bits && pos < limit) fl = F_INEND;   // ran out of input before its boundary
  end = pos;
  out = o;
  nm = m;
  nt = trips;
  flag = fl;
}

__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t n = __shfl_up_sync(0xffffffffu, v, d);
    if (lane >= d) v += n;
  }
  return v;
}

// -DBIODB_DUO_TIMING: cycles per phase, summed over blocks (lane 0 of each warp), read back through
// biodb_debug_inflate_counters slots of inflate_duo_cycles(): decoder [0] round 1, [1] waiting for the resolver,
// [2] rounds >= 2, [3] headers + tables, [4] waiting for input (TMA), [5] total; resolver [8] waiting for the decoder,
// [9] replay, [10] matches, [11] walker + flush, [12] total.
#ifdef BIODB_DUO_TIMING
__device__ unsigned long long g_duo_cycles[16];
#define T_DECL(n) long long _t_##n = 0
#define T_ON(n) _t_##n -= clock64()
#define T_OFF(n) _t_##n += clock64()
#define T_PUT(n, slot) atomicAdd(&g_duo_cycles[slot], (unsigned long long)_t_##n)
#else
#define T_DECL(n)
#define T_ON(n)
#define T_OFF(n)
#define T_PUT(n, slot)
#endif

// ====================================================================================== decoder warp ====
__device__ __noinline__ void duo_decoder(const InflateArgs& a, DuoSmem* s, const uint32_t blk, const int lane,
                                         unsigned long long* counters) {
  const uint64_t poff = a.payload_off[blk];
  const uint32_t csize = a.cdata_size[blk];
  const uint32_t isize = a.isize[blk];

  uint32_t sbase = smem_u32(s);
  asm volatile("mov.u32 %0, %0;" : "+r"(sbase));           // opaque: keeps the addresses in registers
  const uint32_t in_ring = sbase + (uint32_t)offsetof(DuoSmem, in_ring);
  const uint32_t lutd = sbase + (uint32_t)offsetof(DuoSmem, lut_dist);
  const uint32_t mbar = sbase + (uint32_t)offsetof(DuoSmem, mbar);
  DCtx ctx;
  ctx.in_ring = in_ring;
  ctx.lutl = sbase + (uint32_t)offsetof(DuoSmem, lut_lit);
  ctx.lutd = lutd;
  ctx.auxtab = sbase + (uint32_t)offsetof(DuoSmem, auxtab);
  ctx.subl = sbase + (uint32_t)offsetof(DuoSmem, sub_lit);
  ctx.code_dist = &s->code_dist;
  ctx.sorted_dist = s->sorted_dist;
  ctx.tok = a.tok + (size_t)blk * (DUO_TOK_TRIPS * 32) + lane;

  // ---- staging of the compressed payload (TMA) -------------------------------------------------------------
  const uint8_t* pay = a.comp + poff;
  const uint32_t skip = (uint32_t)(((uintptr_t)pay) & 15);
  const uint8_t* src = pay - skip;                         // 16-byte aligned start of the staged stream
  const uint32_t staged = skip + csize;                    // bytes from src that matter
  uint32_t n_chunks = (staged + CH - 1) / CH;
  if (n_chunks == 0) n_chunks = 1;
  uint32_t last_bytes = (staged - (n_chunks - 1) * CH + 15) & ~15u;
  if (last_bytes == 0) last_bytes = 16;
  uint32_t issued = 0, waited = 0;
  auto issue = [&](uint32_t k) {
    __syncwarp();        // every lane has finished reading the slot being overwritten (calls are warp-uniform)
    if (lane == 0) {
      const uint32_t bytes = (k + 1 == n_chunks) ? last_bytes : (uint32_t)CH;
      const uint32_t bar = mbar + 8 * (k % NCH);
      mbar_expect_tx(bar, bytes);
      tma_load(in_ring + (k % NCH) * CH, src + (size_t)k * CH, bytes, bar);
    }
    issued = k + 1;
  };
  auto wait_chunk = [&](uint32_t k) {
    mbar_wait(mbar + 8 * (k % NCH), (k / NCH) & 1);
    waited = k + 1;
    if (k % NCH == 0) {      // slot 0 has new bytes: refresh the guard word behind the ring
      if (lane == 0) sts32(in_ring + PIN_RING, lds32(in_ring));
      __syncwarp();
    }
  };
  // make staged bytes [lo, hi) readable; bytes before lo are not needed any more (the decoder only moves forward)
  auto ensure_input = [&](uint32_t lo, uint32_t hi) {
    const uint32_t c0 = lo / CH;
    uint32_t c1 = (hi - 1) / CH;
    if (c1 >= n_chunks) c1 = n_chunks - 1;
    while (issued < n_chunks && issued < c0 + NCH) issue(issued);
    while (waited <= c1 && waited < issued) wait_chunk(waited);
  };

  // ---- messages to the resolver ---------------------------------------------------------------------------------
  bool owe = false;         // a message is out whose consumption has not been waited for
  auto wait_free = [&]() {
    if (owe) {
      bar_await(BAR_FREE);
      owe = false;
    }
  };
  auto send = [&](uint32_t kind, uint32_t x, uint32_t y) {     // msg_lane[] (if any) is already written
    if (lane == 0) { s->msg_kind = kind; s->msg_a = x; s->msg_b = y; }
    bar_signal(BAR_FULL);
    owe = true;
  };

  // all bit positions are relative to src
  uint32_t pos = skip * 8;
  const uint32_t total_bits = staged * 8;
  ctx.total_bits = total_bits;
  int status = 0;
  uint32_t n_super = 0, n_rounds = 0, n_dblocks = 0;
  uint32_t produced = 0;    // bytes handed to the resolver so far

  T_DECL(r1); T_DECL(wf); T_DECL(r2); T_DECL(hd); T_DECL(in); T_DECL(tot);
  T_ON(tot);
  bool last = false;
  while (!last && status == 0) {
    // ---- block header (warp-uniform, from a 64-bit register bit buffer) ----------------------------------
    T_ON(in);
    ensure_input(pos >> 3, (pos >> 3) + HDR_BYTES);
    T_OFF(in);
    T_ON(hd);
    ++n_dblocks;
    uint64_t bb;
    int bc;
    uint32_t hw = pos >> 5;
    {
      const uint32_t lo = lds32(in_ring + ((hw & (PIN_WORDS - 1)) << 2));
      const uint32_t hi = lds32(in_ring + (((hw + 1) & (PIN_WORDS - 1)) << 2));
      bb = (((uint64_t)hi << 32) | lo) >> (pos & 31);
      bc = 64 - (int)(pos & 31);
      hw += 2;
    }
#define HFILL() do { if (bc <= 32) { bb |= (uint64_t)lds32(in_ring + ((hw & (PIN_WORDS - 1)) << 2)) << bc; bc += 32; ++hw; } } while (0)
#define HDROP(n) do { bb >>= (n); bc -= (n); } while (0)
#define HPOS() (hw * 32 - (uint32_t)bc)
    last = bb & 1;
    const int btype = (int)((bb >> 1) & 3);
    HDROP(3);
    if (btype == 3) { status = STATUS_RETRY; T_OFF(hd); break; }

    if (btype == 0) {
      T_OFF(hd);
      // ---- stored block: the resolver copies the bytes out of the staging ring, piece by piece -----------
      pos = (HPOS() + 7) & ~7u;
      ensure_input(pos >> 3, (pos >> 3) + 4);
      const uint32_t lw = fetch32(in_ring, pos);
      const uint32_t len = lw & 0xffff, nlen = lw >> 16;
      pos += 32;
      if (pos > total_bits || (len ^ 0xffff) != nlen || pos + len * 8 > total_bits || produced + len > isize) {
        status = STATUS_RETRY;
        break;
      }
      uint32_t left = len;
      while (left) {
        const uint32_t piece = left < (uint32_t)STORE_PIECE ? left : (uint32_t)STORE_PIECE;
        const uint32_t b0 = pos >> 3;
        ensure_input(b0, b0 + piece);
        wait_free();
        send(MSG_STORED, b0, piece);
        wait_free();                       // the staging slots of this piece may be recycled only after the copy
        produced += piece;
        pos += piece * 8;
        left -= piece;
      }
      continue;
    }

    if (btype == 1) {
      // ---- fixed Huffman code (RFC 1951 §3.2.6) ---------------------------------------
      pos = HPOS();
      for (int i = lane; i < 288; i += 32) s->lens[i] = i < 144 ? 8 : i < 256 ? 9 : i < 280 ? 7 : 8;
      __syncwarp();
      build_table_par<LIT_BITS>(s->lens, 288, s->lut_lit, s->sorted_lit, &s->code_lit, KIND_LITLEN, lane, s->scratch, s->sub_lit, SUB_CAP);
      __syncwarp();
      s->lens[lane] = 5;
      __syncwarp();
      build_table_par<DIST_BITS>(s->lens, 32, s->lut_dist, s->sorted_dist, &s->code_dist, KIND_DIST, lane, s->scratch);
    } else {
      // ---- dynamic Huffman code (RFC 1951 §3.2.7) --------------------------------------
      HFILL();
      const int hlit = (int)(bb & 31) + 257;
      const int hdist = (int)((bb >> 5) & 31) + 1;
      const int hclen = (int)((bb >> 10) & 15) + 4;
      HDROP(14);
      if (hlit > 286 || hdist > 30) { status = STATUS_RETRY; break; }
      if (lane < 19) s->lens[lane] = 0;
      __syncwarp();
      for (int i = 0; i < hclen; ++i) {
        HFILL();
        // order of code-length code lengths, RFC 1951 §3.2.7, packed 5 bits each
        const uint64_t ord_lo = 16ull | 17ull << 5 | 18ull << 10 | 0ull << 15 | 8ull << 20 | 7ull << 25 | 9ull << 30 |
                                6ull << 35 | 10ull << 40 | 5ull << 45 | 11ull << 50 | 4ull << 55;
        const uint64_t ord_hi = 12ull | 3ull << 5 | 13ull << 10 | 2ull << 15 | 14ull << 20 | 1ull << 25 | 15ull << 30;
        const int sym = i < 12 ? (int)((ord_lo >> (5 * i)) & 31) : (int)((ord_hi >> (5 * (i - 12))) & 31);
        s->lens[sym] = (uint8_t)(bb & 7);
        HDROP(3);
      }
      __syncwarp();
      int r = build_table_par<CL_BITS>(s->lens, 19, s->lut_dist, s->sorted_dist, &s->code_dist, KIND_CODELEN, lane, s->scratch);
      if (r != 0) { status = STATUS_RETRY; break; }
      const int total = hlit + hdist;
      __syncwarp();
      int idx = 0;
      int prev = 0;
      while (idx < total) {
        HFILL();
        const uint32_t e = lds16(lutd + (((uint32_t)bb & ((1u << CL_BITS) - 1)) << 1));
        const int cl = e >> 12;
        if (cl == 0 || ((e >> 8) & 3) == K_SPECIAL) { status = STATUS_RETRY; break; }
        const int sym = e & 31;
        HDROP(cl);
        if (sym < 16) {
          if (lane == 0) s->lens[32 + idx] = (uint8_t)sym;
          prev = sym;
          ++idx;
        } else {
          int rep, val;
          if (sym == 16) {
            if (idx == 0) { status = STATUS_RETRY; break; }
            rep = 3 + (int)(bb & 3);
            HDROP(2);
            val = prev;
          } else if (sym == 17) {
            rep = 3 + (int)(bb & 7);
            HDROP(3);
            val = 0;
          } else {
            rep = 11 + (int)(bb & 127);
            HDROP(7);
            val = 0;
          }
          if (idx + rep > total) { status = STATUS_RETRY; break; }
          for (int k = lane; k < rep; k += 32) s->lens[32 + idx + k] = (uint8_t)val;
          prev = val;
          idx += rep;
        }
      }
      if (status) break;
      pos = HPOS();
      if (pos > total_bits) { status = STATUS_RETRY; break; }
      __syncwarp();
      if (s->lens[32 + 256] == 0) { status = STATUS_RETRY; break; }   // no end-of-block code
      __syncwarp();
      r = build_table_par<LIT_BITS>(s->lens + 32, hlit, s->lut_lit, s->sorted_lit, &s->code_lit, KIND_LITLEN, lane, s->scratch,
                                    s->sub_lit, SUB_CAP);
      if (r != 0) { status = STATUS_RETRY; break; }
      r = build_table_par<DIST_BITS>(s->lens + 32 + hlit, hdist, s->lut_dist, s->sorted_dist, &s->code_dist, KIND_DIST, lane, s->scratch);
      if (r < 0) { status = STATUS_RETRY; break; }
    }
    __syncwarp();
    T_OFF(hd);
#undef HFILL
#undef HDROP
#undef HPOS

    // ---- the codes of the block, one super-chunk of 32 sub-sequences at a time -------------------------------
    bool eob = false;
    while (!eob) {
      if (s->r_bad) { status = STATUS_RETRY; break; }        // the resolver met a distance that reaches before the block
      const uint32_t base = pos;
      T_ON(in);
      ensure_input(base >> 3, (base >> 3) + SUPER_BYTES + 24);
      T_OFF(in);
      const uint32_t lim = base + (uint32_t)(lane + 1) * SUB_BITS;
      uint32_t t = base + (uint32_t)lane * SUB_BITS;
      uint32_t e_, out_, nm_, nt_, fl_;
      // round 1: where does the chain cross into each sub-sequence?  (bit positions only)
      T_ON(r1);
      lane_decode<false>(ctx, true, t, lim, e_, out_, nm_, nt_, fl_);
      T_OFF(r1);
      ++n_super;
      // round 2: every lane from where its predecessor ended, recording tokens — the token area must be free
      T_ON(wf);
      wait_free();
      T_OFF(wf);
      T_ON(r2);
      {
        uint32_t tn = __shfl_up_sync(0xffffffffu, e_, 1);
        if (lane == 0) tn = base;
        t = tn;
        lane_decode<true>(ctx, true, t, lim, e_, out_, nm_, nt_, fl_);
      }
      uint32_t rounds = 2;
      // repair the chain: lane L must start where lane L-1 ended
      uint32_t kstop = 32;      // first lane of the consistent prefix that stopped (end of block / fault), or 32
      uint32_t vcut = 32;       // lanes of the consistent prefix when the rounds ran out
      while (true) {
        uint32_t tn = __shfl_up_sync(0xffffffffu, e_, 1);
        if (lane == 0) tn = base;
        const bool need = tn != t;
        const uint32_t needm = __ballot_sync(0xffffffffu, need);
        const uint32_t stopm = __ballot_sync(0xffffffffu, fl_ != 0);
        const uint32_t vp = needm ? (uint32_t)__ffs(needm) - 1 : 32;      // lanes [0, vp) form a consistent chain
        const uint32_t vstop = stopm & (vp >= 32 ? 0xffffffffu : ((1u << vp) - 1));
        if (vstop) { kstop = (uint32_t)__ffs(vstop) - 1; break; }
        if (!needm) break;
        if (rounds >= (uint32_t)MAX_ROUNDS) { vcut = vp; break; }         // vp >= 1: lane 0 never needs a repair
        {
          uint32_t e2_, o2_, m2_, t2_, f2_;
          if (need) t = tn;
          lane_decode<true>(ctx, need, t, lim, e2_, o2_, m2_, t2_, f2_);
          if (need) { e_ = e2_; out_ = o2_; nm_ = m2_; nt_ = t2_; fl_ = f2_; }
        }
        ++rounds;
      }
      T_OFF(r2);
      n_rounds += rounds;
      // commit the longest prefix of lanes that fits the output ring and the match list
      const uint32_t ncand = kstop < 32 ? kstop + 1 : vcut;
      const uint32_t inc_out = warp_incl_scan(out_, lane);
      const uint32_t inc_nm = warp_incl_scan(nm_, lane);
      const bool fits = (uint32_t)lane < ncand && inc_out <= (uint32_t)OUT_BUDGET && inc_nm <= (uint32_t)MLIST;
      const uint32_t k = (uint32_t)__popc(__ballot_sync(0xffffffffu, fits));   // >= 1: lane 0 always fits
      const uint32_t kl = k - 1;
      const uint32_t chunk_out = __shfl_sync(0xffffffffu, inc_out, kl);
      const uint32_t newpos = __shfl_sync(0xffffffffu, e_, kl);
      const uint32_t stop_flag = (kl == kstop) ? __shfl_sync(0xffffffffu, fl_, kl) : 0;
      if (stop_flag == F_ERR || stop_flag == F_INEND || produced + chunk_out > isize) { status = STATUS_RETRY; break; }
      s->msg_lane[lane] = out_ | (nm_ << 16) | (nt_ << 24);
      send(MSG_CHUNK, k, 0);
      produced += chunk_out;
      pos = newpos;
      eob = stop_flag == F_EOB;
    }
  }

  // drain any TMA chunk still in flight before the CTA (and its shared memory) retires
  while (waited < issued) wait_chunk(waited);

  if (status == 0 && (produced != isize || pos > total_bits)) status = STATUS_RETRY;
  T_ON(wf);
  wait_free();
  T_OFF(wf);
  send(MSG_DONE, (uint32_t)status, 0);
  T_OFF(tot);
  if (lane == 0) {
    T_PUT(r1, 0); T_PUT(wf, 1); T_PUT(r2, 2); T_PUT(hd, 3); T_PUT(in, 4); T_PUT(tot, 5);
    atomicAdd(&counters[1], (unsigned long long)n_super);
    atomicAdd(&counters[2], (unsigned long long)n_rounds);
    atomicAdd(&counters[5], (unsigned long long)n_dblocks);
  }
}

// ===================================================================================== resolver warp ====
__device__ __noinline__ void duo_resolver(const InflateArgs& a, DuoSmem* s, const uint32_t blk, const int lane,
                                          unsigned long long* counters) {
  const uint32_t isize = a.isize[blk];
  const uint64_t obase = a.out_off[blk];
  uint8_t* gout = a.out + obase;
  const uint32_t oa = (uint32_t)(((uintptr_t)gout) & POM);   // ring index of output byte 0

  uint32_t sbase = smem_u32(s);
  asm volatile("mov.u32 %0, %0;" : "+r"(sbase));
  const uint32_t ring = sbase + (uint32_t)offsetof(DuoSmem, out_ring);
  const uint32_t in_ring = sbase + (uint32_t)offsetof(DuoSmem, in_ring);
  const uint32_t mld = sbase + (uint32_t)offsetof(DuoSmem, m_ld);
  const uint32_t mpos = sbase + (uint32_t)offsetof(DuoSmem, m_pos);
  const uint16_t* tok = a.tok + (size_t)blk * (DUO_TOK_TRIPS * 32) + lane;

  uint32_t o = oa;          // oa + bytes produced: ring index is (o & POM)
  uint32_t flushed = 0;     // bytes already stored to HBM
#define OPOS() (o - oa)
  Walker<POUT> wk;
  wk.init(a.walk, ring, oa, isize, obase, blk);

  auto flush_to = [&](uint32_t fe) {
    // copy ring bytes [flushed, fe) to HBM; 16-byte vector stores where the global address allows
    uint32_t f = flushed;
    uint32_t head = (16 - ((oa + f) & 15)) & 15;
    if (head > fe - f) head = fe - f;
    if (head) {
      if ((uint32_t)lane < head) gout[f + lane] = (uint8_t)lds8(ring + ((oa + f + lane) & POM));
      f += head;
    }
    const uint32_t n16 = (fe - f) >> 4;
    for (uint32_t i = lane; i < n16; i += 32) {
      uint4 v = lds128(ring + ((oa + f + 16 * i) & POM));
      __stcs(reinterpret_cast<uint4*>(gout + f + 16 * i), v);
    }
    f += n16 << 4;
    const uint32_t tail = fe - f;
    if ((uint32_t)lane < tail) gout[f + lane] = (uint8_t)lds8(ring + ((oa + f + lane) & POM));
    flushed = fe;
    __syncwarp();
  };
  // after new bytes are complete in the ring: let the record walker see them, flush whole 128-byte lines
  auto produced = [&]() {
    wk.walk_upto(OPOS(), false, lane);
    const uint32_t fe = OPOS() - (o & (FLUSH_ALIGN - 1));
    if (fe > flushed && fe <= OPOS()) flush_to(fe);
  };
  auto release = [&]() { bar_signal(BAR_FREE); };   // message, tokens and stored input bytes are consumed

  uint32_t n_far = 0, n_matches = 0;
  bool bad = false;
  int dstatus = 0;
  T_DECL(wm); T_DECL(rp); T_DECL(mt); T_DECL(fw); T_DECL(rt);
  T_ON(rt);
  while (true) {
    T_ON(wm);
    bar_await(BAR_FULL);
    T_OFF(wm);
    const uint32_t kind = s->msg_kind, ma = s->msg_a, mb = s->msg_b;
    if (kind == MSG_DONE) { dstatus = (int)ma; break; }
    if (kind == MSG_STORED) {
      const uint32_t b0 = ma, piece = mb;
      if (!bad)
        for (uint32_t i = lane; i < piece; i += 32)
          sts8(ring + ((o + i) & POM), lds8(in_ring + ((b0 + i) & (PIN_RING - 1))));
      release();
      o += piece;
      if (!bad) produced();
      continue;
    }
    // ---- MSG_CHUNK: ma lanes of the decoder's super-chunk are committed -----------------------------------
    const uint32_t k = ma;
    const uint32_t mine = (uint32_t)lane < k ? s->msg_lane[lane] : 0;
    const uint32_t out_ = mine & 0xffff, nm_ = (mine >> 16) & 0xff, nt_ = mine >> 24;
    const uint32_t inc_out = warp_incl_scan(out_, lane);
    const uint32_t inc_nm = warp_incl_scan(nm_, lane);
    const uint32_t chunk_out = __shfl_sync(0xffffffffu, inc_out, 31);
    const uint32_t n_match = __shfl_sync(0xffffffffu, inc_nm, 31);
    const uint32_t max_nt = __reduce_max_sync(0xffffffffu, nt_);
    const uint32_t opos0 = OPOS();
    T_ON(rp);
    if (!bad) {
      // replay: literals into the ring, matches into the list
      uint32_t oo = o + (inc_out - out_);               // ring-relative position of this lane's next byte
      uint32_t slot = inc_nm - nm_;
      uint32_t len = 0;
      for (uint32_t t0 = 0; t0 < max_nt; t0 += 8) {
        uint32_t tk[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) tk[j] = (t0 + j < nt_) ? (uint32_t)__ldcg(tok + (size_t)(t0 + j) * 32) : TOK_EOB;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint32_t v = tk[j];
          if (v & TOK_DIST) {
            sts32(mld + (slot << 2), (len - 3) | ((v & 0x7fffu) << 8));
            sts16(mpos + (slot << 1), oo - oa);
            ++slot;
            oo += len;
          } else if (v & TOK_LEN) {
            len = (v & 0xff) + 3;
          } else if (!(v & TOK_EOB)) {
            sts8(ring + (oo & POM), v);
            ++oo;
          }
        }
      }
    }
    release();
    T_OFF(rp);
    T_ON(mt);
    if (!bad) {
      // LZ77 copies, 32 list entries at a time (one per lane, handed around by shuffles)
      const uint32_t opos_end = opos0 + chunk_out;
      const uint32_t ring_lo = opos_end > (uint32_t)POUT ? opos_end - (uint32_t)POUT : 0;   // oldest byte still in the ring
      for (uint32_t j0 = 0; j0 < n_match; j0 += 32) {
        const uint32_t j = j0 + lane;
        const bool valid = j < n_match;
        uint32_t ld = 0, myp = 0x10000;
        if (valid) { ld = lds32(mld + (j << 2)); myp = lds16(mpos + (j << 1)); }
        const uint32_t mylen = (ld & 255) + 3, mydist = (ld >> 8) + 1;
        if (__any_sync(0xffffffffu, mydist > myp)) { bad = true; break; }      // distance too far back
        const uint32_t mysp = myp - mydist;
        const bool far = valid && mysp < ring_lo;
        // (a) sources older than the ring, therefore already flushed (and dist > len): read the block's own output
        //     back from L2.  They depend on nothing in flight, so all of them go at once: the bytes of these matches
        //     are numbered consecutively and dealt out to the lanes, 32 bytes per trip.
        const uint32_t farm = __ballot_sync(0xffffffffu, far);
        if (farm) {
          n_far += (uint32_t)__popc(farm);
          const uint32_t flen = far ? mylen : 0;
          const uint32_t incl = warp_incl_scan(flen, lane);
          const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
          for (uint32_t b0 = 0; b0 < total; b0 += 32) {
            const uint32_t b = b0 + lane;
            uint32_t kk = 0;                     // number of lanes whose bytes all come before byte b
#pragma unroll
            for (uint32_t step = 16; step; step >>= 1) {
              const uint32_t v = __shfl_sync(0xffffffffu, incl, (kk + step - 1) & 31);
              if (v <= b) kk += step;
            }
            kk &= 31;
            const uint32_t off = b - (__shfl_sync(0xffffffffu, incl, kk) - __shfl_sync(0xffffffffu, flen, kk));
            const uint32_t srcp = __shfl_sync(0xffffffffu, mysp, kk) + off;
            const uint32_t dstp = __shfl_sync(0xffffffffu, myp, kk) + off;
            if (b < total) sts8(ring + ((oa + dstp) & POM), __ldcg(gout + srcp));
          }
          __syncwarp();
        }
        // (b) sources inside the ring: in stream order, the whole warp on one match
        uint32_t nearm = __ballot_sync(0xffffffffu, valid && !far);
        while (nearm) {
          const uint32_t kk = (uint32_t)__ffs(nearm) - 1;
          nearm &= nearm - 1;
          const uint32_t len = __shfl_sync(0xffffffffu, mylen, kk), dist = __shfl_sync(0xffffffffu, mydist, kk);
          const uint32_t dr = oa + __shfl_sync(0xffffffffu, myp, kk);           // ring-relative destination
          const uint32_t sr = dr - dist;
          if (dist >= len) {
            if ((uint32_t)lane < len) sts8(ring + ((dr + lane) & POM), lds8(ring + ((sr + lane) & POM)));
            for (uint32_t i = lane + 32; i < len; i += 32) sts8(ring + ((dr + i) & POM), lds8(ring + ((sr + i) & POM)));
          } else {
            uint32_t m = (uint32_t)lane % dist;
            const uint32_t step = 32u % dist;
            for (uint32_t i = lane; i < len; i += 32) {
              sts8(ring + ((dr + i) & POM), lds8(ring + ((sr + m) & POM)));
              m += step;
              if (m >= dist) m -= dist;
            }
          }
          __syncwarp();
        }
      }
      if (bad && lane == 0) s->r_bad = 1;      // the decoder stops at its next super-chunk
      n_matches += n_match;
    }
    T_OFF(mt);
    o += chunk_out;
    T_ON(fw);
    if (!bad) produced();
    T_OFF(fw);
  }
  T_OFF(rt);

  const int status = (dstatus != 0 || bad) ? STATUS_RETRY : 0;
  if (status == 0) {
    wk.walk_upto(isize, true, lane);
    if (OPOS() > flushed) flush_to(OPOS());
  }
  if (lane == 0) {
    a.status[blk] = status;
    if (status == 0) wk.store(0);
    if (status) atomicAdd(&counters[0], 1ull);
    atomicAdd(&counters[3], (unsigned long long)n_far);
    atomicAdd(&counters[4], (unsigned long long)n_matches);
    T_PUT(wm, 8); T_PUT(rp, 9); T_PUT(mt, 10); T_PUT(fw, 11); T_PUT(rt, 12);
  }
#undef OPOS
}

}  // namespace

// diagnostics (biodb_debug_inflate_counters), same slots as inflate_par.cu's
__device__ unsigned long long g_duo_counters[8];

__global__ void __launch_bounds__(64, BIODB_DUO_MIN_CTAS) inflate_duo_kernel(InflateArgs a) {
  __shared__ DuoSmem sm;
  DuoSmem* s = &sm;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t blk = blockIdx.x;
  if (blk >= a.n_blocks) return;
  if (threadIdx.x == 0) s->r_bad = 0;
  if (threadIdx.x < NCH) mbar_init(smem_u32(&s->mbar[threadIdx.x]), 1);
  if (warp == 0) {
    uint32_t b, eb;
    s->auxtab[lane] = 0;
    s->auxtab[32 + lane] = 0;
    if (lane < 29) { len_base((uint32_t)lane, b, eb); s->auxtab[lane] = b; }
    if (lane < 30) { dist_base((uint32_t)lane, b, eb); s->auxtab[32 + lane] = b; }
  }
  if (threadIdx.x == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  __syncthreads();
  if (warp == 0) duo_decoder(a, s, blk, lane, g_duo_counters);
  else duo_resolver(a, s, blk, lane, g_duo_counters);
}

int inflate_duo_resident_blocks(int device) {
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inflate_duo_kernel, 64, 0) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  return per_sm * sms;
}

cudaError_t inflate_duo_counters(unsigned long long* out8, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out8, g_duo_counters, sizeof(g_duo_counters));
  if (e == cudaSuccess && reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyToSymbol(g_duo_counters, z, sizeof(z));
  }
  return e;
}

// phase cycle counters of a -DBIODB_DUO_TIMING build (zeros otherwise)
cudaError_t inflate_duo_cycles(unsigned long long* out16, int reset) {
#ifdef BIODB_DUO_TIMING
  cudaError_t e = cudaMemcpyFromSymbol(out16, g_duo_cycles, sizeof(g_duo_cycles));
  if (e == cudaSuccess && reset) {
    unsigned long long z[16] = {0};
    e = cudaMemcpyToSymbol(g_duo_cycles, z, sizeof(z));
  }
  return e;
#else
  for (int i = 0; i < 16; ++i) out16[i] = 0;
  return cudaSuccess;
#endif
}

size_t inflate_duo_token_bytes(uint32_t n_blocks) { return (size_t)n_blocks * DUO_TOK_TRIPS * 32 * sizeof(uint16_t); }

cudaError_t launch_inflate_duo(const InflateArgs& a, cudaStream_t st) {
  if (a.n_blocks == 0) return cudaSuccess;
  inflate_duo_kernel<<<a.n_blocks, 64, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace biodb
