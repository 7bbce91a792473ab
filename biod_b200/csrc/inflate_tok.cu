// BGZF block inflater for sm_100a in two kernels: DECODE turns the DEFLATE bit stream of every block into tokens,
// RESOLVE turns the tokens into bytes.
//
// Replaces decompressBgzfBlock (bio/core/bgzf/block.d:127-216), i.e. libz's inflateInit2(-15) / inflate(Z_FINISH) /
// inflateEnd on one <=64 KiB raw-DEFLATE payload.  The algorithm is RFC 1951; nothing here is derived from zlib.
//
// The lane-parallel idea: Huffman decoding is serial by nature — the start of a code is known only once the code before
// it is decoded — but Huffman codes self-synchronise.  The bit stream of a DEFLATE block is cut into "super-chunks" of
// 32 sub-sequences; lane L decodes sub-sequence L from its nominal first bit, which is almost never the start of a
// code, yet after a few codes the lane is on the true code chain, and the bit where it crosses into sub-sequence L+1
// is that sub-sequence's true start.  A second round from the true starts is the real decode; lanes whose predecessor
// had not synchronised are repaired in further rounds (2.24 rounds per super-chunk of BAM data), at most five.
//
// Why two kernels.  Huffman decoding is one long dependency chain per warp: measured on the B200, a warp of the decode
// loop issues one instruction every ~10 cycles whatever else the SM does, so the throughput of an SM is the number of
// RESIDENT DECODING WARPS times that rate until the schedulers' integer pipes saturate.  Round 1's predecessors — one
// warp per block doing everything (11 KB of shared memory, 19 warps per SM, every code decoded 3.24 times: 93 GB/s of
// inflated bytes), then a decoder and a resolver warp per block in one kernel (16-18 blocks per SM, the resolver asleep
// half of the time: 108 GB/s) — ran at 55-65 % of the issue slots for that reason (git history).  Here
//   * inflate_decode_kernel (one warp per block) keeps only what decoding needs in shared memory — the TMA staging ring,
//     the Huffman tables — ~7 KB, so 28 blocks are resident per SM.  It decodes "super-chunks" of 32 sub-sequences of
//     SUB_BITS bits lane-parallel (above): round 1 only looks for the points where the
//     sub-sequences synchronise (code lengths alone: the LUT entry carries the extra-bit count); from round 2 on a lane
//     decodes from where its predecessor ended and RECORDS what it decodes as 16-bit tokens — literal / length /
//     distance, two per loop trip: its code and, behind a literal or a short distance, the literal/length code that
//     follows — straight into the block's record stream in global memory (128 contiguous bytes per warp store).  When the
//     chain of lanes is consistent it appends the super-chunk's header (lanes committed, per-lane byte / match / token
//     counts) and goes on; it never waits for anybody.  Sub-sequences shrink when a super-chunk's output would overflow
//     the resolver's ring (streams of long matches) and grow back; a stream whose codes have one length never
//     synchronises and is handed to the warp-serial kernel after sixteen super-chunks that got nowhere.
//   * inflate_resolve_kernel (one warp per block, ~5 KB of shared memory: 32 blocks per SM) walks the record stream:
//     scans the counts into output offsets, REPLAYS the tokens (literals into the shared-memory output ring, matches
//     into a list: ~10 instructions per token instead of a Huffman decode), copies the LZ77 matches (in-ring sources in
//     stream order, older sources read back from L2 all at once), lets the record-chain walker (records.cu) look at the
//     new bytes and flushes whole 128-byte lines to HBM.  Stored blocks are copied by this kernel straight from the
//     compressed payload.
// The price is the record stream: ~2 bytes per token written and read once (~2.4 x the block's own bytes for BAM data),
// on a kernel pair that uses a few per cent of the HBM bandwidth.
// Anything unusual — an invalid code, a distance too far back, a stream that does not end exactly at ISIZE, input that
// runs out, a record stream that outgrows its arena — makes the block a STATUS_RETRY, redone by the warp-serial kernel
// (inflate.cu), which also produces zlib's exact error code.
#include "inflate_common.cuh"

namespace biodb {

namespace {

#ifndef BIODB_TOK_SUB_BITS
#define BIODB_TOK_SUB_BITS 224
#endif
#ifndef BIODB_TOK_MLIST
#define BIODB_TOK_MLIST 128
#endif
#ifndef BIODB_TOK_MAX_ROUNDS
#define BIODB_TOK_MAX_ROUNDS 5
#endif
#ifndef BIODB_TOK_DECODE_CTAS
#define BIODB_TOK_DECODE_CTAS 28
#endif
constexpr int SUB_BITS = BIODB_TOK_SUB_BITS;   // bits of one lane's sub-sequence
constexpr int SUPER_BYTES = SUB_BITS * 4;      // compressed bytes of one nominal super-chunk
constexpr int NCH = 8;                         // chunks in the staging ring
constexpr int PIN_RING = 2048;
constexpr int CH = PIN_RING / NCH;             // bytes per TMA chunk
constexpr int PIN_WORDS = PIN_RING / 4;
constexpr int POUT = 4096;
constexpr uint32_t POM = POUT - 1;
// Output bytes one super-chunk may produce.  The resolver's ring must keep, besides them, the unflushed tail
// (< FLUSH_ALIGN), the longest match (258) for the "older than the ring => already flushed" rule, and ~1.1 KB of history
// for the walker.
constexpr int OUT_BUDGET = POUT - 1536;
constexpr int LANE_CAP = 512;                  // a lane stops taking codes once it has produced this many bytes ...
constexpr int MLIST = BIODB_TOK_MLIST;         // matches one super-chunk may hold
constexpr int TOK_TRIPS = TOK_MAX_TRIPS;       // ... or TOK_TRIPS - 1 tokens
constexpr int LANE_MCAP = TOK_TRIPS / 2;       // (hence at most this many matches)
constexpr int FLUSH_ALIGN = 128;
constexpr int SUB_CAP = LIT_BITS >= 10 ? 320 : 352;
constexpr int STORE_PIECE = 2048;              // stored blocks are copied in pieces of this many bytes
constexpr int HDR_BYTES = 640;                 // >= longest dynamic block header
constexpr int MAX_ROUNDS = BIODB_TOK_MAX_ROUNDS;   // decode rounds per super-chunk before the consistent prefix is committed as it is
static_assert(LANE_CAP > SUB_BITS, "literals alone never reach the cap, so it is checked between codes only");
static_assert(LANE_CAP + SUB_BITS + 257 <= OUT_BUDGET, "one lane must always fit");
static_assert(LANE_MCAP <= MLIST, "one lane must always fit");
static_assert(TOK_TRIPS >= 64 && TOK_TRIPS <= 255, "token count of a lane travels in 8 bits");
static_assert((NCH - 1) * CH >= HDR_BYTES + 16 && (NCH - 1) * CH >= SUPER_BYTES + 32 && CH % 16 == 0, "staging ring too small");
static_assert(STORE_PIECE <= OUT_BUDGET, "");

// tokens (16 bits): bit 15 = a distance (15 bits of distance - 1); else bits 13-14 = literal (the byte itself) / length
// (8 bits of length - 3) / end of block / none.  A decode trip of a lane yields two: its code, and the literal/length
// code behind it if the first was a literal (else TOK_NONE).
constexpr uint32_t TOK_LEN = K_LEN << 13, TOK_EOB = K_EOB << 13, TOK_NONE = 3u << 13, TOK_DIST = 0x8000;
// lane stop reasons
constexpr uint32_t F_EOB = 1, F_ERR = 2, F_INEND = 3;

// The record stream of a block: 16-bit words in its arena of TOK_ARENA_WORDS, records back to back, each a multiple of 32
// words (64 bytes).
//   REC_CHUNK : head[32] = {REC_CHUNK, lanes committed, token rows, ...}; lane[32] x u32 = bytes | matches << 16 |
//               rows << 24 of every lane; then `rows` rows of 32 x u32 (row t = the two tokens of trip t of lanes 0..31)
//   REC_STORED: head[32] = {REC_STORED, -, -, -, offset lo, offset hi, length lo, length hi}: `length` bytes at `offset`
//               of the block's compressed payload
//   REC_DONE  : head[32] = {REC_DONE, status != 0}
enum { REC_CHUNK = 1, REC_STORED = 2, REC_DONE = 3 };
constexpr uint32_t REC_HEAD = 32, REC_LANES = 64, ROW_WORDS = 64;
constexpr uint32_t ARENA = TOK_ARENA_WORDS;
static_assert(ARENA % 64 == 0, "");

struct __align__(16) DecSmem {
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
};

struct __align__(16) ResSmem {
  alignas(16) uint8_t out_ring[POUT];  // (flushed with 16-byte shared-memory loads)
  uint32_t m_ld[MLIST];                // matches of the super-chunk: (length-3) | (distance-1) << 8
  uint16_t m_pos[MLIST];               // and their block-relative output offset
};

struct DCtx {             // shared-space addresses and limits every decoder lane needs
  uint32_t in_ring, lutl, lutd, auxtab, subl;
  uint32_t total_bits;
  const Code* code_dist;
  const uint16_t* sorted_dist;
  uint16_t* tok;          // row 0 of the super-chunk being decoded + lane
};

// 32 bits of the staged stream starting at bit `pos`
__device__ __forceinline__ uint32_t fetch32(uint32_t in_ring, uint32_t pos) {
  const uint32_t a = in_ring + ((pos >> 3) & (uint32_t)(PIN_RING - 4));
  const uint32_t lo = lds32(a);
  const uint32_t hi = lds32(a + 4);         // the word after the last one of the ring is a copy of word 0 (guard)
  return __funnelshift_r(lo, hi, pos);
}

// Every lane with `active` decodes the codes that start in [t, limit) of its own sub-sequence; all 32 lanes of the
// decoder warp must call this together.  One loop trip decodes one Huffman code per lane, whichever kind the lane needs
// next — a literal/length code or the distance code of the length it met in the previous trip — and, when that code
// was a plain literal (nine codes in ten of BAM data are) or a distance that left room, the literal/length code behind
// it as well: both come out of the same 32-bit window (<= 17 + 10 + 5 bits), so the second one costs a table lookup but
// no second fetch, vote or loop turn.  Everything is straight-line, select-based code, so that the lanes stay converged.
// RECORD = false: only the bit position moves (round 1: where do the sub-sequences synchronise?).
// RECORD = true: the lane also counts the bytes and matches it produces and writes one row entry (two tokens) per trip.
template <bool RECORD>
__device__ __forceinline__ void lane_decode(const DCtx& c, bool active, uint32_t t, uint32_t limit, uint32_t& end,
                                            uint32_t& out, uint32_t& nm, uint32_t& nt, uint32_t& flag) {
  const uint32_t lim = limit < c.total_bits ? limit : c.total_bits;
  constexpr uint32_t LITMSK = ((1u << LIT_BITS) - 1) << 1;
  uint32_t pos = t, o = 0, m = 0, fl = 0, trips = 0;
  uint32_t st = 0;          // 0: the next code is a literal/length code, 1: a distance code
  uint32_t len = 0;
  uint32_t lut = c.lutl, msk = LITMSK;
  uint32_t run = (active && pos < lim) ? 1u : 0u;
  uint32_t* tp = reinterpret_cast<uint32_t*>(c.tok);
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
    const uint32_t eb = is_len ? ENTRY_EXTRA_BITS(e) : 0u;
    // the literal/length code behind a plain literal or behind a distance: taken if the first code left room for it in
    // the window (<= 17 bits used, 10 + 5 to come), it starts inside the lane's range and the first-level table knows it
    const uint32_t lit1 = (st | kraw) ? 0u : 1u;
    const uint32_t adv1 = cl + eb;
    const uint32_t bits2 = bits >> adv1;
    const uint32_t e2 = lds16(c.lutl + ((bits2 << 1) & LITMSK));
    const uint32_t k2 = (e2 >> 8) & 3, cl2 = e2 >> 12;
    uint32_t ok2 = (lit1 | st) & (adv1 <= 17 ? 1u : 0u) & (k2 != K_SPECIAL ? 1u : 0u) & (pos + adv1 < lim ? 1u : 0u);
    if (RECORD) ok2 &= (o + lit1 + (st ? len : 0u) < (uint32_t)LANE_CAP ? 1u : 0u);     // (the byte cap is checked between codes)
    const uint32_t eb2 = k2 == K_LEN ? ENTRY_EXTRA_BITS(e2) : 0u;
    const uint32_t adv = adv1 + (ok2 ? cl2 + eb2 : 0u);
    pos += run ? adv : 0u;
    if (RECORD) {
      const uint32_t val = lds32(c.auxtab + (((st << 5) | (e & 31)) << 2)) + ((bits >> cl) & ~(0xffffffffu << eb));
      const uint32_t val2 = lds32(c.auxtab + ((e2 & 31) << 2)) + ((bits2 >> cl2) & ~(0xffffffffu << eb2));
      const uint32_t tk1 = (st ? TOK_DIST : (kraw << 13)) | (is_len ? val - (st ? 1u : 3u) : (e & 0xffu));
      const uint32_t tk2 = ok2 ? ((k2 << 13) | (k2 == K_LEN ? val2 - 3u : (e2 & 0xffu))) : TOK_NONE;
      if (run) *tp = tk1 | (tk2 << 16);
      tp += 32;
      trips += run;
      const uint32_t mat = st & run;                               // a distance: the match is complete
      o += (lit1 & run) + (mat ? len : 0u) + ((ok2 & run & (k2 == K_LIT ? 1u : 0u)));
      m += mat;
      len = ok2 ? val2 : val;
    }
    const uint32_t klast = ok2 ? k2 : (st ? 0u : kraw);            // kind of the last literal/length code of the trip
    const uint32_t eob = (klast == K_EOB ? 1u : 0u) & run;
    fl = eob ? F_EOB : fl;
    st = ok2 ? (k2 == K_LEN ? 1u : 0u) : ((st ^ 1u) & is_len);     // length -> distance next; anything else -> literal/length
    lut = st ? c.lutd : c.lutl;
    msk = st ? ((1u << DIST_BITS) - 1) << 1 : LITMSK;
    // a lane stops only between codes of the literal/length alphabet: at its boundary, or (RECORD) when it has produced
    // LANE_CAP bytes or TOK_TRIPS - 1 rows (hence at most LANE_MCAP matches) — the next lane continues from there
    uint32_t stop = pos >= lim ? 1u : 0u;
    if (RECORD) stop |= (o >= (uint32_t)LANE_CAP ? 1u : 0u) | (trips >= (uint32_t)TOK_TRIPS - 1 ? 1u : 0u);
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

}  // namespace

// diagnostics (biodb_debug_inflate_counters): the slots kernels.h lists
__device__ unsigned long long g_tok_counters[8];

// ===================================================================================== decode kernel ====
__global__ void __launch_bounds__(32, BIODB_TOK_DECODE_CTAS) inflate_decode_kernel(InflateArgs a) {
  __shared__ DecSmem sm;
  DecSmem* s = &sm;
  const int lane = threadIdx.x;
  const uint32_t blk = blockIdx.x;
  if (blk >= a.n_blocks) return;
  const uint64_t poff = a.payload_off[blk];
  const uint32_t csize = a.cdata_size[blk];
  const uint32_t isize = a.isize[blk];
  uint16_t* const arena = a.tok + (size_t)blk * ARENA;

  uint32_t sbase = smem_u32(s);
  asm volatile("mov.u32 %0, %0;" : "+r"(sbase));           // opaque: keeps the addresses in registers
  const uint32_t in_ring = sbase + (uint32_t)offsetof(DecSmem, in_ring);
  const uint32_t lutd = sbase + (uint32_t)offsetof(DecSmem, lut_dist);
  const uint32_t mbar = sbase + (uint32_t)offsetof(DecSmem, mbar);
  DCtx ctx;
  ctx.in_ring = in_ring;
  ctx.lutl = sbase + (uint32_t)offsetof(DecSmem, lut_lit);
  ctx.lutd = lutd;
  ctx.auxtab = sbase + (uint32_t)offsetof(DecSmem, auxtab);
  ctx.subl = sbase + (uint32_t)offsetof(DecSmem, sub_lit);
  ctx.code_dist = &s->code_dist;
  ctx.sorted_dist = s->sorted_dist;
  ctx.tok = arena;

  if (lane < NCH) mbar_init(mbar + 8 * lane, 1);
  if (lane == 0) asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  {
    uint32_t b, eb;
    s->auxtab[lane] = 0;
    s->auxtab[32 + lane] = 0;
    if (lane < 29) { len_base((uint32_t)lane, b, eb); s->auxtab[lane] = b; }
    if (lane < 30) { dist_base((uint32_t)lane, b, eb); s->auxtab[32 + lane] = b; }
  }
  __syncwarp();

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

  // all bit positions are relative to src
  uint32_t pos = skip * 8;
  const uint32_t total_bits = staged * 8;
  ctx.total_bits = total_bits;
  int status = 0;
  uint32_t n_super = 0, n_rounds = 0, n_dblocks = 0, n_stuck = 0;
  int why = 0;                      // 1: the arena ran out, 2: the stream does not synchronise (counters 6 / 7)
  uint32_t produced = 0;    // bytes handed to the resolver so far
  uint32_t cur = 0;         // next free word of the arena
  uint32_t sub_bits = SUB_BITS;   // bits per lane of the next super-chunk (adapts to the stream, see below)
  // room a record may need, plus the closing REC_DONE
  constexpr uint32_t CHUNK_MAX = REC_HEAD + REC_LANES + (uint32_t)TOK_TRIPS * ROW_WORDS;

  bool last = false;
  while (!last && status == 0) {
    // ---- block header (warp-uniform, from a 64-bit register bit buffer) ----------------------------------
    ensure_input(pos >> 3, (pos >> 3) + HDR_BYTES);
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
    if (btype == 3) { status = STATUS_RETRY; break; }

    if (btype == 0) {
      // ---- stored block: the resolve kernel copies it straight from the payload ------------------------------
      pos = (HPOS() + 7) & ~7u;
      ensure_input(pos >> 3, (pos >> 3) + 4);
      const uint32_t lw = fetch32(in_ring, pos);
      const uint32_t len = lw & 0xffff, nlen = lw >> 16;
      pos += 32;
      if (pos > total_bits || (len ^ 0xffff) != nlen || pos + len * 8 > total_bits || produced + len > isize ||
          cur + 2 * REC_HEAD > ARENA) {
        status = STATUS_RETRY;
        break;
      }
      if (lane == 0) {
        const uint32_t off = (pos >> 3) - skip;              // within the block's payload
        uint16_t* h = arena + cur;
        h[0] = REC_STORED; h[4] = (uint16_t)off; h[5] = (uint16_t)(off >> 16); h[6] = (uint16_t)len; h[7] = (uint16_t)(len >> 16);
      }
      cur += REC_HEAD;
      produced += len;
      // the decoder only skips the bytes, but the staging ring has to pass them: its chunks are issued and waited for
      // strictly in order (the mbarrier parities count every chunk), at most NCH in flight.  Not behind the LAST block
      // of the stream — what a level-0 writer produces: one final stored block per BGZF block — where nothing more is
      // read (the chunks still in flight are drained at the end).
      if (!last)
        for (uint32_t b = pos >> 3, e = (pos >> 3) + len; b < e; b += (NCH - 2) * CH)
          ensure_input(b, (b + (NCH - 2) * CH < e ? b + (NCH - 2) * CH : e));
      pos += len * 8;
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
#undef HFILL
#undef HDROP
#undef HPOS

    // ---- the codes of the block, one super-chunk of 32 sub-sequences at a time -------------------------------
    bool eob = false;
    while (!eob) {
      if (cur + CHUNK_MAX + REC_HEAD > ARENA) { status = STATUS_RETRY; why = 1; break; }   // the record stream outgrew its arena
      const uint32_t base = pos;
      ensure_input(base >> 3, (base >> 3) + SUPER_BYTES + 24);
      const uint32_t lim = base + (uint32_t)(lane + 1) * sub_bits;
      uint32_t t = base + (uint32_t)lane * sub_bits;
      uint32_t e_, out_, nm_, nt_, fl_;
      // round 1: where does the chain cross into each sub-sequence?  (bit positions only)
      lane_decode<false>(ctx, true, t, lim, e_, out_, nm_, nt_, fl_);
      ++n_super;
      // round 2: every lane from where its predecessor ended, recording tokens into the rows of this record
      ctx.tok = arena + cur + REC_HEAD + REC_LANES + 2 * lane;
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
      n_rounds += rounds;
      // commit the longest prefix of lanes that fits the resolver's output ring and match list
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
      // A stream whose codes do not synchronise — fixed-length codes, e.g. a fixed-Huffman block of bytes below 144 —
      // gains one lane per round: the rounds run out with a consistent prefix of a few lanes and a few dozen bytes.  That
      // is serial work, which the warp-serial kernel does three times cheaper than this loop; give such a block up
      // after sixteen super-chunks in a row that ended that way, instead of grinding through it and flooding the record
      // stream with rows of lanes that are thrown away.  (Streams of long matches do not synchronise either — a lane
      // that starts on a distance code reads it as a length code for ever — but there the resolver's ring is what limits
      // a commit, a few lanes fill it, and nothing is lost: those never count here.)
      if (vcut <= 8 && k == ncand && chunk_out < 512) {
        if (++n_stuck >= 16) { status = STATUS_RETRY; why = 2; break; }
      } else {
        n_stuck = 0;
      }
      const uint32_t rows = __reduce_max_sync(0xffffffffu, (uint32_t)lane < k ? nt_ : 0u);
      reinterpret_cast<uint32_t*>(arena + cur + REC_HEAD)[lane] = (uint32_t)lane < k ? (out_ | (nm_ << 16) | (nt_ << 24)) : 0u;
      if (lane == 0) {
        uint16_t* h = arena + cur;
        h[0] = REC_CHUNK; h[1] = (uint16_t)k; h[2] = (uint16_t)rows;
      }
      cur += REC_HEAD + REC_LANES + rows * ROW_WORDS;
      // Streams that expand a lot (long matches, one-bit codes) fill the resolver's ring with fewer than 32 lanes: the
      // other lanes' decoding — and their rows of the record stream — would be thrown away super-chunk after super-chunk.
      // Shorter sub-sequences make 32 lanes fit again; they grow back when the output gets small.
      if (k < ncand) sub_bits = max(64u, (sub_bits * k / 32u) & ~7u);
      else if (k == 32 && 2 * chunk_out < (uint32_t)OUT_BUDGET && sub_bits < (uint32_t)SUB_BITS) sub_bits = min((uint32_t)SUB_BITS, 2 * sub_bits);
      produced += chunk_out;
      pos = newpos;
      eob = stop_flag == F_EOB;
    }
  }

  // drain any TMA chunk still in flight before the CTA (and its shared memory) retires
  while (waited < issued) wait_chunk(waited);

  if (status == 0 && (produced != isize || pos > total_bits)) status = STATUS_RETRY;
  if (lane == 0) {
    uint16_t* h = arena + cur;                 // (room for this record is part of every capacity check above)
    h[0] = REC_DONE; h[1] = status ? 1 : 0;
    atomicAdd(&g_tok_counters[1], (unsigned long long)n_super);
    atomicAdd(&g_tok_counters[2], (unsigned long long)n_rounds);
    atomicAdd(&g_tok_counters[5], (unsigned long long)n_dblocks);
    if (why) atomicAdd(&g_tok_counters[5 + why], 1ull);
  }
}

// ==================================================================================== resolve kernel ====
__global__ void __launch_bounds__(32, 32) inflate_resolve_kernel(InflateArgs a) {
  __shared__ ResSmem sm;
  ResSmem* s = &sm;
  const int lane = threadIdx.x;
  const uint32_t blk = blockIdx.x;
  if (blk >= a.n_blocks) return;
  const uint32_t isize = a.isize[blk];
  const uint64_t obase = a.out_off[blk];
  uint8_t* gout = a.out + obase;
  const uint8_t* pay = a.comp + a.payload_off[blk];
  const uint32_t oa = (uint32_t)(((uintptr_t)gout) & POM);   // ring index of output byte 0
  const uint16_t* arena = a.tok + (size_t)blk * ARENA;

  uint32_t sbase = smem_u32(s);
  asm volatile("mov.u32 %0, %0;" : "+r"(sbase));
  const uint32_t ring = sbase + (uint32_t)offsetof(ResSmem, out_ring);
  const uint32_t mld = sbase + (uint32_t)offsetof(ResSmem, m_ld);
  const uint32_t mpos = sbase + (uint32_t)offsetof(ResSmem, m_pos);

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

  uint32_t n_far = 0, n_matches = 0;
  bool bad = false;
  int dstatus = 0;
  uint32_t cur = 0;
  // The head of the next record and its lane words are requested while the current record is still being worked on: the
  // record stream comes from HBM / L2, and three dependent loads per record would otherwise sit on the critical path.
  uint2 head = __ldcg(reinterpret_cast<const uint2*>(arena));                        // {kind | lanes << 16, rows | ...}
  uint32_t lane_word = __ldcg(reinterpret_cast<const uint32_t*>(arena + REC_HEAD) + lane);
  while (!bad) {
    const uint32_t kind = head.x & 0xffff;
    if (kind == REC_DONE) { dstatus = (int)(head.x >> 16); break; }
    if (kind == REC_STORED) {
      const uint32_t off = (uint32_t)__ldcg(arena + cur + 4) | ((uint32_t)__ldcg(arena + cur + 5) << 16);
      const uint32_t len = (uint32_t)__ldcg(arena + cur + 6) | ((uint32_t)__ldcg(arena + cur + 7) << 16);
      cur += REC_HEAD;
      head = __ldcg(reinterpret_cast<const uint2*>(arena + cur));
      lane_word = __ldcg(reinterpret_cast<const uint32_t*>(arena + cur + REC_HEAD) + lane);
      if (OPOS() + len > isize) { bad = true; break; }
      for (uint32_t p0 = 0; p0 < len; p0 += STORE_PIECE) {
        const uint32_t piece = len - p0 < (uint32_t)STORE_PIECE ? len - p0 : (uint32_t)STORE_PIECE;
        const uint8_t* g = pay + off + p0;
        // bytes up to the next word boundary of the ring, then whole words (the source read as aligned words and
        // funnel-shifted into place), then the tail
        uint32_t hb = (4u - (o & 3u)) & 3u;
        if (hb > piece) hb = piece;
        if ((uint32_t)lane < hb) sts8(ring + ((o + lane) & POM), __ldg(g + lane));
        const uint32_t nw = (piece - hb) >> 2;
        const uintptr_t ga = (uintptr_t)(g + hb);
        const uint32_t sh = (uint32_t)(ga & 3) * 8;
        const uint32_t* gw = reinterpret_cast<const uint32_t*>(ga & ~(uintptr_t)3);
        for (uint32_t k = lane; k < nw; k += 32) {
          const uint32_t lo = __ldg(gw + k), hi = sh ? __ldg(gw + k + 1) : 0u;
          sts32(ring + ((o + hb + 4 * k) & POM), __funnelshift_r(lo, hi, sh));
        }
        const uint32_t done = hb + 4 * nw;
        if (done + lane < piece) sts8(ring + ((o + done + lane) & POM), __ldg(g + done + lane));
        __syncwarp();
        o += piece;
        produced();
      }
      continue;
    }
    if (kind != REC_CHUNK) { bad = true; break; }             // (cannot happen: the decode kernel wrote the stream)
    // ---- a super-chunk: k lanes of the decoder are committed -----------------------------------------------
    const uint32_t k = head.x >> 16, rows = head.y & 0xffff;
    const uint32_t mine = (uint32_t)lane < k ? lane_word : 0u;
    const uint32_t* tok = reinterpret_cast<const uint32_t*>(arena + cur + REC_HEAD + REC_LANES) + lane;
    cur += REC_HEAD + REC_LANES + rows * ROW_WORDS;
    head = __ldcg(reinterpret_cast<const uint2*>(arena + cur));                      // (the arena ends with room for this)
    lane_word = __ldcg(reinterpret_cast<const uint32_t*>(arena + cur + REC_HEAD) + lane);
    const uint32_t out_ = mine & 0xffff, nm_ = (mine >> 16) & 0xff, nt_ = mine >> 24;
    const uint32_t inc_out = warp_incl_scan(out_, lane);
    const uint32_t inc_nm = warp_incl_scan(nm_, lane);
    const uint32_t chunk_out = __shfl_sync(0xffffffffu, inc_out, 31);
    const uint32_t n_match = __shfl_sync(0xffffffffu, inc_nm, 31);
    const uint32_t opos0 = OPOS();
    if (opos0 + chunk_out > isize || chunk_out > (uint32_t)OUT_BUDGET || n_match > (uint32_t)MLIST) { bad = true; break; }
    {
      // replay: literals into the ring, matches into the list
      uint32_t oo = o + (inc_out - out_);               // ring-relative position of this lane's next byte
      uint32_t slot = inc_nm - nm_;
      uint32_t len = 0;
      // rows eight at a time, the next eight requested before the current ones are replayed
      uint32_t tk[8], nx[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) nx[j] = ((uint32_t)j < nt_) ? __ldcs(tok + (size_t)j * 32) : (TOK_NONE | (TOK_NONE << 16));
      for (uint32_t t0 = 0; t0 < rows; t0 += 8) {
#pragma unroll
        for (int j = 0; j < 8; ++j) tk[j] = nx[j];
        if (t0 + 8 < rows) {
#pragma unroll
          for (int j = 0; j < 8; ++j) nx[j] = (t0 + 8 + j < nt_) ? __ldcs(tok + (size_t)(t0 + 8 + j) * 32) : (TOK_NONE | (TOK_NONE << 16));
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const uint32_t v = h ? tk[j] >> 16 : tk[j] & 0xffffu;
            const uint32_t kind = (v >> 13) & 3;
            if (v & TOK_DIST) {
              sts32(mld + (slot << 2), (len - 3) | ((v & 0x7fffu) << 8));
              sts16(mpos + (slot << 1), oo - oa);
              ++slot;
              oo += len;
            } else if (kind == K_LEN) {
              len = (v & 0xff) + 3;
            } else if (kind == K_LIT) {
              sts8(ring + (oo & POM), v);
              ++oo;
            }
          }
        }
      }
    }
    __syncwarp();
    {
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
            if (len > 32)
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
      n_matches += n_match;
    }
    if (bad) break;
    o += chunk_out;
    produced();
  }

  const int status = (dstatus != 0 || bad || OPOS() != isize) ? STATUS_RETRY : 0;
  if (status == 0) {
    wk.walk_upto(isize, true, lane);
    if (OPOS() > flushed) flush_to(OPOS());
  }
  if (lane == 0) {
    a.status[blk] = status;
    if (status == 0) wk.store(0);
    if (status) atomicAdd(&g_tok_counters[0], 1ull);
    atomicAdd(&g_tok_counters[3], (unsigned long long)n_far);
    atomicAdd(&g_tok_counters[4], (unsigned long long)n_matches);
  }
#undef OPOS
}

int inflate_tok_resident_blocks(int device) {
  int per_sm = 0, sms = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, inflate_decode_kernel, 32, 0) != cudaSuccess) return 0;
  if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device) != cudaSuccess) return 0;
  return per_sm * sms;
}

cudaError_t inflate_tok_counters(unsigned long long* out8, int reset) {
  cudaError_t e = cudaMemcpyFromSymbol(out8, g_tok_counters, sizeof(g_tok_counters));
  if (e == cudaSuccess && reset) {
    unsigned long long z[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    e = cudaMemcpyToSymbol(g_tok_counters, z, sizeof(z));
  }
  return e;
}

// (+ 512 bytes: the resolver requests the head and lane words of the record behind the last one before it looks at it)
size_t inflate_tok_token_bytes(uint32_t n_blocks) { return (size_t)n_blocks * ARENA * sizeof(uint16_t) + 512; }

cudaError_t launch_inflate_tok(const InflateArgs& a, cudaStream_t st) {
  if (a.n_blocks == 0) return cudaSuccess;
  inflate_decode_kernel<<<a.n_blocks, 32, 0, st>>>(a);
  inflate_resolve_kernel<<<a.n_blocks, 32, 0, st>>>(a);
  return cudaGetLastError();
}

}  // namespace biodb
