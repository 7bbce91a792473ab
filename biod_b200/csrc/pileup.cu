// Coordinate-sorted pileup builder for sm_100a.
//
// Replaces PileupRange.popFront / initNewReference (bio/std/hts/bam/pileup.d:345-424), the
// PileupRead CIGAR cursor (pileup.d:175-222), current_base / current_base_quality
// (pileup.d:115-134) and the filters of pileupInstance / pileupColumns (pileup.d:480-519).
//
// The reference sweeps positions sequentially, carrying a list of live reads.  Here the same
// columns are built data-parallel for one reference ("group") of start-sorted reads:
//   premax[j]   = max end over reads <= j                       (max-scan)
//   islands     = maximal position runs with coverage > 0       (flag + add-scan); with
//                 skip_zero_coverage=false the whole group is one island (pileup.d:389-392)
//   column c    = island column base + (position - island start)
//   coverage    = +1/-1 difference array over columns            (atomics + add-scan)
//   col_off     = exclusive scan of coverage
//   hi[c], lo[c]= window of candidate reads for column c        (atomicMax marks + max-scans)
//   entries     = one warp per column: ballot the candidates that are live at the position,
//                 rank = popc of lower lanes -> file order inside the column, exactly the order
//                 the reference's stable compaction keeps (pileup.d:351-359,381-383).
#include <cuda_runtime.h>
#include <stdint.h>

#include "pileup.h"
#include "scan.cuh"

namespace biodb {

namespace {

constexpr int32_t DEAD = INT32_MIN;

__device__ __forceinline__ uint32_t ld32u(const uint8_t* p) {
  uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t lo = __ldg(w);
  if (sh == 0) return lo;
  uint32_t hi = __ldg(w + 1);
  return __funnelshift_r(lo, hi, sh);
}
__device__ __forceinline__ uint32_t consume(uint32_t raw) { return (0x3C1A7u >> ((raw & 0xF) * 2)) & 3; }  // cigar.d:116

__device__ __forceinline__ const uint8_t* record_body(const ReadsView& v, uint32_t j) {
  return (j < v.n_carry ? v.carry_data : v.u) + v.rec_off[j] + 4;
}

// ---- group discovery ------------------------------------------------------------------------
__global__ void find_groups_kernel(ReadsView v, uint32_t* boundaries, uint32_t* n_boundaries, uint32_t cap) {
  uint32_t j = blockIdx.x * blockDim.x + threadIdx.x + 1;
  if (j >= v.n) return;
  if (v.ref_id[j] != v.ref_id[j - 1]) {
    uint32_t k = atomicAdd(n_boundaries, 1u);
    if (k < cap) boundaries[k] = j;
  }
}

// ---- per-read preparation: liveness, CIGAR validity, sortedness ------------------------------
// info[0] = error status, info[1] = index+1 of the last live read, info[2] = index+1 of first read with
// end >= start_from (for the prefix-drop rule, pileup.d:482-489)
__global__ void prep_kernel(ReadsView v, uint32_t g0, uint32_t g1, uint32_t drop_before, int32_t* eend, uint4* rinfo,
                            int32_t* info) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  int32_t pos = v.pos[j], end = v.end_pos[j];
  bool live = (int32_t)((uint32_t)end - (uint32_t)pos) > 0 && j >= drop_before;   // basesCovered() > 0 (pileup.d:481)
  if (live) {
    // PileupRead constructor (pileup.d:175-192): first reference-consuming op that is not N; query-consuming
    // ops in front of it (S, I) advance the query offset
    const uint8_t* rec = record_body(v, j);
    uint32_t lname = v.bin_mq_nl[j] & 0xFF, nc = v.flag_nc[j] & 0xFFFF;
    const uint8_t* cg = rec + 32 + lname;
    bool ok = false, skipped_n = false;
    uint32_t qoff0 = 0, k = 0, first_t = 0;
    for (; k < nc; ++k) {
      uint32_t raw = ld32u(cg + 4 * k);
      uint32_t t = consume(raw);
      if (t & 2) {
        if ((raw & 0xF) != 3) { ok = true; first_t = t; break; }
        if (raw >> 4) skipped_n = true;
      } else if (t & 1) {
        qoff0 += raw >> 4;
      }
    }
    // "simple" read: that op is M/=/X and no other reference-consuming op follows, so the cursor at reference
    // offset k is just query offset qoff0 + k
    bool simple = ok && first_t == 3;
    for (uint32_t q = k + 1; simple && q < nc; ++q)
      if (consume(ld32u(cg + 4 * q)) & 2) simple = false;
    if (!ok || skipped_n) atomicCAS(&info[0], 0, -7);          // BIODB_ERR_CIGAR
    if (pos < 0) atomicCAS(&info[0], 0, -8);
    if (j > g0) {
      int32_t pp = v.pos[j - 1], pe = v.end_pos[j - 1];
      bool plive = (int32_t)((uint32_t)pe - (uint32_t)pp) > 0;
      if (plive && pp > pos) atomicCAS(&info[0], 0, -8);       // BIODB_ERR_UNSORTED
    }
    atomicMax(&info[1], (int32_t)(j + 1));
    const uint64_t seq = (uint64_t)(uintptr_t)(cg + 4 * nc);
    const uint32_t gidx = j < v.n_carry ? v.carry_gidx[j] : (uint32_t)(v.first_index + (j - v.n_carry));
    rinfo[j] = make_uint4((uint32_t)seq, (uint32_t)(seq >> 32), (qoff0 & 0x7fffffffu) | (simple ? 0x80000000u : 0u), gidx);
  }
  eend[j] = live ? end : DEAD;
}

__global__ void first_kept_kernel(ReadsView v, uint32_t g0, uint32_t g1, uint64_t start_from, uint32_t* first) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  int32_t pos = v.pos[j], end = v.end_pos[j];
  bool live = (int32_t)((uint32_t)end - (uint32_t)pos) > 0;
  if (live && (uint64_t)(int64_t)end >= start_from) atomicMin(first, j);
}

// ---- islands -----------------------------------------------------------------------------------
__global__ void island_flag_kernel(ReadsView v, uint32_t g0, uint32_t g1, const int32_t* eend, const int32_t* pm,
                                   int skip_zero, uint32_t* flag) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  uint32_t f = 0;
  if (eend[j] != DEAD) {
    int32_t prev = (j == g0) ? DEAD : pm[j - 1];
    if (prev == DEAD || (skip_zero && v.pos[j] >= prev)) f = 1;
  }
  flag[j] = f;
}

__global__ void island_table_kernel(ReadsView v, uint32_t g0, uint32_t g1, const uint32_t* flag, const uint32_t* iid1,
                                    const int32_t* pm, IslandTable t) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  uint32_t id1 = iid1[j];
  if (id1 == 0) return;                    // dead reads ahead of the first island
  if (flag[j]) { t.start[id1 - 1] = v.pos[j]; t.first[id1 - 1] = j; }
  if (j + 1 == g1 || flag[j + 1]) t.end[id1 - 1] = pm[j];
}

__global__ void island_cols_kernel(IslandTable t, uint32_t n_islands, int64_t clo, int64_t chi, uint32_t* ncol) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n_islands) return;
  int64_t cs = t.start[i], ce = t.end[i];
  if (cs < clo) cs = clo;
  if (ce > chi) ce = chi;
  int64_t n = ce - cs;
  if (n < 0) n = 0;
  t.cs[i] = cs;
  ncol[i] = (uint32_t)(n > 0x7fffffff ? 0x7fffffff : n);
}

// ---- per-read scatter into the column arrays ------------------------------------------------------
__global__ void scatter_kernel(ReadsView v, uint32_t g0, uint32_t g1, const int32_t* eend, const int32_t* pm,
                               const uint32_t* iid1, IslandTable t, const uint32_t* colbase, const uint32_t* ncol,
                               ColumnScratch c) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  int32_t e = eend[j];
  if (e == DEAD) return;
  uint32_t i = iid1[j] - 1;
  int64_t cs = t.cs[i];
  uint32_t nc = ncol[i];
  if (nc == 0) return;
  int64_t ce = cs + nc;
  int64_t pos = v.pos[j];
  if (pos >= ce || (int64_t)e <= cs) return;           // no column of this batch sees the read
  uint32_t base = colbase[i];
  int64_t a = pos < cs ? cs : pos;
  int64_t b = (int64_t)e < ce ? (int64_t)e : ce;
  atomicAdd(&c.diff[base + (uint32_t)(a - cs)], 1);
  atomicAdd(&c.diff[base + (uint32_t)(b - cs)], -1);
  if (pos >= cs) atomicAdd(&c.nstart[base + (uint32_t)(pos - cs)], 1u);
  atomicMax(&c.hi[base + (uint32_t)(a - cs)], j + 1);
  // frontier read: extends the covered prefix, so it is the first live read from max(prev frontier, pos) on
  int32_t prev = (j == g0) ? DEAD : pm[j - 1];
  if (prev == DEAD || e > prev) {
    int64_t f = (prev == DEAD || (int64_t)prev < pos) ? pos : (int64_t)prev;
    if (f < cs) f = cs;
    if (f < ce) atomicMax(&c.lo[base + (uint32_t)(f - cs)], j + 1);
  }
}

__global__ void colpos_kernel(IslandTable t, const uint32_t* colbase, uint32_t n_islands, uint32_t n_col, uint64_t* col_pos) {
  uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  uint32_t lo = 0, hi = n_islands;       // last island with colbase <= c
  while (hi - lo > 1) {
    uint32_t m = (lo + hi) >> 1;
    if (colbase[m] <= c) lo = m; else hi = m;
  }
  col_pos[c] = (uint64_t)(t.cs[lo] + (int64_t)(c - colbase[lo]));
}

// ---- column entries -------------------------------------------------------------------------------
struct Entry { uint8_t base, qual; uint32_t qoff; };

constexpr int ENT_WARPS = 8;
#ifndef BIODB_CHUNK
#define BIODB_CHUNK 32
#endif
constexpr int CHUNK = BIODB_CHUNK;   // columns per warp (<= 32)
constexpr int CAND_CAP = 128;        // candidate reads gathered per round (a chunk with more takes several rounds)

// 4-bit code -> IUPAC character without a memory lookup: "=ACMGRSV" "TWYHKDBN" packed little-endian (base.d:85)
__device__ __forceinline__ uint32_t base_char(uint32_t code) {
  // two byte permutes over the 8-byte halves of the table, selected by bit 3 of the code
  const uint32_t lo = __byte_perm(0x4D43413Du, 0x56535247u, code & 7);     // "=ACM" "GRSV"
  const uint32_t hi = __byte_perm(0x48595754u, 0x4E42444Bu, code & 7);     // "TWYH" "KDBN"
  return ((code & 8) ? hi : lo) & 0xFF;
}

// CIGAR cursor of a read at reference offset k from its position (PileupRead ctor + k x incrementPosition,
// pileup.d:175-222) and the base/quality there (pileup.d:115-134, read.d:364-383, base.d:85).  The cursor is kept alive
// across the columns a lane visits for one read: columns come in ascending order, so the walk only ever moves
// forward — amortised O(1) per column instead of a CIGAR walk from the start for every column.
struct Cursor {
  uint32_t i, raw, t;   // current reference-consuming operation: index, packed word, consume bits
  uint32_t k0, q;       // reference offset (from the read's position) and query offset where it starts
};
__device__ __forceinline__ void cursor_init(Cursor& c, const uint8_t* cg, uint32_t nc) {
  c.i = 0; c.raw = 0; c.t = 0; c.k0 = 0; c.q = 0;
  for (; c.i < nc; ++c.i) {
    c.raw = ld32u(cg + 4 * c.i);
    c.t = consume(c.raw);
    if (c.t & 2) { if ((c.raw & 0xF) != 3) break; }
    else if (c.t & 1) c.q += c.raw >> 4;
  }
}
__device__ __forceinline__ Entry cursor_eval(Cursor& c, const uint8_t* cg, uint32_t nc, const uint8_t* seq, const uint8_t* ql,
                                             int32_t lseq, uint32_t k) {
  Entry en;
  en.base = '-';
  en.qual = 255;
  while (c.i < nc) {
    const uint32_t len = c.raw >> 4;
    const uint32_t eff = len ? len : 1;          // a zero-length op still holds the cursor for one position
    const uint32_t r = k - c.k0;
    if (r < eff) {
      uint32_t qoff = c.q;
      if (c.t == 3) {
        qoff += r;
        if (qoff < (uint32_t)lseq) {
          const uint32_t byte = __ldg(seq + (qoff >> 1));
          en.base = (uint8_t)base_char((qoff & 1) ? (byte & 0xF) : (byte >> 4));
          en.qual = __ldg(ql + qoff);
        } else {
          // SEQ shorter than the CIGAR says (e.g. SEQ '*'): BioD's lazy accessor would raise a RangeError only
          // if asked for this base; the eager builder marks it with base 0x00 / quality 255 instead.
          en.base = 0;
        }
      }
      en.qoff = qoff;
      return en;
    }
    c.k0 += eff;
    if (c.t & 1) c.q += eff;
    for (++c.i; c.i < nc; ++c.i) {
      c.raw = ld32u(cg + 4 * c.i);
      c.t = consume(c.raw);
      if (c.t & 2) break;
      if (c.t & 1) c.q += c.raw >> 4;
    }
  }
  en.qoff = c.q;
  return en;
}

// Read-stationary column builder.  A warp owns CHUNK consecutive columns; lane ci keeps column ci's position and
// its running output offset.  The candidate reads of the chunk ([lo of its first column, hi of its last)) are
// taken 32 at a time, one per lane, their cursor data loaded ONCE; then for every column of the chunk the lanes
// whose read is live there are balloted, ranked (popc of lower lanes = file order, pileup.d:351-359,381-383)
// and write read_idx / base / qual at consecutive slots.
// MAQ: the entries feed maq_kernel (maq.cu) instead of leaving the device: per entry base | strand << 7 — or 0xFF for a
// base MaqSnpCaller drops, quality below minimum_base_quality or '-' (maq.d:401-404) — and min(base quality, mapping
// quality) (:407-409); no read_idx.
template <bool COUNTS, bool WANT_Q, bool MAQ>
__global__ void __launch_bounds__(ENT_WARPS * 32) entries_kernel(ReadsView v, const int32_t* __restrict__ eend,
                                                                 const uint4* __restrict__ rinfo, ColumnScratch c,
                                                                 ColumnOutput o, uint32_t n_col, const uint32_t* __restrict__ redo,
                                                                 int32_t* info) {
  const uint32_t lane = threadIdx.x & 31;
  const uint32_t warp = blockIdx.x * ENT_WARPS + (threadIdx.x >> 5);
  const uint32_t lt = (1u << lane) - 1;
  const uint32_t c0 = warp * CHUNK;
  if (c0 >= n_col) return;
  if (redo && !redo[warp]) return;          // this chunk was done by entries_tile_kernel
  const uint32_t ncols = min((uint32_t)CHUNK, n_col - c0);
  const uint64_t chunk_off = o.col_off[c0];
  int32_t my_p = 0;               // BAM positions are int32 (read.d:93)
  uint32_t my_off = 0;            // running offset of column `lane` relative to chunk_off
  if (lane < ncols) {
    my_p = (int32_t)o.col_pos[c0 + lane];
    my_off = (uint32_t)(o.col_off[c0 + lane] - chunk_off);
  }
  const int32_t p_first = __shfl_sync(0xffffffffu, my_p, 0), p_last = __shfl_sync(0xffffffffu, my_p, ncols - 1);
  uint32_t lo = c.lo[c0];
  const uint32_t hi = c.hi[c0 + ncols - 1];
  uint32_t cnt[6] = {0, 0, 0, 0, 0, 0};     // COUNTS: A, C, G, T, other, deletion of column `lane`
  if (hi == 0) {
    if (COUNTS && lane < ncols)
      for (int q = 0; q < 6; ++q) o.counts[(size_t)(c0 + lane) * 6 + q] = 0;
    return;
  }
  lo = lo ? lo - 1 : 0;
  // The candidate reads of the chunk — alive somewhere between its first and last position — are first gathered, in
  // file order, into a per-warp list: [lo, hi) also holds reads that ended long ago and, with long N-skips, a few
  // live reads hundreds of indices before the rest; walking that range 32 reads at a time would run the column loop
  // for windows with one live lane.
  __shared__ uint32_t s_list[ENT_WARPS][CAND_CAP];
  uint32_t* list = s_list[threadIdx.x >> 5];
  uint32_t jn = lo;
  while (jn < hi) {
   uint32_t n_cand = 0;
   while (jn < hi && n_cand + 32 <= (uint32_t)CAND_CAP) {
     const uint32_t j = jn + lane;
     bool cand = false;
     if (j < hi) {
       const int32_t e = eend[j];
       // a chunk may span several islands, so positions are not contiguous: test against its first / last position
       cand = e != DEAD && e > p_first && v.pos[j] <= p_last;
     }
     const uint32_t b = __ballot_sync(0xffffffffu, cand);
     if (cand) list[n_cand + __popc(b & lt)] = j;
     n_cand += __popc(b);
     jn += 32;
   }
   __syncwarp();
   for (uint32_t w0 = 0; w0 < n_cand; w0 += 32) {
    const bool cand = w0 + lane < n_cand;
    const uint32_t j = cand ? list[w0 + lane] : 0;
    int32_t e = DEAD, pos = 0, lseq = 0;
    uint4 ri = make_uint4(0, 0, 0, 0);
    uint32_t mapq = 0, rev = 0;
    if (cand) {
      e = eend[j];
      pos = v.pos[j];
      ri = rinfo[j];
      lseq = v.l_seq[j];
      if (MAQ) {
        mapq = (v.bin_mq_nl[j] >> 8) & 0xff;        // read.d:958-962
        rev = (v.flag_nc[j] >> 20) & 1;             // flag 0x10: is_reverse_strand
      }
    }
    const uint8_t* seq = (const uint8_t*)(uintptr_t)(((uint64_t)ri.y << 32) | ri.x);
    const uint8_t* ql = seq + (((uint32_t)lseq + 1) >> 1);
    const bool simple = (ri.z & 0x80000000u) != 0;
    const uint32_t qoff0 = ri.z & 0x7fffffffu;
    Cursor cur;
    uint32_t cur_nc = 0xffffffffu;          // not initialised yet for this read
    for (uint32_t ci = 0; ci < ncols; ++ci) {
      const int32_t p = __shfl_sync(0xffffffffu, my_p, ci);
      const bool live = cand && pos <= p && e > p;
      const uint32_t m = __ballot_sync(0xffffffffu, live);
      if (m == 0) continue;
      const uint32_t coff = COUNTS ? 0 : __shfl_sync(0xffffffffu, my_off, ci);
      uint32_t base = 0xFFFFFFFFu;
      if (live) {
        const uint64_t slot = chunk_off + coff + __popc(m & lt);
        const uint32_t k = (uint32_t)(p - pos);
        uint32_t qual = 255, qoff;
        base = '-';
        if (simple) {
          // single M/=/X run: query offset is linear in the column (pileup.d:195-203)
          qoff = qoff0 + k;
          if (qoff < (uint32_t)lseq) {
            const uint32_t byte = __ldg(seq + (qoff >> 1));
            base = base_char((qoff & 1) ? (byte & 0xF) : (byte >> 4));
            qual = __ldg(ql + qoff);
          } else {
            base = 0;      // SEQ shorter than the CIGAR says: see cursor_eval
          }
        } else {
          if (cur_nc == 0xffffffffu) {
            cur_nc = v.flag_nc[j] & 0xFFFF;
            cursor_init(cur, seq - 4 * cur_nc, cur_nc);
          }
          Entry en = cursor_eval(cur, seq - 4 * cur_nc, cur_nc, seq, ql, lseq, k);
          base = en.base;
          qual = en.qual;
          qoff = en.qoff;
        }
        if (MAQ) {
          const bool ok = qual >= (uint32_t)o.maq_min_base_quality && base != '-';
          o.base[slot] = ok ? (uint8_t)(base | (rev << 7)) : (uint8_t)0xFF;
          o.qual[slot] = (uint8_t)(qual < mapq ? qual : mapq);
        } else if (!COUNTS) {
          o.read_idx[slot] = ri.w;
          o.base[slot] = (uint8_t)base;
          o.qual[slot] = (uint8_t)qual;
          if (WANT_Q) o.qoff[slot] = qoff;
        }
      }
      if (COUNTS) {
        const uint32_t a = __popc(__ballot_sync(0xffffffffu, base == 'A')), cc = __popc(__ballot_sync(0xffffffffu, base == 'C'));
        const uint32_t g = __popc(__ballot_sync(0xffffffffu, base == 'G')), t = __popc(__ballot_sync(0xffffffffu, base == 'T'));
        const uint32_t d = __popc(__ballot_sync(0xffffffffu, base == '-'));
        if (lane == ci) {
          cnt[0] += a; cnt[1] += cc; cnt[2] += g; cnt[3] += t; cnt[5] += d;
          cnt[4] += __popc(m) - a - cc - g - t - d;
        }
      } else if (lane == ci) {
        my_off += __popc(m);
      }
    }
   }
   __syncwarp();
  }
  if (COUNTS && lane < ncols)
    for (int q = 0; q < 6; ++q) o.counts[(size_t)(c0 + lane) * 6 + q] = cnt[q];
}

// ---- column entries, column-stationary ---------------------------------------------------------------------------
// The same entries, built the other way round: a warp owns TCHUNK consecutive columns, LANE = COLUMN, and the chunk's
// candidate reads are taken one after the other (file order) with their cursor data broadcast to the warp.  A lane whose
// position the read covers appends one entry to its own column — its running count IS the entry's rank, so the ballot,
// the rank popcount and the per-column shuffles of entries_kernel disappear (3x fewer instructions per entry) — and
// neighbouring lanes read neighbouring bases / qualities of the read (coalesced).  The chunk's entries are contiguous in
// the output ([col_off[c0], col_off[c0 + TCHUNK))), so they are staged in shared memory and written out in one coalesced
// sweep.  A chunk with more entries than the tile holds is left to entries_kernel (flagged in `redo`).
// MODE 0: read_idx + base + qual.  MODE 1 (compact_reads): base + qual, and per column the sequential encoding of its
// read list — last read, 64-bit window mask, number of stragglers (reads more than 63 records older than the last) —
// without ever writing read_idx.  MODE 2 (MAQ): base | strand << 7 (0xFF = dropped) + min(quality, mapq).
constexpr int TCHUNK = 32;
constexpr int TILE_WARPS = 4;
constexpr int TILE_CAP = 1280;          // entries staged per warp: 32 columns x coverage 40

template <int MODE>
__global__ void __launch_bounds__(TILE_WARPS * 32) entries_tile_kernel(ReadsView v, const int32_t* __restrict__ eend,
                                                                       const uint4* __restrict__ rinfo, ColumnScratch c,
                                                                       ColumnOutput o, uint32_t n_col, uint32_t* last_read,
                                                                       uint64_t* live_mask, uint32_t* nstrag, uint32_t* redo,
                                                                       int32_t* info) {
  const uint32_t lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t warp = blockIdx.x * TILE_WARPS + wib;
  const uint32_t c0 = warp * TCHUNK;
  if (c0 >= n_col) return;
  const uint32_t ncols = min((uint32_t)TCHUNK, n_col - c0);
  const uint64_t chunk_off = o.col_off[c0];
  const uint32_t n_chunk = (uint32_t)(o.col_off[c0 + ncols] - chunk_off);
  const bool mine = lane < ncols;
  int32_t p = 0;
  uint32_t cnt = 0;                       // next entry of column `lane`, relative to chunk_off
  if (mine) {
    p = (int32_t)o.col_pos[c0 + lane];
    cnt = (uint32_t)(o.col_off[c0 + lane] - chunk_off);
  }
  if (n_chunk > (uint32_t)TILE_CAP) {     // (warp-uniform) too deep for the tile: entries_kernel redoes this chunk
    if (lane == 0) redo[warp] = 1;
    if (MODE == 1 && mine) { last_read[c0 + lane] = 0; live_mask[c0 + lane] = 0; nstrag[c0 + lane] = 0; }
    return;
  }
  if (lane == 0) redo[warp] = 0;
  const int32_t p_first = __shfl_sync(0xffffffffu, p, 0), p_last = __shfl_sync(0xffffffffu, p, ncols - 1);
  uint32_t lo = c.lo[c0];
  const uint32_t hi = c.hi[c0 + ncols - 1];
  __shared__ uint32_t s_ridx[MODE == 0 ? TILE_WARPS * TILE_CAP : 1];
  __shared__ __align__(16) uint8_t s_base[TILE_WARPS][TILE_CAP];
  __shared__ __align__(16) uint8_t s_qual[TILE_WARPS][TILE_CAP];
  uint32_t* t_ridx = s_ridx + (MODE == 0 ? wib * TILE_CAP : 0);
  uint8_t* t_base = s_base[wib];
  uint8_t* t_qual = s_qual[wib];
  // MODE 1: the read list of column `lane` as (last read, mask of the 64 records up to it, count of older ones)
  uint32_t m_last = 0, m_ns = 0, m_any = 0;
  uint64_t m_mask = 0;
  lo = lo ? lo - 1 : 0;
  for (uint32_t j0 = lo; j0 < hi; j0 += 32) {
    // 32 reads of the range at a time: their cursor data into the lanes' registers, handed round by shuffles
    const uint32_t jl = j0 + lane;
    int32_t re = DEAD, rpos = 0, rlseq = 0;
    uint4 ri = make_uint4(0, 0, 0, 0);
    uint32_t rmq = 0;
    bool cand = false;
    if (jl < hi) {
      re = eend[jl];
      rpos = v.pos[jl];
      cand = re != DEAD && re > p_first && rpos <= p_last;
      if (cand) {
        ri = rinfo[jl];
        rlseq = v.l_seq[jl];
        if (MODE == 2) rmq = ((v.bin_mq_nl[jl] >> 8) & 0xff) | (((v.flag_nc[jl] >> 20) & 1) << 8);
      }
    }
    uint32_t cm = __ballot_sync(0xffffffffu, cand);
    while (cm) {
      const uint32_t k = (uint32_t)__ffs(cm) - 1;
      cm &= cm - 1;
      const int32_t e = __shfl_sync(0xffffffffu, re, k), pos = __shfl_sync(0xffffffffu, rpos, k);
      const int32_t lseq = __shfl_sync(0xffffffffu, rlseq, k);
      const uint32_t sx = __shfl_sync(0xffffffffu, ri.x, k), sy = __shfl_sync(0xffffffffu, ri.y, k);
      const uint32_t sz = __shfl_sync(0xffffffffu, ri.z, k), gidx = __shfl_sync(0xffffffffu, ri.w, k);
      const uint32_t mq = MODE == 2 ? __shfl_sync(0xffffffffu, rmq, k) : 0;
      const bool live = mine && pos <= p && e > p;
      if (!live) continue;
      const uint8_t* seq = (const uint8_t*)(uintptr_t)(((uint64_t)sy << 32) | sx);
      const uint8_t* ql = seq + (((uint32_t)lseq + 1) >> 1);
      const uint32_t kk = (uint32_t)(p - pos);
      uint32_t base = '-', qual = 255;
      if (sz & 0x80000000u) {
        // single M/=/X run: the query offset is linear in the column (pileup.d:195-203)
        const uint32_t qoff = (sz & 0x7fffffffu) + kk;
        if (qoff < (uint32_t)lseq) {
          const uint32_t byte = __ldg(seq + (qoff >> 1));
          base = base_char((qoff & 1) ? (byte & 0xF) : (byte >> 4));
          qual = __ldg(ql + qoff);
        } else {
          base = 0;        // SEQ shorter than the CIGAR says: see cursor_eval
        }
      } else {
        // any other CIGAR: every lane walks it from the start to its own offset (5 % of the reads, a few operations)
        const uint32_t nc = v.flag_nc[j0 + k] & 0xFFFF;
        Cursor cur;
        cursor_init(cur, seq - 4 * nc, nc);
        const Entry en = cursor_eval(cur, seq - 4 * nc, nc, seq, ql, lseq, kk);
        base = en.base;
        qual = en.qual;
      }
      const uint32_t slot = cnt++;
      if (MODE == 2) {
        const bool ok = qual >= (uint32_t)o.maq_min_base_quality && base != '-';
        t_base[slot] = ok ? (uint8_t)(base | ((mq >> 8) << 7)) : (uint8_t)0xFF;
        t_qual[slot] = (uint8_t)(qual < (mq & 0xff) ? qual : (mq & 0xff));
      } else {
        t_base[slot] = (uint8_t)base;
        t_qual[slot] = (uint8_t)qual;
        if (MODE == 0) t_ridx[slot] = gidx;
      }
      if (MODE == 1) {
        // reads come in file order: the new one becomes the last, the window moves up by the index gap
        const uint32_t sh = m_any ? gidx - m_last : 0;
        m_ns += sh >= 64 ? (uint32_t)__popcll(m_mask) : (sh ? (uint32_t)__popcll(m_mask >> (64 - sh)) : 0u);
        m_mask = (sh >= 64 ? 0ull : m_mask << sh) | 1ull;
        m_last = gidx;
        m_any = 1;
      }
    }
  }
  __syncwarp();
  if (MODE == 1 && mine) { last_read[c0 + lane] = m_last; live_mask[c0 + lane] = m_mask; nstrag[c0 + lane] = m_ns; }
  // write-out: the tile is one contiguous stretch of the entry arrays
  if (MODE == 0)
    for (uint32_t i = lane; i < n_chunk; i += 32) o.read_idx[chunk_off + i] = t_ridx[i];
  {
    uint8_t* gb = o.base + chunk_off;
    uint8_t* gq = o.qual + chunk_off;
    // bytes up to the first 16-byte boundary of the global arrays, then 16 bytes per lane, then the tail
    const uint32_t head = min(n_chunk, (uint32_t)((16 - (chunk_off & 15)) & 15));
    if (lane < head) { gb[lane] = t_base[lane]; gq[lane] = t_qual[lane]; }
    const uint32_t n16 = (n_chunk - head) >> 4;
    for (uint32_t i = lane; i < n16; i += 32) {
      uint4 vb, vq;
      const uint8_t* sb = t_base + head + 16 * i;
      const uint8_t* sq = t_qual + head + 16 * i;
      if ((head & 3) == 0) {
        vb = make_uint4(((const uint32_t*)sb)[0], ((const uint32_t*)sb)[1], ((const uint32_t*)sb)[2], ((const uint32_t*)sb)[3]);
        vq = make_uint4(((const uint32_t*)sq)[0], ((const uint32_t*)sq)[1], ((const uint32_t*)sq)[2], ((const uint32_t*)sq)[3]);
      } else {
        uint32_t wb[4], wq[4];
#pragma unroll
        for (int w = 0; w < 4; ++w) {
          wb[w] = sb[4 * w] | (sb[4 * w + 1] << 8) | (sb[4 * w + 2] << 16) | ((uint32_t)sb[4 * w + 3] << 24);
          wq[w] = sq[4 * w] | (sq[4 * w + 1] << 8) | (sq[4 * w + 2] << 16) | ((uint32_t)sq[4 * w + 3] << 24);
        }
        vb = make_uint4(wb[0], wb[1], wb[2], wb[3]);
        vq = make_uint4(wq[0], wq[1], wq[2], wq[3]);
      }
      *reinterpret_cast<uint4*>(gb + head + 16 * i) = vb;
      *reinterpret_cast<uint4*>(gq + head + 16 * i) = vq;
    }
    const uint32_t done = head + (n16 << 4);
    if (done + lane < n_chunk) { gb[done + lane] = t_base[done + lane]; gq[done + lane] = t_qual[done + lane]; }
  }
}

// compact_reads without read_idx: the stragglers of the (few) columns that have any, by walking the column's candidate
// reads once more.  One thread per column; strag_off = exclusive scan of nstrag.
__global__ void strag_walk_kernel(ReadsView v, const int32_t* __restrict__ eend, const uint4* __restrict__ rinfo,
                                  const uint32_t* __restrict__ lo, const uint32_t* __restrict__ hi,
                                  const uint64_t* __restrict__ col_pos, uint32_t n_col, const uint32_t* __restrict__ strag_off,
                                  uint32_t* strag_col, uint32_t* strag_idx) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  const uint32_t s0 = strag_off[c], ns = strag_off[c + 1] - s0;
  if (ns == 0) return;
  const int32_t p = (int32_t)col_pos[c];
  uint32_t j = lo[c] ? lo[c] - 1 : 0, k = 0;
  for (; j < hi[c] && k < ns; ++j) {
    const int32_t e = eend[j];
    if (e != DEAD && v.pos[j] <= p && e > p) {       // live at the column: the first ns of them are the stragglers
      strag_col[s0 + k] = c;
      strag_idx[s0 + k] = rinfo[j].w;
      ++k;
    }
  }
}

// ---- compact read lists (biodb_pileup_params.compact_reads) -----------------------------------------------
// The reads of a column are file-ordered, and all but a few long-spanning ones lie within a short window of record
// indices.  One thread per column turns the column's read_idx list into: the index of its last read, a 64-bit mask
// (bit d = read `last - d` is in the column) and the number of leading reads older than that window ("stragglers",
// copied out verbatim by compact_strag_kernel).  12 bytes per column instead of 4 bytes per entry cross PCIe.
__global__ void compact_mask_kernel(const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ read_idx,
                                    uint32_t n_col, uint32_t* last_read, uint64_t* mask, uint32_t* nstrag,
                                    const uint32_t* __restrict__ redo) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c > n_col) return;
  if (c == n_col) { nstrag[c] = 0; return; }
  if (redo && !redo[c / CHUNK]) return;       // entries_tile_kernel wrote this column's list itself
  const uint64_t b = col_off[c], e = col_off[c + 1];
  uint32_t last = 0, ns = 0;
  uint64_t m = 0;
  if (e > b) {
    last = read_idx[e - 1];
    for (uint64_t i = e; i-- > b;) {
      const uint32_t d = last - read_idx[i];
      if (d >= 64) { ns = (uint32_t)(i - b + 1); break; }     // ascending order: everything before is older still
      m |= 1ull << d;
    }
  }
  last_read[c] = last;
  mask[c] = m;
  nstrag[c] = ns;
}

__global__ void compact_strag_kernel(const uint64_t* __restrict__ col_off, const uint32_t* __restrict__ read_idx,
                                     uint32_t n_col, const uint32_t* __restrict__ strag_off, uint32_t* strag_col,
                                     uint32_t* strag_idx) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  const uint32_t s0 = strag_off[c], ns = strag_off[c + 1] - s0;
  const uint64_t b = col_off[c];
  for (uint32_t i = 0; i < ns; ++i) {
    strag_col[s0 + i] = c;
    strag_idx[s0 + i] = read_idx[b + i];
  }
}

// Bases of the column entries, two per byte (4-bit codes of "=ACMGRSVTWYHKDBN", first entry in the high nibble — the
// packing of BAM's own SEQ field, read.d:364-383).  The two byte values that are not a base — '-' inside a deletion /
// reference skip, 0 for a base asked past l_seq — become code 0 and are listed apart ("special" entries): pass 1
// packs and counts them per thread block, pass 2 (only when there are any) writes the list in entry order.
__device__ __forceinline__ uint32_t base_code(uint32_t ch) {
  // (ch & 31) is unique over the 16 IUPAC characters (= A B C D G H K M N R S T V W Y -> 29 1 2 3 4 7 8 11 13 14 18 19
  // 20 22 23 25): nibble k of lo (k < 16) / nibble k-16 of hi holds the code of the character with (ch & 31) == k
  const uint64_t lo = 0xf30c00b400d2e10ull, hi = 0xa097086500ull;
  const uint32_t k = ch & 31;
  return (uint32_t)(((k & 16) ? hi : lo) >> ((k & 15) * 4)) & 15;
}
constexpr int PACK_THREADS = 256;
__global__ void __launch_bounds__(PACK_THREADS) pack_base_kernel(const uint8_t* __restrict__ base, uint64_t n_entries,
                                                                 uint8_t* base4, uint32_t* block_special) {
  const uint64_t i = (uint64_t)blockIdx.x * PACK_THREADS + threadIdx.x;     // entries 2i, 2i+1
  uint32_t nsp = 0;
  if (2 * i < n_entries) {
    const uint32_t b0 = base[2 * i];
    const uint32_t b1 = 2 * i + 1 < n_entries ? base[2 * i + 1] : 'N';
    const bool s0 = b0 == '-' || b0 == 0, s1 = 2 * i + 1 < n_entries && (b1 == '-' || b1 == 0);
    base4[i] = (uint8_t)(((s0 ? 0 : base_code(b0)) << 4) | ((s1 || 2 * i + 1 >= n_entries) ? 0 : base_code(b1)));
    nsp = (s0 ? 1u : 0u) + (s1 ? 1u : 0u);
  }
  __shared__ uint32_t total;
  if (threadIdx.x == 0) total = 0;
  __syncthreads();
  if (nsp) atomicAdd(&total, nsp);
  __syncthreads();
  if (threadIdx.x == 0) block_special[blockIdx.x] = total;
}
__global__ void __launch_bounds__(PACK_THREADS) special_scatter_kernel(const uint8_t* __restrict__ base, uint64_t n_entries,
                                                                       const uint32_t* __restrict__ block_off,
                                                                       uint32_t* special_entry, uint8_t* special_base) {
  const uint32_t n_here = block_off[blockIdx.x + 1] - block_off[blockIdx.x];
  if (n_here == 0) return;
  // rare path: thread 0 of a block that has special entries walks its 512 entries in order
  if (threadIdx.x != 0) return;
  uint32_t out = block_off[blockIdx.x];
  const uint64_t e0 = (uint64_t)blockIdx.x * PACK_THREADS * 2;
  const uint64_t e1 = e0 + PACK_THREADS * 2 < n_entries ? e0 + PACK_THREADS * 2 : n_entries;
  for (uint64_t e = e0; e < e1; ++e) {
    const uint32_t b = base[e];
    if (b == '-' || b == 0) {
      special_entry[out] = (uint32_t)e;
      special_base[out] = (uint8_t)b;
      ++out;
    }
  }
}

// Column positions as runs of consecutive positions (one run per stretch of non-zero coverage, usually one per batch)
__global__ void run_flag_kernel(const uint64_t* __restrict__ col_pos, uint32_t n_col, uint32_t* flag) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  flag[c] = (c == 0 || col_pos[c] != col_pos[c - 1] + 1) ? 1u : 0u;
}
__global__ void run_scatter_kernel(const uint64_t* __restrict__ col_pos, uint32_t n_col, const uint32_t* __restrict__ flag,
                                   const uint32_t* __restrict__ incl, uint64_t* run_pos, uint32_t* run_first_col) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  if (flag[c]) {
    run_pos[incl[c] - 1] = col_pos[c];
    run_first_col[incl[c] - 1] = c;
  }
  if (c + 1 == n_col) run_first_col[incl[c]] = n_col;
}

// ---- carry to the next batch ------------------------------------------------------------------------
__global__ void carry_flag_kernel(uint32_t g0, uint32_t g1, const int32_t* eend, int64_t limit, uint32_t* flag) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  int32_t e = eend[j];
  flag[j - g0] = (e != DEAD && (int64_t)e > limit) ? 1u : 0u;
}

__global__ void carry_size_kernel(ReadsView v, uint32_t g0, uint32_t g1, const int32_t* block_size, const uint32_t* flag,
                                  uint64_t* bytes) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  // 4-byte prefix + body, padded to 4 so that unaligned-load helpers never straddle into the next record's page
  bytes[j - g0] = flag[j - g0] ? (((uint64_t)block_size[j] + 4 + 3) & ~3ull) : 0;
}

__global__ void carry_index_kernel(uint32_t g0, uint32_t g1, const uint32_t* flag, const uint32_t* slot_incl, uint32_t* cidx) {
  uint32_t j = g0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  if (flag[j - g0]) cidx[slot_incl[j - g0] - 1] = j;
}

__global__ void carry_copy_kernel(ReadsView v, uint32_t g0, const int32_t* block_size, const uint32_t* cidx, uint32_t n_carry_out,
                                  const uint64_t* byte_excl, CarryOut out) {
  // one warp per carried read
  uint32_t s = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (s >= n_carry_out) return;
  uint32_t j = cidx[s];
  uint64_t dst = byte_excl[j - g0];
  if (lane == 0) {
    out.pos[s] = v.pos[j];
    out.end_pos[s] = v.end_pos[j];
    out.ref_id[s] = v.ref_id[j];
    out.rec_off[s] = dst;
    out.bin_mq_nl[s] = v.bin_mq_nl[j];
    out.flag_nc[s] = v.flag_nc[j];
    out.l_seq[s] = v.l_seq[j];
    out.block_size[s] = block_size[j];
    out.gidx[s] = j < v.n_carry ? v.carry_gidx[j] : (uint32_t)(v.first_index + (j - v.n_carry));
  }
  const uint8_t* src = (j < v.n_carry ? v.carry_data : v.u) + v.rec_off[j];
  uint32_t n = (uint32_t)block_size[j] + 4;
  for (uint32_t i = lane; i < n; i += 32) out.data[dst + i] = src[i];
}

// shard bookkeeping: number of records of the batch that start before slice offset x (rec_off ascends)
__global__ void count_below_kernel(const uint64_t* __restrict__ rec_off, uint32_t n, uint64_t x, uint64_t* out) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t m = (lo + hi) >> 1;
    if (rec_off[m] < x) lo = m + 1; else hi = m;
  }
  *out = lo;
}

// Exact halos (include/biod_b200.h, biodb_pileup_shard_reach): for each of the nk later shard boundaries (ascending
// (ref, pos) keys), the smallest slice offset of an own record [first, n) that reaches across it — same reference and
// end_position beyond the boundary's position.  A read that does not reach boundary k reaches no later one.
__global__ void reach_kernel(const int32_t* __restrict__ ref_id, const int32_t* __restrict__ pos,
                             const int32_t* __restrict__ end_pos, const uint64_t* __restrict__ rec_off, uint32_t first, uint32_t n,
                             const int32_t* __restrict__ kref, const int64_t* __restrict__ kpos, uint32_t nk,
                             unsigned long long* reach) {
  const uint32_t i = first + blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int32_t e = end_pos[i], r = ref_id[i];
  if (e <= pos[i] || r < 0) return;                 // covers no position: not a read of any pileup
  for (uint32_t k = 0; k < nk; ++k) {
    const int32_t br = kref[k];
    if (br == r) {
      if ((int64_t)e > kpos[k]) atomicMin(&reach[k], (unsigned long long)rec_off[i]);
      else break;
    } else if (br < 0 || br > r) {
      break;                                        // boundary on a later reference (or among the unmapped reads)
    }
  }
}

template <typename K, typename... A>
inline void launch1d(K k, uint32_t n, cudaStream_t st, A... a) {
  if (n == 0) return;
  k<<<(n + 255) / 256, 256, 0, st>>>(a...);
  ++g_kernel_launches;
}

}  // namespace

void pileup_find_groups(const ReadsView& v, uint32_t* boundaries, uint32_t* n_boundaries, uint32_t cap, cudaStream_t st) {
  cudaMemsetAsync(n_boundaries, 0, sizeof(uint32_t), st);
  if (v.n > 1) launch1d(find_groups_kernel, v.n - 1, st, v, boundaries, n_boundaries, cap);
}

void pileup_count_below(const uint64_t* rec_off, uint32_t n, uint64_t x, uint64_t* out, cudaStream_t st) {
  count_below_kernel<<<1, 1, 0, st>>>(rec_off, n, x, out);
  ++g_kernel_launches;
}

void pileup_reach(const int32_t* ref_id, const int32_t* pos, const int32_t* end_pos, const uint64_t* rec_off, uint32_t first,
                  uint32_t n, const int32_t* kref, const int64_t* kpos, uint32_t nk, unsigned long long* reach, cudaStream_t st) {
  if (first >= n || nk == 0) return;
  launch1d(reach_kernel, n - first, st, ref_id, pos, end_pos, rec_off, first, n, kref, kpos, nk, reach);
}

void pileup_first_kept(const ReadsView& v, uint32_t g0, uint32_t g1, uint64_t start_from, uint32_t* first, cudaStream_t st) {
  cudaMemsetAsync(first, 0xff, sizeof(uint32_t), st);
  launch1d(first_kept_kernel, g1 - g0, st, v, g0, g1, start_from, first);
}

// Phase 1: liveness, premax, islands, island column counts.  After it (sync) the host reads
// n_islands = tmp32[tiles(n)] (flag scan total) and n_col = tmp32b total.
void pileup_phase1(const ReadsView& v, uint32_t g0, uint32_t g1, uint32_t drop_before, int skip_zero, int64_t clo,
                   int64_t chi, GroupScratch& s, cudaStream_t st) {
  const uint32_t n = g1 - g0;
  launch1d(prep_kernel, n, st, v, g0, g1, drop_before, s.eend, s.rinfo, s.info);
  device_scan<true>(s.eend + g0, s.pm + g0, n, s.tmp_i32, OpMax(), DEAD, st);
  launch1d(island_flag_kernel, n, st, v, g0, g1, s.eend, s.pm, skip_zero, s.flag);
  device_scan<true>(s.flag + g0, s.iid1 + g0, n, s.tmp_u32, OpAdd(), 0u, st);
  launch1d(island_table_kernel, n, st, v, g0, g1, s.flag, s.iid1, s.pm, s.islands);
  // n_islands is only known on the device: run the island kernels over the upper bound n and let
  // entries beyond n_islands be harmless (ncol for them is computed from stale table rows, so zero them first)
  s.clo = clo;
  s.chi = chi;
}

void pileup_island_cols(uint32_t n_islands, GroupScratch& s, cudaStream_t st) {
  launch1d(island_cols_kernel, n_islands, st, s.islands, n_islands, s.clo, s.chi, s.ncol);
  device_scan<false>(s.ncol, s.colbase, n_islands, s.tmp_u32b, OpAdd(), 0u, st);
}

// Phase 2: columns.  cols.* must have room for n_col+1 and be zeroed (diff, nstart, hi, lo).
void pileup_phase2(const ReadsView& v, uint32_t g0, uint32_t g1, uint32_t n_islands, uint32_t n_col, GroupScratch& s,
                   ColumnScratch& c, ColumnOutput& o, cudaStream_t st) {
  const uint32_t n = g1 - g0;
  cudaMemsetAsync(c.diff, 0, (size_t)(n_col + 1) * 4, st);
  cudaMemsetAsync(c.nstart, 0, (size_t)(n_col + 1) * 4, st);
  cudaMemsetAsync(c.hi, 0, (size_t)(n_col + 1) * 4, st);
  cudaMemsetAsync(c.lo, 0, (size_t)(n_col + 1) * 4, st);
  launch1d(scatter_kernel, n, st, v, g0, g1, s.eend, s.pm, s.iid1, s.islands, s.colbase, s.ncol, c);
  device_scan<true>(c.diff, c.diff, (uint64_t)n_col + 1, s.tmp_i32, OpAdd(), 0, st);            // coverage
  device_scan<false>(c.diff, o.col_off, (uint64_t)n_col + 1, s.tmp_u64, OpAdd(), (uint64_t)0, st);  // col_off
  device_scan<true>(c.hi, c.hi, (uint64_t)n_col, s.tmp_u32, OpMax(), 0u, st);
  device_scan<true>(c.lo, c.lo, (uint64_t)n_col, s.tmp_u32b, OpMax(), 0u, st);
  launch1d(colpos_kernel, n_col, st, s.islands, s.colbase, n_islands, n_col, o.col_pos);
}

void pileup_entries(const ReadsView& v, uint32_t n_col, GroupScratch& s, ColumnScratch& c, ColumnOutput& o, cudaStream_t st,
                    const uint32_t* redo) {
  if (n_col == 0) return;
  static_assert(CHUNK == TCHUNK, "redo[] is indexed by the chunks of both kernels");
  uint32_t warps = (n_col + CHUNK - 1) / CHUNK;
  uint32_t grid = (warps + ENT_WARPS - 1) / ENT_WARPS;
  if (o.maq_min_base_quality >= 0) entries_kernel<false, false, true><<<grid, ENT_WARPS * 32, 0, st>>>(v, s.eend, s.rinfo, c, o, n_col, redo, s.info);
  else if (o.counts) entries_kernel<true, false, false><<<grid, ENT_WARPS * 32, 0, st>>>(v, s.eend, s.rinfo, c, o, n_col, redo, s.info);
  else if (o.qoff) entries_kernel<false, true, false><<<grid, ENT_WARPS * 32, 0, st>>>(v, s.eend, s.rinfo, c, o, n_col, redo, s.info);
  else entries_kernel<false, false, false><<<grid, ENT_WARPS * 32, 0, st>>>(v, s.eend, s.rinfo, c, o, n_col, redo, s.info);
  ++g_kernel_launches;
}

// Column-stationary entries (entries_tile_kernel) for the plain, compact and MAQ forms; chunks too deep for the tile are
// flagged in redo[] and done by entries_kernel (pileup_entries with the same arguments and redo).  mode: 0 / 1 / 2.
uint32_t pileup_tile_chunks(uint32_t n_col) { return (n_col + TCHUNK - 1) / TCHUNK; }
void pileup_entries_tile(const ReadsView& v, uint32_t n_col, GroupScratch& s, ColumnScratch& c, ColumnOutput& o, int mode,
                         uint32_t* last_read, uint64_t* live_mask, uint32_t* nstrag, uint32_t* redo, cudaStream_t st) {
  if (n_col == 0) return;
  const uint32_t warps = pileup_tile_chunks(n_col);
  const uint32_t grid = (warps + TILE_WARPS - 1) / TILE_WARPS;
  if (mode == 0) entries_tile_kernel<0><<<grid, TILE_WARPS * 32, 0, st>>>(v, s.eend, s.rinfo, c, o, n_col, last_read, live_mask, nstrag, redo, s.info);
  else if (mode == 1) entries_tile_kernel<1><<<grid, TILE_WARPS * 32, 0, st>>>(v, s.eend, s.rinfo, c, o, n_col, last_read, live_mask, nstrag, redo, s.info);
  else entries_tile_kernel<2><<<grid, TILE_WARPS * 32, 0, st>>>(v, s.eend, s.rinfo, c, o, n_col, last_read, live_mask, nstrag, redo, s.info);
  ++g_kernel_launches;
}

void pileup_strag_walk(const ReadsView& v, uint32_t n_col, GroupScratch& s, const uint32_t* lo, const uint32_t* hi,
                       const uint64_t* col_pos, const uint32_t* strag_off, uint32_t* strag_col, uint32_t* strag_idx,
                       cudaStream_t st) {
  launch1d(strag_walk_kernel, n_col, st, v, s.eend, s.rinfo, lo, hi, col_pos, n_col, strag_off, strag_col, strag_idx);
}

// Compact read lists, step 1: per-column last read, window mask and straggler offsets (exclusive scan; the total is
// strag_off[n_col], read by the host after a sync).  Step 2 copies the stragglers.
void pileup_compact_masks(uint32_t n_col, const ColumnOutput& o, uint32_t* last_read, uint64_t* mask, uint32_t* nstrag,
                          uint32_t* strag_off, GroupScratch& s, cudaStream_t st, const uint32_t* redo) {
  launch1d(compact_mask_kernel, n_col + 1, st, o.col_off, o.read_idx, n_col, last_read, mask, nstrag, redo);
  device_scan<false>(nstrag, strag_off, (uint64_t)n_col + 1, s.tmp_u32, OpAdd(), 0u, st);
}

void pileup_compact_stragglers(uint32_t n_col, const ColumnOutput& o, const uint32_t* strag_off, uint32_t* strag_col,
                               uint32_t* strag_idx, cudaStream_t st) {
  launch1d(compact_strag_kernel, n_col, st, o.col_off, o.read_idx, n_col, strag_off, strag_col, strag_idx);
}

// Packed bases: pass 1 (base4 + special counts per block, then their exclusive scan: block_off[n_blocks] = total).
uint32_t pileup_pack_blocks(uint64_t n_entries) { return (uint32_t)((n_entries + 2 * PACK_THREADS - 1) / (2 * PACK_THREADS)); }
void pileup_pack_bases(uint64_t n_entries, const uint8_t* base, uint8_t* base4, uint32_t* block_special, uint32_t* block_off,
                       GroupScratch& s, uint32_t* scan_tmp, cudaStream_t st) {
  const uint32_t nb = pileup_pack_blocks(n_entries);
  if (nb == 0) return;
  pack_base_kernel<<<nb, PACK_THREADS, 0, st>>>(base, n_entries, base4, block_special);
  ++g_kernel_launches;
  cudaMemsetAsync(block_special + nb, 0, 4, st);
  device_scan<false>(block_special, block_off, (uint64_t)nb + 1, scan_tmp, OpAdd(), 0u, st);
}
void pileup_pack_specials(uint64_t n_entries, const uint8_t* base, const uint32_t* block_off, uint32_t* special_entry,
                          uint8_t* special_base, cudaStream_t st) {
  const uint32_t nb = pileup_pack_blocks(n_entries);
  if (nb == 0) return;
  special_scatter_kernel<<<nb, PACK_THREADS, 0, st>>>(base, n_entries, block_off, special_entry, special_base);
  ++g_kernel_launches;
}

// Position runs: flag + inclusive scan (n_runs = incl[n_col-1], read by the host after a sync), then the scatter.
void pileup_position_runs_scan(uint32_t n_col, const ColumnOutput& o, uint32_t* flag, uint32_t* incl, GroupScratch& s,
                               cudaStream_t st) {
  launch1d(run_flag_kernel, n_col, st, o.col_pos, n_col, flag);
  device_scan<true>(flag, incl, (uint64_t)n_col, s.tmp_u32b, OpAdd(), 0u, st);
}
void pileup_position_runs_scatter(uint32_t n_col, const ColumnOutput& o, const uint32_t* flag, const uint32_t* incl,
                                  uint64_t* run_pos, uint32_t* run_first_col, cudaStream_t st) {
  launch1d(run_scatter_kernel, n_col, st, o.col_pos, n_col, flag, incl, run_pos, run_first_col);
}

// Carry: live reads of [g0,g1) with end > limit, compacted in order into `out`.
// After the call (sync) the host reads the count = tmp_u32 total and bytes = tmp_u64 total.
void pileup_carry(const ReadsView& v, uint32_t g0, uint32_t g1, const int32_t* block_size, int64_t limit, GroupScratch& s,
                  CarryOut& out, cudaStream_t st) {
  const uint32_t n = g1 - g0;
  if (n == 0) return;
  launch1d(carry_flag_kernel, n, st, g0, g1, s.eend, limit, s.cflag);
  launch1d(carry_size_kernel, n, st, v, g0, g1, block_size, s.cflag, s.cbytes);
  device_scan<true>(s.cflag, s.cslot, n, s.tmp_u32, OpAdd(), 0u, st);
  device_scan<false>(s.cbytes, s.cbytes, n, s.tmp_u64, OpAdd(), (uint64_t)0, st);
}

void pileup_carry_copy(const ReadsView& v, uint32_t g0, uint32_t g1, const int32_t* block_size, uint32_t n_carry_out,
                       GroupScratch& s, CarryOut& out, cudaStream_t st) {
  const uint32_t n = g1 - g0;
  if (n == 0 || n_carry_out == 0) return;
  uint32_t* cidx = s.ncol;     // the island column counts are dead once the entries are written
  launch1d(carry_index_kernel, n, st, g0, g1, s.cflag, s.cslot, cidx);
  uint64_t threads = (uint64_t)n_carry_out * 32;
  carry_copy_kernel<<<(uint32_t)((threads + 255) / 256), 256, 0, st>>>(v, g0, block_size, cidx, n_carry_out, s.cbytes, out);
  ++g_kernel_launches;
}

}  // namespace biodb
