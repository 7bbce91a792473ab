// BAM record scanner for sm_100a.
//
// Replaces the per-record loop of BamReadRange.readNext (bio/std/hts/bam/readrange.d:118-173):
// `block_size:i32` + body framing, and the fixed-field getters of BamRead
// (bam/read.d:907-1003), CigarOperation (bam/cigar.d:103-131) and basesCovered
// (read.d:255-262) -> end_position (read.d:1380-1383), as SoA tables.
//
// The record chain is sequential (each block_size gives the next offset).  It is walked
// speculatively in parallel, one warp per BGZF block, assuming a record starts at the block
// boundary — true for files written by BioD (bam/writer.d:259-267) and htslib unless a record
// is larger than a block — then verified (`out[b-1] == start[b]`) and, only where the
// speculation failed (records straddling blocks, e.g. mg1655_chunk.bam), re-walked in order.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace biodb {

namespace {

constexpr int WARPS = 4;

// unaligned little-endian 32-bit load built from two aligned ones
__device__ __forceinline__ uint32_t ld32u(const uint8_t* p) {
  uintptr_t a = (uintptr_t)p;
  const uint32_t* w = (const uint32_t*)(a & ~(uintptr_t)3);
  uint32_t sh = (uint32_t)(a & 3) * 8;
  uint32_t lo = __ldg(w);
  if (sh == 0) return lo;
  uint32_t hi = __ldg(w + 1);
  return __funnelshift_r(lo, hi, sh);
}

// consume bits of CIGAR_TYPE (bam/cigar.d:116): bit0 query-consuming, bit1 reference-consuming
__device__ __forceinline__ uint32_t cigar_consume(uint32_t raw) { return (0x3C1A7u >> ((raw & 0xF) * 2)) & 3; }


// Walk the record chain from absolute offset `p` while record starts lie inside [*, end).
// Returns the stop reason; *out = offset where the walk stopped.
__device__ int walk_block(const uint8_t* u, uint64_t u_len, uint64_t p, uint64_t blk_start, uint64_t end,
                          uint16_t* rel, uint32_t* cnt_out, uint32_t* ncig_out, uint64_t* out, uint32_t cnt0 = 0,
                          uint32_t ncig0 = 0) {
  uint32_t cnt = cnt0, ncig = ncig0;
  int why = WALK_OK;
  while (p < end) {
    if (p + 4 > u_len) { why = WALK_TAIL; break; }
    const uint8_t* r = u + p;
    int32_t bs = (int32_t)ld32u(r);
    if (bs < 32) { why = WALK_BAD_SIZE; break; }
    if (p + 4 + (uint64_t)bs > u_len) { why = WALK_TAIL; break; }
    uint32_t bin_mq_nl = ld32u(r + 12), flag_nc = ld32u(r + 16);
    int32_t l_seq = (int32_t)ld32u(r + 20);
    uint32_t lname = bin_mq_nl & 0xFF, nc = flag_nc & 0xFFFF;
    uint64_t need = 32ull + lname + 4ull * nc + (l_seq > 0 ? ((uint64_t)l_seq + 1) / 2 + (uint64_t)l_seq : 0);
    if (l_seq < 0 || need > (uint64_t)bs) { why = WALK_BAD_FIELDS; break; }
    if (cnt < (uint32_t)SCAN_SLOTS) rel[cnt] = (uint16_t)(p - blk_start);
    ++cnt;
    ncig += nc;
    p += 4 + (uint64_t)bs;
  }
  *cnt_out = cnt;
  *ncig_out = ncig;
  *out = p;
  return why;
}

__global__ void __launch_bounds__(WARPS * 32) scan_walk_kernel(const uint8_t* __restrict__ u, uint64_t u_len,
                                                               const uint64_t* __restrict__ block_uoff,
                                                               uint32_t n_blocks, ScanWorkspace ws) {
  uint32_t b = blockIdx.x * WARPS + (threadIdx.x >> 5);
  if (b >= n_blocks || (threadIdx.x & 31) != 0) return;
  uint64_t start = block_uoff[b], end = block_uoff[b + 1];
  uint32_t cnt, ncig;
  uint64_t out;
  int why = walk_block(u, u_len, start, start, end, ws.rel + (size_t)b * SCAN_SLOTS, &cnt, &ncig, &out);
  ws.cnt[b] = cnt;
  ws.ncig[b] = ncig;
  ws.out[b] = out;
  ws.in[b] = start;
  ws.bad[b] = why;
}

// Verify the speculation, repair it sequentially where it failed, then scan the per-block counts.
__global__ void __launch_bounds__(1024) scan_resolve_kernel(const uint8_t* __restrict__ u, uint64_t u_len,
                                                           const uint64_t* __restrict__ block_uoff, uint32_t n_blocks,
                                                           int final_slice, ScanWorkspace ws, RecordArrays ra,
                                                           uint64_t* __restrict__ result) {
  constexpr uint32_t FAIL_WORDS = 1024;           // bitmap of inconsistent blocks (32768 blocks)
  __shared__ uint32_t s_failbits[FAIL_WORDS];
  __shared__ uint32_t s_nfail;
  __shared__ uint64_t s_scan[2][256];
  __shared__ uint64_t s_carry[2];
  __shared__ uint32_t s_last;     // last block whose records count (the chain stops inside or after it)
  __shared__ int s_why;           // why the chain stopped there
  const uint32_t t = threadIdx.x;
  if (n_blocks == 0) {
    if (t == 0) { result[0] = 0; result[1] = 0; result[2] = 0; result[3] = 0; ra.cigar_off[0] = 0; }
    return;
  }
  // phase 0 (parallel): a fused walk stops when the header of its last record straddles the block end — finish
  // those walks from where they stopped (one record each, common in htsjdk-written files)
  for (uint32_t b = t; b < n_blocks; b += blockDim.x) {
    if (ws.bad[b] == WALK_INCOMPLETE) {
      uint32_t cnt, ncig;
      uint64_t out;
      ws.bad[b] = walk_block(u, u_len, ws.out[b], block_uoff[b], block_uoff[b + 1], ws.rel + (size_t)b * SCAN_SLOTS, &cnt, &ncig,
                             &out, ws.cnt[b], ws.ncig[b]);
      ws.cnt[b] = cnt; ws.ncig[b] = ncig; ws.out[b] = out;
    }
  }
  for (uint32_t w = t; w < FAIL_WORDS; w += blockDim.x) s_failbits[w] = 0;
  if (t == 0) { s_nfail = 0; s_carry[0] = 0; s_carry[1] = 0; }
  __syncthreads();
  // phase 1 (parallel): block b is inconsistent if the chain of block b-1 does not end exactly where block b's
  // walk entered, or if that chain stopped early
  for (uint32_t b = 1 + t; b < n_blocks; b += blockDim.x) {
    if (ws.out[b - 1] != ws.in[b] || ws.bad[b - 1] != WALK_OK) {
      if (b < FAIL_WORDS * 32) atomicOr(&s_failbits[b >> 5], 1u << (b & 31));
      atomicAdd(&s_nfail, 1u);
    }
  }
  __syncthreads();
  // phase 2 (one thread, rare): repair the inconsistent blocks in order; a repaired chain usually rejoins the
  // speculated one at once, so only the flagged blocks are visited
  if (t == 0) {
    uint32_t last = n_blocks - 1;
    int why = ws.bad[n_blocks - 1];
    if (s_nfail) {
      auto next_fail = [&](uint32_t from) -> uint32_t {       // first flagged block >= from
        if (n_blocks > FAIL_WORDS * 32) {                      // bitmap too small: plain scan
          for (uint32_t b = from; b < n_blocks; ++b)
            if (b > 0 && (ws.out[b - 1] != ws.in[b] || ws.bad[b - 1] != WALK_OK)) return b;
          return n_blocks;
        }
        for (uint32_t w = from >> 5; w < FAIL_WORDS && w * 32 < n_blocks; ++w) {
          uint32_t bits = s_failbits[w];
          if (w == (from >> 5)) bits &= ~0u << (from & 31);
          if (bits) return w * 32 + (uint32_t)__ffs(bits) - 1;
        }
        return n_blocks;
      };
      uint32_t b = next_fail(1);
      bool stopped = false;
      while (b < n_blocks && !stopped) {
        if (ws.bad[b - 1] != WALK_OK) { last = b - 1; why = ws.bad[b - 1]; stopped = true; break; }
        uint64_t in = ws.out[b - 1];
        // walk forward from b until the chain rejoins the speculated one
        for (; b < n_blocks; ++b) {
          const uint64_t start = block_uoff[b], end = block_uoff[b + 1];
          if (in == ws.in[b]) break;             // rejoined: block b was walked from the right entry
          if (in >= end) {                       // the whole block lies inside a straddling record
            ws.cnt[b] = 0; ws.ncig[b] = 0; ws.out[b] = in; ws.in[b] = in; ws.bad[b] = WALK_OK;
            continue;
          }
          uint32_t cnt, ncig;
          uint64_t out;
          ws.bad[b] = walk_block(u, u_len, in, start, end, ws.rel + (size_t)b * SCAN_SLOTS, &cnt, &ncig, &out);
          ws.cnt[b] = cnt; ws.ncig[b] = ncig; ws.out[b] = out; ws.in[b] = in;
          if (ws.bad[b] != WALK_OK) { last = b; why = ws.bad[b]; stopped = true; break; }
          in = out;
        }
        if (stopped || b >= n_blocks) break;
        b = next_fail(b + 1);
      }
      if (!stopped) { last = n_blocks - 1; why = ws.bad[n_blocks - 1]; }
    }
    s_last = last;
    s_why = why;
  }
  __syncthreads();
  const uint32_t last = s_last;
  // exclusive scans of cnt / ncig; blocks past `last` hold no records.  Both 32-bit counts ride in one u64
  // (records < 2^32 and cigar words < 2^32 per slice) through a shuffle-based block scan.
  for (uint32_t base = 0; base < n_blocks + 1; base += blockDim.x) {
    const uint32_t b = base + t;
    uint64_t c = 0, g = 0;
    if (b < n_blocks) {
      if (b > last) { ws.cnt[b] = 0; ws.ncig[b] = 0; }
      c = ws.cnt[b];
      g = ws.ncig[b];
    }
    const uint32_t lane = t & 31, wid = t >> 5;
    uint64_t vc = c, vg = g;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint64_t oc = __shfl_up_sync(0xffffffffu, vc, d), og = __shfl_up_sync(0xffffffffu, vg, d);
      if (lane >= (uint32_t)d) { vc += oc; vg += og; }
    }
    if (lane == 31) { s_scan[0][wid] = vc; s_scan[1][wid] = vg; }
    __syncthreads();
    uint64_t pc = 0, pg = 0, tc = 0, tg = 0;
    for (uint32_t w = 0; w < blockDim.x / 32; ++w) {
      const uint64_t xc = s_scan[0][w], xg = s_scan[1][w];
      if (w < wid) { pc += xc; pg += xg; }
      tc += xc;
      tg += xg;
    }
    if (b <= n_blocks) {
      ws.rec_base[b] = s_carry[0] + pc + vc - c;
      ws.cig_base[b] = s_carry[1] + pg + vg - g;
    }
    __syncthreads();
    if (t == 0) { s_carry[0] += tc; s_carry[1] += tg; }
    __syncthreads();
  }
  if (t == 0) {
    const uint64_t n = ws.rec_base[n_blocks], ncg = ws.cig_base[n_blocks];
    const uint64_t tail = ws.out[last];
    const int why = s_why;
    int status = 0;
    if (why == WALK_BAD_SIZE || why == WALK_BAD_FIELDS) status = -4;          // BIODB_ERR_TRUNCATED
    else if (why == WALK_TAIL && final_slice) {
      // readrange.d:139-150: fewer than 4 bytes left ends the range silently; a cut body throws (:169)
      if (tail + 4 <= u_len) status = -4;
    }
    if (n > ra.capacity || ncg > ra.cigar_capacity) status = -10;             // BIODB_ERR_NOMEM
    result[0] = n;
    result[1] = tail;
    result[2] = ncg;
    result[3] = (uint64_t)(int64_t)status;
    if (n <= ra.capacity) ra.cigar_off[n] = ncg;
  }
}

// Field extraction: one warp per BGZF block, one lane per record.
__global__ void __launch_bounds__(WARPS * 32) scan_extract_kernel(const uint8_t* __restrict__ u,
                                                                  const uint64_t* __restrict__ block_uoff,
                                                                  uint32_t n_blocks, ScanWorkspace ws, RecordArrays ra) {
  const uint32_t b = blockIdx.x * WARPS + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (b >= n_blocks) return;
  const uint32_t cnt = ws.cnt[b];
  if (cnt == 0) return;
  const uint64_t rbase = ws.rec_base[b];
  uint64_t cbase = ws.cig_base[b];
  if (rbase + cnt > ra.capacity) return;
  const uint64_t start = block_uoff[b];
  const uint16_t* rel = ws.rel + (size_t)b * SCAN_SLOTS;
  for (uint32_t r0 = 0; r0 < cnt; r0 += 32) {
    const uint32_t r = r0 + lane;
    const bool live = r < cnt;
    uint64_t off = 0;
    uint32_t nc = 0, lname = 0, flag_nc = 0;
    int32_t pos = 0;
    const uint8_t* rec = u;
    if (live) {
      off = start + rel[r];
      rec = u + off;
      int32_t bs = (int32_t)ld32u(rec);
      int32_t ref_id = (int32_t)ld32u(rec + 4);
      pos = (int32_t)ld32u(rec + 8);
      uint32_t bin_mq_nl = ld32u(rec + 12);
      flag_nc = ld32u(rec + 16);
      int32_t l_seq = (int32_t)ld32u(rec + 20);
      lname = bin_mq_nl & 0xFF;
      nc = flag_nc & 0xFFFF;
      const uint64_t i = rbase + r;
      ra.rec_off[i] = off;
      ra.block_size[i] = bs;
      ra.ref_id[i] = ref_id;
      ra.pos[i] = pos;
      ra.bin_mq_nl[i] = bin_mq_nl;
      ra.flag_nc[i] = flag_nc;
      ra.l_seq[i] = l_seq;
    }
    // exclusive prefix of n_cigar over the 32 records of this step
    uint32_t incl = nc;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      uint32_t v = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= (uint32_t)d) incl += v;
    }
    const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
    if (live) {
      const uint64_t coff = cbase + (incl - nc);
      ra.cigar_off[rbase + r] = coff;
      uint32_t covered = 0;
      const uint8_t* cg = rec + 36 + lname;
      const bool fits = coff + nc <= ra.cigar_capacity;
      for (uint32_t k = 0; k < nc; ++k) {
        uint32_t raw = ld32u(cg + 4 * k);
        if (fits) ra.cigar[coff + k] = raw;
        if (cigar_consume(raw) & 2) covered += raw >> 4;
      }
      if ((flag_nc >> 16) & 0x4) covered = 0;     // is_unmapped (read.d:257-259)
      ra.end_pos[rbase + r] = (int32_t)((uint32_t)pos + covered);
    }
    cbase += total;
  }
}

}  // namespace

namespace {
__global__ void copy_bytes_kernel(uint8_t* __restrict__ dst, const uint8_t* __restrict__ src, size_t bytes) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 15) == 0) {
    const uint4* s4 = (const uint4*)src;
    uint4* d4 = (uint4*)dst;
    for (size_t n = bytes >> 4; i < n; i += stride) d4[i] = s4[i];
  } else if ((((uintptr_t)dst | (uintptr_t)src | bytes) & 3) == 0) {
    const uint32_t* s4 = (const uint32_t*)src;
    uint32_t* d4 = (uint32_t*)dst;
    for (size_t n = bytes >> 2; i < n; i += stride) d4[i] = s4[i];
  } else {
    for (; i < bytes; i += stride) dst[i] = src[i];
  }
}
}  // namespace

cudaError_t launch_copy_bytes(void* dst, const void* src, size_t bytes, cudaStream_t st) {
  if (bytes == 0) return cudaSuccess;
  size_t blocks = (bytes / 16 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 2048) blocks = 2048;
  copy_bytes_kernel<<<(unsigned)blocks, 256, 0, st>>>((uint8_t*)dst, (const uint8_t*)src, bytes);
  ++g_kernel_launches;
  return cudaGetLastError();
}

size_t scan_workspace_bytes(uint32_t n_blocks) {
  size_t nb = n_blocks + 1;
  size_t bytes = 0;
  bytes += (size_t)nb * SCAN_SLOTS * sizeof(uint16_t);
  bytes += nb * (2 * sizeof(uint32_t) + 4 * sizeof(uint64_t) + sizeof(int32_t)) + 256;
  return (bytes + 255) & ~(size_t)255;
}

ScanWorkspace carve_scan_workspace(void* base, uint32_t n_blocks) {
  ScanWorkspace ws;
  size_t nb = n_blocks + 1;
  uint8_t* p = (uint8_t*)base;
  ws.out = (uint64_t*)p; p += nb * 8;
  ws.in = (uint64_t*)p; p += nb * 8;
  ws.rec_base = (uint64_t*)p; p += nb * 8;
  ws.cig_base = (uint64_t*)p; p += nb * 8;
  ws.cnt = (uint32_t*)p; p += nb * 4;
  ws.ncig = (uint32_t*)p; p += nb * 4;
  ws.bad = (int32_t*)p; p += nb * 4;
  p = (uint8_t*)(((uintptr_t)p + 15) & ~(uintptr_t)15);
  ws.rel = (uint16_t*)p;
  return ws;
}

cudaError_t launch_scan_records(const uint8_t* u, uint64_t u_len, const uint64_t* block_uoff, uint32_t n_blocks,
                                uint32_t n_walk, int final_slice, const RecordArrays& out, uint64_t* result,
                                const ScanWorkspace& ws, cudaStream_t st) {
  uint32_t grid = (n_blocks + WARPS - 1) / WARPS;
  if (n_walk > n_blocks) n_walk = n_blocks;
  if (n_walk) scan_walk_kernel<<<(n_walk + WARPS - 1) / WARPS, WARPS * 32, 0, st>>>(u, u_len, block_uoff, n_walk, ws);
  g_kernel_launches += (n_blocks ? 2 : 1) + (n_walk ? 1 : 0);
  scan_resolve_kernel<<<1, 1024, 0, st>>>(u, u_len, block_uoff, n_blocks, final_slice, ws, out, result);
  if (n_blocks) scan_extract_kernel<<<grid, WARPS * 32, 0, st>>>(u, block_uoff, n_blocks, ws, out);
  return cudaGetLastError();
}

}  // namespace biodb
