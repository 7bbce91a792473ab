// Region filter over the record tables of a batch: BamReadFilter of
// bio/std/hts/bam/randomaccessmanager.d:366-462, for one region [beg, end) — or several sorted, non-overlapping
// regions — of one reference.
//
// The reference walks the reads of the index's chunks one by one: reads of earlier references are skipped, the first
// read of a later reference or with position >= end ends the range, and a read is kept when it starts inside the
// region or reaches into it (position + basesCovered() > beg; zero-length reads that start at or before beg are not
// kept).  Here: one pass flags the reads and finds the first read that ends the range (atomicMin), two device-wide
// scans number the kept reads and their CIGAR words, and one gather writes the compacted SoA tables the C ABI
// hands out.  The raw record bytes stay where they are: rec_off keeps pointing into the batch's slice.
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "scan.cuh"

namespace biodb {

namespace {

__global__ void region_stop_kernel(const int32_t* __restrict__ ref_id, const int32_t* __restrict__ pos, uint32_t n,
                                   uint32_t ref, uint32_t end, uint32_t* first_stop) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  // D compares the int fields against the region's uint fields as unsigned (ref_id -1 is "later than any reference")
  const uint32_t cur = (uint32_t)ref_id[j];
  if (cur > ref || (cur == ref && (uint32_t)pos[j] >= end)) atomicMin(first_stop, j);
}

__global__ void region_keep_kernel(const int32_t* __restrict__ ref_id, const int32_t* __restrict__ pos,
                                   const int32_t* __restrict__ end_pos, const uint32_t* __restrict__ flag_nc, uint32_t n,
                                   uint32_t ref, uint32_t beg, const uint32_t* __restrict__ first_stop,
                                   uint32_t* __restrict__ keep, uint32_t* __restrict__ ccnt,
                                   const uint32_t* __restrict__ regs, uint32_t n_regs) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n) return;
  bool k = j < *first_stop && (uint32_t)ref_id[j] == ref;
  if (k && n_regs > 1) {
    // several regions (randomaccessmanager.d:396-450): the filter's region pointer only moves forward, when a read starts
    // at or beyond the current region's end — so a read is judged against the first region that ends behind its position
    const uint32_t p = (uint32_t)pos[j];
    uint32_t lo = 0, hi = n_regs;
    while (lo < hi) {
      const uint32_t mid = (lo + hi) >> 1;
      if (regs[2 * mid + 1] > p) hi = mid; else lo = mid + 1;
    }
    beg = regs[2 * min(lo, n_regs - 1)];                                 // (lo < n_regs: the read is in front of first_stop)
  }
  if (k) k = (uint32_t)pos[j] > beg || (uint32_t)end_pos[j] > beg;       // end_pos = position + basesCovered()
  keep[j] = k ? 1u : 0u;
  ccnt[j] = k ? (flag_nc[j] & 0xFFFFu) : 0u;
}

__global__ void region_gather_kernel(RecordArrays in, RecordArrays out, uint32_t n, const uint32_t* __restrict__ keep,
                                     const uint32_t* __restrict__ slot_incl, const uint32_t* __restrict__ coff_excl,
                                     const uint32_t* __restrict__ totals /* [0] kept, [1] cigar words */) {
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j == 0) out.cigar_off[totals[0]] = totals[1];
  if (j >= n || !keep[j]) return;
  const uint32_t s = slot_incl[j] - 1;
  out.rec_off[s] = in.rec_off[j];
  out.block_size[s] = in.block_size[j];
  out.ref_id[s] = in.ref_id[j];
  out.pos[s] = in.pos[j];
  out.end_pos[s] = in.end_pos[j];
  out.bin_mq_nl[s] = in.bin_mq_nl[j];
  out.flag_nc[s] = in.flag_nc[j];
  out.l_seq[s] = in.l_seq[j];
  const uint32_t c0 = coff_excl[j], nc = in.flag_nc[j] & 0xFFFFu;
  out.cigar_off[s] = c0;
  const uint32_t* src = in.cigar + in.cigar_off[j];
  for (uint32_t k = 0; k < nc; ++k) out.cigar[c0 + k] = src[k];
}

__global__ void region_totals_kernel(const uint32_t* a, const uint32_t* b, uint32_t* totals) {
  totals[0] = *a;
  totals[1] = *b;
}

}  // namespace

size_t region_scratch_elems(uint64_t n) { return 4 * (size_t)(n + 8) + 2 * (scan_temp_elems(n) + 8) + 16; }

cudaError_t launch_region_filter(const RecordArrays& in, uint64_t n64, uint32_t ref, uint32_t beg, uint32_t end,
                                 const RecordArrays& out, uint32_t* scratch, uint32_t* info, cudaStream_t st,
                                 const uint32_t* regs, uint32_t n_regs) {
  const uint32_t n = (uint32_t)n64;
  uint32_t* keep = scratch;
  uint32_t* slot = keep + (n + 8);
  uint32_t* ccnt = slot + (n + 8);
  uint32_t* coff = ccnt + (n + 8);
  uint32_t* tmp_a = coff + (n + 8);
  uint32_t* tmp_b = tmp_a + scan_temp_elems(n) + 8;
  cudaMemsetAsync(info, 0xff, 4, st);                                // info[0] = first read that ends the range
  if (n) {
    const uint32_t grid = (n + 255) / 256;
    region_stop_kernel<<<grid, 256, 0, st>>>(in.ref_id, in.pos, n, ref, end, info);
    region_keep_kernel<<<grid, 256, 0, st>>>(in.ref_id, in.pos, in.end_pos, in.flag_nc, n, ref, beg, info, keep, ccnt, regs, n_regs);
    g_kernel_launches += 2;
  }
  device_scan<true>(keep, slot, n, tmp_a, OpAdd(), 0u, st);          // tmp_a[tiles] = number of kept reads
  device_scan<false>(ccnt, coff, n, tmp_b, OpAdd(), 0u, st);         // tmp_b[tiles] = their CIGAR words
  const uint32_t tiles = (uint32_t)((n + SCAN_TILE - 1) / SCAN_TILE);
  region_totals_kernel<<<1, 1, 0, st>>>(tmp_a + tiles, tmp_b + tiles, info + 1);    // info[1], info[2]
  region_gather_kernel<<<(n + 255) / 256 + 1, 256, 0, st>>>(in, out, n, keep, slot, coff, info + 1);
  g_kernel_launches += 2;
  return cudaGetLastError();
}

}  // namespace biodb
