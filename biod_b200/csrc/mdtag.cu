// Reference bases from MD tags on the device (PileupRangeUsingMdTag, bam/pileup.d:522-654).
//
// The reference asks one read at a time for its dna() string (md/reconstruct.d:38-214) and hands out one character of
// it per column.  Here the work is split three ways:
//   md_len_kernel     one thread per read: the LENGTH of dna(read) (the whole MD / CIGAR / SEQ walk without output) —
//                     all the host-side chain (md_chain.h) needs to know which read serves which positions;
//   md_replay_kernel  one thread per chain segment: replays the provider's dna() into reference_base[] of the columns
//                     the segment covers (every other column keeps 'N');
//   md_keep_kernel    the dna() strings of the (at most two) providers that the next batch may still ask for are
//                     written out, so that they outlive the record bytes of this batch.
// The walk itself (DnaWalk, md_walk.h) is the same code the CPU tests check (tests/test_md_chain.py).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"
#include "md_walk.h"
#include "pileup.h"

namespace biodb {

namespace {

constexpr int32_t DEAD = INT32_MIN;

struct Source {             // where the dna() of a read comes from
  const uint8_t* body;      // raw record (after block_size), or nullptr
  int64_t block_size;
  const uint8_t* str;       // an already materialised dna() string, or nullptr
  uint32_t str_len;
};

__device__ __forceinline__ Source locate(const ReadsView& v, const int32_t* block_size, const MdKeep& keep, uint64_t id) {
  Source s{nullptr, 0, nullptr, 0};
  const uint32_t n_new = v.n - v.n_carry;
  uint32_t j = 0xffffffffu;
  if (id >= v.first_index && id - v.first_index < (uint64_t)n_new) {
    j = v.n_carry + (uint32_t)(id - v.first_index);
  } else if (id < v.first_index) {
    // carried reads keep file order: binary search for the index
    const uint32_t key = (uint32_t)id;
    uint32_t lo = 0, hi = v.n_carry;
    while (lo < hi) {
      const uint32_t m = (lo + hi) >> 1;
      if (v.carry_gidx[m] < key) lo = m + 1; else hi = m;
    }
    if (lo < v.n_carry && v.carry_gidx[lo] == key) j = lo;
  }
  if (j != 0xffffffffu) {
    s.body = (j < v.n_carry ? v.carry_data : v.u) + v.rec_off[j] + 4;
    s.block_size = block_size[j];
    return s;
  }
  for (int k = 0; k < 2; ++k)
    if (keep.id[k] == id) {
      s.str = keep.data[k];
      s.str_len = keep.len[k];
      return s;
    }
  return s;
}

__global__ void md_len_kernel(ReadsView v, const int32_t* __restrict__ block_size, const int32_t* __restrict__ eend,
                              uint32_t a0, uint32_t g1, int32_t* __restrict__ dna_len) {
  const uint32_t j = a0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= g1) return;
  int32_t n = 0;
  if (eend[j] != DEAD) {
    const uint8_t* body = (j < v.n_carry ? v.carry_data : v.u) + v.rec_off[j] + 4;
    DnaWalk w;
    w.init(body, block_size[j]);
    while (w.next() >= 0) ++n;
  }
  dna_len[j] = n;
}

__global__ void md_replay_kernel(ReadsView v, const int32_t* __restrict__ block_size, const MdSeg* __restrict__ segs,
                                 uint32_t n_segs, MdKeep keep, const uint64_t* __restrict__ col_pos, uint32_t n_col,
                                 uint8_t* __restrict__ ref_base) {
  const uint32_t k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n_segs) return;
  const MdSeg sg = segs[k];
  const int64_t p0 = sg.first, p1 = sg.first + sg.count;
  // first column of this batch at or after the segment's first position (columns are sorted by position)
  uint32_t lo = 0, hi = n_col;
  while (lo < hi) {
    const uint32_t m = (lo + hi) >> 1;
    if ((int64_t)col_pos[m] < p0) lo = m + 1; else hi = m;
  }
  if (lo >= n_col) return;
  int64_t p = (int64_t)col_pos[lo];
  if (p >= p1) return;
  const Source src = locate(v, block_size, keep, sg.read);
  int64_t skip = sg.offset + (p - p0);
  uint32_t c = lo;
  if (src.body) {
    DnaWalk w;
    w.init(src.body, src.block_size);
    for (; skip > 0; --skip)
      if (w.next() < 0) return;
    // a segment lies inside one stretch of consecutive positions, so its columns are consecutive too
    for (; p < p1 && c < n_col && (int64_t)col_pos[c] == p; ++p, ++c) {
      const int ch = w.next();
      if (ch < 0) return;
      ref_base[c] = (uint8_t)ch;
    }
  } else if (src.str) {
    for (; p < p1 && c < n_col && (int64_t)col_pos[c] == p && skip < (int64_t)src.str_len; ++p, ++c, ++skip)
      ref_base[c] = src.str[skip];
  }
}

__global__ void md_keep_kernel(ReadsView v, const int32_t* __restrict__ block_size, MdKeep old_keep, MdKeep new_keep) {
  const uint32_t k = threadIdx.x;
  if (k >= 2 || new_keep.id[k] == ~0ull) return;
  uint8_t* out = const_cast<uint8_t*>(new_keep.data[k]);
  const uint32_t cap = new_keep.len[k];
  const Source src = locate(v, block_size, old_keep, new_keep.id[k]);
  uint32_t n = 0;
  if (src.body) {
    DnaWalk w;
    w.init(src.body, src.block_size);
    for (int ch; n < cap && (ch = w.next()) >= 0; ++n) out[n] = (uint8_t)ch;
  } else if (src.str) {
    for (; n < cap && n < src.str_len; ++n) out[n] = src.str[n];
  }
  for (; n < cap; ++n) out[n] = 'N';        // (unreachable with a consistent chain: the length was computed from the same walk)
}

}  // namespace

void md_dna_lengths(const ReadsView& v, const int32_t* block_size, const int32_t* eend, uint32_t a0, uint32_t g1,
                    int32_t* dna_len, cudaStream_t st) {
  if (g1 <= a0) return;
  const uint32_t n = g1 - a0;
  md_len_kernel<<<(n + 127) / 128, 128, 0, st>>>(v, block_size, eend, a0, g1, dna_len);
  ++g_kernel_launches;
}

void md_replay(const ReadsView& v, const int32_t* block_size, const MdSeg* segs, uint32_t n_segs, const MdKeep& keep,
               const uint64_t* col_pos, uint32_t n_col, uint8_t* ref_base, cudaStream_t st) {
  if (n_segs == 0 || n_col == 0) return;
  md_replay_kernel<<<(n_segs + 127) / 128, 128, 0, st>>>(v, block_size, segs, n_segs, keep, col_pos, n_col, ref_base);
  ++g_kernel_launches;
}

void md_keep(const ReadsView& v, const int32_t* block_size, const MdKeep& old_keep, const MdKeep& new_keep, cudaStream_t st) {
  if (new_keep.id[0] == ~0ull && new_keep.id[1] == ~0ull) return;
  md_keep_kernel<<<1, 32, 0, st>>>(v, block_size, old_keep, new_keep);
  ++g_kernel_launches;
}

}  // namespace biodb
