// MAQ genotype likelihoods per pileup column on the device.
//
// One thread per column restates ErrorModelCoefficients.computeLikelihoods (bio/std/hts/snpcallers/maq.d:138-248) over
// the column's valid read bases, then the first two entries of GenotypeLikelihoodInfo's score order (:258-277) — what
// makeCall needs (:460-486).  Floating point follows the reference's types: fsum / bsum are doubles, the per-genotype
// sums are floats fed from doubles; the reference's `real` (x87) constant C = 10 / ln 10 is a double here, so a score can
// differ from the reference's in its last float bit (the tests allow 2e-6 relative).  Two choices the reference leaves
// open are pinned: bases of equal quality keep their column order (the reference sorts with Phobos'
// unstable sort, maq.d:151); of more than 255 valid bases the first 255 are used (the reference draws a random sample
// with an unpredictable seed, :142-147).
#include "maq.h"

#include "kernels.h"

namespace biodb {

namespace {

constexpr int MAQ_MAX = 255;

// nucleotide index of a base character: A C G T -> 0..3, anything else 4 (never a key of the sums, maq.d:90,179)
__device__ __forceinline__ uint32_t nuc_index(uint32_t ch) {
  ch &= 0x5f;                      // upper case
  return ch == 'A' ? 0u : ch == 'C' ? 1u : ch == 'G' ? 2u : ch == 'T' ? 3u : 4u;
}

__global__ void __launch_bounds__(128) maq_kernel(const uint64_t* __restrict__ col_off, const uint8_t* __restrict__ base_s,
                                                  const uint8_t* __restrict__ qual_m, uint32_t n_col, MaqDevTables t,
                                                  MaqColumns out) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  const uint64_t a = col_off[c], b = col_off[c + 1];
  uint16_t key[MAQ_MAX];           // quality << 8 | index among the valid bases: sorting the keys is the stable sort by quality
  uint8_t bw[MAQ_MAX];             // nucleotide index * 2 + strand
  uint32_t n = 0, nv = 0;
  for (uint64_t e = a; e < b; ++e) {
    const uint32_t bs = base_s[e];
    if (bs == 0xFF) continue;      // below minimum_base_quality, or '-' (maq.d:401-404)
    ++nv;
    if (n < (uint32_t)MAQ_MAX) {
      key[n] = (uint16_t)(((uint32_t)qual_m[e] << 8) | n);
      bw[n] = (uint8_t)(nuc_index(bs & 0x7f) * 2 + (bs >> 7));
      ++n;
    }
  }
  out.n_valid[c] = (uint16_t)(nv > 0xffff ? 0xffff : nv);
  if (n == 0) {
    out.gt0[c] = 255; out.gt1[c] = 255; out.s0[c] = 0.f; out.s1[c] = 0.f;
    return;
  }
  for (uint32_t i = 1; i < n; ++i) {                 // insertion sort, ascending
    const uint16_t k = key[i];
    uint32_t j = i;
    while (j > 0 && key[j - 1] > k) { key[j] = key[j - 1]; --j; }
    key[j] = k;
  }
  // maq.d:158-171, from the highest quality down.  w: bases of the same nucleotide AND strand seen so far (8 byte
  // counters in one word), cn: bases of the same nucleotide.
  unsigned long long w = 0;
  uint32_t cn = 0;
  double fsum0 = 0, fsum1 = 0, fsum2 = 0, fsum3 = 0, bsum0 = 0, bsum1 = 0, bsum2 = 0, bsum3 = 0;
  for (uint32_t r = n; r-- > 0;) {
    const uint32_t k = key[r];
    uint32_t q = k >> 8;
    q = q < 4 ? 4 : q > 63 ? 63 : q;
    const uint32_t slot = bw[k & 0xff], bi = slot >> 1;
    if (bi >= 4) continue;
    const uint32_t wv = (uint32_t)(w >> (8 * slot)) & 0xff, cv = (cn >> (8 * bi)) & 0xff;
    const double f = t.fk[wv];
    const double fb = f * t.beta[(size_t)q << 16 | (size_t)n << 8 | cv];
    if (bi == 0) { fsum0 += f; bsum0 += fb; }
    if (bi == 1) { fsum1 += f; bsum1 += fb; }
    if (bi == 2) { fsum2 += f; bsum2 += fb; }
    if (bi == 3) { fsum3 += f; bsum3 += fb; }
    w += 1ull << (8 * slot);
    cn += 1u << (8 * bi);
  }
  const double bsum[4] = {bsum0, bsum1, bsum2, bsum3};
  const uint32_t cnt[4] = {cn & 0xff, (cn >> 8) & 0xff, (cn >> 16) & 0xff, cn >> 24};
  (void)fsum0; (void)fsum1; (void)fsum2; (void)fsum3;      // (tmp3 of the reference sums fsum and never uses it)
  const double C = 4.342944819032518;                       // 10 / ln 10
  // scores in genotype-code order (first * 5 + second; homozygotes i|i, heterozygotes j|i with j > i): the order
  // GenotypeLikelihoodInfo walks them in; the best two under "a later one replaces only if strictly smaller"
  float best = 0.f, second = 0.f;
  uint32_t gbest = 255, gsecond = 255;
  auto offer = [&](uint32_t g, float s) {
    if (s < 0.f) s = 0.f;                                   // maq.d:236-241
    if (gbest == 255 || s < best) { second = best; gsecond = gbest; best = s; gbest = g; }
    else if (gsecond == 255 || s < second) { second = s; gsecond = g; }
  };
#pragma unroll
  for (int first = 0; first < 4; ++first) {
    // heterozygotes first|i for i < first come before the homozygote first|first in code order
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (i >= first) continue;
      float tmp1 = 0.f;
      int tmp2 = 0;
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k != i && k != first) { tmp1 = (float)((double)tmp1 + bsum[k]); tmp2 += (int)cnt[k]; }
      const double lh = t.lhet[(size_t)(cnt[i] + cnt[first]) << 8 | cnt[first]];
      const float s = tmp2 > 0 ? (float)((double)tmp1 - C * lh) : (float)(-C * lh);
      offer((uint32_t)(first * 5 + i), s);
    }
    float tmp1 = 0.f;
    int tmp2 = 0;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (k != first) { tmp1 = (float)((double)tmp1 + bsum[k]); tmp2 += (int)cnt[k]; }
    offer((uint32_t)(first * 6), tmp2 > 0 ? tmp1 : 0.f);
  }
  out.gt0[c] = (uint8_t)gbest;
  out.gt1[c] = (uint8_t)gsecond;
  out.s0[c] = best;
  out.s1[c] = second;
}

// Base5 code of a reference base character (bio/core/base.d:163-182)
__device__ __forceinline__ uint32_t base5_code(uint32_t ch) {
  const uint32_t k = nuc_index(ch);
  return k;        // 4 = N
}

__global__ void maq_flag_kernel(MaqColumns m, const uint8_t* __restrict__ ref_base, uint32_t n_col, float min_q, uint32_t* flag) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col) return;
  const uint32_t g = m.gt0[c];
  const uint32_t r = base5_code(ref_base ? ref_base[c] : 'N');
  // makeCall: gts.count >= 2 (maq.d:466); findSNPs: is_variant && quality > minimum_call_quality (:519)
  flag[c] = (g != 255 && g != r * 6 && (m.s1[c] - m.s0[c]) > min_q) ? 1u : 0u;
}

__global__ void maq_gather_kernel(MaqColumns m, const uint8_t* __restrict__ ref_base, const uint64_t* __restrict__ col_pos,
                                  uint32_t n_col, const uint32_t* __restrict__ flag, const uint32_t* __restrict__ incl,
                                  uint32_t* call_col, uint64_t* call_pos, uint8_t* call_gt, uint8_t* call_ref, float* call_qual) {
  const uint32_t c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= n_col || !flag[c]) return;
  const uint32_t k = incl[c] - 1;
  call_col[k] = c;
  call_pos[k] = col_pos[c];
  call_gt[k] = m.gt0[c];
  call_ref[k] = ref_base ? ref_base[c] : (uint8_t)'N';
  call_qual[k] = m.s1[c] - m.s0[c];
}

}  // namespace

void maq_columns(const uint64_t* col_off, const uint8_t* base_s, const uint8_t* qual_m, uint32_t n_col, MaqDevTables t,
                 MaqColumns out, cudaStream_t st) {
  if (n_col == 0) return;
  maq_kernel<<<(n_col + 127) / 128, 128, 0, st>>>(col_off, base_s, qual_m, n_col, t, out);
  ++g_kernel_launches;
}

void maq_call_flags(const MaqColumns& m, const uint8_t* ref_base, uint32_t n_col, float min_q, uint32_t* flag, cudaStream_t st) {
  if (n_col == 0) return;
  maq_flag_kernel<<<(n_col + 255) / 256, 256, 0, st>>>(m, ref_base, n_col, min_q, flag);
  ++g_kernel_launches;
}

void maq_call_gather(const MaqColumns& m, const uint8_t* ref_base, const uint64_t* col_pos, uint32_t n_col, const uint32_t* flag,
                     const uint32_t* incl, uint32_t* call_col, uint64_t* call_pos, uint8_t* call_gt, uint8_t* call_ref,
                     float* call_qual, cudaStream_t st) {
  if (n_col == 0) return;
  maq_gather_kernel<<<(n_col + 255) / 256, 256, 0, st>>>(m, ref_base, col_pos, n_col, flag, incl, call_col, call_pos, call_gt,
                                                         call_ref, call_qual);
  ++g_kernel_launches;
}

}  // namespace biodb
