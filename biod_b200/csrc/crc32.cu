// CRC-32 (IEEE 802.3, the polynomial zlib and the BGZF footer use) of every inflated block, on the device.
//
// BioD asserts `block.crc32 == crc32(0, uncompressed)` only in debug builds (bio/core/bgzf/block.d:187; -release
// compiles it out, Makefile:33); options.verify_crc turns the same check on here.  One warp per block: each lane
// runs a byte-wise table CRC over its contiguous slice, then the 32 partial CRCs are merged with the identity
// crc(A||B) = crc(A) * x^(8|B|)  xor  crc(B)  in GF(2)[x]/P  (carry-less multiply by shift-and-xor).
#include <cuda_runtime.h>
#include <stdint.h>

#include "kernels.h"

namespace biodb {
namespace {

constexpr uint32_t POLY = 0xEDB88320u;   // reflected

// a * b mod P with reflected bit order (bit 31 = x^0)
__device__ __forceinline__ uint32_t gf_mul(uint32_t a, uint32_t b) {
  uint32_t r = 0;
#pragma unroll 4
  for (int i = 0; i < 32; ++i) {
    if (a & 0x80000000u) r ^= b;
    a <<= 1;
    b = (b >> 1) ^ ((b & 1) ? POLY : 0);
  }
  return r;
}
// x^(8n) mod P
__device__ uint32_t gf_pow_x8(uint32_t n) {
  uint32_t result = 0x80000000u;      // 1
  uint32_t base = 0x00800000u;        // x^8
  while (n) {
    if (n & 1) result = gf_mul(result, base);
    base = gf_mul(base, base);
    n >>= 1;
  }
  return result;
}

__global__ void __launch_bounds__(128) crc32_kernel(const uint8_t* __restrict__ out, const uint64_t* __restrict__ out_off,
                                                    const uint32_t* __restrict__ isize, uint32_t n_blocks,
                                                    uint32_t* __restrict__ crc) {
  __shared__ uint32_t table[256];
  for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) {
    uint32_t c = i;
    for (int k = 0; k < 8; ++k) c = (c >> 1) ^ ((c & 1) ? POLY : 0);
    table[i] = c;
  }
  __syncthreads();
  const uint32_t b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= n_blocks) return;
  const uint32_t n = isize[b];
  const uint8_t* p = out + out_off[b];
  const uint32_t per = (n + 31) / 32;
  const uint32_t lo = min(n, lane * per), hi = min(n, lo + per);
  // raw CRC register (no pre/post inversion) of the lane's slice starting from 0
  uint32_t c = 0;
  for (uint32_t i = lo; i < hi; ++i) c = table[(c ^ p[i]) & 0xFF] ^ (c >> 8);
  // shift it past the bytes that follow the slice, then fold
  c = gf_mul(c, gf_pow_x8(n - hi));
#pragma unroll
  for (int d = 16; d; d >>= 1) c ^= __shfl_xor_sync(0xffffffffu, c, d);
  // the standard CRC starts from 0xFFFFFFFF: that initial register contributes 0xFFFFFFFF * x^(8n); final inversion
  if (lane == 0) crc[b] = ~(c ^ gf_mul(0xFFFFFFFFu, gf_pow_x8(n)));
}

}  // namespace

cudaError_t launch_crc32(const uint8_t* out, const uint64_t* out_off, const uint32_t* isize, uint32_t n_blocks,
                         uint32_t* crc, cudaStream_t st) {
  if (n_blocks == 0) return cudaSuccess;
  crc32_kernel<<<(n_blocks + 3) / 4, 128, 0, st>>>(out, out_off, isize, n_blocks, crc);
  ++g_kernel_launches;
  return cudaGetLastError();
}

}  // namespace biodb
