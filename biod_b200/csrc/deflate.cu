// BGZF compression on the device (first part): bgzfCompress (bio/core/bgzf/compress.d:43-103)
// and the block cutting of BgzfOutputStream (bgzf/outputstream.d:50-223: a new block every BGZF_BLOCK_SIZE = 0xFF00
// bytes, the 28-byte EOF block at close) for a buffer that is complete when the call is made.
//
//   deflate_blocks_kernel  one thread per BGZF block: raw DEFLATE of its chunk (deflate_enc.h) into a 64 KiB slot
//   crc32 (crc32.cu)       CRC-32 of every chunk, for the footer
//   bgzf_pack_kernel       one CTA per block: header (BSIZE), payload, footer (CRC32, ISIZE) packed back to back at
//                          the offsets an exclusive scan of the block sizes gives
// The compressed bytes differ from zlib's (the reference only asks that they come back: outputstream.d:225-247); they
// are valid DEFLATE, which the tests check with zlib and with this library's own inflate kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>

#include "deflate_enc.h"
#include "runtime.h"
#include "scan.cuh"

namespace biodb {

namespace {

constexpr uint32_t BGZF_CHUNK = 0xFF00;       // BGZF_BLOCK_SIZE (bgzf/constants.d:61)
constexpr uint32_t SLOT = 65536;              // BGZF_MAX_BLOCK_SIZE (:60)
constexpr uint32_t SLAB_BLOCKS = 4096;        // blocks per round trip to the device (256 MiB of slots)

__global__ void __launch_bounds__(64) deflate_blocks_kernel(const uint8_t* __restrict__ in, uint64_t in_len, uint32_t n_blocks,
                                                            uint8_t* __restrict__ slots, uint64_t* __restrict__ in_off,
                                                            uint32_t* __restrict__ isize, uint32_t* __restrict__ total,
                                                            uint16_t* __restrict__ htabs, int level) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint16_t* htab = htabs + (size_t)b * DEFL_HASH_SIZE;      // 8 KiB of scratch per block (global: L2-resident)
  const uint64_t off = (uint64_t)b * BGZF_CHUNK;
  const uint32_t n = (uint32_t)(in_len - off < BGZF_CHUNK ? in_len - off : BGZF_CHUNK);
  const uint32_t len = deflate_block(in + off, n, slots + (size_t)b * SLOT + 18, SLOT - 26, htab, level);
  in_off[b] = off;
  isize[b] = n;
  total[b] = len + 26;                        // header 18 + payload + footer 8 (compress.d:88)
}

__global__ void __launch_bounds__(128) bgzf_pack_kernel(const uint8_t* __restrict__ slots, const uint32_t* __restrict__ total,
                                                        const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ crc,
                                                        const uint32_t* __restrict__ isize, uint8_t* __restrict__ out) {
  const uint32_t b = blockIdx.x;
  const uint32_t t = total[b];
  uint8_t* dst = out + out_off[b];
  const uint8_t* src = slots + (size_t)b * SLOT;
  if (threadIdx.x == 0) {
    const uint8_t head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};     // BLOCK_HEADER_START (constants.d:30-38)
    for (int k = 0; k < 16; ++k) dst[k] = head[k];
    dst[16] = (uint8_t)(t - 1);                                                            // BSIZE = block length - 1
    dst[17] = (uint8_t)((t - 1) >> 8);
    const uint32_t c = crc[b], n = isize[b];
    for (int k = 0; k < 4; ++k) { dst[t - 8 + k] = (uint8_t)(c >> (8 * k)); dst[t - 4 + k] = (uint8_t)(n >> (8 * k)); }
  }
  for (uint32_t i = 18 + threadIdx.x; i < t - 8; i += blockDim.x) dst[i] = src[i];
}

}  // namespace

}  // namespace biodb

using namespace biodb;

extern "C" {

size_t biodb_bgzf_compress_bound(size_t len) {
  const size_t nb = (len + BGZF_CHUNK - 1) / BGZF_CHUNK;
  return nb * (size_t)SLOT + 28;
}

// Host-only: the encoder of deflate_enc.h compiled for the CPU, for the tests (raw DEFLATE of one chunk).
int64_t biodb_debug_deflate_block(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, int32_t level) {
  if ((!in && n) || !out) return -1;
  uint16_t htab[DEFL_HASH_SIZE];
  return (int64_t)deflate_block(in, n, out, cap, htab, level);
}

biodb_status biodb_bgzf_compress(int32_t device, const void* data, size_t len, int32_t level, int32_t add_eof, void* out,
                                 size_t cap, size_t* out_len) {
  static const uint8_t EOF_BLOCK[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if ((!data && len) || !out || !out_len || level < -1 || level > 9) return BIODB_ERR_ARG;   // compress.d:46-48
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BIODB_ERR_CUDA;           // no CPU fallback
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return BIODB_ERR_CUDA;
  cudaStream_t st = nullptr;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return BIODB_ERR_CUDA;
  biodb_status rc = BIODB_OK;
  size_t written = 0;
  {
    DevBuf d_in, d_slots, d_off, d_isize, d_total, d_crc, d_ooff, d_tmp, d_out, d_htab;
    const uint8_t* src = (const uint8_t*)data;
    const size_t n_all = (len + BGZF_CHUNK - 1) / BGZF_CHUNK;
    for (size_t b0 = 0; b0 < n_all && rc == BIODB_OK; b0 += SLAB_BLOCKS) {
      const uint32_t nb = (uint32_t)std::min<size_t>(SLAB_BLOCKS, n_all - b0);
      const size_t in0 = b0 * BGZF_CHUNK, in_len = std::min<size_t>((size_t)nb * BGZF_CHUNK, len - in0);
      uint64_t tot = 0;
      bool ok = d_in.ensure(in_len + 64, st) == cudaSuccess && d_slots.ensure((size_t)nb * SLOT, st) == cudaSuccess &&
                d_off.ensure((size_t)nb * 8, st) == cudaSuccess && d_isize.ensure((size_t)nb * 4, st) == cudaSuccess &&
                d_total.ensure((size_t)(nb + 1) * 4, st) == cudaSuccess && d_crc.ensure((size_t)nb * 4, st) == cudaSuccess &&
                d_ooff.ensure((size_t)(nb + 1) * 8, st) == cudaSuccess &&
                d_tmp.ensure((scan_temp_elems(nb + 1) + 8) * 8, st) == cudaSuccess &&
                d_htab.ensure((size_t)nb * DEFL_HASH_SIZE * 2, st) == cudaSuccess;
      ok = ok && cudaMemcpyAsync(d_in.p, src + in0, in_len, cudaMemcpyHostToDevice, st) == cudaSuccess;
      if (ok) {
        deflate_blocks_kernel<<<(nb + 63) / 64, 64, 0, st>>>(d_in.as<uint8_t>(), in_len, nb, d_slots.as<uint8_t>(),
                                                              d_off.as<uint64_t>(), d_isize.as<uint32_t>(),
                                                              d_total.as<uint32_t>(), d_htab.as<uint16_t>(), level);
        ++g_kernel_launches;
        ok = cudaMemsetAsync(d_total.as<uint32_t>() + nb, 0, 4, st) == cudaSuccess &&
             launch_crc32(d_in.as<uint8_t>(), d_off.as<uint64_t>(), d_isize.as<uint32_t>(), nb, d_crc.as<uint32_t>(), st) == cudaSuccess;
      }
      if (ok) {
        device_scan<false>(d_total.as<uint32_t>(), d_ooff.as<uint64_t>(), (uint64_t)nb + 1, d_tmp.as<uint64_t>(), OpAdd(),
                           (uint64_t)0, st);
        ok = cudaMemcpyAsync(&tot, d_ooff.as<uint64_t>() + nb, 8, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
      }
      if (ok && written + tot + (add_eof ? 28 : 0) > cap) { rc = BIODB_ERR_NOMEM; break; }
      ok = ok && d_out.ensure((size_t)tot + 64, st) == cudaSuccess;
      if (ok) {
        bgzf_pack_kernel<<<nb, 128, 0, st>>>(d_slots.as<uint8_t>(), d_total.as<uint32_t>(), d_ooff.as<uint64_t>(),
                                             d_crc.as<uint32_t>(), d_isize.as<uint32_t>(), d_out.as<uint8_t>());
        ++g_kernel_launches;
        ok = cudaMemcpyAsync((uint8_t*)out + written, d_out.p, (size_t)tot, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
      }
      if (!ok) { rc = BIODB_ERR_CUDA; break; }
      written += (size_t)tot;
    }
    cudaStreamSynchronize(st);
  }
  cudaStreamDestroy(st);
  if (rc != BIODB_OK) return rc;
  if (add_eof) {                                                     // BgzfOutputStream.close -> addEofBlock (outputstream.d:218-221)
    if (written + 28 > cap) return BIODB_ERR_NOMEM;
    memcpy((uint8_t*)out + written, EOF_BLOCK, 28);
    written += 28;
  }
  *out_len = written;
  return BIODB_OK;
}

}  // extern "C"
