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
#include <string>
#include <vector>

#include "bai_build.h"
#include "deflate_enc.h"
#include "runtime.h"
#include "scan.cuh"

namespace biodb {

namespace {

constexpr uint32_t BGZF_CHUNK = 0xFF00;       // BGZF_BLOCK_SIZE (bgzf/constants.d:61)
constexpr uint32_t SLOT = 65536;              // BGZF_MAX_BLOCK_SIZE (:60)
constexpr uint32_t SLAB_BLOCKS = 4096;        // blocks per round trip to the device (256 MiB of slots)
constexpr size_t BLOCK_SCRATCH = (DEFL_HASH_SIZE * 2 + sizeof(DeflWork) + 255) & ~(size_t)255;   // per block, in global memory

__global__ void __launch_bounds__(64) deflate_blocks_kernel(const uint8_t* __restrict__ in, uint64_t in_len, uint32_t n_blocks,
                                                            uint8_t* __restrict__ slots, uint64_t* __restrict__ in_off,
                                                            uint32_t* __restrict__ isize, uint32_t* __restrict__ total,
                                                            uint16_t* __restrict__ htabs, int level) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  // scratch of the block (global memory, L2-resident): the hash table, then the encoder's work area
  uint16_t* htab = (uint16_t*)((uint8_t*)htabs + (size_t)b * BLOCK_SCRATCH);
  DeflWork* work = (DeflWork*)(htab + DEFL_HASH_SIZE);
  const uint64_t off = (uint64_t)b * BGZF_CHUNK;
  const uint32_t n = (uint32_t)(in_len - off < BGZF_CHUNK ? in_len - off : BGZF_CHUNK);
  const uint32_t len = deflate_block(in + off, n, slots + (size_t)b * SLOT + 18, SLOT - 26, htab, level, work);
  in_off[b] = off;
  isize[b] = n;
  total[b] = len + 26;                        // header 18 + payload + footer 8 (compress.d:88)
}

// the same for chunks of the caller's choosing (BamWriter ends a block where a record would not fit any more):
// chunk b is in[chunk_off[b] - chunk_off[0], chunk_off[b + 1] - chunk_off[0])
__global__ void __launch_bounds__(64) deflate_chunks_kernel(const uint8_t* __restrict__ in, const uint64_t* __restrict__ chunk_off,
                                                            uint32_t n_blocks, uint8_t* __restrict__ slots,
                                                            uint64_t* __restrict__ in_off, uint32_t* __restrict__ isize,
                                                            uint32_t* __restrict__ total, uint16_t* __restrict__ htabs, int level) {
  const uint32_t b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= n_blocks) return;
  uint16_t* htab = (uint16_t*)((uint8_t*)htabs + (size_t)b * BLOCK_SCRATCH);
  DeflWork* work = (DeflWork*)(htab + DEFL_HASH_SIZE);
  const uint64_t off = chunk_off[b] - chunk_off[0];
  const uint32_t n = (uint32_t)(chunk_off[b + 1] - chunk_off[b]);
  const uint32_t len = deflate_block(in + off, n, slots + (size_t)b * SLOT + 18, SLOT - 26, htab, level, work);
  in_off[b] = off;
  isize[b] = n;
  total[b] = len + 26;
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
  DeflWork work;
  return (int64_t)deflate_block(in, n, out, cap, htab, level, &work);
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
                d_htab.ensure((size_t)nb * BLOCK_SCRATCH, st) == cudaSuccess;
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

// ---- BamWriter (bio/std/hts/bam/writer.d:67-300) over the device compressor ---------------------------------------------

namespace {

// Chunks [chunk_off[b], chunk_off[b+1]) of `data` (each 1 .. 0xFF00 bytes) -> BGZF blocks appended to *out.
biodb_status compress_chunks(int32_t device, const uint8_t* data, const uint64_t* chunk_off, size_t n_all, int32_t level,
                             std::vector<uint8_t>* out) {
  if (n_all == 0) return BIODB_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BIODB_ERR_CUDA;           // no CPU fallback
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return BIODB_ERR_CUDA;
  cudaStream_t st = nullptr;
  if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return BIODB_ERR_CUDA;
  biodb_status rc = BIODB_OK;
  {
    DevBuf d_in, d_coff, d_slots, d_off, d_isize, d_total, d_crc, d_ooff, d_tmp, d_out, d_htab;
    for (size_t b0 = 0; b0 < n_all && rc == BIODB_OK; b0 += SLAB_BLOCKS) {
      const uint32_t nb = (uint32_t)std::min<size_t>(SLAB_BLOCKS, n_all - b0);
      const uint64_t in0 = chunk_off[b0], in_len = chunk_off[b0 + nb] - in0;
      uint64_t tot = 0;
      bool ok = d_in.ensure((size_t)in_len + 64, st) == cudaSuccess && d_coff.ensure((size_t)(nb + 1) * 8, st) == cudaSuccess &&
                d_slots.ensure((size_t)nb * SLOT, st) == cudaSuccess && d_off.ensure((size_t)nb * 8, st) == cudaSuccess &&
                d_isize.ensure((size_t)nb * 4, st) == cudaSuccess && d_total.ensure((size_t)(nb + 1) * 4, st) == cudaSuccess &&
                d_crc.ensure((size_t)nb * 4, st) == cudaSuccess && d_ooff.ensure((size_t)(nb + 1) * 8, st) == cudaSuccess &&
                d_tmp.ensure((scan_temp_elems(nb + 1) + 8) * 8, st) == cudaSuccess &&
                d_htab.ensure((size_t)nb * BLOCK_SCRATCH, st) == cudaSuccess;
      ok = ok && cudaMemcpyAsync(d_in.p, data + in0, (size_t)in_len, cudaMemcpyHostToDevice, st) == cudaSuccess &&
           cudaMemcpyAsync(d_coff.p, chunk_off + b0, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, st) == cudaSuccess;
      if (ok) {
        deflate_chunks_kernel<<<(nb + 63) / 64, 64, 0, st>>>(d_in.as<uint8_t>(), d_coff.as<uint64_t>(), nb, d_slots.as<uint8_t>(),
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
      ok = ok && d_out.ensure((size_t)tot + 64, st) == cudaSuccess;
      if (ok) {
        bgzf_pack_kernel<<<nb, 128, 0, st>>>(d_slots.as<uint8_t>(), d_total.as<uint32_t>(), d_ooff.as<uint64_t>(),
                                             d_crc.as<uint32_t>(), d_isize.as<uint32_t>(), d_out.as<uint8_t>());
        ++g_kernel_launches;
        const size_t at = out->size();
        out->resize(at + (size_t)tot);
        ok = cudaMemcpyAsync(out->data() + at, d_out.p, (size_t)tot, cudaMemcpyDeviceToHost, st) == cudaSuccess &&
             cudaStreamSynchronize(st) == cudaSuccess;
      }
      if (!ok) rc = BIODB_ERR_CUDA;
    }
    cudaStreamSynchronize(st);
  }
  cudaStreamDestroy(st);
  return rc;
}

inline uint32_t rd32(const uint8_t* p) { return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24); }

// reg2bin (bam/bai/bin.d:82-92)
inline uint16_t reg2bin(int32_t beg, int32_t end) {
  if (end == beg) end = beg + 1;
  --end;
  if (beg >> 14 == end >> 14) return (uint16_t)(((1 << 15) - 1) / 7 + (beg >> 14));
  if (beg >> 17 == end >> 17) return (uint16_t)(((1 << 12) - 1) / 7 + (beg >> 17));
  if (beg >> 20 == end >> 20) return (uint16_t)(((1 << 9) - 1) / 7 + (beg >> 20));
  if (beg >> 23 == end >> 23) return (uint16_t)(((1 << 6) - 1) / 7 + (beg >> 23));
  if (beg >> 26 == end >> 26) return (uint16_t)(((1 << 3) - 1) / 7 + (beg >> 26));
  return 0;
}

}  // namespace

struct biodb_writer {
  int32_t device = -1, level = -1;
  // BgzfOutputStream (bgzf/outputstream.d:50-223): the bytes written so far and where its blocks begin
  std::vector<uint8_t> bytes;
  std::vector<uint64_t> cuts{0};       // starts of the blocks that are complete; the current block starts at cuts.back()
  size_t stream_cur = 0;               // _current_size of the stream: bytes in the current block
  // BamWriter (bam/writer.d)
  size_t rec_cur = 0;                  // _current_size of the writer: record bytes it believes the current block holds
  int32_t n_refs = 0;
  bool header_done = false;
  std::vector<uint8_t> out;            // the finished file
  std::vector<uint8_t> index;          // its BAI index (biodb_writer_index)
  std::string err;
  // what the index needs to know of every record written (writer.d:150-195 parses it back out of the blocks)
  struct Rec { uint64_t at; uint32_t size; int32_t ref_id, pos, end_pos; uint32_t bin; bool unmapped; };
  std::vector<Rec> recs;

  void flush_current_block() {         // outputstream.d:136-161
    if (stream_cur == 0) return;
    cuts.push_back(bytes.size());
    stream_cur = 0;
  }
  void write(const uint8_t* p, size_t size) {     // writeBlock, outputstream.d:107-132
    if (size + stream_cur >= BGZF_CHUNK) {
      while (size + stream_cur >= BGZF_CHUNK) {
        const size_t room = BGZF_CHUNK - stream_cur;
        bytes.insert(bytes.end(), p, p + room);
        p += room;
        size -= room;
        stream_cur = BGZF_CHUNK;
        flush_current_block();
      }
      bytes.insert(bytes.end(), p, p + size);
      stream_cur = size;
    } else {
      bytes.insert(bytes.end(), p, p + size);
      stream_cur += size;
    }
  }
  void write_i32(int32_t v) {
    const uint8_t b[4] = {(uint8_t)v, (uint8_t)(v >> 8), (uint8_t)(v >> 16), (uint8_t)(v >> 24)};
    write(b, 4);
  }
};

extern "C" {

biodb_status biodb_writer_begin(int32_t device, int32_t level, biodb_writer** out) {
  if (!out || level < -1 || level > 9) return BIODB_ERR_ARG;
  biodb_writer* w = new biodb_writer;
  w->device = device;
  w->level = level;
  w->write((const uint8_t*)"BAM\1", 4);                                  // writer.d:90
  *out = w;
  return BIODB_OK;
}
void biodb_writer_end(biodb_writer* w) { delete w; }
const char* biodb_writer_error(const biodb_writer* w) { return w ? w->err.c_str() : ""; }

// writeSamHeader + writeReferenceSequenceInfo (writer.d:139-181): l_text, text, n_ref, (l_name, name, NUL, l_ref)*, then
// the current block is flushed so that the records start a block of their own
biodb_status biodb_writer_header(biodb_writer* w, const char* text, size_t text_len, int32_t n_refs, const char* const* names,
                                 const int32_t* lengths) {
  if (!w || (!text && text_len) || n_refs < 0 || (n_refs && (!names || !lengths)) || w->header_done) return BIODB_ERR_ARG;
  w->write_i32((int32_t)text_len);
  w->write((const uint8_t*)text, text_len);
  w->write_i32(n_refs);
  for (int32_t i = 0; i < n_refs; ++i) {
    const size_t ln = strlen(names[i]);
    w->write_i32((int32_t)(ln + 1));
    w->write((const uint8_t*)names[i], ln);
    const uint8_t nul = 0;
    w->write(&nul, 1);
    w->write_i32(lengths[i]);
  }
  w->n_refs = n_refs;
  w->header_done = true;
  w->flush_current_block();
  return BIODB_OK;
}

// writeRecord (writer.d:244-268) for every record of `records` (block_size prefix + body, back to back): the bin is
// recalculated (read.d:1028-1030), a record that would not fit into the current block starts a new one.
biodb_status biodb_writer_records(biodb_writer* w, const uint8_t* records, size_t len) {
  if (!w || (!records && len)) return BIODB_ERR_ARG;
  size_t p = 0;
  std::vector<uint8_t> rec;
  while (p < len) {
    if (len - p < 4) { w->err = "truncated record"; return BIODB_ERR_TRUNCATED; }
    const int32_t bs = (int32_t)rd32(records + p);
    if (bs < 32 || (size_t)bs > len - p - 4) { w->err = "truncated record"; return BIODB_ERR_TRUNCATED; }
    rec.assign(records + p, records + p + 4 + (size_t)bs);
    uint8_t* r = rec.data() + 4;
    const int32_t ref_id = (int32_t)rd32(r), pos = (int32_t)rd32(r + 4);
    if (!(ref_id == -1 || (ref_id >= 0 && ref_id < w->n_refs))) {
      w->err = "Read reference ID is out of range";
      return BIODB_ERR_ARG;
    }
    const uint32_t lname = r[8], nc = (uint32_t)r[12] | ((uint32_t)r[13] << 8), flag = (uint32_t)r[14] | ((uint32_t)r[15] << 8);
    if (32ull + lname + 4ull * nc > (uint64_t)bs || lname == 0) { w->err = "malformed record"; return BIODB_ERR_TRUNCATED; }
    uint32_t covered = 0;                                                  // basesCovered (read.d:255-262)
    if (!(flag & 4))
      for (uint32_t k = 0; k < nc; ++k) {
        const uint32_t raw = rd32(r + 32 + lname + 4 * k);
        if ((0x3C1A7u >> ((raw & 0xF) * 2)) & 2) covered += raw >> 4;
      }
    const uint16_t bin = reg2bin(pos, (int32_t)((uint32_t)pos + covered));
    r[10] = (uint8_t)bin;
    r[11] = (uint8_t)(bin >> 8);
    r[32 + lname - 1] = 0;                                                 // read.d:616-617 (the byte doubles as a flag in memory)
    const size_t read_size = rec.size();                                   // size_in_bytes (read.d:609-611)
    if (read_size + w->rec_cur > BGZF_CHUNK) {
      w->flush_current_block();
      w->recs.push_back(biodb_writer::Rec{w->bytes.size(), (uint32_t)read_size, ref_id, pos, (int32_t)((uint32_t)pos + covered), bin, (flag & 4) != 0});
      w->write(rec.data(), rec.size());
      w->rec_cur = read_size;
    } else {
      w->recs.push_back(biodb_writer::Rec{w->bytes.size(), (uint32_t)read_size, ref_id, pos, (int32_t)((uint32_t)pos + covered), bin, (flag & 4) != 0});
      w->write(rec.data(), rec.size());
      w->rec_cur += read_size;
    }
    p += 4 + (size_t)bs;
  }
  return BIODB_OK;
}

biodb_status biodb_writer_flush(biodb_writer* w) {                         // writer.d:271-273 (ends the current block)
  if (!w) return BIODB_ERR_ARG;
  w->flush_current_block();
  return BIODB_OK;
}

// Host-only view of what has been written: the uncompressed bytes and the starts of the BGZF blocks they will become
// (n_cuts entries; a last block still open runs to *len).  For the CPU tests of the block layout.
biodb_status biodb_writer_layout(const biodb_writer* w, const uint8_t** data, size_t* len, const uint64_t** cuts, size_t* n_cuts) {
  if (!w || !data || !len || !cuts || !n_cuts) return BIODB_ERR_ARG;
  *data = w->bytes.data();
  *len = w->bytes.size();
  *cuts = w->cuts.data();
  *n_cuts = w->cuts.size();
  return BIODB_OK;
}

// finish (writer.d:276-280): every block compressed on the device, the EOF block appended; *data stays valid until
// biodb_writer_end.
biodb_status biodb_writer_finish(biodb_writer* w, const uint8_t** data, size_t* len) {
  static const uint8_t EOF_BLOCK[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (!w || !data || !len) return BIODB_ERR_ARG;
  w->flush_current_block();
  w->out.clear();
  const biodb_status rc = compress_chunks(w->device, w->bytes.data(), w->cuts.data(), w->cuts.size() - 1, w->level, &w->out);
  if (rc != BIODB_OK) return rc;
  w->out.insert(w->out.end(), EOF_BLOCK, EOF_BLOCK + 28);
  *data = w->out.data();
  *len = w->out.size();
  return BIODB_OK;
}

// Host-only test hook: take `data` for the finished file (a BGZF stream with exactly the writer's block layout, e.g.
// compressed by zlib), so that biodb_writer_index can be checked without a device.
biodb_status biodb_writer_debug_set_output(biodb_writer* w, const uint8_t* data, size_t len) {
  if (!w || (!data && len)) return BIODB_ERR_ARG;
  w->flush_current_block();
  w->out.assign(data, data + len);
  return BIODB_OK;
}

// The BAI index of the finished file — what BamWriter builds while writing coordinate-sorted output (writer.d:139-195,
// 171-175: IndexBuilder with check_bins, fed with every record and the virtual offsets it got in the file).  Call after
// biodb_writer_finish.  BIODB_ERR_UNSORTED if the records were not in coordinate order.
biodb_status biodb_writer_index(biodb_writer* w, const uint8_t** data, size_t* len) {
  if (!w || !data || !len || w->out.empty()) return BIODB_ERR_ARG;
  // where the blocks of the layout begin in the file: the BSIZE chain of the finished stream
  const size_t nb = w->cuts.size() - 1;
  std::vector<uint64_t> cb;
  size_t p = 0;
  while (p + 18 <= w->out.size() && cb.size() <= nb) {
    cb.push_back(p);
    p += ((size_t)w->out[p + 16] | ((size_t)w->out[p + 17] << 8)) + 1;
  }
  if (cb.size() != nb + 1) { w->err = "the finished file does not have the writer's block layout"; return BIODB_ERR_FORMAT; }
  auto voffset = [&](uint64_t x) -> uint64_t {                    // of byte x of the uncompressed stream
    const size_t i = (size_t)(std::upper_bound(w->cuts.begin(), w->cuts.end(), x) - w->cuts.begin()) - 1;   // cuts[i] <= x
    return (cb[i] << 16) | (x - w->cuts[i]);                       // (x == the end of the stream: the EOF block, offset 0)
  };
  BaiBuilder b;
  b.begin(w->n_refs, true);
  for (const biodb_writer::Rec& r : w->recs) {
    if (!b.put(r.ref_id, r.pos, r.end_pos, r.bin, r.unmapped, voffset(r.at), voffset(r.at + r.size))) {
      w->err = b.err;
      return b.err.rfind("BAM file is not", 0) == 0 ? BIODB_ERR_UNSORTED : BIODB_ERR_FORMAT;
    }
  }
  b.finish();
  w->index.swap(b.out);
  *data = w->index.data();
  *len = w->index.size();
  return BIODB_OK;
}

}  // extern "C"
