// BGZF compression on the device (first part): bgzfCompress (bio/core/bgzf/compress.d:43-103)
// and the block cutting of BgzfOutputStream (bgzf/outputstream.d:50-223: a new block every BGZF_BLOCK_SIZE = 0xFF00
// bytes, the 28-byte EOF block at close) for a buffer that is complete when the call is made.
//
//   deflate_warp_kernel    one warp per BGZF block (a persistent grid draws block numbers from a counter): raw DEFLATE of
//                          its chunk into a 64 KiB slot —
//                            parse   the 32 lanes look at 32 consecutive positions at once: hash probe of a 2-way table in
//                                    shared memory, match lengths of up to three candidates, lazy step, the greedy chain
//                                    through the window by pointer doubling; tokens go to a scratch list in global memory,
//                                    their statistics to shared-memory counters (deflate_enc.h states the parse; its host
//                                    loop writes the same tokens);
//                            codes   symbols ranked by the whole warp, the two-queue Huffman construction and the
//                                    code-length header by lane 0 (a few hundred steps in shared memory);
//                            emit    32 tokens at a time: bit lengths scanned, codes OR-ed into a shared-memory staging
//                                    window, whole words stored to the slot
//   crc32 (crc32.cu)       CRC-32 of every chunk, for the footer
//   bgzf_pack_kernel       one CTA per block: header (BSIZE), payload, footer (CRC32, ISIZE) packed back to back at
//                          the offsets an exclusive scan of the block sizes gives
// Host side: slabs of blocks run through two sets of buffers and two streams, so that the copies of one slab (caller's
// memory -> pinned -> device, and back) overlap the kernels of the other.
// The compressed bytes differ from zlib's (the reference only asks that they come back: outputstream.d:225-247); they
// are valid DEFLATE, which the tests check with zlib and with this library's own inflate kernels.
#include <cuda_runtime.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "bai_build.h"
#include "deflate_enc.h"
#include "runtime.h"
#include "scan.cuh"

namespace biodb {

namespace {

constexpr uint32_t BGZF_CHUNK = 0xFF00;       // BGZF_BLOCK_SIZE (bgzf/constants.d:61)
constexpr uint32_t SLOT = 65536;              // BGZF_MAX_BLOCK_SIZE (:60)
constexpr uint32_t PAYLOAD_AT = 32;           // where the DEFLATE payload stands inside its slot (word aligned)
constexpr uint32_t SLAB_BLOCKS = 2048;        // blocks per round trip to the device (128 MiB of slots)
constexpr uint32_t TOK_STRIDE = 65536;        // tokens of one block (at most 65 280 literals + end of block)
constexpr uint32_t STAGE_WORDS = 192;         // header of a dynamic block: at most ~570 bytes
constexpr uint32_t FULL = 0xffffffffu;

struct EncSmem {
  union {
    uint16_t htab[DEFL_HASH_SIZE * DEFL_WAYS];          // while parsing
    struct {                                            // afterwards
      DeflWork work;
      uint32_t stage[STAGE_WORDS];
    } post;
  };
  uint32_t hist[288 + 32];                              // literal / length and distance code counts (32-bit for atomicAdd)
};

__device__ __forceinline__ uint32_t load32u(const uint8_t* a) {     // 4 bytes at any address, little endian
  const uintptr_t ua = (uintptr_t)a;
  const uint32_t* w = (const uint32_t*)(ua & ~(uintptr_t)3);
  return __funnelshift_r(__ldg(w), __ldg(w + 1), (uint32_t)(ua & 3) * 8);
}

// enters position p (hash h) of the lanes with `v` into the table the way the host's loop over rising positions does:
// a bucket keeps its two latest positions
__device__ __forceinline__ void table_insert(uint16_t* htab, bool v, uint32_t h, uint32_t p, uint32_t base, int lane) {
  const uint32_t same = __match_any_sync(FULL, v ? h : (0x10000u | (uint32_t)lane));
  if (v && (same >> lane) == 1u) {                                  // the highest lane of its group
    const uint32_t rest = same & ((1u << lane) - 1);
    uint32_t* bucket = reinterpret_cast<uint32_t*>(htab + h * 2);
    const uint32_t w1 = rest ? base + (31u - (uint32_t)__clz((int)rest)) : (*bucket & 0xffffu);
    *bucket = p | (w1 << 16);
  }
  __syncwarp();
}

// used symbols of freq[0, n) by rising (frequency, symbol) into order[]; keys: scratch of n words.  Returns their number.
__device__ uint32_t warp_sort_symbols(const uint16_t* freq, uint32_t n, uint16_t* order, uint32_t* keys, int lane) {
  uint32_t used = 0;
  for (uint32_t b0 = 0; b0 < n; b0 += 32) {
    const uint32_t k = b0 + lane;
    const bool u = k < n && freq[k] != 0;
    const uint32_t m = __ballot_sync(FULL, u);
    if (u) keys[used + __popc(m & ((1u << lane) - 1))] = ((uint32_t)freq[k] << 16) | k;
    used += __popc(m);
  }
  __syncwarp();
  for (uint32_t e = lane; e < used; e += 32) {
    const uint32_t key = keys[e];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < used; ++j) rank += keys[j] < key;
    order[rank] = (uint16_t)(key & 0xffff);
  }
  __syncwarp();
  return used;
}

__device__ __forceinline__ uint32_t warp_incl_scan_u32(uint32_t v, int lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const uint32_t t = __shfl_up_sync(FULL, v, d);
    if (lane >= d) v += t;
  }
  return v;
}

// chunk_off == nullptr: block b is in[b * 0xFF00, ...) of in_len bytes; else chunk b is
// in[chunk_off[b] - chunk_off[0], chunk_off[b + 1] - chunk_off[0]) (BamWriter ends a block where a record would not fit).
__global__ void __launch_bounds__(32) deflate_warp_kernel(const uint8_t* __restrict__ in, uint64_t in_len,
                                                          const uint64_t* __restrict__ chunk_off, uint32_t n_blocks,
                                                          uint8_t* __restrict__ slots, uint64_t* __restrict__ in_off,
                                                          uint32_t* __restrict__ isize, uint32_t* __restrict__ total,
                                                          uint32_t* __restrict__ toks, uint32_t* __restrict__ counter, int level) {
  __shared__ EncSmem s;
  const int lane = threadIdx.x;
  const uint32_t lt = (1u << lane) - 1;
  uint32_t* tok = toks + (size_t)blockIdx.x * TOK_STRIDE;
  while (true) {
    uint32_t b = 0;
    if (lane == 0) b = atomicAdd(counter, 1u);
    b = __shfl_sync(FULL, b, 0);
    if (b >= n_blocks) break;
    uint64_t off;
    uint32_t n;
    if (chunk_off) {
      off = chunk_off[b] - chunk_off[0];
      n = (uint32_t)(chunk_off[b + 1] - chunk_off[b]);
    } else {
      off = (uint64_t)b * BGZF_CHUNK;
      n = (uint32_t)(in_len - off < BGZF_CHUNK ? in_len - off : BGZF_CHUNK);
    }
    const uint8_t* src = in + off;
    uint8_t* out = slots + (size_t)b * SLOT + PAYLOAD_AT;
    uint32_t* out32 = reinterpret_cast<uint32_t*>(out);
    const uint32_t cap = SLOT - PAYLOAD_AT - 8;
    if (lane == 0) { in_off[b] = off; isize[b] = n; }
    bool stored = level == 0 || n < 8;
    uint32_t size = 0;
    if (!stored) {
      // ---- parse --------------------------------------------------------------------------------------------
      for (uint32_t k = lane; k < DEFL_HASH_SIZE * DEFL_WAYS / 2; k += 32) reinterpret_cast<uint32_t*>(s.htab)[k] = 0xffffffffu;
      for (uint32_t k = lane; k < 288 + 32; k += 32) s.hist[k] = 0;
      __syncwarp();
      uint32_t i = 0, nt = 0, extra = 0;
      while (i < n) {
        const uint32_t p = i + lane;
        const bool inb = p < n, valid = p + 4 <= n;
        const uint32_t f4 = inb ? load32u(src + p) : 0u;
        const uint32_t h = defl_hash(f4);
        const uint32_t same = __match_any_sync(FULL, valid ? h : (0x10000u | (uint32_t)lane));
        uint32_t c0 = DEFL_NONE, c1 = DEFL_NONE, c2 = DEFL_NONE;
        if (valid) {
          const uint32_t e = *reinterpret_cast<const uint32_t*>(s.htab + h * 2);    // as the table stood before this window
          c1 = e & 0xffffu;
          c2 = e >> 16;
          const uint32_t lower = same & lt;
          if (lower) c0 = i + (31u - (uint32_t)__clz((int)lower));
        }
        __syncwarp();
        if (valid && (same >> lane) == 1u) {
          const uint32_t rest = same & lt;
          const uint32_t w1 = rest ? i + (31u - (uint32_t)__clz((int)rest)) : c1;
          *reinterpret_cast<uint32_t*>(s.htab + h * 2) = p | (w1 << 16);
        }
        // match lengths of the candidates, 4 bytes a step, all three side by side
        const uint32_t maxl = valid ? (n - p < DEFL_WCAP ? n - p : DEFL_WCAP) : 0u;
        uint32_t m0 = 0, m1 = 0, m2 = 0;
        bool a0 = c0 != DEFL_NONE, a1 = c1 != DEFL_NONE && p - c1 <= 32768, a2 = c2 != DEFL_NONE && p - c2 <= 32768;
        for (uint32_t k = 0; k < DEFL_WCAP; k += 4) {
          if (!__any_sync(FULL, a0 || a1 || a2)) break;
          if (a0 || a1 || a2) {
            const uint32_t wp = k ? load32u(src + p + k) : f4;
            if (a0) {
              const uint32_t x = load32u(src + c0 + k) ^ wp;
              if (x) { m0 = k + (((uint32_t)__ffs((int)x) - 1u) >> 3); a0 = false; } else { m0 = k + 4; a0 = m0 < maxl; }
            }
            if (a1) {
              const uint32_t x = load32u(src + c1 + k) ^ wp;
              if (x) { m1 = k + (((uint32_t)__ffs((int)x) - 1u) >> 3); a1 = false; } else { m1 = k + 4; a1 = m1 < maxl; }
            }
            if (a2) {
              const uint32_t x = load32u(src + c2 + k) ^ wp;
              if (x) { m2 = k + (((uint32_t)__ffs((int)x) - 1u) >> 3); a2 = false; } else { m2 = k + 4; a2 = m2 < maxl; }
            }
          }
        }
        m0 = min(m0, maxl); m1 = min(m1, maxl); m2 = min(m2, maxl);
        uint32_t len = m0, q = c0;
        if (m1 > len) { len = m1; q = c1; }
        if (m2 > len) { len = m2; q = c2; }
        if (len < DEFL_MIN_MATCH) len = 0;
        const uint32_t dist = p - q;
        // one step of lazy evaluation, then the chain of tokens from lane 0 through the window
        uint32_t nx = __shfl_down_sync(FULL, len, 1);
        if (lane == 31) nx = 0;
        const uint32_t eff = (len && nx > len) ? 0u : len;
        uint32_t J = inb ? min(32u, (uint32_t)lane + (eff ? eff : 1u)) : 32u;
        uint32_t M = 1u << lane;
#pragma unroll
        for (int r = 0; r < 5; ++r) {
          const uint32_t Mj = __shfl_sync(FULL, M, J & 31), Jj = __shfl_sync(FULL, J, J & 31);
          if (J < 32) { M |= Mj; J = Jj; }
        }
        uint32_t sel = __shfl_sync(FULL, M, 0);
        if (n - i < 32) sel &= (1u << (n - i)) - 1;
        const uint32_t ls = 31u - (uint32_t)__clz((int)sel);              // the last token of the window (sel has bit 0)
        const uint32_t eff_ls = __shfl_sync(FULL, eff, ls);
        uint32_t len_ls = eff_ls;
        if (eff_ls == DEFL_WCAP) {                                        // it may go on: the whole warp follows it
          const uint32_t pp = i + ls, qq = pp - __shfl_sync(FULL, dist, ls);
          const uint32_t ml = n - pp < 258 ? n - pp : 258;
          while (len_ls < ml) {
            const uint32_t o = len_ls + 4 * lane;
            const uint32_t x = o < ml ? (load32u(src + qq + o) ^ load32u(src + pp + o)) : 0u;
            const uint32_t mism = __ballot_sync(FULL, x != 0);
            if (mism) {
              const int fl = __ffs((int)mism) - 1;
              const uint32_t xf = __shfl_sync(FULL, x, fl);
              len_ls = min(ml, len_ls + 4u * fl + (((uint32_t)__ffs((int)xf) - 1u) >> 3));
              break;
            }
            len_ls = min(ml, len_ls + 128u);
          }
        }
        const bool issel = (sel >> lane) & 1u;
        if (issel) {
          uint32_t t;
          if (eff) {
            const uint32_t L = (uint32_t)lane == ls ? len_ls : eff;
            t = DEFL_TOK_MATCH | ((L - 3) << 16) | (dist - 1);
            uint32_t ci, eb, ev;
            defl_len_code(L, &ci, &eb, &ev);
            atomicAdd(&s.hist[257 + ci], 1u);
            extra += eb;
            defl_dist_code(dist, &ci, &eb, &ev);
            atomicAdd(&s.hist[288 + ci], 1u);
            extra += eb;
          } else {
            t = f4 & 0xffu;
            atomicAdd(&s.hist[t], 1u);
          }
          tok[nt + __popc(sel & lt)] = t;
        }
        nt += __popc(sel);
        const uint32_t exitp = i + ls + (eff_ls ? len_ls : 1u);
        __syncwarp();
        for (uint32_t b0 = i + 32; b0 < exitp; b0 += 32) {               // what a long match skipped
          const uint32_t pp = b0 + lane;
          const bool v = pp < exitp && pp + 4 <= n;
          const uint32_t hh = v ? defl_hash(load32u(src + pp)) : 0u;
          table_insert(s.htab, v, hh, pp, b0, lane);
        }
        i = exitp;
      }
      if (lane == 0) tok[nt] = DEFL_TOK_EOB;
      ++nt;
      extra = __reduce_add_sync(FULL, extra);
      __syncwarp();
      // ---- codes --------------------------------------------------------------------------------------------
      DeflWork* w = &s.post.work;
      for (uint32_t k = lane; k < 286; k += 32) w->cnt.ll[k] = (uint16_t)s.hist[k];
      if (lane < 30) w->cnt.dd[lane] = (uint16_t)s.hist[288 + lane];
      if (lane == 0) { w->cnt.ll[256] = 1; w->cnt.extra = extra; }
      for (uint32_t k = lane; k < STAGE_WORDS; k += 32) s.post.stage[k] = 0;
      __syncwarp();
      uint32_t used = warp_sort_symbols(w->cnt.ll, 286, w->order, w->weight, lane);
      if (lane == 0) defl_code_lengths_sorted(w->cnt.ll, 286, 15, w->ll_len, w, used);
      __syncwarp();
      used = warp_sort_symbols(w->cnt.dd, 30, w->order, w->weight, lane);
      uint32_t hb = 0, dyn_size = 0;
      if (lane == 0) {
        defl_code_lengths_sorted(w->cnt.dd, 30, 15, w->dd_len, w, used);
        uint64_t bits = 0;
        const bool dynamic = defl_plan(w, &bits);
        dyn_size = (uint32_t)((bits + 7) / 8);
        if (dyn_size < n + 5 && dyn_size <= cap) {
          DeflBits hbits{reinterpret_cast<uint8_t*>(s.post.stage), 0, STAGE_WORDS * 4, 0, 0};
          defl_write_header(hbits, w, dynamic);
          hb = hbits.n * 8 + hbits.bits;
          hbits.flush();
        }
      }
      hb = __shfl_sync(FULL, hb, 0);
      size = __shfl_sync(FULL, dyn_size, 0);
      __syncwarp();
      if (hb == 0) {
        stored = true;
      } else {
        // ---- emit -------------------------------------------------------------------------------------------
        uint32_t* stage = s.post.stage;
        uint32_t wbase = hb >> 5, bitpos = hb;
        for (uint32_t k = lane; k < wbase; k += 32) out32[k] = stage[k];
        const uint32_t part0 = stage[wbase];
        __syncwarp();
        for (uint32_t k = lane; k < STAGE_WORDS; k += 32) stage[k] = 0;
        __syncwarp();
        if (lane == 0) stage[0] = part0;
        __syncwarp();
        for (uint32_t t0 = 0; t0 < nt; t0 += 32) {
          const uint32_t k = t0 + lane;
          uint32_t nb = 0;
          uint64_t v = 0;
          if (k < nt) v = defl_token_bits(tok[k], w->ll_code, w->ll_len, w->dd_code, w->dd_len, &nb);
          const uint32_t incl = warp_incl_scan_u32(nb, lane);
          const uint32_t tot = __shfl_sync(FULL, incl, 31);
          if (nb) {
            const uint32_t start = bitpos + incl - nb, wi = (start >> 5) - wbase, sh = start & 31;
            const uint64_t lo = v << sh;
            const uint32_t x0 = (uint32_t)lo, x1 = (uint32_t)(lo >> 32), x2 = sh ? (uint32_t)(v >> (64 - sh)) : 0u;
            if (x0) atomicOr(&stage[wi], x0);
            if (x1) atomicOr(&stage[wi + 1], x1);
            if (x2) atomicOr(&stage[wi + 2], x2);
          }
          __syncwarp();
          bitpos += tot;
          const uint32_t fw = (bitpos >> 5) - wbase;                       // whole words: at most 49
          for (uint32_t j = lane; j < fw; j += 32) out32[wbase + j] = stage[j];
          const uint32_t part = stage[fw];
          __syncwarp();
          for (uint32_t j = lane; j < fw + 3; j += 32) stage[j] = 0;
          __syncwarp();
          if (lane == 0) stage[0] = part;
          __syncwarp();
          wbase += fw;
        }
        if (lane == 0 && (bitpos & 31)) out32[wbase] = stage[0];
        size = (bitpos + 7) >> 3;
      }
      __syncwarp();
    }
    if (stored) {                                                          // deflate_stored, by the whole warp
      if (lane == 0) {
        out[0] = 1;
        out[1] = (uint8_t)n;
        out[2] = (uint8_t)(n >> 8);
        out[3] = (uint8_t)~n;
        out[4] = (uint8_t)(~n >> 8);
      }
      for (uint32_t k = lane; k < n; k += 32) out[5 + k] = src[k];
      size = n + 5;
    }
    if (lane == 0) total[b] = size + 26;                                   // header 18 + payload + footer 8 (compress.d:88)
    __syncwarp();
  }
}

__global__ void __launch_bounds__(128) bgzf_pack_kernel(const uint8_t* __restrict__ slots, const uint32_t* __restrict__ total,
                                                        const uint64_t* __restrict__ out_off, const uint32_t* __restrict__ crc,
                                                        const uint32_t* __restrict__ isize, uint8_t* __restrict__ out) {
  const uint32_t b = blockIdx.x;
  const uint32_t t = total[b];
  uint8_t* dst = out + out_off[b];
  const uint8_t* src = slots + (size_t)b * SLOT + PAYLOAD_AT;
  if (threadIdx.x == 0) {
    const uint8_t head[16] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0};     // BLOCK_HEADER_START (constants.d:30-38)
    for (int k = 0; k < 16; ++k) dst[k] = head[k];
    dst[16] = (uint8_t)(t - 1);                                                            // BSIZE = block length - 1
    dst[17] = (uint8_t)((t - 1) >> 8);
    const uint32_t c = crc[b], n = isize[b];
    for (int k = 0; k < 4; ++k) { dst[t - 8 + k] = (uint8_t)(c >> (8 * k)); dst[t - 4 + k] = (uint8_t)(n >> (8 * k)); }
  }
  for (uint32_t i = threadIdx.x; i < t - 26; i += blockDim.x) dst[18 + i] = src[i];
}

// copies of the size the slabs have are worth a few threads
void par_memcpy(void* dst, const void* src, size_t n) {
  const size_t piece = (size_t)8 << 20;
  if (n < 2 * piece) { memcpy(dst, src, n); return; }
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const size_t nt = std::min<size_t>({(size_t)8, (size_t)hw, (n + piece - 1) / piece});
  const size_t per = ((n + nt - 1) / nt + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (size_t t = 1; t < nt; ++t) {
    const size_t a = t * per;
    if (a >= n) break;
    th.emplace_back([=] { memcpy((uint8_t*)dst + a, (const uint8_t*)src + a, std::min(per, n - a)); });
  }
  memcpy(dst, src, std::min(per, n));
  for (std::thread& t : th) t.join();
}

// What one device keeps between calls: two sets of buffers and streams (slab k uses set k & 1).
struct EncSet {
  cudaStream_t st = nullptr;
  DevBuf d_in, d_coff, d_slots, d_off, d_isize, d_total, d_crc, d_ooff, d_tmp, d_out, d_tok, d_counter;
  PinBuf h_in, h_out, h_coff, h_tot;
  uint32_t nb = 0;
  uint64_t in_len = 0;
};
struct EncCtx {
  std::mutex mu;
  EncSet set[2];
  int grid = 0;
  uint64_t kernel_us = 0, calls = 0;     // (statistics of the last call, biodb_debug_deflate_stats)
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};
EncCtx* enc_ctx(int device) {
  static std::mutex mu;
  static EncCtx* ctx[64] = {};
  std::lock_guard<std::mutex> g(mu);
  if (device < 0 || device >= 64) return nullptr;
  if (!ctx[device]) ctx[device] = new EncCtx;
  return ctx[device];
}

// Chunks of `data` -> BGZF blocks handed to `sink(bytes, n)` in order.  chunk_off == nullptr: one block per 0xFF00 bytes
// of data[0, len); else n_all chunks [chunk_off[b], chunk_off[b + 1]) (each 1 .. 0xFF00 bytes).
template <typename Sink>
biodb_status compress_slabs(int32_t device, const uint8_t* data, size_t len, const uint64_t* chunk_off, size_t n_all, int32_t level,
                            Sink&& sink) {
  if (n_all == 0) return BIODB_OK;
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BIODB_ERR_CUDA;           // no CPU fallback
  if (device >= 0 && cudaSetDevice(device) != cudaSuccess) return BIODB_ERR_CUDA;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return BIODB_ERR_CUDA;
  EncCtx* cx = enc_ctx(dev);
  if (!cx) return BIODB_ERR_CUDA;
  std::lock_guard<std::mutex> guard(cx->mu);
  if (!cx->grid) {
    int per_sm = 0, sms = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, deflate_warp_kernel, 32, 0) != cudaSuccess ||
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || per_sm < 1 || sms < 1)
      return BIODB_ERR_CUDA;
    for (EncSet& s : cx->set)
      if (!s.st && cudaStreamCreateWithFlags(&s.st, cudaStreamNonBlocking) != cudaSuccess) return BIODB_ERR_CUDA;
    if ((!cx->ev0 && cudaEventCreate(&cx->ev0) != cudaSuccess) || (!cx->ev1 && cudaEventCreate(&cx->ev1) != cudaSuccess))
      return BIODB_ERR_CUDA;
    cx->grid = per_sm * sms;                               // (set last: a context that failed half-way is set up again)
  }
  cx->kernel_us = 0;
  ++cx->calls;
  const size_t n_slabs = (n_all + SLAB_BLOCKS - 1) / SLAB_BLOCKS;
  // input -> pinned -> device, the encoder, CRC-32, the scan of the block sizes and the total back: all asynchronous
  auto submit = [&](size_t k) -> bool {
    EncSet& s = cx->set[k & 1];
    const size_t b0 = k * SLAB_BLOCKS;
    const uint32_t nb = (uint32_t)std::min<size_t>(SLAB_BLOCKS, n_all - b0);
    const uint64_t in0 = chunk_off ? chunk_off[b0] : (uint64_t)b0 * BGZF_CHUNK;
    const uint64_t in1 = chunk_off ? chunk_off[b0 + nb] : std::min<uint64_t>((uint64_t)(b0 + nb) * BGZF_CHUNK, len);
    const uint64_t in_len = in1 - in0;
    const uint32_t grid = (uint32_t)std::min<size_t>((size_t)cx->grid, nb);
    cudaStream_t st = s.st;
    s.nb = nb;
    s.in_len = in_len;
    bool ok = s.d_in.ensure((size_t)in_len + 64, st) == cudaSuccess && s.h_in.ensure((size_t)in_len + 64) == cudaSuccess &&
              s.d_slots.ensure((size_t)nb * SLOT, st) == cudaSuccess && s.d_off.ensure((size_t)nb * 8, st) == cudaSuccess &&
              s.d_isize.ensure((size_t)nb * 4, st) == cudaSuccess && s.d_total.ensure((size_t)(nb + 1) * 4, st) == cudaSuccess &&
              s.d_crc.ensure((size_t)nb * 4, st) == cudaSuccess && s.d_ooff.ensure((size_t)(nb + 1) * 8, st) == cudaSuccess &&
              s.d_tmp.ensure((scan_temp_elems(nb + 1) + 8) * 8, st) == cudaSuccess &&
              s.d_tok.ensure((size_t)grid * TOK_STRIDE * 4, st) == cudaSuccess && s.d_counter.ensure(64, st) == cudaSuccess &&
              s.h_tot.ensure(64) == cudaSuccess;
    if (!ok) return false;
    par_memcpy(s.h_in.p, data + in0, (size_t)in_len);
    ok = cudaMemcpyAsync(s.d_in.p, s.h_in.p, (size_t)in_len, cudaMemcpyHostToDevice, st) == cudaSuccess;
    const uint64_t* d_coff = nullptr;
    if (ok && chunk_off) {
      ok = s.d_coff.ensure((size_t)(nb + 1) * 8, st) == cudaSuccess && s.h_coff.ensure((size_t)(nb + 1) * 8) == cudaSuccess;
      if (ok) {
        memcpy(s.h_coff.p, chunk_off + b0, (size_t)(nb + 1) * 8);
        ok = cudaMemcpyAsync(s.d_coff.p, s.h_coff.p, (size_t)(nb + 1) * 8, cudaMemcpyHostToDevice, st) == cudaSuccess;
        d_coff = s.d_coff.as<uint64_t>();
      }
    }
    ok = ok && cudaMemsetAsync(s.d_counter.p, 0, 4, st) == cudaSuccess;
    if (!ok) return false;
    if (k == 0) cudaEventRecord(cx->ev0, st);            // (the statistics time the encoder kernel of the first slab)
    deflate_warp_kernel<<<grid, 32, 0, st>>>(s.d_in.as<uint8_t>(), in_len, d_coff, nb, s.d_slots.as<uint8_t>(), s.d_off.as<uint64_t>(),
                                             s.d_isize.as<uint32_t>(), s.d_total.as<uint32_t>(), s.d_tok.as<uint32_t>(),
                                             s.d_counter.as<uint32_t>(), level);
    ++g_kernel_launches;
    if (k == 0) cudaEventRecord(cx->ev1, st);
    ok = cudaMemsetAsync(s.d_total.as<uint32_t>() + nb, 0, 4, st) == cudaSuccess &&
         launch_crc32(s.d_in.as<uint8_t>(), s.d_off.as<uint64_t>(), s.d_isize.as<uint32_t>(), nb, s.d_crc.as<uint32_t>(), st) == cudaSuccess;
    if (!ok) return false;
    device_scan<false>(s.d_total.as<uint32_t>(), s.d_ooff.as<uint64_t>(), (uint64_t)nb + 1, s.d_tmp.as<uint64_t>(), OpAdd(), (uint64_t)0, st);
    return cudaMemcpyAsync(s.h_tot.p, s.d_ooff.as<uint64_t>() + nb, 8, cudaMemcpyDeviceToHost, st) == cudaSuccess;
  };
  // once the total is known: pack, copy back, hand out
  auto finish = [&](size_t k) -> biodb_status {
    EncSet& s = cx->set[k & 1];
    cudaStream_t st = s.st;
    if (cudaStreamSynchronize(st) != cudaSuccess) return BIODB_ERR_CUDA;
    if (k == 0) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, cx->ev0, cx->ev1) == cudaSuccess) cx->kernel_us = (uint64_t)(ms * 1000.0f);
    }
    const uint64_t tot = *s.h_tot.as<uint64_t>();
    if (s.d_out.ensure((size_t)tot + 64, st) != cudaSuccess || s.h_out.ensure((size_t)tot + 64) != cudaSuccess) return BIODB_ERR_NOMEM;
    bgzf_pack_kernel<<<s.nb, 128, 0, st>>>(s.d_slots.as<uint8_t>(), s.d_total.as<uint32_t>(), s.d_ooff.as<uint64_t>(),
                                           s.d_crc.as<uint32_t>(), s.d_isize.as<uint32_t>(), s.d_out.as<uint8_t>());
    ++g_kernel_launches;
    if (cudaMemcpyAsync(s.h_out.p, s.d_out.p, (size_t)tot, cudaMemcpyDeviceToHost, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return BIODB_ERR_CUDA;
    return sink(s.h_out.as<uint8_t>(), (size_t)tot);
  };
  biodb_status rc = submit(0) ? BIODB_OK : BIODB_ERR_CUDA;
  for (size_t k = 0; k < n_slabs && rc == BIODB_OK; ++k) {
    if (k + 1 < n_slabs && !submit(k + 1)) rc = BIODB_ERR_CUDA;
    const biodb_status f = finish(k);
    if (rc == BIODB_OK) rc = f;
  }
  for (EncSet& s : cx->set) cudaStreamSynchronize(s.st);
  return rc;
}

}  // namespace

}  // namespace biodb

using namespace biodb;

extern "C" {

size_t biodb_bgzf_compress_bound(size_t len) {
  const size_t nb = (len + BGZF_CHUNK - 1) / BGZF_CHUNK;
  return nb * (size_t)SLOT + 28;
}

// Host-only: the encoder of deflate_enc.h as the host states it, for the tests (raw DEFLATE of one chunk; the device
// writes the same bytes).
int64_t biodb_debug_deflate_block(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, int32_t level) {
  if ((!in && n) || !out) return -1;
  return (int64_t)deflate_block_host(in, n, out, cap, level);
}

// Statistics of the last biodb_bgzf_compress / biodb_writer_finish on `device`: out[0] = microseconds the encoder kernel
// of the first slab took, out[1] = CTAs of its persistent grid, out[2] = calls so far.
biodb_status biodb_debug_deflate_stats(int32_t device, uint64_t* out) {
  if (!out) return BIODB_ERR_ARG;
  int dev = device;
  if (dev < 0 && cudaGetDevice(&dev) != cudaSuccess) return BIODB_ERR_CUDA;
  EncCtx* cx = enc_ctx(dev);
  if (!cx) return BIODB_ERR_ARG;
  std::lock_guard<std::mutex> g(cx->mu);
  out[0] = cx->kernel_us;
  out[1] = (uint64_t)cx->grid;
  out[2] = cx->calls;
  out[3] = 0;
  return BIODB_OK;
}

biodb_status biodb_bgzf_compress(int32_t device, const void* data, size_t len, int32_t level, int32_t add_eof, void* out,
                                 size_t cap, size_t* out_len) {
  static const uint8_t EOF_BLOCK[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if ((!data && len) || !out || !out_len || level < -1 || level > 9) return BIODB_ERR_ARG;   // compress.d:46-48
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) return BIODB_ERR_CUDA;           // no CPU fallback
  size_t written = 0;
  const size_t n_all = (len + BGZF_CHUNK - 1) / BGZF_CHUNK;
  const biodb_status rc = compress_slabs(device, (const uint8_t*)data, len, nullptr, n_all, level, [&](const uint8_t* p, size_t n) {
    if (written + n + (add_eof ? 28 : 0) > cap) return BIODB_ERR_NOMEM;
    par_memcpy((uint8_t*)out + written, p, n);
    written += n;
    return BIODB_OK;
  });
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
  out->reserve(out->size() + (size_t)((chunk_off[n_all] - chunk_off[0]) / 2));
  return compress_slabs(device, data, 0, chunk_off, n_all, level, [&](const uint8_t* p, size_t n) {
    out->insert(out->end(), p, p + n);
    return BIODB_OK;
  });
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
  std::vector<uint8_t> bytes;          // stream bytes from offset bytes_base on (blocks already compressed are dropped)
  uint64_t bytes_base = 0;
  uint64_t stream_len() const { return bytes_base + bytes.size(); }
  std::vector<uint64_t> cuts{0};       // starts (stream offsets) of the blocks that are complete; the current block starts at cuts.back()
  size_t cuts_done = 0;                // blocks [0, cuts_done) have been compressed (biodb_writer_drain / _finish)
  std::vector<uint64_t> block_at{0};   // file offset of every compressed block, then the length of the file so far
  bool drained = false, finished = false;
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
    cuts.push_back(stream_len());
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
      w->recs.push_back(biodb_writer::Rec{w->stream_len(), (uint32_t)read_size, ref_id, pos, (int32_t)((uint32_t)pos + covered), bin, (flag & 4) != 0});
      w->write(rec.data(), rec.size());
      w->rec_cur = read_size;
    } else {
      w->recs.push_back(biodb_writer::Rec{w->stream_len(), (uint32_t)read_size, ref_id, pos, (int32_t)((uint32_t)pos + covered), bin, (flag & 4) != 0});
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
  if (w->bytes_base) return BIODB_ERR_ARG;                              // (blocks were already handed out: biodb_writer_drain)
  *data = w->bytes.data();
  *len = w->bytes.size();
  *cuts = w->cuts.data();
  *n_cuts = w->cuts.size();
  return BIODB_OK;
}

// Compress the complete blocks [cuts_done, upto) into w->out (appended), note where each lands in the file, drop their
// uncompressed bytes.
static biodb_status writer_compress(biodb_writer* w, size_t upto) {
  if (upto <= w->cuts_done) return BIODB_OK;
  std::vector<uint64_t> rel(upto - w->cuts_done + 1);
  for (size_t k = 0; k < rel.size(); ++k) rel[k] = w->cuts[w->cuts_done + k] - w->bytes_base;
  const size_t at = w->out.size();
  const biodb_status rc = compress_chunks(w->device, w->bytes.data(), rel.data(), rel.size() - 1, w->level, &w->out);
  if (rc != BIODB_OK) return rc;
  // the BSIZE chain of what was just written: the file offsets of these blocks
  uint64_t file0 = w->block_at.back();
  w->block_at.pop_back();
  size_t p = at;
  while (p + 18 <= w->out.size()) {
    w->block_at.push_back(file0 + (p - at));
    p += ((size_t)w->out[p + 16] | ((size_t)w->out[p + 17] << 8)) + 1;
  }
  w->block_at.push_back(file0 + (w->out.size() - at));
  const uint64_t drop = w->cuts[upto] - w->bytes_base;
  w->bytes.erase(w->bytes.begin(), w->bytes.begin() + (ptrdiff_t)drop);
  w->bytes_base = w->cuts[upto];
  w->cuts_done = upto;
  return BIODB_OK;
}

// Streaming use (bgzf/outputstream.d:136-173 hands every full block to the task pool as it completes): if at least
// min_blocks complete blocks have accumulated, compress them now and hand their BGZF bytes out — *data is valid until the
// next call on this writer; *len = 0 if there was not enough to do.  The uncompressed bytes of those blocks are released:
// a writer that is drained regularly holds min_blocks blocks, not the file.  After a drain, biodb_writer_finish returns
// only what has not been handed out yet.
biodb_status biodb_writer_drain(biodb_writer* w, uint32_t min_blocks, const uint8_t** data, size_t* len) {
  if (!w || !data || !len || w->finished) return BIODB_ERR_ARG;
  *data = nullptr;
  *len = 0;
  const size_t ready = w->cuts.size() - 1 - w->cuts_done;
  if (ready == 0 || ready < min_blocks) return BIODB_OK;
  w->out.clear();
  w->drained = true;
  const biodb_status rc = writer_compress(w, w->cuts.size() - 1);
  if (rc != BIODB_OK) return rc;
  *data = w->out.data();
  *len = w->out.size();
  return BIODB_OK;
}

// finish (writer.d:276-280): every block not yet handed out compressed on the device, the EOF block appended; *data stays
// valid until biodb_writer_end.  Without earlier drains that is the whole file.
biodb_status biodb_writer_finish(biodb_writer* w, const uint8_t** data, size_t* len) {
  static const uint8_t EOF_BLOCK[28] = {31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (!w || !data || !len || w->finished) return BIODB_ERR_ARG;
  w->flush_current_block();
  w->out.clear();
  const biodb_status rc = writer_compress(w, w->cuts.size() - 1);
  if (rc != BIODB_OK) return rc;
  w->out.insert(w->out.end(), EOF_BLOCK, EOF_BLOCK + 28);
  w->finished = true;
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
  w->block_at.clear();
  size_t p = 0;
  while (p + 18 <= len && w->block_at.size() < w->cuts.size() - 1) {      // the blocks of the layout (the EOF block is not one)
    w->block_at.push_back(p);
    p += ((size_t)data[p + 16] | ((size_t)data[p + 17] << 8)) + 1;
  }
  w->block_at.push_back(p);
  w->finished = true;
  return BIODB_OK;
}

// The BAI index of the finished file — what BamWriter builds while writing coordinate-sorted output (writer.d:139-195,
// 171-175: IndexBuilder with check_bins, fed with every record and the virtual offsets it got in the file).  Call after
// biodb_writer_finish.  BIODB_ERR_UNSORTED if the records were not in coordinate order.
biodb_status biodb_writer_index(biodb_writer* w, const uint8_t** data, size_t* len) {
  if (!w || !data || !len || !w->finished) return BIODB_ERR_ARG;
  // where the blocks of the layout begin in the file: noted while they were compressed (block_at)
  const size_t nb = w->cuts.size() - 1;
  const std::vector<uint64_t>& cb = w->block_at;
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
