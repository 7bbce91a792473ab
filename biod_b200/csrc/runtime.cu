// Host runtime + C ABI of libbiod_b200.so (see include/biod_b200.h and runtime.h).
#include "runtime.h"
#include "bai.h"
#include "bai_build.h"
#include "md_chain.h"
#include "md_walk.h"

#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <cstdio>
#include <thread>
#include <cstdlib>
#include <cstring>

using namespace biodb;

#define CUDA_TRY(expr)                                                                        \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return this->fail(BIODB_ERR_CUDA, 0, 0, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

namespace biodb {

cudaError_t DevBuf::ensure(size_t bytes, cudaStream_t st, size_t keep) {
  if (bytes <= cap) return cudaSuccess;
  size_t ncap = std::max(bytes, cap + cap / 2);
  ncap = (ncap + 255) & ~(size_t)255;
  void* np = nullptr;
  cudaError_t e = cudaMalloc(&np, ncap);
  if (e != cudaSuccess) return e;
  if (p) {
    if (keep) {
      e = cudaMemcpyAsync(np, p, std::min(keep, cap), cudaMemcpyDeviceToDevice, st);
      if (e == cudaSuccess) e = cudaStreamSynchronize(st);
      if (e != cudaSuccess) { cudaFree(np); return e; }
    } else {
      cudaStreamSynchronize(st);
    }
    cudaFree(p);
  }
  p = np;
  cap = ncap;
  return cudaSuccess;
}

cudaError_t PinBuf::ensure(size_t bytes) {
  if (bytes <= cap) return cudaSuccess;
  size_t ncap = std::max(bytes, cap + cap / 2);
  void* np = nullptr;
  cudaError_t e = cudaHostAlloc(&np, ncap, cudaHostAllocMapped | cudaHostAllocPortable);
  if (e != cudaSuccess) return e;
  if (p) cudaFreeHost(p);
  p = np;
  cap = ncap;
  return cudaSuccess;
}

static inline uint16_t le16(const uint8_t* p) { return (uint16_t)(p[0] | (p[1] << 8)); }
static inline uint32_t le32(const uint8_t* p) {
  return (uint32_t)p[0] | ((uint32_t)p[1] << 8) | ((uint32_t)p[2] << 16) | ((uint32_t)p[3] << 24);
}

static void set_error(biodb_error* e, int status, int zerr, uint64_t off, const std::string& msg) {
  e->status = status;
  e->zlib_errnum = zerr;
  e->file_offset = off;
  snprintf(e->message, sizeof e->message, "%s", msg.c_str());
}

// BGZF member header — the checks and messages of fillBgzfBufferFromStream
// (bio/core/bgzf/inputstream.d:54-199).  1 = block, 0 = clean end of stream, <0 = BgzfException.
// d[0 .. len) are the file's bytes from absolute offset `abs` on; pos is relative to d.
int parse_bgzf_header(const uint8_t* d, uint64_t len, uint64_t pos, BlockInfo* b, biodb_error* e, uint64_t abs) {
  auto bad = [&](const std::string& why) {
    set_error(e, BIODB_ERR_BGZF, 0, abs + pos,
              "Error reading BGZF block starting from offset " + std::to_string(abs + pos) + ": " + why);
    return (int)BIODB_ERR_BGZF;
  };
  static const char* kShort = "stream error: not enough data in stream";
  if (pos >= len || len - pos < 4) return 0;
  const uint8_t* h = d + pos;
  if (h[0] != 0x1f || h[1] != 0x8b || h[2] != 0x08 || h[3] != 0x04) return bad("wrong BGZF magic");
  if (pos + 12 > len) return bad(kShort);
  const uint32_t xlen = le16(h + 10);
  uint64_t q = pos + 12;
  uint32_t seen = 0, bsize = 0;
  bool have = false;
  while (seen < xlen) {
    if (q + 4 > len) return bad(kShort);
    const uint8_t s1 = d[q], s2 = d[q + 1];
    const uint32_t slen = le16(d + q + 2);
    if (s1 == 'B' && s2 == 'C') {
      if (slen != 2) return bad("wrong BC subfield length: " + std::to_string(slen) + "; expected 2");
      if (have) return bad("duplicate field with block size");
      if (q + 6 > len) return bad(kShort);
      bsize = le16(d + q + 4);
      have = true;
    }
    q += 4 + slen;
    seen += 4 + slen;
  }
  if (seen != xlen)
    return bad("total length of subfields in bytes (" + std::to_string(seen) + ") is not equal to gzip_extra_length (" +
               std::to_string(xlen) + ")");
  if (!have) return bad("block size was not found in any subfield");
  const int64_t cdata = (int64_t)bsize - (int64_t)xlen - 19;
  if (cdata > 65536)
    return bad("compressed data size is more than 65536 bytes, which is not allowed by current BAM specification");
  if (cdata < 0 || q + (uint64_t)cdata + 8 > len) return bad(kShort);
  b->coffset = abs + pos;
  b->payload = abs + q;
  b->bsize = bsize;
  b->cdata_size = (uint32_t)cdata;
  b->crc32 = le32(d + q + cdata);
  b->isize = le32(d + q + cdata + 4);
  return 1;
}

// ---------------------------------------------------------------------------------- Pass ----

Pass::~Pass() {
  join_prefetch();
  if (st) {
    cudaStreamSynchronize(st);
    cudaStreamDestroy(st);
  }
  if (h2d_st) {
    cudaStreamSynchronize(h2d_st);
    cudaStreamDestroy(h2d_st);
  }
  for (cudaEvent_t e : {ev_begin, ev_end, ev_a, ev_b, pf_done, comp_free[0], comp_free[1]})
    if (e) cudaEventDestroy(e);
  for (const Timed& t : timed) { cudaEventDestroy(t.a); cudaEventDestroy(t.b); }
  for (cudaEvent_t e : ev_pool) cudaEventDestroy(e);
}

void Pass::mark_begin() {
  if (began) return;
  began = true;
  launches0 = g_kernel_launches;
  cudaEventRecord(ev_begin, st);
}
void Pass::mark_end() {
  if (!began) return;
  cudaEventRecord(ev_end, st);
  cudaEventSynchronize(ev_end);
  collect_timing();
  float ms = 0;
  if (cudaEventElapsedTime(&ms, ev_begin, ev_end) == cudaSuccess) stats.total_ms = ms;
  stats.kernel_launches = g_kernel_launches - launches0;
  stats.n_records = n_records_total;
}
static cudaEvent_t pool_event(std::vector<cudaEvent_t>& pool) {
  if (!pool.empty()) { cudaEvent_t e = pool.back(); pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
void Pass::stage_begin() {
  stage_a = pool_event(ev_pool);
  cudaEventRecord(stage_a, st);
}
void Pass::stage_end(double* acc) {
  cudaEvent_t b = pool_event(ev_pool);
  cudaEventRecord(b, st);
  timed.push_back(Timed{stage_a, b, acc});
  stage_a = nullptr;
}
void Pass::collect_timing() {
  for (const Timed& t : timed) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, t.a, t.b) == cudaSuccess) *t.acc += ms;
    ev_pool.push_back(t.a);
    ev_pool.push_back(t.b);
  }
  timed.clear();
}

biodb_status Pass::fail(int status, int zerr, uint64_t off, const std::string& msg) {
  set_error(&r->err, status, zerr, off, msg);
  finished = true;
  return (biodb_status)status;
}

biodb_status Pass::init(biodb_reader* rd, uint64_t coffset, uint32_t uoffset) {
  r = rd;
  if (cudaSetDevice(rd->device) != cudaSuccess) return fail(BIODB_ERR_CUDA, 0, 0, "cudaSetDevice failed");
  CUDA_TRY(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreate(&ev_begin));
  CUDA_TRY(cudaEventCreate(&ev_end));
  CUDA_TRY(cudaEventCreate(&ev_a));
  CUDA_TRY(cudaEventCreate(&ev_b));
  CUDA_TRY(cudaStreamCreateWithFlags(&h2d_st, cudaStreamNonBlocking));
  CUDA_TRY(cudaEventCreateWithFlags(&pf_done, cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&comp_free[0], cudaEventDisableTiming));
  CUDA_TRY(cudaEventCreateWithFlags(&comp_free[1], cudaEventDisableTiming));
  next_coffset = coffset;
  first_skip = uoffset;
  CUDA_TRY(h_result.ensure(64));
  CUDA_TRY(d_result.ensure(64));
  return BIODB_OK;
}

void Pass::rewind(uint64_t coffset, uint32_t uoffset) {
  stop_coffset = ~0ull;
  stop_uoffset = 0;
  next_coffset = coffset;
  first_skip = uoffset;
  entry_search = false;
  join_prefetch();
  pf_c0 = pf_c1 = 0;
  pre_valid = false;
  if (h2d_st) cudaStreamSynchronize(h2d_st);
  supplier_done = false;
  memset(&pending, 0, sizeof pending);
  finished = false;
  n_records_total = 0;
  blocks.clear();
  u_len = carry_tail_len = 0;
  segs.clear();
  next_segs.clear();
  n = n_cigar = tail = 0;
  final_slice = false;
  memset(&stats, 0, sizeof stats);
  began = false;
}

RecordArrays Pass::arrays(uint64_t front) const {
  RecordArrays a;
  a.rec_off = d_rec[0].as<uint64_t>() + front;
  a.block_size = d_rec[1].as<int32_t>() + front;
  a.ref_id = d_rec[2].as<int32_t>() + front;
  a.pos = d_rec[3].as<int32_t>() + front;
  a.end_pos = d_rec[4].as<int32_t>() + front;
  a.bin_mq_nl = d_rec[5].as<uint32_t>() + front;
  a.flag_nc = d_rec[6].as<uint32_t>() + front;
  a.l_seq = d_rec[7].as<int32_t>() + front;
  a.cigar_off = d_rec[8].as<uint64_t>() + front;
  a.cigar = d_rec[9].as<uint32_t>();
  a.capacity = rec_capacity;
  a.cigar_capacity = cigar_capacity;
  return a;
}

uint64_t Pass::voffset_of(uint64_t x) const { return voffset_in(segs, x); }

uint64_t voffset_in(const std::vector<Seg>& segs, uint64_t x) {
  // VirtualOffset (virtualoffset.d:43); an offset that is exactly a block end reports the next
  // block with uoffset 0 (inputstream.d:443-445,516-524).
  size_t lo = 0, hi = segs.size();
  while (lo < hi) {
    size_t m = (lo + hi) / 2;
    if (segs[m].ustart <= x) lo = m + 1; else hi = m;
  }
  if (lo == 0) return 0;
  const Seg& s = segs[lo - 1];
  if (x - s.ustart >= s.len) return s.cend << 16;
  return (s.coffset << 16) | (uint64_t)(s.within + (x - s.ustart));
}

// Host walk of the BSIZE chain: up to max_blocks block headers from next_coffset on (inputstream.d:54-199, 386-424).
static void walk_headers(const biodb_reader* r, uint64_t stop_coffset, uint32_t stop_uoffset, uint32_t max_blocks,
                         std::vector<BlockInfo>& blocks, uint64_t& next_coffset, bool& supplier_done, biodb_error& pending) {
  while (blocks.size() < max_blocks && !supplier_done && !pending.status) {
    if (next_coffset > stop_coffset || (next_coffset == stop_coffset && stop_uoffset == 0)) { supplier_done = true; break; }
    BlockInfo b;
    int rc = const_cast<biodb_reader*>(r)->header_at(next_coffset, &b, &pending);
    if (rc < 0) break;
    if (rc == 0 || b.isize == 0) { supplier_done = true; break; }     // EOF block ends the stream (inputstream.d:393-394)
    if (b.isize > 65536) {                                             // block.d:150-152
      set_error(&pending, BIODB_ERR_FORMAT, 0, b.coffset, "Uncompressed block size must be within 65536 bytes");
      break;
    }
    blocks.push_back(b);
    next_coffset = b.coffset + b.bsize + 1;
  }
}

int Pass::join_prefetch() {
  return pf_job.valid() ? pf_job.get() : 0;
}

const uint8_t* Pass::stage_host(uint64_t c0, uint64_t c1, cudaError_t* err) {
  *err = cudaSuccess;
  if (r->file) return r->file + c0;
  // streamed file: into the slab that is not the source of the copy issued last (that copy has been waited for by the
  // time a third range is staged: every batch synchronises its stream)
  slab_cur ^= 1;
  PinBuf& sl = h_slab[slab_cur];
  const size_t n = (size_t)(c1 - c0);
  if (n + 64 > sl.cap && (*err = sl.ensure(n + n / 8 + 65536)) != cudaSuccess) return nullptr;
  if (!r->read_bytes(c0, n, sl.p, 4)) { *err = cudaErrorUnknown; return nullptr; }
  return sl.as<uint8_t>();
}

biodb_status Pass::next(uint32_t max_blocks, uint64_t front_slots) {
  n = n_cigar = 0;
  if (finished) {
    if (pending.status) { r->err = pending; return (biodb_status)pending.status; }
    if (front_slots == 0) return BIODB_EOF;
    // nothing left in the file, but the caller still holds carried reads: hand out an empty final batch
    if (rec_front < front_slots || rec_capacity == 0) {
      rec_front = std::max(rec_front, front_slots);
      rec_capacity = std::max<uint64_t>(rec_capacity, 1024);
      cigar_capacity = std::max<uint64_t>(cigar_capacity, 1024);
      static const size_t esz0[10] = {8, 4, 4, 4, 4, 4, 4, 4, 8, 4};
      for (int a = 0; a < 9; ++a) CUDA_TRY(d_rec[a].ensure((size_t)(rec_capacity + rec_front + 2) * esz0[a], st));
      CUDA_TRY(d_rec[9].ensure((size_t)(cigar_capacity + 2) * 4, st));
    }
    u_len = 0;
    tail = 0;
    final_slice = true;
    return BIODB_OK;
  }
  if (!began && r->opts.pin_input == 2 && !r->d_file.p)
    r->ensure_pinned(next_coffset, stop_coffset == ~0ull ? r->flen : std::min<uint64_t>(r->flen, stop_coffset + 65536 + 64));
  mark_begin();
  // ---- 1. walk the BSIZE chain on the host (18 bytes per block) ----------------------------------
  blocks.clear();
  if (pre_valid && pre_max == max_blocks) {
    // already walked while the previous batch was being inflated
    blocks.swap(pre_blocks);
    next_coffset = pre_next_coffset;
    supplier_done = pre_supplier_done;
    pending = pre_pending;
  } else {
    walk_headers(r, stop_coffset, stop_uoffset, max_blocks, blocks, next_coffset, supplier_done, pending);
  }
  pre_valid = false;
  const uint32_t nb = (uint32_t)blocks.size();
  bool last_batch = supplier_done || pending.status != 0;
  if (nb == 0 && carry_tail_len == 0 && front_slots == 0) {
    finished = true;
    if (pending.status) { r->err = pending; return (biodb_status)pending.status; }
    return BIODB_EOF;
  }
  // ---- 2. per-block tables ---------------------------------------------------------------------------
  const bool has_carry = carry_tail_len > 0;
  const uint32_t nsb = nb + (has_carry ? 1 : 0);          // scan blocks (the carried tail is a pseudo block)
  const size_t tab_bytes = (size_t)nb * (8 + 8 + 4 + 4) + (size_t)(nsb + 2) * 8 + 64;
  CUDA_TRY(h_tab.ensure(tab_bytes));
  CUDA_TRY(d_tab.ensure(tab_bytes, st));
  uint8_t* hp = h_tab.as<uint8_t>();
  uint64_t* h_payload = (uint64_t*)hp;
  uint64_t* h_outoff = h_payload + nb;
  uint64_t* h_buoff = h_outoff + nb;
  uint32_t* h_cdata = (uint32_t*)(h_buoff + nsb + 2);
  uint32_t* h_isize = h_cdata + nb;
  const uint64_t c0 = nb ? blocks[0].coffset : 0;
  const uint64_t c1 = nb ? blocks[nb - 1].coffset + blocks[nb - 1].bsize + 1 : 0;
  segs.swap(next_segs);
  next_segs.clear();
  uint64_t off = carry_tail_len;
  uint32_t k = 0;
  if (has_carry) h_buoff[k++] = 0;
  for (uint32_t i = 0; i < nb; ++i) {
    h_payload[i] = blocks[i].payload - c0;
    h_outoff[i] = off;
    h_cdata[i] = blocks[i].cdata_size;
    h_isize[i] = blocks[i].isize;
    h_buoff[k++] = off;
    segs.push_back(Seg{off, blocks[i].coffset, 0, blocks[i].isize, blocks[i].coffset + blocks[i].bsize + 1});
    off += blocks[i].isize;
  }
  h_buoff[k] = off;
  if (!has_carry && nb) h_buoff[0] += first_skip;         // the first record of the file sits behind the header
  first_skip = 0;
  u_len = off;
  if (nb && stop_uoffset && blocks[nb - 1].coffset == stop_coffset) {
    // the stream ends inside this block (end of a BAI chunk): the block is inflated whole, but the record scan sees
    // only its first stop_uoffset bytes
    u_len = h_outoff[nb - 1] + std::min<uint32_t>(stop_uoffset, blocks[nb - 1].isize);
    h_buoff[k] = u_len;
  }
  uint8_t* dp = d_tab.as<uint8_t>();
  const uint64_t* d_payload = (const uint64_t*)dp;
  const uint64_t* d_outoff = d_payload + nb;
  const uint64_t* d_buoff = d_outoff + nb;
  const uint32_t* d_cdata = (const uint32_t*)(d_buoff + nsb + 2);
  const uint32_t* d_isize = d_cdata + nb;
  // ---- 3. device buffers -------------------------------------------------------------------------------
  const bool resident = r->d_file.p != nullptr;
  if (const int jr = join_prefetch()) return fail(jr == 1 ? BIODB_ERR_IO : BIODB_ERR_CUDA, 0, pf_c0, "cannot read the file");
  if (!resident && nb) {
    if (pf_c1 && pf_c0 == c0) {
      // the start of this batch was prefetched while the previous one was computed: switch buffers, wait for the copy
      // and top up whatever the guess fell short of
      comp_cur ^= 1;
      CUDA_TRY(cudaStreamWaitEvent(st, pf_done, 0));
      if ((size_t)(c1 - c0) + 256 > d_comp2[comp_cur].cap)
        CUDA_TRY(d_comp2[comp_cur].ensure((size_t)(c1 - c0) + 256, st, (size_t)(pf_c1 - c0)));
      if (pf_c1 < c1) {
        cudaError_t he;
        const uint8_t* hp = stage_host(pf_c1, c1, &he);
        if (!hp) return fail(he == cudaErrorUnknown ? BIODB_ERR_IO : BIODB_ERR_CUDA, 0, pf_c1, "cannot read the file");
        CUDA_TRY(cudaMemcpyAsync(d_comp2[comp_cur].as<uint8_t>() + (pf_c1 - c0), hp, (size_t)(c1 - pf_c1),
                                 cudaMemcpyHostToDevice, st));
        if (!r->file) CUDA_TRY(cudaStreamSynchronize(st));      // (a top-up of a streamed file: rare, and its slab is reused next)
        stats.h2d_bytes += c1 - pf_c1;
      }
    } else {
      CUDA_TRY(cudaStreamSynchronize(h2d_st));     // a stale prefetch must not land later
      CUDA_TRY(d_comp2[comp_cur].ensure((size_t)(c1 - c0) + 256, st));
      cudaError_t he;
      const uint8_t* hp = stage_host(c0, c1, &he);
      if (!hp) return fail(he == cudaErrorUnknown ? BIODB_ERR_IO : BIODB_ERR_CUDA, 0, c0, "cannot read the file");
      CUDA_TRY(cudaMemcpyAsync(d_comp2[comp_cur].p, hp, (size_t)(c1 - c0), cudaMemcpyHostToDevice, st));
      stats.h2d_bytes += c1 - c0;
    }
    pf_c1 = 0;
  }
  CUDA_TRY(d_u.ensure((size_t)off + 256, st));
  CUDA_TRY(d_status.ensure((size_t)nb * 8 + 16, st));      // statuses, then (verify_crc) the computed CRCs
  CUDA_TRY(h_status.ensure((size_t)nb * 8 + 16));
  CUDA_TRY(d_ws.ensure(scan_workspace_bytes(nsb), st));
  if (has_carry)
    CUDA_TRY(launch_copy_bytes(d_u.p, d_carry_tail.p, carry_tail_len, st));
  CUDA_TRY(cudaMemcpyAsync(d_tab.p, h_tab.p, tab_bytes, cudaMemcpyHostToDevice, st));
  if (nb) {
    const uint8_t* comp = resident ? r->d_file.as<uint8_t>() + c0 : d_comp2[comp_cur].as<uint8_t>();
    InflateArgs ia{comp, d_payload, d_cdata, d_outoff, d_isize, d_u.as<uint8_t>(), d_status.as<int32_t>(), nb, WalkOut{}, nullptr};
    if (const size_t tb = inflate_token_bytes(nb)) {
      CUDA_TRY(d_tok.ensure(tb, st));
      ia.tok = d_tok.as<uint16_t>();
    }
    if (!raw_mode) {
      // fused record-chain walk: the inflate warps follow the block_size chain of their own block
      ScanWorkspace w0 = carve_scan_workspace(d_ws.p, nsb);
      ia.walk = WalkOut{w0.rel, w0.cnt, w0.ncig, w0.out, w0.in, w0.bad, has_carry ? 1u : 0u,
                        has_carry ? 0u : (uint32_t)(h_buoff[0]), u_len, (int32_t)r->ref_names.size(),
                        (entry_search && !has_carry) ? 1 : 0};
      entry_search = false;
    }
    stage_begin();
    CUDA_TRY(launch_inflate(ia, st));
    stage_end(&stats.inflate_ms);
    stats.inflate_launches += 1;
    stats.n_blocks += nb;
    stats.compressed_bytes += c1 - c0;
    stats.uncompressed_bytes += off - (has_carry ? carry_tail_len : 0);
    if (!last_batch) {
      // the GPU is busy for milliseconds now: walk the next batch's block headers meanwhile
      pre_blocks.clear();
      pre_next_coffset = next_coffset;
      pre_supplier_done = supplier_done;
      pre_pending = pending;
      walk_headers(r, stop_coffset, stop_uoffset, max_blocks, pre_blocks, pre_next_coffset, pre_supplier_done, pre_pending);
      pre_max = max_blocks;
      pre_valid = true;
    }
    if (!resident) {
      CUDA_TRY(cudaEventRecord(comp_free[comp_cur], st));
      // prefetch the next batch's bytes — its blocks are known by now — to the other device buffer on the copy stream
      if (!last_batch && !pre_blocks.empty()) {
        const uint64_t n0 = pre_blocks.front().coffset;
        const uint64_t n1 = pre_blocks.back().coffset + pre_blocks.back().bsize + 1;
        const int nxt = comp_cur ^ 1;
        // the other buffer was last read by the inflate kernel of the previous batch, which precedes this batch's
        // kernels on the compute stream
        CUDA_TRY(cudaStreamWaitEvent(h2d_st, comp_free[nxt], 0));
        CUDA_TRY(d_comp2[nxt].ensure((size_t)(n1 - n0) + (size_t)((n1 - n0) >> 4) + 65536 + 256, h2d_st));
        if (r->file) {
          CUDA_TRY(cudaMemcpyAsync(d_comp2[nxt].p, r->file + n0, (size_t)(n1 - n0), cudaMemcpyHostToDevice, h2d_st));
          CUDA_TRY(cudaEventRecord(pf_done, h2d_st));
        } else {
          // a streamed file: pread into the slab that is not the source of the copy issued last, then the copy — on a
          // thread of its own, so that reading the next batch overlaps everything this batch still has to do (joined
          // at the top of the next call)
          slab_cur ^= 1;
          PinBuf* sl = &h_slab[slab_cur];
          const size_t nbytes = (size_t)(n1 - n0);
          if (nbytes + 64 > sl->cap) CUDA_TRY(sl->ensure(nbytes + nbytes / 8 + 65536));
          void* dst = d_comp2[nxt].p;
          pf_job = std::async(std::launch::async, [this, sl, dst, n0, nbytes]() -> int {
            cudaSetDevice(r->device);
            if (!r->read_bytes(n0, nbytes, sl->p, 8)) return 1;
            if (cudaMemcpyAsync(dst, sl->p, nbytes, cudaMemcpyHostToDevice, h2d_st) != cudaSuccess) return 2;
            return cudaEventRecord(pf_done, h2d_st) == cudaSuccess ? 0 : 2;
          });
        }
        stats.h2d_bytes += n1 - n0;
        pf_c0 = n0;
        pf_c1 = n1;
      }
    }
    const bool check_crc = r->opts.verify_crc != 0;
    if (check_crc) {
      CUDA_TRY(launch_crc32(d_u.as<uint8_t>(), d_outoff, d_isize, nb, d_status.as<uint32_t>() + nb, st));
    }
    CUDA_TRY(launch_copy_bytes(h_status.p, d_status.p, (size_t)nb * (check_crc ? 8 : 4), st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int32_t* hs = h_status.as<int32_t>();
    if (check_crc) {
      // block.d:187 `assert(block.crc32 == crc32(0, uncompressed))` — a mismatch is reported like corrupt data
      const uint32_t* hc = h_status.as<uint32_t>() + nb;
      for (uint32_t i = 0; i < nb; ++i)
        if (hs[i] == 0 && hc[i] != blocks[i].crc32) hs[i] = -1003;
    }
    for (uint32_t i = 0; i < nb; ++i) {
      if (hs[i] == -1003) {
        set_error(&pending, BIODB_ERR_ZLIB, -3, blocks[i].coffset, "CRC32 of the inflated BGZF block does not match its footer");
        u_len = h_outoff[i];
        blocks.resize(i);
        last_batch = true;
        break;
      }
      if (hs[i] != 0) {
        // the blocks before the faulty one are still delivered; the ZlibException surfaces when the
        // iteration reaches it (test/unittests.d:140-142)
        set_error(&pending, BIODB_ERR_ZLIB, hs[i], blocks[i].coffset,
                  std::string("zlib error ") + std::to_string(hs[i]) + (hs[i] == -3 ? " (data error)" : " (buffer error)"));
        u_len = h_outoff[i];
        blocks.resize(i);
        last_batch = true;
        break;
      }
    }
  }
  const uint32_t nb2 = (uint32_t)blocks.size();
  const uint32_t nsb2 = nb2 + (has_carry ? 1 : 0);
  if (nb2 != nb) {
    // truncated slice: rewrite the boundary table tail
    h_buoff[nsb2] = u_len;
    CUDA_TRY(cudaMemcpyAsync((void*)(d_buoff + nsb2), &h_buoff[nsb2], 8, cudaMemcpyHostToDevice, st));
    while (!segs.empty() && segs.back().ustart >= u_len) segs.pop_back();
  }
  final_slice = last_batch;
  end_coffset_last = segs.empty() ? next_coffset : segs.back().cend;
  if (raw_mode) {
    tail = u_len;
    if (last_batch) finished = true;
    return BIODB_OK;
  }
  // ---- 4. record scan (grow the tables and retry when they are too small) ------------------------------
  const int eof_semantics = (last_batch && pending.status == 0) ? 1 : 0;
  uint64_t want_rec = std::max<uint64_t>(rec_capacity, u_len / 96 + 1024);
  uint64_t want_cig = std::max<uint64_t>(cigar_capacity, want_rec * 2);
  // same carving as the one handed to the fused walk (the layout depends on the block count it was carved for)
  ScanWorkspace ws = carve_scan_workspace(d_ws.p, nsb);
  ws_cur = ws;
  ws_carry = has_carry ? 1u : 0u;
  for (int attempt = 0; attempt < 8; ++attempt) {
    if (want_rec != rec_capacity || want_cig != cigar_capacity || rec_front < front_slots) {
      rec_capacity = want_rec;
      cigar_capacity = want_cig;
      rec_front = std::max(rec_front, front_slots);
      static const size_t esz[10] = {8, 4, 4, 4, 4, 4, 4, 4, 8, 4};
      for (int a = 0; a < 9; ++a) CUDA_TRY(d_rec[a].ensure((size_t)(rec_capacity + rec_front + 2) * esz[a], st));
      CUDA_TRY(d_rec[9].ensure((size_t)(cigar_capacity + 2) * 4, st));
    }
    RecordArrays ra = arrays(front_slots);
    stage_begin();
    // only the carried-tail pseudo block still needs the stand-alone walk kernel
    CUDA_TRY(launch_scan_records(d_u.as<uint8_t>(), u_len, d_buoff, nsb2, (has_carry || nb == 0) ? 1u : 0u, eof_semantics, ra,
                                 d_result.as<uint64_t>(), ws, st));
    stage_end(&stats.scan_ms);
    CUDA_TRY(launch_copy_bytes(h_result.p, d_result.p, 32, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    const uint64_t* res = h_result.as<uint64_t>();
    if ((int64_t)res[3] == BIODB_ERR_NOMEM) {
      want_rec = std::max<uint64_t>(res[0] + res[0] / 8 + 1024, rec_capacity);
      want_cig = std::max<uint64_t>(res[2] + res[2] / 8 + 1024, cigar_capacity);
      continue;
    }
    n = res[0];
    tail = res[1];
    n_cigar = res[2];
    if ((int64_t)res[3] == BIODB_ERR_TRUNCATED && pending.status == 0) {
      set_error(&pending, BIODB_ERR_TRUNCATED, 0, end_coffset_last, "not enough data in stream (truncated or malformed BAM record)");
      last_batch = true;
      final_slice = true;
    }
    break;
  }
  // ---- 5. carry the cut record at the end of the slice into the next one ---------------------------------
  carry_tail_len = 0;
  if (!last_batch && tail < u_len) {
    carry_tail_len = u_len - tail;
    CUDA_TRY(d_carry_tail.ensure((size_t)carry_tail_len + 256, st));
    CUDA_TRY(launch_copy_bytes(d_carry_tail.p, d_u.as<uint8_t>() + tail, carry_tail_len, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    for (const Seg& s : segs) {
      uint64_t a = std::max(s.ustart, tail), b = s.ustart + s.len;
      if (a < b) next_segs.push_back(Seg{a - tail, s.coffset, (uint32_t)(s.within + (a - s.ustart)), (uint32_t)(b - a), s.cend});
    }
  }
  n_records_total += n;
  if (last_batch) finished = true;
  return BIODB_OK;
}

}  // namespace biodb

int biodb_reader::header_at(uint64_t pos, biodb::BlockInfo* b, biodb_error* e) {
  if (file) return parse_bgzf_header(file, flen, pos, b, e, 0);
  if (hdr_map) return parse_bgzf_header(hdr_map, flen, pos, b, e, 0);
  // streamed, no mapping: a BGZF member is at most 64 KiB of payload behind a header of at most 64 KiB of extra fields, so a window
  // that reaches WIN_NEED bytes past pos (or the end of the file) shows parse_bgzf_header everything it may look at
  constexpr uint64_t WIN_NEED = 192 * 1024, WIN_SIZE = 8ull << 20;
  std::lock_guard<std::mutex> lk(hdr_mu);
  if (pos >= flen) return 0;
  const uint64_t need_end = std::min(flen, pos + WIN_NEED);
  if (pos < hdr_off || need_end > hdr_off + hdr_win.size()) {
    const uint64_t n = std::min(flen - pos, WIN_SIZE);
    hdr_win.resize((size_t)n);
    hdr_off = pos;
    if (!read_bytes(pos, (size_t)n, hdr_win.data())) {
      set_error(e, BIODB_ERR_IO, 0, pos, "read error at file offset " + std::to_string(pos));
      hdr_win.clear();
      return (int)BIODB_ERR_IO;
    }
  }
  return parse_bgzf_header(hdr_win.data(), hdr_win.size(), pos - hdr_off, b, e, hdr_off);
}

bool biodb_reader::read_bytes(uint64_t off, size_t len, void* dst, int threads) const {
  if (off > flen || len > flen - off) return false;
  if (file) { memcpy(dst, file + off, len); return true; }
  auto part = [&](uint64_t o, size_t n, uint8_t* d) -> bool {
    while (n) {
      const ssize_t got = pread(fd, d, n, (off_t)o);
      if (got <= 0) return false;
      o += (uint64_t)got; d += got; n -= (size_t)got;
    }
    return true;
  };
  if (threads <= 1 || len < (8u << 20)) return part(off, len, (uint8_t*)dst);
  std::vector<std::thread> th;
  std::vector<char> ok((size_t)threads, 1);
  const size_t per = (len + threads - 1) / threads;
  for (int t = 0; t < threads; ++t) {
    const size_t a = std::min(len, (size_t)t * per), z = std::min(len, a + per);
    th.emplace_back([&, t, a, z] { ok[(size_t)t] = part(off + a, z - a, (uint8_t*)dst + a) ? 1 : 0; });
  }
  for (auto& x : th) x.join();
  for (char c : ok) if (!c) return false;
  return true;
}

void biodb_reader::ensure_pinned(uint64_t lo, uint64_t hi) {
  if (opts.pin_input != 2 || !file || lo >= hi) return;
  const uint64_t page = 4096;
  const uint64_t base = (uint64_t)(uintptr_t)file;
  // page-align in address space, clip to the buffer's own pages
  uint64_t a = ((base + lo) & ~(page - 1)), b = ((base + std::min(hi, flen) + page - 1) & ~(page - 1));
  const uint64_t first = base & ~(page - 1), last = (base + flen + page - 1) & ~(page - 1);
  a = std::max(a, first);
  b = std::min(b, last);
  std::lock_guard<std::mutex> lk(pool_mu);
  std::vector<std::pair<uint64_t, uint64_t>> gaps;
  uint64_t cur = a;
  for (const auto& r : pinned) {
    if (r.second <= cur) continue;
    if (r.first >= b) break;
    if (r.first > cur) gaps.push_back({cur, r.first});
    cur = std::max(cur, r.second);
  }
  if (cur < b) gaps.push_back({cur, b});
  for (const auto& g : gaps) {
    void* p = (void*)(uintptr_t)g.first;
    const size_t n = (size_t)(g.second - g.first);
    bool ok = cudaHostRegister(p, n, cudaHostRegisterReadOnly) == cudaSuccess ||
              cudaHostRegister(p, n, cudaHostRegisterDefault) == cudaSuccess;
    cudaGetLastError();
    if (ok) pinned.push_back(g);          // (a range that cannot be locked is simply copied from pageable memory)
  }
  std::sort(pinned.begin(), pinned.end());
}

biodb_status biodb_reader::build_block_index() {
  if (!block_index.empty()) return BIODB_OK;
  uint64_t pos = reads_start_coffset;
  biodb_error e{};
  while (true) {
    BlockInfo b;
    int rc = header_at(pos, &b, &e);
    if (rc < 0) { err = e; return (biodb_status)e.status; }
    if (rc == 0 || b.isize == 0) break;
    block_index.push_back(pos);
    pos = b.coffset + b.bsize + 1;
  }
  data_end_coffset = pos;
  return BIODB_OK;
}

// =========================================================================================== C ABI ====

static thread_local biodb_error g_open_error;

extern "C" {

const char* biodb_version(void) { return "biod_b200 0.1 (sm_100a)"; }

void biodb_default_options(biodb_options* o) {
  memset(o, 0, sizeof *o);
  o->device = -1;
  o->blocks_per_batch = 0;      // = three full waves of the inflate kernel on the device (8436 on a B200)
}

const biodb_error* biodb_open_error(void) { return &g_open_error; }
const biodb_error* biodb_last_error(const biodb_reader* r) { return r ? &r->err : &g_open_error; }

static biodb_status open_common(biodb_reader* r) {
  // BamReader constructor (bam/reader.d:100-124): magic, header text, reference table — the blocks
  // that hold them are inflated on the device, then parsed here.
  Pass p;
  p.raw_mode = true;
  biodb_status s = p.init(r, 0, 0);
  if (s != BIODB_OK) return s;
  std::vector<uint8_t> u;
  biodb::PinBuf host;
  bool done = false;
  auto fetch = [&]() -> biodb_status {   // append the next few blocks to u
    biodb_status st = p.next(4, 0);
    if (st != BIODB_OK) { done = true; return st; }
    size_t fresh = (size_t)p.u_len;   // raw_mode: the slice is the plain byte stream
    size_t carry = 0;
    if (fresh) {
      if (host.ensure(fresh) != cudaSuccess) return BIODB_ERR_CUDA;
      if (cudaMemcpy(host.p, p.d_u.p, fresh, cudaMemcpyDeviceToHost) != cudaSuccess) return BIODB_ERR_CUDA;
      u.insert(u.end(), host.as<uint8_t>() + carry, host.as<uint8_t>() + fresh);
    }
    if (p.finished) done = true;
    return BIODB_OK;
  };
  // remember where each block's bytes start so the first record's virtual offset can be computed
  std::vector<BlockInfo> seen;
  std::vector<uint64_t> ustart;
  biodb_status deferred = BIODB_OK;
  auto need = [&](size_t upto) -> biodb_status {
    while (u.size() < upto && !done) {
      size_t before = u.size();
      biodb_status st = fetch();
      if (st != BIODB_OK && st != BIODB_EOF) { deferred = st; break; }
      uint64_t o = before;
      for (const BlockInfo& b : p.blocks) { seen.push_back(b); ustart.push_back(o); o += b.isize; }
      if (st == BIODB_EOF) break;
    }
    if (u.size() >= upto) return BIODB_OK;
    if (deferred != BIODB_OK) return deferred;         // the error lies inside the header: raise it now
    set_error(&r->err, BIODB_ERR_TRUNCATED, 0, 0, "not enough data in stream");
    return BIODB_ERR_TRUNCATED;
  };
  need(4);
  if (u.size() < 4 || memcmp(u.data(), "BAM\1", 4) != 0) {
    if (deferred != BIODB_OK && u.size() < 4) return deferred;
    set_error(&r->err, BIODB_ERR_FORMAT, 0, 0, "Invalid file format: expected BAM\\1");   // reader.d:111-113
    return BIODB_ERR_FORMAT;
  }
  size_t q = 4;
  if ((s = need(q + 4)) != BIODB_OK) return s;
  int32_t l_text = (int32_t)le32(u.data() + q);
  q += 4;
  if (l_text < 0) { set_error(&r->err, BIODB_ERR_FORMAT, 0, 0, "negative l_text"); return BIODB_ERR_FORMAT; }
  if ((s = need(q + (size_t)l_text + 4)) != BIODB_OK) return s;
  r->text.assign((const char*)u.data() + q, (size_t)l_text);
  q += (size_t)l_text;
  int32_t n_ref = (int32_t)le32(u.data() + q);
  q += 4;
  for (int32_t i = 0; i < n_ref; ++i) {
    if ((s = need(q + 4)) != BIODB_OK) return s;
    int32_t l_name = (int32_t)le32(u.data() + q);      // referenceinfo.d:57-62
    q += 4;
    if (l_name < 0) { set_error(&r->err, BIODB_ERR_FORMAT, 0, 0, "negative l_name"); return BIODB_ERR_FORMAT; }
    if ((s = need(q + (size_t)l_name + 4)) != BIODB_OK) return s;
    std::string nm((const char*)u.data() + q, (size_t)l_name);
    if (!nm.empty() && nm.back() == '\0') nm.pop_back();
    r->ref_names.push_back(nm);
    q += (size_t)l_name;
    r->ref_lens.push_back((int32_t)le32(u.data() + q));
    q += 4;
  }
  // virtualTell() right after the header (reader.d:121-123)
  need(q + 1);   // make the block that holds the first record known, if there is one
  size_t bi = 0;
  while (bi + 1 < seen.size() && ustart[bi + 1] <= q) ++bi;
  if (seen.empty()) return BIODB_ERR_FORMAT;
  if (q - ustart[bi] >= seen[bi].isize) {
    r->reads_start_coffset = seen[bi].coffset + seen[bi].bsize + 1;
    r->reads_start_uoffset = 0;
  } else {
    r->reads_start_coffset = seen[bi].coffset;
    r->reads_start_uoffset = (uint32_t)(q - ustart[bi]);
  }
  r->reads_start_vo = (r->reads_start_coffset << 16) | r->reads_start_uoffset;
  memset(&r->err, 0, sizeof r->err);   // errors of later blocks surface during iteration (unittests.d:139)
  return BIODB_OK;
}

// a reader that did not make it through finish_open: what biodb_close would release
static void drop_reader(biodb_reader* r) {
  if (r->hdr_map) munmap((void*)r->hdr_map, (size_t)r->flen);
  if (r->fd >= 0) close(r->fd);
  if (r->registered) cudaHostUnregister((void*)r->file);
  delete r;
}

static biodb_status finish_open(biodb_reader* r, const biodb_options* opts, biodb_reader** out) {
  if (opts) r->opts = *opts; else biodb_default_options(&r->opts);
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
    set_error(&g_open_error, BIODB_ERR_CUDA, 0, 0, "no CUDA device: biod_b200 has no CPU fallback");
    drop_reader(r);
    return BIODB_ERR_CUDA;
  }
  if (r->opts.device < 0) { if (cudaGetDevice(&r->device) != cudaSuccess) r->device = 0; }
  else r->device = r->opts.device;
  if (r->opts.blocks_per_batch <= 0) {
    // one CTA per BGZF block: a batch of a whole number of waves leaves no SM idle behind a partial last wave.
    // 0 = three waves; -k = k waves (short passes — one shard of many — pipeline better in smaller batches)
    const int waves = r->opts.blocks_per_batch < 0 ? -r->opts.blocks_per_batch : 3;
    int slots = inflate_resident_blocks(r->device);
    r->opts.blocks_per_batch = slots > 0 ? waves * slots : 8192;
  }
  biodb_status s = open_common(r);
  if (s != BIODB_OK) {
    g_open_error = r->err;
    if (!g_open_error.status) set_error(&g_open_error, s, 0, 0, "open failed");
    drop_reader(r);
    return s;
  }
  if (r->opts.resident_input && r->flen) {
    bool ok = r->d_file.ensure((size_t)r->flen + 256) == cudaSuccess;
    if (ok && r->file) ok = cudaMemcpy(r->d_file.p, r->file, (size_t)r->flen, cudaMemcpyHostToDevice) == cudaSuccess;
    if (ok && !r->file) {                       // streamed file: through one pinned slab
      biodb::PinBuf slab;
      const size_t step = 256u << 20;
      ok = slab.ensure(std::min<size_t>(step, (size_t)r->flen) + 64) == cudaSuccess;
      for (uint64_t o = 0; ok && o < r->flen; o += step) {
        const size_t n = (size_t)std::min<uint64_t>(step, r->flen - o);
        ok = r->read_bytes(o, n, slab.p, 4) &&
             cudaMemcpy(r->d_file.as<uint8_t>() + o, slab.p, n, cudaMemcpyHostToDevice) == cudaSuccess;
      }
    }
    if (!ok) {
      set_error(&g_open_error, BIODB_ERR_CUDA, 0, 0, "cannot make the compressed file resident in device memory");
      drop_reader(r);
      return BIODB_ERR_CUDA;
    }
  }
  *out = r;
  return BIODB_OK;
}

biodb_status biodb_open_memory(const void* data, size_t len, const biodb_options* opts, biodb_reader** out) {
  if (!data || !out) return BIODB_ERR_ARG;
  biodb_reader* r = new biodb_reader;
  r->file = (const uint8_t*)data;
  r->flen = len;
  if (opts && opts->pin_input == 1 && len)
    r->registered = cudaHostRegister((void*)data, len, cudaHostRegisterReadOnly) == cudaSuccess ||
                    cudaHostRegister((void*)data, len, cudaHostRegisterDefault) == cudaSuccess;
  cudaGetLastError();
  return finish_open(r, opts, out);
}

biodb_status biodb_open(const char* path, const biodb_options* opts, biodb_reader** out) {
  if (!path || !out) return BIODB_ERR_ARG;
  const int fd = open(path, O_RDONLY | O_CLOEXEC);
  struct stat sb;
  if (fd < 0 || fstat(fd, &sb) != 0 || !S_ISREG(sb.st_mode)) {
    if (fd >= 0) close(fd);
    set_error(&g_open_error, BIODB_ERR_IO, 0, 0, std::string("cannot open ") + path);
    return BIODB_ERR_IO;
  }
  // the file is streamed (runtime.h): nothing but the header blocks is read here, whatever its size (off_t is 64 bits)
  biodb_reader* r = new biodb_reader;
  r->fd = fd;
  r->file = nullptr;
  r->flen = (uint64_t)sb.st_size;
  if (r->flen) {
    // walking the BSIZE chain through a window would copy the whole file a second time: look at the block headers
    // through a mapping instead (a few bytes per block; the bulk of the file still goes pread -> pinned slab -> device)
    void* m = mmap(nullptr, (size_t)r->flen, PROT_READ, MAP_SHARED, fd, 0);
    if (m != MAP_FAILED) r->hdr_map = (const uint8_t*)m;
  }
  return finish_open(r, opts, out);
}

void biodb_pileup_destroy_pooled(void* p);
static void reads_destroy_pooled(void* p);

void biodb_close(biodb_reader* r) {
  if (!r) return;
  for (void* p : r->pileup_pool) biodb_pileup_destroy_pooled(p);
  for (void* p : r->reads_pool) reads_destroy_pooled(p);
  if (r->registered) cudaHostUnregister((void*)r->file);
  for (const auto& g : r->pinned) cudaHostUnregister((void*)(uintptr_t)g.first);
  if (r->hdr_map) munmap((void*)r->hdr_map, (size_t)r->flen);
  if (r->fd >= 0) close(r->fd);
  delete r;
}

biodb_status biodb_header_text(const biodb_reader* r, const char** text, size_t* len) {
  if (!r || !text || !len) return BIODB_ERR_ARG;
  *text = r->text.data();
  *len = r->text.size();
  return BIODB_OK;
}
int32_t biodb_n_refs(const biodb_reader* r) { return r ? (int32_t)r->ref_names.size() : 0; }
biodb_status biodb_ref_info(const biodb_reader* r, int32_t i, const char** name, int32_t* name_len, int32_t* length) {
  if (!r || i < 0 || (size_t)i >= r->ref_names.size()) return BIODB_ERR_ARG;
  if (name) *name = r->ref_names[i].c_str();
  if (name_len) *name_len = (int32_t)r->ref_names[i].size();
  if (length) *length = r->ref_lens[i];
  return BIODB_OK;
}
uint64_t biodb_reads_start_voffset(const biodb_reader* r) { return r ? r->reads_start_vo : 0; }
int32_t biodb_input_is_pinned(const biodb_reader* r) { return r && (r->registered || !r->pinned.empty()) ? 1 : 0; }

uint64_t biodb_file_size(const biodb_reader* r) { return r ? r->flen : 0; }

}  // extern "C"

// ------------------------------------------------------------------------------------- reads ----

// BamReader.reads iterator.  Batches are read one ahead: while the caller works on batch k (host slot A), batch k+1 is
// computed, copied device->device into a staging area (so that the pass may reuse its buffers for batch k+2 at once)
// and from there to host slot B on a copy stream — the PCIe copy of one batch overlaps the kernels of the next.
struct ReadsSlot {
  PinBuf h_data, h_arr[10], h_vo[2];
  cudaEvent_t done = nullptr;        // the device->host copies of this batch have landed
  uint64_t n = 0, used = 0, n_cigar = 0, first = 0;
  std::vector<Seg> segs;             // block provenance of the slice, for virtual offsets (readrange.d:55-66)
  // region mode: where the chunk of this batch ends and the next one begins (0 = there is none)
  uint64_t chunk_end_vo = 0, next_chunk_beg_vo = 0;
};
struct biodb_index {
  BaiIndex bai;
};
struct biodb_index_builder {
  BaiBuilder b;
  biodb_status failed = BIODB_OK;
  bool finished = false;
};

struct biodb_reads {
  Pass pass;
  // region mode (biodb_reads_begin_region): the chunks the index names are read one after the other, each as a pass of
  // its own from chunk.beg to chunk.end, and every batch goes through the region filter (region.cu)
  bool region = false;                 // chunk mode: the pass reads it->chunks one after the other
  bool filter = false;                 // ... and every batch goes through the region filter
  uint32_t chunk_blocks = 0;           // blocks per batch in chunk mode (0 = the reader's option)
  bool region_done = false;            // a read beyond the region was met: nothing further can overlap it
  std::vector<VoChunk> chunks;
  size_t chunk_i = 0;                  // chunk being read
  uint32_t reg_ref = 0, reg_beg = 0, reg_end = 0;
  std::vector<uint32_t> regs;          // several regions (biodb_reads_begin_regions): (begin, end) pairs, also on the device
  DevBuf d_regs;
  DevBuf d_sel[10], d_scratch, d_info;
  PinBuf h_info;
  uint64_t sel_cap = 0, sel_cig_cap = 0;
  uint64_t sel_n = 0, sel_cig = 0;
  ReadsSlot slot[2];
  int cur = 0;                       // slot the caller holds
  bool inflight = false;             // slot[cur ^ 1] holds the batch read ahead
  bool ahead_done = false;           // the read-ahead hit EOF / an error: ahead_status is returned after the batch in flight
  biodb_status ahead_status = BIODB_OK;
  DevBuf d_stage[11];                // device staging: data + the ten record arrays
  cudaStream_t cs = nullptr;
  cudaEvent_t computed = nullptr, staged = nullptr;
  bool staged_valid = false;
  ~biodb_reads() {
    if (cs) { cudaStreamSynchronize(cs); cudaStreamDestroy(cs); }
    for (cudaEvent_t e : {computed, staged, slot[0].done, slot[1].done})
      if (e) cudaEventDestroy(e);
  }
  void reset() {
    if (cs) cudaStreamSynchronize(cs);
    inflight = ahead_done = staged_valid = false;
    ahead_status = BIODB_OK;
    region = region_done = filter = false;
    chunk_blocks = 0;
    chunks.clear();
    chunk_i = 0;
    regs.clear();
  }
};

// Point the pass at chunk k of a region read: from chunk.beg to chunk.end (virtual offsets).
static void region_seek(biodb_reads* it, size_t k) {
  const VoChunk& c = it->chunks[k];
  it->pass.rewind(c.beg >> 16, (uint32_t)(c.beg & 0xFFFF));
  it->pass.stop_coffset = c.end >> 16;
  it->pass.stop_uoffset = (uint32_t)(c.end & 0xFFFF);
}

// Region mode: the next batch that holds at least one read of the region, filtered and compacted into it->d_sel.
static biodb_status region_next(biodb_reads* it) {
  Pass& p = it->pass;
  cudaStream_t st = p.st;
  while (true) {
    if (it->region_done || it->chunk_i >= it->chunks.size()) return BIODB_EOF;
    biodb_status s = p.next(it->chunk_blocks ? it->chunk_blocks : (uint32_t)p.r->opts.blocks_per_batch, 0);
    if (s == BIODB_EOF) {
      if (++it->chunk_i >= it->chunks.size()) return BIODB_EOF;
      region_seek(it, it->chunk_i);
      continue;
    }
    if (s != BIODB_OK) return s;
    if (p.n == 0) continue;
    if (!it->filter) return BIODB_OK;                     // getReadsBetween: every record of the stretch
    if (p.n > 0xfffffff0ull) return p.fail(BIODB_ERR_NOMEM, 0, 0, "too many records in one batch; lower blocks_per_batch");
    // room for the compacted tables
    if (p.n + 8 > it->sel_cap || p.n_cigar + 8 > it->sel_cig_cap) {
      it->sel_cap = std::max<uint64_t>(it->sel_cap, p.n + p.n / 8 + 1024);
      it->sel_cig_cap = std::max<uint64_t>(it->sel_cig_cap, p.n_cigar + p.n_cigar / 8 + 1024);
      static const size_t esz[10] = {8, 4, 4, 4, 4, 4, 4, 4, 8, 4};
      for (int a = 0; a < 9; ++a)
        if (it->d_sel[a].ensure((size_t)(it->sel_cap + 2) * esz[a], st) != cudaSuccess) return p.fail(BIODB_ERR_CUDA, 0, 0, "allocation failed");
      if (it->d_sel[9].ensure((size_t)(it->sel_cig_cap + 2) * 4, st) != cudaSuccess) return p.fail(BIODB_ERR_CUDA, 0, 0, "allocation failed");
    }
    if (it->d_scratch.ensure(region_scratch_elems(p.n) * 4, st) != cudaSuccess || it->d_info.ensure(64, st) != cudaSuccess ||
        it->h_info.ensure(64) != cudaSuccess)
      return p.fail(BIODB_ERR_CUDA, 0, 0, "allocation failed");
    RecordArrays in = p.arrays(0);
    RecordArrays out{it->d_sel[0].as<uint64_t>(), it->d_sel[1].as<int32_t>(), it->d_sel[2].as<int32_t>(), it->d_sel[3].as<int32_t>(),
                     it->d_sel[4].as<int32_t>(), it->d_sel[5].as<uint32_t>(), it->d_sel[6].as<uint32_t>(), it->d_sel[7].as<int32_t>(),
                     it->d_sel[8].as<uint64_t>(), it->d_sel[9].as<uint32_t>(), it->sel_cap, it->sel_cig_cap};
    p.stage_begin();
    cudaError_t e = launch_region_filter(in, p.n, it->reg_ref, it->reg_beg, it->reg_end, out, it->d_scratch.as<uint32_t>(),
                                         it->d_info.as<uint32_t>(), st, it->regs.size() > 2 ? it->d_regs.as<uint32_t>() : nullptr,
                                         (uint32_t)(it->regs.size() / 2));
    p.stage_end(&p.stats.scan_ms);
    if (e != cudaSuccess || launch_copy_bytes(it->h_info.p, it->d_info.p, 12, st) != cudaSuccess ||
        cudaStreamSynchronize(st) != cudaSuccess)
      return p.fail(BIODB_ERR_CUDA, 0, 0, "region filter failed");
    const uint32_t* hi = it->h_info.as<uint32_t>();
    if (hi[0] != 0xffffffffu) it->region_done = true;
    it->sel_n = hi[1];
    it->sel_cig = hi[2];
    if (it->sel_n) return BIODB_OK;
  }
}

static const size_t REC_ESZ[10] = {8, 4, 4, 4, 4, 4, 4, 4, 8, 4};

// Compute the next batch and start its copies into host slot s.  BIODB_OK, BIODB_EOF or the error met.
static biodb_status reads_produce(biodb_reads* it, int s) {
  Pass& p = it->pass;
  if (!it->cs) {
    if (cudaStreamCreateWithFlags(&it->cs, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&it->computed, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&it->staged, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&it->slot[0].done, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&it->slot[1].done, cudaEventDisableTiming) != cudaSuccess)
      return p.fail(BIODB_ERR_CUDA, 0, 0, "stream / event creation failed");
  }
  // the staging copy of the previous batch still reads the pass's buffers
  if (it->staged_valid && cudaStreamWaitEvent(p.st, it->staged, 0) != cudaSuccess) return p.fail(BIODB_ERR_CUDA, 0, 0, "wait failed");
  const uint64_t first = p.n_records_total;
  biodb_status st;
  if (it->region) {
    st = region_next(it);
    if (st != BIODB_OK) return st;
  } else {
    do {
      st = p.next((uint32_t)p.r->opts.blocks_per_batch, 0);
    } while (st == BIODB_OK && p.n == 0 && !p.finished);   // a slice may hold only part of one huge record
    if (st != BIODB_OK) return st;
    if (p.n == 0) return p.next(1, 0);                      // finished: raises the pending error or EOF
  }
  ReadsSlot& sl = it->slot[s];
  sl.n = it->filter ? it->sel_n : p.n;
  sl.used = p.tail;                                       // bytes of the slice covered by whole records
  sl.n_cigar = it->filter ? it->sel_cig : p.n_cigar;
  sl.first = it->region ? 0 : first;
  sl.segs = p.segs;
  sl.chunk_end_vo = sl.next_chunk_beg_vo = 0;
  if (it->region && it->chunk_i < it->chunks.size()) {
    sl.chunk_end_vo = it->chunks[it->chunk_i].end;
    if (it->chunk_i + 1 < it->chunks.size()) sl.next_chunk_beg_vo = it->chunks[it->chunk_i + 1].beg;
  }
  cudaStream_t cs = it->cs;
  bool ok = cudaEventRecord(it->computed, p.st) == cudaSuccess && cudaStreamWaitEvent(cs, it->computed, 0) == cudaSuccess;
  size_t bytes[11];
  const void* src[11];
  bytes[0] = (size_t)sl.used;
  src[0] = p.d_u.p;
  for (int a = 0; a < 9; ++a) {
    bytes[1 + a] = (size_t)(sl.n + (a == 8 ? 1 : 0)) * REC_ESZ[a];
    src[1 + a] = it->filter ? it->d_sel[a].p : p.d_rec[a].p;
  }
  bytes[10] = (size_t)sl.n_cigar * 4;
  src[10] = it->filter ? it->d_sel[9].p : p.d_rec[9].p;
  // device -> staging (ordered behind the previous batch's staging -> host copies on the same stream)
  for (int k = 0; k < 11 && ok; ++k) {
    // (grow with headroom: the two host slots and the staging area see batches of slightly different sizes, and a
    //  re-allocation of half a gigabyte of pinned memory costs hundreds of milliseconds)
    ok = (bytes[k] + 16 <= it->d_stage[k].cap || it->d_stage[k].ensure(bytes[k] + bytes[k] / 8 + 4096, cs) == cudaSuccess) &&
         (bytes[k] == 0 || cudaMemcpyAsync(it->d_stage[k].p, src[k], bytes[k], cudaMemcpyDeviceToDevice, cs) == cudaSuccess);
  }
  ok = ok && cudaEventRecord(it->staged, cs) == cudaSuccess;
  it->staged_valid = true;
  // staging -> host slot
  for (int k = 0; k < 11 && ok; ++k) {
    PinBuf& h = k == 0 ? sl.h_data : sl.h_arr[k - 1];
    ok = (bytes[k] + 16 <= h.cap || h.ensure(bytes[k] + bytes[k] / 8 + 4096) == cudaSuccess) &&
         (bytes[k] == 0 || cudaMemcpyAsync(h.p, it->d_stage[k].p, bytes[k], cudaMemcpyDeviceToHost, cs) == cudaSuccess);
  }
  ok = ok && cudaEventRecord(sl.done, cs) == cudaSuccess;
  if (!ok) return p.fail(BIODB_ERR_CUDA, 0, 0, "device to host copy failed");
  p.stats.d2h_bytes += sl.used + sl.n * 44 + 8 + sl.n_cigar * 4;
  return BIODB_OK;
}

extern "C" {

static void reads_destroy_pooled(void* p) { delete (biodb_reads*)p; }

biodb_status biodb_reads_begin(biodb_reader* r, biodb_reads** out) {
  if (!r || !out) return BIODB_ERR_ARG;
  {
    std::lock_guard<std::mutex> lk(r->pool_mu);
    if (!r->reads_pool.empty()) {
      biodb_reads* it = (biodb_reads*)r->reads_pool.back();
      r->reads_pool.pop_back();
      it->reset();
      it->pass.rewind(r->reads_start_coffset, r->reads_start_uoffset);
      *out = it;
      return BIODB_OK;
    }
  }
  biodb_reads* it = new biodb_reads;
  biodb_status s = it->pass.init(r, r->reads_start_coffset, r->reads_start_uoffset);
  if (s != BIODB_OK) { delete it; return s; }
  *out = it;
  return BIODB_OK;
}

// ---- IndexBuilder (bai/indexing.d) behind the C ABI: host only ------------------------------------------------------
biodb_status biodb_index_builder_begin(int32_t n_refs, int32_t check_bins, biodb_index_builder** out) {
  if (!out || n_refs < 0) return BIODB_ERR_ARG;
  biodb_index_builder* b = new biodb_index_builder;
  b->b.begin(n_refs, check_bins != 0);
  *out = b;
  return BIODB_OK;
}
biodb_status biodb_index_builder_put(biodb_index_builder* b, uint64_t n, const int32_t* ref_id, const int32_t* pos,
                                     const int32_t* end_pos, const uint32_t* bin_mq_nl, const uint32_t* flag_nc,
                                     const uint64_t* start_voffset, const uint64_t* end_voffset) {
  if (!b || (n && (!ref_id || !pos || !end_pos || !bin_mq_nl || !flag_nc || !start_voffset || !end_voffset))) return BIODB_ERR_ARG;
  if (b->failed) return b->failed;
  for (uint64_t i = 0; i < n; ++i) {
    const bool unm = ((flag_nc[i] >> 16) & 4) != 0;
    if (!b->b.put(ref_id[i], pos[i], end_pos[i], bin_mq_nl[i] >> 16, unm, start_voffset[i], end_voffset[i])) {
      b->failed = b->b.err.rfind("BAM file is not", 0) == 0 ? BIODB_ERR_UNSORTED : BIODB_ERR_FORMAT;
      return b->failed;
    }
  }
  return BIODB_OK;
}
biodb_status biodb_index_builder_finish(biodb_index_builder* b, const uint8_t** data, size_t* len) {
  if (!b || !data || !len) return BIODB_ERR_ARG;
  if (b->failed) return b->failed;
  if (!b->finished) { b->b.finish(); b->finished = true; }
  *data = b->b.out.data();
  *len = b->b.out.size();
  return BIODB_OK;
}
const char* biodb_index_builder_error(const biodb_index_builder* b) { return b ? b->b.err.c_str() : ""; }
void biodb_index_builder_end(biodb_index_builder* b) { delete b; }

biodb_status biodb_index_open(const void* bai, size_t len, biodb_index** out) {
  if (!bai || !out) return BIODB_ERR_ARG;
  biodb_index* ix = new biodb_index;
  std::string msg;
  const int rc = ix->bai.parse((const uint8_t*)bai, len, &msg);
  if (rc != 0) {
    delete ix;
    set_error(&g_open_error, rc == -3 ? BIODB_ERR_FORMAT : BIODB_ERR_TRUNCATED, 0, 0, msg);
    return (biodb_status)g_open_error.status;
  }
  *out = ix;
  return BIODB_OK;
}
void biodb_index_close(biodb_index* ix) { delete ix; }
int32_t biodb_index_n_refs(const biodb_index* ix) { return ix ? (int32_t)ix->bai.refs.size() : 0; }
int32_t biodb_index_last_linear_offset(const biodb_index* ix, int32_t n_refs, uint64_t* out) {
  // reader.d:380-383: the last entry of the last non-empty linear index among references [0, n_refs)
  if (!ix || !out) return 0;
  const int32_t n = std::min<int32_t>(n_refs, (int32_t)ix->bai.refs.size());
  for (int32_t r = n - 1; r >= 0; --r)
    if (!ix->bai.refs[(size_t)r].ioffsets.empty()) { *out = ix->bai.refs[(size_t)r].ioffsets.back(); return 1; }
  return 0;
}
int64_t biodb_index_chunks(const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end, uint64_t* out2, uint64_t cap) {
  if (!ix) return -1;
  std::vector<VoChunk> c;
  if (!ix->bai.region_chunks(ref_id, beg, end, &c)) return -1;
  for (uint64_t k = 0; k < c.size() && k < cap && out2; ++k) { out2[2 * k] = c[k].beg; out2[2 * k + 1] = c[k].end; }
  return (int64_t)c.size();
}

}  // extern "C"

// The stretches of the file a region read goes through: getChunks, then StreamChunksSupplier.moveToNextChunk
// (inputstream.d:262-277): chunks that begin in the same BGZF block are read as one stretch, from the first one's
// start to the last one's end.
biodb_status biodb_index_region_chunks(const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end,
                                       std::vector<biodb::VoChunk>* out) {
  std::vector<VoChunk> c;
  out->clear();
  if (!ix || !ix->bai.region_chunks(ref_id, beg, end, &c)) return BIODB_ERR_ARG;
  for (size_t k = 0; k < c.size();) {
    size_t i = k + 1;
    while (i < c.size() && (c[i].beg >> 16) <= (c[k].beg >> 16)) ++i;
    out->push_back(VoChunk{c[k].beg, c[i - 1].end});
    k = i;
  }
  return BIODB_OK;
}

extern "C" {

biodb_status biodb_reads_begin_region(biodb_reader* r, const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end,
                                      biodb_reads** out) {
  if (!r || !ix || !out) return BIODB_ERR_ARG;
  std::vector<VoChunk> c;
  if (beg >= end) {                                        // reference.d:77
    set_error(&r->err, BIODB_ERR_ARG, 0, 0, "start must be less than end");
    return BIODB_ERR_ARG;
  }
  if (biodb_index_region_chunks(ix, ref_id, beg, end, &c) != BIODB_OK) {      // randomaccessmanager.d:206-208
    set_error(&r->err, BIODB_ERR_ARG, 0, 0, "Invalid reference sequence index");
    return BIODB_ERR_ARG;
  }
  biodb_status s = biodb_reads_begin(r, out);
  if (s != BIODB_OK) return s;
  biodb_reads* it = *out;
  it->region = true;
  it->filter = true;
  it->region_done = false;
  it->chunks.swap(c);
  it->chunk_i = 0;
  it->reg_ref = ref_id;
  it->reg_beg = beg;
  it->reg_end = end;
  if (!it->chunks.empty()) region_seek(it, 0);
  return BIODB_OK;
}

// getReads(BamRegion[]) for the regions of ONE reference (randomaccessmanager.d:286-296 per group of :316-337; the
// caller joins the groups of several references one after the other, as the reference does): the regions are sorted,
// overlapping ones joined (nonOverlapping, algo.d:95-162: prev.end >= next.begin), the chunks of all of them read as
// one stream and filtered by the multi-region BamReadFilter.
biodb_status biodb_reads_begin_regions(biodb_reader* r, const biodb_index* ix, uint32_t ref_id, uint32_t n, const uint32_t* begs,
                                       const uint32_t* ends, biodb_reads** out) {
  if (!r || !ix || !out || !n || !begs || !ends) return BIODB_ERR_ARG;
  std::vector<std::pair<uint32_t, uint32_t>> rg(n), merged;
  for (uint32_t k = 0; k < n; ++k) {
    if (begs[k] >= ends[k]) {                              // enforce(beg < end), randomaccessmanager.d:256
      set_error(&r->err, BIODB_ERR_ARG, 0, 0, "start must be less than end");
      return BIODB_ERR_ARG;
    }
    rg[k] = {begs[k], ends[k]};
  }
  std::sort(rg.begin(), rg.end());
  for (const auto& x : rg) {
    if (!merged.empty() && merged.back().second >= x.first) merged.back().second = std::max(merged.back().second, x.second);
    else merged.push_back(x);
  }
  std::vector<VoChunk> c, cc;
  if (!ix->bai.regions_chunks(ref_id, merged, &c)) {
    set_error(&r->err, BIODB_ERR_ARG, 0, 0, "Invalid reference sequence index");
    return BIODB_ERR_ARG;
  }
  for (size_t k = 0; k < c.size();) {                      // moveToNextChunk: chunks that begin in the same block are one stretch
    size_t i = k + 1;
    while (i < c.size() && (c[i].beg >> 16) <= (c[k].beg >> 16)) ++i;
    cc.push_back(VoChunk{c[k].beg, c[i - 1].end});
    k = i;
  }
  biodb_status s = biodb_reads_begin(r, out);
  if (s != BIODB_OK) return s;
  biodb_reads* it = *out;
  it->region = true;
  it->filter = true;
  it->region_done = false;
  it->chunks.swap(cc);
  it->chunk_i = 0;
  it->reg_ref = ref_id;
  it->reg_beg = merged.front().first;
  it->reg_end = merged.back().second;
  it->regs.clear();
  for (const auto& x : merged) { it->regs.push_back(x.first); it->regs.push_back(x.second); }
  if (it->regs.size() > 2) {
    if (it->d_regs.ensure(it->regs.size() * 4, it->pass.st) != cudaSuccess ||
        cudaMemcpyAsync(it->d_regs.p, it->regs.data(), it->regs.size() * 4, cudaMemcpyHostToDevice, it->pass.st) != cudaSuccess ||
        cudaStreamSynchronize(it->pass.st) != cudaSuccess) {
      biodb_reads_end(it);
      *out = nullptr;
      return BIODB_ERR_CUDA;
    }
  }
  if (!it->chunks.empty()) region_seek(it, 0);
  return BIODB_OK;
}

// Host-only: the chunk list of the above (merged regions -> getGroupChunks), (begin, end) virtual offsets into out2.
int64_t biodb_index_regions_chunks(const biodb_index* ix, uint32_t ref_id, uint32_t n, const uint32_t* begs, const uint32_t* ends,
                                   uint64_t* out2, uint64_t cap) {
  if (!ix || !n || !begs || !ends) return -1;
  std::vector<std::pair<uint32_t, uint32_t>> rg(n), merged;
  for (uint32_t k = 0; k < n; ++k) {
    if (begs[k] >= ends[k]) return -1;
    rg[k] = {begs[k], ends[k]};
  }
  std::sort(rg.begin(), rg.end());
  for (const auto& x : rg) {
    if (!merged.empty() && merged.back().second >= x.first) merged.back().second = std::max(merged.back().second, x.second);
    else merged.push_back(x);
  }
  std::vector<VoChunk> c;
  if (!ix->bai.regions_chunks(ref_id, merged, &c)) return -1;
  for (uint64_t k = 0; k < c.size() && k < cap && out2; ++k) { out2[2 * k] = c[k].beg; out2[2 * k + 1] = c[k].end; }
  return (int64_t)c.size();
}

biodb_status biodb_reads_begin_between(biodb_reader* r, uint64_t from_voffset, uint64_t to_voffset, uint32_t max_blocks,
                                       biodb_reads** out) {
  if (!r || !out) return BIODB_ERR_ARG;
  biodb_status s = biodb_reads_begin(r, out);
  if (s != BIODB_OK) return s;
  biodb_reads* it = *out;
  it->region = true;
  it->filter = false;
  it->chunk_blocks = max_blocks;
  it->chunks.assign(1, VoChunk{from_voffset, to_voffset});
  it->chunk_i = 0;
  if (to_voffset == ~0ull) {
    it->pass.rewind(from_voffset >> 16, (uint32_t)(from_voffset & 0xFFFF));     // to the end of the file
  } else if (to_voffset <= from_voffset) {
    it->chunks.clear();
  } else {
    region_seek(it, 0);
  }
  return BIODB_OK;
}

biodb_status biodb_reads_next(biodb_reads* it, biodb_record_batch* batch) {
  if (!it || !batch) return BIODB_ERR_ARG;
  Pass& p = it->pass;
  memset(batch, 0, sizeof *batch);
  if (!it->inflight && !it->ahead_done) {                 // first call of the pass
    biodb_status st = reads_produce(it, it->cur ^ 1);
    if (st != BIODB_OK) { it->ahead_done = true; it->ahead_status = st; return st; }
    it->inflight = true;
  }
  if (!it->inflight) return it->ahead_status;             // EOF / the error, again
  const int s = it->cur ^ 1;                              // the batch to hand out now
  bool next_inflight = false;
  if (!it->ahead_done) {
    // read ahead into the slot the caller has just released
    biodb_status st = reads_produce(it, it->cur);
    if (st == BIODB_OK) next_inflight = true;
    else { it->ahead_done = true; it->ahead_status = st; }
  }
  ReadsSlot& sl = it->slot[s];
  if (cudaEventSynchronize(sl.done) != cudaSuccess) return p.fail(BIODB_ERR_CUDA, 0, 0, "device to host copy failed");
  it->cur = s;
  it->inflight = next_inflight;
  cudaStreamWaitEvent(p.st, sl.done, 0);                  // the pass's clock stops after the copies of what it handed out
  p.mark_end();
  const uint64_t n = sl.n;
  batch->n = n;
  batch->first_index = sl.first;
  batch->data = sl.h_data.as<uint8_t>();
  batch->data_len = sl.used;
  batch->rec_off = sl.h_arr[0].as<uint64_t>();
  batch->block_size = sl.h_arr[1].as<int32_t>();
  batch->ref_id = sl.h_arr[2].as<int32_t>();
  batch->pos = sl.h_arr[3].as<int32_t>();
  batch->end_pos = sl.h_arr[4].as<int32_t>();
  batch->bin_mq_nl = sl.h_arr[5].as<uint32_t>();
  batch->flag_nc = sl.h_arr[6].as<uint32_t>();
  batch->l_seq = sl.h_arr[7].as<int32_t>();
  batch->cigar_off = sl.h_arr[8].as<uint64_t>();
  batch->cigar = sl.h_arr[9].as<uint32_t>();
  if (p.r->opts.want_offsets) {
    if (sl.h_vo[0].ensure((size_t)n * 8 + 8) != cudaSuccess || sl.h_vo[1].ensure((size_t)n * 8 + 8) != cudaSuccess)
      return p.fail(BIODB_ERR_CUDA, 0, 0, "pinned allocation failed");
    uint64_t* sv = sl.h_vo[0].as<uint64_t>();
    uint64_t* ev = sl.h_vo[1].as<uint64_t>();
    for (uint64_t i = 0; i < n; ++i) {
      sv[i] = voffset_in(sl.segs, batch->rec_off[i]);                                  // readrange.d:64-66
      ev[i] = voffset_in(sl.segs, batch->rec_off[i] + 4 + (uint64_t)batch->block_size[i]);   // readrange.d:55-57
      // after the last record of a chunk BioD's stream has already moved on to the next chunk, whose start is what
      // virtualTell reports (inputstream.d:516-524)
      if (sl.next_chunk_beg_vo && ev[i] == sl.chunk_end_vo) ev[i] = sl.next_chunk_beg_vo;
    }
    batch->start_voffset = sv;
    batch->end_voffset = ev;
  }
  return BIODB_OK;
}

void biodb_reads_end(biodb_reads* it) {
  if (!it) return;
  biodb_reader* r = it->pass.r;
  if (it->pass.st) cudaStreamSynchronize(it->pass.st);
  it->reset();
  {
    std::lock_guard<std::mutex> lk(r->pool_mu);
    if (r->reads_pool.size() < 2) { r->reads_pool.push_back(it); return; }
  }
  delete it;
}

void biodb_reads_stats(const biodb_reads* it, biodb_stats* out) {
  if (it && out) *out = it->pass.stats;
}

float biodb_reads_progress(const biodb_reads* it) {
  if (!it || !it->pass.r || it->pass.r->flen == 0) return 0.f;
  return (float)((double)it->pass.next_coffset / (double)it->pass.r->flen);
}

// ---- device-resident stage API --------------------------------------------------------------------------

biodb_status biodb_dev_inflate(const uint8_t* comp, const uint64_t* payload_off, const uint32_t* cdata_size,
                               const uint64_t* out_off, const uint32_t* isize, uint32_t n_blocks, uint8_t* out,
                               int32_t* status, uint32_t* crc, void* stream) {
  InflateArgs ia{comp, payload_off, cdata_size, out_off, isize, out, status, n_blocks, WalkOut{}, nullptr};
  cudaError_t e = launch_inflate(ia, (cudaStream_t)stream);
  if (e == cudaSuccess && crc) e = launch_crc32(out, out_off, isize, n_blocks, crc, (cudaStream_t)stream);
  if (e != cudaSuccess) {
    set_error(&g_open_error, BIODB_ERR_CUDA, 0, 0, std::string("biodb_dev_inflate: ") + cudaGetErrorString(e));   // biodb_open_error()
    return BIODB_ERR_CUDA;
  }
  return BIODB_OK;
}

// Host-only: MdChain (md_chain.h) over n reads; returns the number of segments, writes at most cap of them as
// (first, count, read, offset) quadruples of int64.  batch_reads > 0 drives it the way the batch pipeline does: after
// every batch_reads reads the columns below the last read's position are drained (a marker quadruple
// (limit, -1, ref, 0) follows the segments of each drain), and a reference is finished as soon as its last read is in.
int64_t biodb_debug_md_chain(const int32_t* ref_id, const int64_t* pos, const int64_t* end, const int64_t* dna_len, uint64_t n,
                             int32_t skip_zero_coverage, uint64_t batch_reads, int64_t* seg4, uint64_t cap) {
  MdChain chain(skip_zero_coverage != 0);
  std::vector<MdSegment> segs;
  if (!batch_reads) {
    for (uint64_t i = 0; i < n; ++i) chain.admit(i, ref_id[i], pos[i], end[i], dna_len[i], &segs);
  } else {
    std::vector<int32_t> p32, e32, l32;
    for (uint64_t i = 0; i < n;) {
      // one reference group of one batch, the unit the pipeline hands to admit_many
      uint64_t j = i + 1;
      while (j < n && ref_id[j] == ref_id[i] && j % batch_reads != 0) ++j;
      p32.assign(pos + i, pos + j);
      e32.assign(end + i, end + j);
      l32.assign(dna_len + i, dna_len + j);
      chain.admit_many(i, ref_id[i], p32.data(), e32.data(), l32.data(), (size_t)(j - i), &segs);
      if (j == n || ref_id[j] != ref_id[i]) {
        chain.finish_reference(&segs);
      } else {
        chain.drain(pos[j - 1], &segs);
        segs.push_back(MdSegment{pos[j - 1], -1, (uint64_t)(int64_t)ref_id[i], 0});
      }
      i = j;
    }
  }
  chain.finish(&segs);
  for (uint64_t k = 0; k < segs.size() && k < cap; ++k) {
    seg4[4 * k] = segs[k].first;
    seg4[4 * k + 1] = segs[k].count;
    seg4[4 * k + 2] = (int64_t)segs[k].read;
    seg4[4 * k + 3] = segs[k].offset;
  }
  return (int64_t)segs.size();
}

// Host-only: DnaWalk (md_walk.h) over one raw record body (the bytes after block_size); returns the length of
// dna(read) and writes at most cap of its characters.
int64_t biodb_debug_md_dna(const uint8_t* body, int64_t block_size, uint8_t* out, uint64_t cap) {
  if (!body || block_size < 32) return -1;
  DnaWalk w;
  w.init(body, block_size);
  int64_t n = 0;
  for (int c; (c = w.next()) >= 0; ++n)
    if ((uint64_t)n < cap && out) out[n] = (uint8_t)c;
  return n;
}

biodb_status biodb_debug_inflate_counters(uint64_t* out8, int32_t reset) {
  if (!out8) return BIODB_ERR_ARG;
  cudaDeviceSynchronize();
  unsigned long long v[8];
  if (inflate_counters(v, reset) != cudaSuccess) return BIODB_ERR_CUDA;
  for (int i = 0; i < 8; ++i) out8[i] = v[i];
  return BIODB_OK;
}

size_t biodb_dev_scan_workspace_bytes(uint32_t n_blocks) { return scan_workspace_bytes(n_blocks); }

biodb_status biodb_dev_scan_records(const uint8_t* u, uint64_t u_len, const uint64_t* block_uoff, uint32_t n_blocks,
                                    int32_t final_slice, biodb_dev_records* o, uint64_t* result, void* workspace,
                                    size_t workspace_bytes, void* stream) {
  if (!o || workspace_bytes < scan_workspace_bytes(n_blocks)) return BIODB_ERR_ARG;
  RecordArrays ra{o->rec_off, o->block_size, o->ref_id, o->pos, o->end_pos, o->bin_mq_nl, o->flag_nc, o->l_seq,
                  o->cigar_off, o->cigar, o->capacity, o->cigar_capacity};
  ScanWorkspace ws = carve_scan_workspace(workspace, n_blocks);
  return launch_scan_records(u, u_len, block_uoff, n_blocks, n_blocks, final_slice, ra, result, ws, (cudaStream_t)stream) == cudaSuccess
             ? BIODB_OK
             : BIODB_ERR_CUDA;
}

}  // extern "C"
