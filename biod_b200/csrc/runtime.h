// Host runtime behind the C ABI: file / block-table handling, device + pinned buffers, the batch
// pipeline (H2D -> inflate -> record scan -> [pileup] -> D2H).  No CPU inflate / decode / pileup
// path exists here by design: without a CUDA device every entry point fails with BIODB_ERR_CUDA.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include <memory>
#include <mutex>
#include <string>
#include <future>
#include <vector>

#include "../../include/biod_b200.h"
#include "kernels.h"
#include "pileup.h"

namespace biodb {

struct DevBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~DevBuf() { if (p) cudaFree(p); }
  // grows (never shrinks); contents are NOT preserved unless keep > 0
  cudaError_t ensure(size_t bytes, cudaStream_t st = 0, size_t keep = 0);
  template <typename T> T* as() const { return (T*)p; }
};
struct PinBuf {
  void* p = nullptr;
  size_t cap = 0;
  ~PinBuf() { if (p) cudaFreeHost(p); }
  cudaError_t ensure(size_t bytes);
  template <typename T> T* as() const { return (T*)p; }
};

struct BlockInfo;
}
namespace biodb {
struct BlockInfo {      // BgzfBlock (bgzf/block.d:42-73)
  uint64_t coffset;     // start_offset
  uint64_t payload;     // file offset of the deflate payload
  uint32_t bsize;       // total size - 1
  uint32_t cdata_size;
  uint32_t crc32;
  uint32_t isize;       // input_size
};

struct Seg {            // where a run of uncompressed bytes of the current slice came from
  uint64_t ustart;      // offset inside the slice
  uint64_t coffset;     // BGZF block start
  uint32_t within;      // offset inside that block's uncompressed data
  uint32_t len;
  uint64_t cend;        // end_offset of the block (start of the next one)
};

}  // namespace biodb

struct biodb_reader {
  biodb_options opts;
  int device = 0;
  const uint8_t* file = nullptr;  // the whole file in host memory (biodb_open_memory), or nullptr: streamed from fd
  uint64_t flen = 0;
  // biodb_open(path): the file is STREAMED, as BioD's reader streams it through n_tasks x 128 KiB of read-ahead
  // (bgzf/inputstream.d:467-478) — never held whole.  Block headers are parsed out of a small sliding window (hdr_win),
  // the compressed bytes of a batch are pread into one of the pass's two pinned slabs while the GPU inflates the batch
  // before (Pass::stage_host), so the pinned memory a reader needs is two batches, whatever the file's size.
  int fd = -1;
  const uint8_t* hdr_map = nullptr;   // read-only mapping of the file, used ONLY to look at the 26 header / footer bytes of
                                      // each block (no copy; the window below is the fallback when mmap is refused)
  std::vector<uint8_t> hdr_win;   // file bytes [hdr_off, hdr_off + hdr_win.size())
  uint64_t hdr_off = 0;
  std::mutex hdr_mu;
  // parse the BGZF member header at file offset pos (in memory or through the window): as parse_bgzf_header
  int header_at(uint64_t pos, biodb::BlockInfo* b, biodb_error* e);
  // copy file bytes [off, off + len) to dst (host memory): memcpy or pread, on `threads` threads for large ranges
  bool read_bytes(uint64_t off, size_t len, void* dst, int threads = 1) const;
  bool registered = false;        // the whole caller buffer is page-locked (options.pin_input == 1)
  // options.pin_input == 2: page-lock on demand, only the byte ranges passes really read (one shard of many GPUs' worth
  // of file must not make every process pin the whole file)
  std::vector<std::pair<uint64_t, uint64_t>> pinned;   // disjoint, sorted [lo, hi) offsets into file
  void ensure_pinned(uint64_t lo, uint64_t hi);
  std::string text;
  std::vector<std::string> ref_names;
  std::vector<int32_t> ref_lens;
  uint64_t reads_start_vo = 0;
  uint64_t reads_start_coffset = 0;
  uint32_t reads_start_uoffset = 0;
  bool reads_start_at_eof = false;
  biodb_error err{};
  biodb::DevBuf d_file;           // options.resident_input: the compressed file in HBM
  // finished passes park their buffers here so that the next pass does not reallocate them
  std::mutex pool_mu;
  std::vector<void*> pileup_pool, reads_pool;
  // coffset of every data block from the first record on (built on demand for sharding)
  std::vector<uint64_t> block_index;
  uint64_t data_end_coffset = 0;
  biodb_status build_block_index();
  // MAQ coefficient tables on the device (maq.h), kept per (depcorr, eta)
  struct MaqCache { float depcorr, eta; biodb::DevBuf fk, beta, lhet; };
  std::vector<std::unique_ptr<MaqCache>> maq_cache;
  // cuts of the file into n shards (biodb_shard_cuts), kept per n
  struct ShardCuts { uint32_t n = 0; std::vector<uint64_t> vo; std::vector<int32_t> ref; std::vector<int64_t> pos; };
  std::vector<ShardCuts> shard_cuts;
};

namespace biodb { struct VoChunk; }
struct biodb_index;
// chunk list of a region read (defined in runtime.cu next to biodb_index)
biodb_status biodb_index_region_chunks(const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end,
                                       std::vector<biodb::VoChunk>* out);

namespace biodb {

// One sequential pass over the BGZF blocks of a reader: owns a CUDA stream and all buffers.
struct Pass {
  biodb_reader* r = nullptr;
  cudaStream_t st = nullptr;
  // position in the file
  uint64_t next_coffset = 0;
  uint64_t stop_coffset = ~0ull;      // shard end: blocks at or beyond it are not read
  uint32_t stop_uoffset = 0;          // > 0: the block AT stop_coffset is read too, but only its first stop_uoffset bytes
                                      // belong to the stream (the skip_end of a BAI chunk, inputstream.d:316-322)
  uint32_t first_skip = 0;          // bytes of the first block that precede the first record
  bool entry_search = false;        // the pass starts at a block boundary that need not be a record boundary: the first
                                    // block searches for the record chain's entry like every later one (Walker)
  bool supplier_done = false;       // EOF block / end of file reached (inputstream.d:393-394)
  biodb_error pending{};            // error to raise once the blocks before it are consumed
  bool finished = false;
  uint64_t n_records_total = 0;
  // host staging
  std::vector<BlockInfo> blocks;    // blocks of the current batch
  // the BSIZE chain of the NEXT batch, walked on the host while the GPU inflates the current one
  std::vector<BlockInfo> pre_blocks;
  bool pre_valid = false;
  uint32_t pre_max = 0;
  uint64_t pre_next_coffset = 0;
  bool pre_supplier_done = false;
  biodb_error pre_pending{};
  PinBuf h_tab;                     // per-block tables (payload_off, out_off, cdata, isize, block_uoff)
  PinBuf h_status, h_result;
  // device
  DevBuf d_comp2[2], d_tab, d_status, d_u, d_carry_tail, d_ws, d_result, d_tok;
  PinBuf h_slab[2];                 // streamed files: the compressed bytes of a batch on their way to the device
  int slab_cur = 0;
  std::future<int> pf_job;          // streamed files: the prefetch (pread into a slab, then the copy) on a thread of its own
  int join_prefetch();              // waits for it; 0 = none / fine, 1 = read error, 2 = CUDA error
  // host address of file bytes [c0, c1) for a host->device copy: the caller's buffer, or (streamed file) the next slab
  const uint8_t* stage_host(uint64_t c0, uint64_t c1, cudaError_t* err);
  // host->device prefetch of the next batch's compressed bytes (input not resident): while batch k is inflated out of
  // d_comp2[cur], the bytes that follow it in the file go to d_comp2[cur^1] on their own stream
  cudaStream_t h2d_st = nullptr;
  cudaEvent_t pf_done = nullptr;       // the prefetch copy has landed
  cudaEvent_t comp_free[2] = {nullptr, nullptr};   // the inflate kernel reading d_comp2[i] has finished
  int comp_cur = 0;
  uint64_t pf_c0 = 0, pf_c1 = 0;       // file range held by d_comp2[comp_cur^1] (pf_c1 == 0: none)
  DevBuf d_rec[10];                 // RecordArrays columns
  uint64_t rec_capacity = 0, cigar_capacity = 0;
  uint64_t rec_front = 0;           // slots reserved in front of the batch records (pileup carry)
  // current slice
  uint64_t u_len = 0;
  uint64_t carry_tail_len = 0;      // bytes of a cut record carried into the next slice
  std::vector<Seg> segs, next_segs;
  uint64_t end_coffset_last = 0;
  uint64_t n = 0, n_cigar = 0, tail = 0;
  bool final_slice = false;
  ScanWorkspace ws_cur{};           // scan workspace of the current slice (rec_base etc.)
  uint32_t ws_carry = 0;            // 1 when scan block 0 is the carried tail
  bool raw_mode = false;            // header pass: inflate only, no record scan, no carry
  // measurement (biodb_stats)
  biodb_stats stats{};
  cudaEvent_t ev_begin = nullptr, ev_end = nullptr, ev_a = nullptr, ev_b = nullptr;
  bool began = false;
  unsigned long long launches0 = 0;
  void mark_begin();                // first operation of the pass
  void mark_end();                  // call with the stream idle-able: records + syncs, updates total_ms
  // stage timing without extra synchronisation: event pairs are recorded around the launches of a stage and
  // turned into milliseconds at the next point where the stream is synchronised anyway
  struct Timed { cudaEvent_t a, b; double* acc; };
  std::vector<cudaEvent_t> ev_pool;
  std::vector<Timed> timed;
  cudaEvent_t stage_a = nullptr;
  void stage_begin();
  void stage_end(double* acc);
  void collect_timing();            // call when the stream is known to be idle

  ~Pass();
  biodb_status init(biodb_reader* rd, uint64_t coffset, uint32_t uoffset);
  void rewind(uint64_t coffset, uint32_t uoffset);   // start a new pass, keeping every buffer
  // Inflate + scan the next batch.  BIODB_OK (n may be 0), BIODB_EOF, or an error.
  biodb_status next(uint32_t max_blocks, uint64_t front_slots);
  RecordArrays arrays(uint64_t front) const;
  uint64_t voffset_of(uint64_t x) const;
  biodb_status fail(int status, int zerr, uint64_t off, const std::string& msg);
};

uint64_t voffset_in(const std::vector<Seg>& segs, uint64_t x);
int parse_bgzf_header(const uint8_t* d, uint64_t len, uint64_t pos, BlockInfo* b, biodb_error* e);

}  // namespace biodb
