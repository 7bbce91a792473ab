// Internal launch interface between the host runtime and the three kernel families.
#pragma once
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace biodb {

// number of kernel launches issued by this library (all streams); read as deltas for biodb_stats
extern unsigned long long g_kernel_launches;

// Small copies done by a kernel instead of a DMA engine: on the compute stream they must not queue behind the
// gigabyte-sized device->host copies of the previous batch.  dst/src may be device memory or mapped pinned host memory.
cudaError_t launch_copy_bytes(void* dst, const void* src, size_t bytes, cudaStream_t st);

// ---- inflate.cu ------------------------------------------------------------------------------
constexpr int SCAN_SLOTS = 2048;   // max records that can start inside one 64 KiB block (65536/37 < 2048)

// Optional fused record-chain walk (see records.cu): the inflate warp follows the block_size chain of its own
// block out of the shared-memory output ring, assuming a record starts at the block boundary.
struct WalkOut {
  uint16_t* rel;      // [n_scan_blocks * SCAN_SLOTS] record starts relative to the block start (nullptr = no walk)
  uint32_t* cnt;
  uint32_t* ncig;
  uint64_t* out;      // absolute offset (in the slice) where the chain stopped
  uint64_t* in;       // absolute offset where it entered
  int32_t* bad;       // stop reason (WALK_*)
  uint32_t sb_offset; // scan-block index of BGZF block 0 (1 when a carried tail forms pseudo block 0)
  uint32_t in0;       // entry offset inside block 0 (bytes of header in front of the first record)
  uint64_t u_len;     // length of the whole slice
  int32_t n_refs;     // reference count of the file (plausibility of refID fields)
  int32_t search0;    // 1 = the chain's entry into block 0 is not known either (a pass that starts at a block boundary of
                      // a file whose records straddle blocks): block 0 searches for it like every later block
};
enum { WALK_OK = 0, WALK_TAIL = 1, WALK_BAD_SIZE = 2, WALK_BAD_FIELDS = 3, WALK_INCOMPLETE = 4 };

struct InflateArgs {
  const uint8_t* comp;          // device: compressed file slice
  const uint64_t* payload_off;  // [n] byte offset of each block's raw-DEFLATE payload inside comp
  const uint32_t* cdata_size;   // [n]
  const uint64_t* out_off;      // [n] byte offset of each block's output inside out
  const uint32_t* isize;        // [n]
  uint8_t* out;
  int32_t* status;              // [n] 0 / Z_DATA_ERROR / Z_BUF_ERROR
  uint32_t n_blocks;
  WalkOut walk;
  uint16_t* tok;                // device: the blocks' record streams (inflate_tok.cu), inflate_token_bytes(n_blocks) bytes;
                                // nullptr = launch_inflate allocates it for the call (stream-ordered)
};
// inflate_tok.cu: token rows (decode trips) one lane of the decode kernel may record per super-chunk
constexpr int TOK_MAX_TRIPS = 128;
// inflate_tok.cu: 16-bit words of one block's record stream (token rows + super-chunk headers).  A block of 64 KiB has
// at most 65536 tokens; rows hold two tokens per lane and are padded to the longest lane: BAM data needs ~57 K words,
// an all-literal block ~85 K, a block dense in 3-byte matches ~130 K.  A stream that outgrows the arena goes to the
// warp-serial kernel.
constexpr uint32_t TOK_ARENA_WORDS = 160 * 1024;
// inflate.cu: the kernels chosen by BIODB_INFLATE (default: the decode + resolve pair of inflate_tok.cu, then the
// warp-serial kernel on the blocks they gave up on; "serial": the warp-serial kernel alone)
cudaError_t launch_inflate(const InflateArgs& a, cudaStream_t st);
// bytes of InflateArgs::tok for n_blocks blocks (0 when the selected kernel needs none)
size_t inflate_token_bytes(uint32_t n_blocks);
// inflate_tok.cu: decode kernel + resolve kernel; blocks they cannot finish get status STATUS_RETRY (inflate_common.cuh)
cudaError_t launch_inflate_tok(const InflateArgs& a, cudaStream_t st);
size_t inflate_tok_token_bytes(uint32_t n_blocks);
int inflate_tok_resident_blocks(int device);
// diagnostics since the last reset: [0] blocks given up (redone by the warp-serial kernel), [1] super-chunks,
// [2] decode rounds, [3] matches read back from L2, [4] matches, [5] DEFLATE blocks
cudaError_t inflate_tok_counters(unsigned long long* out8, int reset);
cudaError_t inflate_counters(unsigned long long* out8, int reset);
size_t inflate_smem_bytes();
// BGZF blocks (CTAs of the decode kernel) resident on the whole device at once; 0 if unknown
int inflate_resident_blocks(int device);

// ---- crc32.cu ---------------------------------------------------------------------------------
// CRC-32 of each block's inflated bytes (options.verify_crc; BioD asserts it in debug builds, block.d:187)
cudaError_t launch_crc32(const uint8_t* out, const uint64_t* out_off, const uint32_t* isize, uint32_t n_blocks,
                         uint32_t* crc, cudaStream_t st);

// ---- records.cu ------------------------------------------------------------------------------
struct RecordArrays {
  uint64_t* rec_off;
  int32_t* block_size;
  int32_t* ref_id;
  int32_t* pos;
  int32_t* end_pos;
  uint32_t* bin_mq_nl;
  uint32_t* flag_nc;
  int32_t* l_seq;
  uint64_t* cigar_off;   // [cap+1]
  uint32_t* cigar;
  uint64_t capacity;
  uint64_t cigar_capacity;
};
struct ScanWorkspace {            // device scratch, sized by scan_workspace_bytes(n_blocks)
  uint16_t* rel;                  // [n_blocks*SCAN_SLOTS] record starts relative to the block start
  uint32_t* cnt;                  // [n_blocks] records whose prefix starts in the block
  uint32_t* ncig;                 // [n_blocks] cigar words of those records
  uint64_t* out;                  // [n_blocks] where the chain leaves the block (absolute offset in u)
  uint64_t* in;                   // [n_blocks] where the chain entered (speculated, then true)
  uint64_t* rec_base;             // [n_blocks+1] exclusive scan of cnt
  uint64_t* cig_base;             // [n_blocks+1]
  int32_t* bad;                   // [n_blocks] malformed-record flag met during the walk
};
size_t scan_workspace_bytes(uint32_t n_blocks);
ScanWorkspace carve_scan_workspace(void* base, uint32_t n_blocks);
// result (device, 4 x u64): n_records, tail offset, n_cigar_words, status
// n_walk: number of leading scan blocks whose chain still has to be walked by scan_walk_kernel (all of them,
// or only the carried-tail pseudo block when the inflate kernel already walked the rest).
cudaError_t launch_scan_records(const uint8_t* u, uint64_t u_len, const uint64_t* block_uoff, uint32_t n_blocks,
                                uint32_t n_walk, int final_slice, const RecordArrays& out, uint64_t* result,
                                const ScanWorkspace& ws, cudaStream_t st);

// ---- region.cu -------------------------------------------------------------------------------
// BamReadFilter (randomaccessmanager.d:366-462) for one region + compaction of the record tables: `out` receives the
// kept records in order (cigar_off rebased to its own cigar array).  scratch: region_scratch_elems(n) u32;
// info (device, 4 x u32): [0] index of the first read that ends the range (0xffffffff = none), [1] reads kept,
// [2] their CIGAR words.
size_t region_scratch_elems(uint64_t n);
// regs / n_regs (getReads(BamRegion[])): n_regs > 1 sorted, non-overlapping regions of `ref` as (begin, end) pairs on the
// device; [beg, end) is then the first region's begin and the last one's end.
cudaError_t launch_region_filter(const RecordArrays& in, uint64_t n, uint32_t ref, uint32_t beg, uint32_t end,
                                 const RecordArrays& out, uint32_t* scratch, uint32_t* info, cudaStream_t st,
                                 const uint32_t* regs = nullptr, uint32_t n_regs = 0);

}  // namespace biodb
