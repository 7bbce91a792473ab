// BAI index construction on the host (last part): IndexBuilder of bio/std/hts/bam/bai/indexing.d:
// 56-351 restated — put() per read in file order (:281-322), updateLinearIndex (:133-164), updateChunks (:219-246),
// dumpCurrentReference (:186-216) with the metadata pseudo-bin 37450, finish (:325-339).  A sequential scan over
// (reference, position, end, bin, unmapped flag, start / end virtual offset) of every read with O(1) state per
// reference: it stays on the host; the reads and their offsets come from the GPU passes.
// Restatement-defined: the bins of a reference are written in ascending id order (D iterates an associative array:
// the order is the hash table's), and positions beyond the 2^29 the linear index can hold are clamped to its last window.
#pragma once
#include <stdint.h>

#include <map>
#include <string>
#include <vector>

namespace biodb {

struct BaiBuilder {
  struct Chunk { uint64_t beg, end; };
  static constexpr size_t LINEAR_SIZE = 37449 - 4680 + 1;      // BAI_MAX_BIN_ID - BAI_MAX_NONLEAF_BIN_ID + 1 (:262)

  std::vector<uint8_t> out;
  std::string err;
  int32_t n_refs = 0;
  bool check_bins = false;
  // state of IndexBuilder
  std::vector<uint64_t> linear = std::vector<uint64_t>(LINEAR_SIZE, 0);
  size_t linear_write_length = 0;
  struct Prev { int32_t ref_id = -1, position = 0, end_position = 0; uint32_t bin = 0; bool is_unmapped = false;
                uint64_t start_vo = 0, end_vo = 0; uint64_t index = 0; } prev;
  uint64_t no_coord = 0, beg_vo = ~0ull, end_vo = 0, unmapped = 0, mapped = 0;
  bool first_read = true;
  std::map<uint32_t, std::vector<Chunk>> chunks;
  uint64_t current_chunk_beg = 0;
  uint64_t n_put = 0;

  void w32(uint32_t v) { for (int k = 0; k < 4; ++k) out.push_back((uint8_t)(v >> (8 * k))); }
  void w64(uint64_t v) { for (int k = 0; k < 8; ++k) out.push_back((uint8_t)(v >> (8 * k))); }

  void begin(int32_t number_of_references, bool check) {         // :258-268
    n_refs = number_of_references;
    check_bins = check;
    out.assign({'B', 'A', 'I', 1});
    w32((uint32_t)n_refs);
  }
  static size_t to_linear(int64_t position) {                     // toLinearIndexOffset (:51-53)
    const int64_t w = position < 0 ? 0 : position / 16384;
    return (size_t)(w < (int64_t)LINEAR_SIZE ? w : (int64_t)LINEAR_SIZE - 1);
  }
  static uint32_t reg2bin(int32_t beg, int32_t end) {             // bai/bin.d:82-92
    if (end == beg) end = beg + 1;
    --end;
    if (beg >> 14 == end >> 14) return ((1 << 15) - 1) / 7 + (beg >> 14);
    if (beg >> 17 == end >> 17) return ((1 << 12) - 1) / 7 + (beg >> 17);
    if (beg >> 20 == end >> 20) return ((1 << 9) - 1) / 7 + (beg >> 20);
    if (beg >> 23 == end >> 23) return ((1 << 6) - 1) / 7 + (beg >> 23);
    if (beg >> 26 == end >> 26) return ((1 << 3) - 1) / 7 + (beg >> 26);
    return 0;
  }
  void write_empty_reference() { w32(0); w32(0); }                 // :101-104
  void update_linear_index() {                                     // :133-164
    size_t beg, end;
    if (prev.is_unmapped) {
      end = beg = to_linear(prev.position);
    } else {
      beg = to_linear(prev.position);
      end = to_linear((int64_t)prev.position + (prev.end_position - prev.position) - 1);
    }
    for (size_t i = beg; i < end + 1; ++i)
      if (linear[i] == 0) linear[i] = prev.start_vo;
    if (end + 1 > linear_write_length) linear_write_length = end + 1;
  }
  void update_chunks() {                                           // :219-246
    const uint64_t current_chunk_end = prev.end_vo;
    std::vector<Chunk>& cs = chunks[prev.bin];
    if (cs.empty() || (cs.back().end >> 16) != (current_chunk_beg >> 16)) cs.push_back(Chunk{current_chunk_beg, current_chunk_end});
    else cs.back().end = current_chunk_end;
    current_chunk_beg = current_chunk_end;
  }
  void dump_current_reference() {                                  // :186-216, :166-184
    w32((uint32_t)(chunks.size() + 1));                            // + the metadata bin
    for (const auto& kv : chunks) {
      if (kv.second.empty()) continue;
      w32(kv.first);
      w32((uint32_t)kv.second.size());
      for (const Chunk& c : kv.second) { w64(c.beg); w64(c.end); }
    }
    w32(37450);
    w32(2);
    w64(beg_vo);
    w64(end_vo);
    w64(mapped);
    w64(unmapped);
    w32((uint32_t)linear_write_length);
    uint64_t last = 0;
    for (size_t i = 0; i < linear_write_length; ++i) {
      uint64_t v = linear[i];
      if (v == 0) v = last; else last = v;
      w64(v);
    }
    std::fill(linear.begin(), linear.end(), 0);
    linear_write_length = 0;
    chunks.clear();
    current_chunk_beg = prev.end_vo;
    beg_vo = end_vo = current_chunk_beg;
    unmapped = mapped = 0;
  }
  // put(BamReadBlock) (:281-322).  false: the input is not sorted / a bin is wrong (message in err).
  bool put(int32_t ref_id, int32_t position, int32_t end_position, uint32_t bin, bool is_unmapped, uint64_t start_vo,
           uint64_t read_end_vo) {
    const uint64_t index = n_put++;
    // checkThatInputIsSorted (:248-262)
    if (!first_read && ref_id != -1 && !(prev.ref_id < ref_id) && !(ref_id == prev.ref_id && position >= prev.position)) {
      err = "BAM file is not coordinate-sorted: read " + std::to_string(index) + " (" + std::to_string(ref_id) + ":" +
            std::to_string(position) + ") must be after read " + std::to_string(prev.index) + " (" +
            std::to_string(prev.ref_id) + ":" + std::to_string(prev.position) + ")";
      return false;
    }
    bool ok = true;
    if (ref_id >= 0 && position >= 0) {
      if (first_read) {
        prev = Prev{ref_id, position, end_position, bin, is_unmapped, start_vo, read_end_vo, index};
        first_read = false;
        current_chunk_beg = start_vo;
        for (int32_t i = 0; i < ref_id; ++i) write_empty_reference();
      } else {
        if (check_bins) {                                          // checkThatBinIsCorrect (:236-246)
          const uint32_t expected = reg2bin(position, end_position);
          if (bin != expected) {
            err = "Bin in read " + std::to_string(index) + " is set incorrectly (" + std::to_string(bin) +
                  " instead of expected " + std::to_string(expected) + ")";
            ok = false;
          }
        }
        if (ok) {
          if (ref_id > prev.ref_id) {
            update_linear_index();
            update_chunks();
            dump_current_reference();
            for (int32_t i = prev.ref_id + 1; i < ref_id; ++i) write_empty_reference();
          }
          if (ref_id == prev.ref_id) {
            update_linear_index();
            if (bin != prev.bin) update_chunks();
          }
          prev = Prev{ref_id, position, end_position, bin, is_unmapped, start_vo, read_end_vo, index};
        }
      }
    }
    // updateMetadata (:117-131), at scope exit
    if (ref_id == -1) {
      ++no_coord;
    } else {
      if (is_unmapped) ++unmapped; else ++mapped;
      if (beg_vo == ~0ull) beg_vo = start_vo;
      end_vo = read_end_vo;
    }
    return ok;
  }
  void finish() {                                                  // :325-339
    if (!first_read) {
      update_linear_index();
      update_chunks();
      dump_current_reference();
    }
    for (int32_t i = prev.ref_id + 1; i < n_refs; ++i) write_empty_reference();
    w64(no_coord);
  }
};

}  // namespace biodb
