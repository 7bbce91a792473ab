// BAI index on the host: which stretches of the file hold the reads of a region.
//
// Reference behaviour restated: BaiFile.parse (bio/std/hts/bam/baifile.d:126-169), Index.getMinimumOffset (:77-82),
// Bin.canOverlapWith (bam/bai/bin.d:49-78), RandomAccessManager.getChunks / appendChunks (bam/randomaccessmanager.d:
// 211-244) and the merge of overlapping chunks (bio/core/utils/algo.d:95-162).  A few kilobytes of integer tables per
// reference, read once per file: this stays on the host; what it selects — BGZF blocks to inflate, records to scan
// and filter — is the GPU's work (runtime.cu: region passes).
#pragma once
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <string>
#include <utility>
#include <vector>

namespace biodb {

struct VoChunk { uint64_t beg, end; };          // a pair of virtual offsets (bgzf/chunk.d:28-41)

struct BaiRef {
  std::vector<uint32_t> bin_id;
  std::vector<uint32_t> bin_first;              // chunks of bin k: chunks[bin_first[k] .. bin_first[k+1])
  std::vector<VoChunk> chunks;
  std::vector<uint64_t> ioffsets;               // linear index, one per 16 kbp window
};

struct BaiIndex {
  std::vector<BaiRef> refs;

  // 0 on success; -3 (format) / -4 (truncated) with a message otherwise
  int parse(const uint8_t* d, size_t n, std::string* msg) {
    size_t o = 0;
    auto u32 = [&](uint32_t* v) {
      if (n - o < 4) return false;
      *v = (uint32_t)d[o] | ((uint32_t)d[o + 1] << 8) | ((uint32_t)d[o + 2] << 16) | ((uint32_t)d[o + 3] << 24);
      o += 4;
      return true;
    };
    auto u64 = [&](uint64_t* v) {
      uint32_t lo, hi;
      if (!u32(&lo) || !u32(&hi)) return false;
      *v = (uint64_t)lo | ((uint64_t)hi << 32);
      return true;
    };
    const char* cut = "not enough data in stream";
    // a count read as a negative int32 is a corrupt index, never "an empty table" (BioD: uninitializedArray with a
    // negative length throws); a count larger than what the file still holds ends as "not enough data" below
    auto bad_count = [&](uint32_t v, size_t) { return (int32_t)v < 0; };
    const char* neg = "Invalid BAI file: negative table size";
    if (n < 4) { *msg = cut; return -4; }
    if (memcmp(d, "BAI\1", 4) != 0) { *msg = "Invalid file format: expected BAI\\1"; return -3; }
    o = 4;
    uint32_t n_ref;
    if (!u32(&n_ref)) { *msg = cut; return -4; }
    if (bad_count(n_ref, 8)) { *msg = neg; return -3; }
    refs.clear();
    for (int32_t r = 0; r < (int32_t)n_ref; ++r) {
      BaiRef ref;
      uint32_t n_bin;
      if (!u32(&n_bin)) { *msg = cut; return -4; }
      if (bad_count(n_bin, 8)) { *msg = neg; return -3; }
      for (int32_t b = 0; b < (int32_t)n_bin; ++b) {
        uint32_t id, n_chunk;
        if (!u32(&id) || !u32(&n_chunk)) { *msg = cut; return -4; }
        if (bad_count(n_chunk, 16)) { *msg = neg; return -3; }
        ref.bin_id.push_back(id);
        ref.bin_first.push_back((uint32_t)ref.chunks.size());
        for (int32_t c = 0; c < (int32_t)n_chunk; ++c) {
          VoChunk ch;
          if (!u64(&ch.beg) || !u64(&ch.end)) { *msg = cut; return -4; }
          ref.chunks.push_back(ch);
        }
      }
      ref.bin_first.push_back((uint32_t)ref.chunks.size());
      uint32_t n_intv;
      if (!u32(&n_intv)) { *msg = cut; return -4; }
      if (bad_count(n_intv, 8)) { *msg = neg; return -3; }
      for (int32_t k = 0; k < (int32_t)n_intv; ++k) {
        uint64_t v;
        if (!u64(&v)) { *msg = cut; return -4; }
        ref.ioffsets.push_back(v);
      }
      refs.push_back(std::move(ref));
    }
    return 0;
  }

  // bin.d:49-78.  D mixes uint and int here: `id - magic` and both comparisons are unsigned.
  static bool bin_overlaps(uint32_t id, int32_t begin, int32_t end) {
    if (id == 0) return true;
    if (id > 37449u) return false;              // BAI_MAX_BIN_ID (bam/constants.d:35)
    if (begin < 0) begin = 0;
    int32_t first = 4681, b = begin >> 14, e = end >> 14;
    for (;;) {
      const uint32_t delta = id - (uint32_t)first;
      if ((uint32_t)b <= delta && delta <= (uint32_t)e) return true;
      first >>= 3;
      if (first == 0) return false;
      b >>= 3;
      e >>= 3;
    }
  }

  // getChunks(BamRegion(ref_id, beg, end)): false = "Invalid reference sequence index"
  bool region_chunks(uint32_t ref_id, uint32_t beg, uint32_t end, std::vector<VoChunk>* out) const {
    out->clear();
    if (ref_id >= refs.size()) return false;
    const BaiRef& r = refs[ref_id];
    const int32_t pos = std::max<int32_t>(0, (int32_t)beg);
    const int32_t w = std::min<int32_t>(pos / 16384, (int32_t)r.ioffsets.size() - 1);    // BAI_LINEAR_INDEX_WINDOW_SIZE
    const uint64_t min_offset = w == -1 ? 0 : r.ioffsets[(size_t)w];
    std::vector<VoChunk> all;
    for (size_t k = 0; k < r.bin_id.size(); ++k) {
      if (!bin_overlaps(r.bin_id[k], (int32_t)beg, (int32_t)end)) continue;
      for (uint32_t c = r.bin_first[k]; c < r.bin_first[k + 1]; ++c) {
        const VoChunk& ch = r.chunks[c];
        if (ch.end > min_offset) all.push_back(VoChunk{std::max(ch.beg, min_offset), ch.end});
      }
    }
    std::sort(all.begin(), all.end(), [](const VoChunk& a, const VoChunk& b) { return a.beg != b.beg ? a.beg < b.beg : a.end < b.end; });
    for (const VoChunk& ch : all) {
      if (!out->empty() && out->back().end >= ch.beg) out->back().end = std::max(out->back().end, ch.end);
      else out->push_back(ch);
    }
    return true;
  }

  // getGroupChunks (randomaccessmanager.d:246-296): the chunks of every bin that the bin bitset of ALL the regions (one
  // reference; sorted, non-overlapping, each begin < end) marks, cut at the linear index's minimum offset for the first
  // region's start, sorted and merged.  The reference's bitset has BAI_MAX_BIN_ID = 37449 entries and is indexed with
  // every bin id of the index, so the pseudo-bin 37450 of samtools' and BioD's own indexes is out of its range (an
  // error in D): bins beyond the bitset are passed over here, as Bin.canOverlapWith does for a single region.
  bool regions_chunks(uint32_t ref_id, const std::vector<std::pair<uint32_t, uint32_t>>& regions, std::vector<VoChunk>* out) const {
    out->clear();
    if (ref_id >= refs.size() || regions.empty()) return false;
    const BaiRef& r = refs[ref_id];
    std::vector<char> bitset(37449, 0);
    bitset[0] = 1;
    for (const auto& rg : regions) {
      const uint32_t beg = rg.first, end = rg.second - 1;
      uint32_t k;
      for (k = 1 + (beg >> 26); k <= 1 + (end >> 26); ++k) bitset[k] = 1;
      for (k = 9 + (beg >> 23); k <= 9 + (end >> 23); ++k) bitset[k] = 1;
      for (k = 73 + (beg >> 20); k <= 73 + (end >> 20); ++k) bitset[k] = 1;
      for (k = 585 + (beg >> 17); k <= 585 + (end >> 17); ++k) bitset[k] = 1;
      for (k = 4681 + (beg >> 14); k <= 4681 + (end >> 14) && k < 37449; ++k) bitset[k] = 1;
    }
    const int32_t pos = std::max<int32_t>(0, (int32_t)regions.front().first);
    const int32_t w = std::min<int32_t>(pos / 16384, (int32_t)r.ioffsets.size() - 1);
    const uint64_t min_offset = w == -1 ? 0 : r.ioffsets[(size_t)w];
    std::vector<VoChunk> all;
    for (size_t k = 0; k < r.bin_id.size(); ++k) {
      if (r.bin_id[k] >= bitset.size() || !bitset[r.bin_id[k]]) continue;
      for (uint32_t c = r.bin_first[k]; c < r.bin_first[k + 1]; ++c) {
        const VoChunk& ch = r.chunks[c];
        if (ch.end > min_offset) all.push_back(VoChunk{std::max(ch.beg, min_offset), ch.end});
      }
    }
    std::sort(all.begin(), all.end(), [](const VoChunk& a, const VoChunk& b) { return a.beg != b.beg ? a.beg < b.beg : a.end < b.end; });
    for (const VoChunk& ch : all) {
      if (!out->empty() && out->back().end >= ch.beg) out->back().end = std::max(out->back().end, ch.end);
      else out->push_back(ch);
    }
    return true;
  }
};

}  // namespace biodb
