// Internal interface of the pileup builder (pileup.cu) used by the host runtime.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace biodb {

// The work table: carried-over reads of earlier batches (indices [0,n_carry)) followed by the
// records of the current batch, all in file order.  rec_off of a carried read is relative to
// carry_data, of a batch read relative to u.
struct ReadsView {
  const int32_t* pos;
  const int32_t* end_pos;
  const int32_t* ref_id;
  const uint64_t* rec_off;
  const uint32_t* bin_mq_nl;
  const uint32_t* flag_nc;
  const int32_t* l_seq;
  const uint32_t* carry_gidx;   // [n_carry] file-order index of carried reads
  const uint8_t* carry_data;
  const uint8_t* u;
  uint32_t n_carry;
  uint32_t n;                   // carried + batch reads
  uint64_t first_index;         // file-order index of batch read 0
};

struct IslandTable {
  int32_t* start;   // position of the island's first read
  int32_t* end;     // max end over the island
  uint32_t* first;  // index of its first read
  int64_t* cs;      // first emitted position after clipping
};

struct GroupScratch {            // sized by read capacity
  int32_t* eend;                 // end_pos of live reads, INT32_MIN for filtered ones
  int32_t* pm;                   // running max of eend
  uint4* rinfo;                  // per live read: {seq pointer lo, hi, query offset at the first column | simple<<31, file-order index}
  uint32_t* flag;                // island-start flags
  uint32_t* iid1;                // 1-based island id
  IslandTable islands;
  uint32_t* ncol;
  uint32_t* colbase;
  uint32_t* cflag;
  uint32_t* cslot;
  uint64_t* cbytes;
  int32_t* tmp_i32;              // scan tile aggregates (each sized for max(reads, columns+1))
  uint32_t* tmp_u32;
  uint32_t* tmp_u32b;
  uint64_t* tmp_u64;
  int32_t* info;                 // [4]: status, last live read + 1
  int64_t clo, chi;
};

struct ColumnScratch {           // sized by column capacity + 1
  int32_t* diff;                 // difference array, then coverage
  uint32_t* nstart;
  uint32_t* hi;
  uint32_t* lo;
};

struct ColumnOutput {
  uint64_t* col_pos;             // [n_col]
  uint64_t* col_off;             // [n_col+1]
  uint32_t* read_idx;            // [n_entries]
  uint8_t* base;
  uint8_t* qual;
  uint32_t* qoff;                // or nullptr
  uint32_t* counts;              // counts_only: [n_col*6] A,C,G,T,other,deletion; entries are then not written
  int32_t maq_min_base_quality;  // >= 0: MAQ mode — base / qual hold the caller's view of the entries (see entries_kernel)
};

struct CarryOut {
  int32_t* pos;
  int32_t* end_pos;
  int32_t* ref_id;
  uint64_t* rec_off;
  uint32_t* bin_mq_nl;
  uint32_t* flag_nc;
  int32_t* l_seq;
  int32_t* block_size;
  uint32_t* gidx;
  uint8_t* data;
};

// ---- mdtag.cu: reference bases from MD tags (row N1) ------------------------------------------------------------
struct MdSeg {                   // = MdSegment of md_chain.h: reference_base[P] = dna(read)[offset + (P - first)], P in [first, first+count)
  int64_t first, count;
  uint64_t read;                 // file-order index of the provider
  int64_t offset;
};
struct MdKeep {                  // dna() strings of providers that outlive the record bytes of their batch
  uint64_t id[2];                // file-order index, ~0 = empty slot
  uint32_t len[2];
  const uint8_t* data[2];
};
// dna_len[j] = length of dna(read j) for the live reads of [a0, g1), 0 for the others
void md_dna_lengths(const ReadsView& v, const int32_t* block_size, const int32_t* eend, uint32_t a0, uint32_t g1,
                    int32_t* dna_len, cudaStream_t st);
// ref_base[c] (preset to 'N') of the columns covered by the segments
void md_replay(const ReadsView& v, const int32_t* block_size, const MdSeg* segs, uint32_t n_segs, const MdKeep& keep,
               const uint64_t* col_pos, uint32_t n_col, uint8_t* ref_base, cudaStream_t st);
// materialise the dna() of new_keep.id[k] (len[k] characters) into new_keep.data[k]
void md_keep(const ReadsView& v, const int32_t* block_size, const MdKeep& old_keep, const MdKeep& new_keep, cudaStream_t st);

void pileup_count_below(const uint64_t* rec_off, uint32_t n, uint64_t x, uint64_t* out, cudaStream_t st);
void pileup_reach(const int32_t* ref_id, const int32_t* pos, const int32_t* end_pos, const uint64_t* rec_off, uint32_t first,
                  uint32_t n, const int32_t* kref, const int64_t* kpos, uint32_t nk, unsigned long long* reach, cudaStream_t st);
void pileup_find_groups(const ReadsView& v, uint32_t* boundaries, uint32_t* n_boundaries, uint32_t cap, cudaStream_t st);
void pileup_first_kept(const ReadsView& v, uint32_t g0, uint32_t g1, uint64_t start_from, uint32_t* first, cudaStream_t st);
void pileup_phase1(const ReadsView& v, uint32_t g0, uint32_t g1, uint32_t drop_before, int skip_zero, int64_t clo,
                   int64_t chi, GroupScratch& s, cudaStream_t st);
void pileup_island_cols(uint32_t n_islands, GroupScratch& s, cudaStream_t st);
void pileup_phase2(const ReadsView& v, uint32_t g0, uint32_t g1, uint32_t n_islands, uint32_t n_col, GroupScratch& s,
                   ColumnScratch& c, ColumnOutput& o, cudaStream_t st);
// redo (may be null): only the chunks of 32 columns flagged there are done (the others were done by the tile kernel)
void pileup_entries(const ReadsView& v, uint32_t n_col, GroupScratch& s, ColumnScratch& c, ColumnOutput& o, cudaStream_t st,
                    const uint32_t* redo = nullptr);
// column-stationary form (pileup.cu: entries_tile_kernel); mode 0 explicit, 1 compact (last_read / live_mask / nstrag
// written directly, no read_idx), 2 MAQ.  redo[pileup_tile_chunks(n_col)]: chunks left to pileup_entries.
uint32_t pileup_tile_chunks(uint32_t n_col);
void pileup_entries_tile(const ReadsView& v, uint32_t n_col, GroupScratch& s, ColumnScratch& c, ColumnOutput& o, int mode,
                         uint32_t* last_read, uint64_t* live_mask, uint32_t* nstrag, uint32_t* redo, cudaStream_t st);
void pileup_strag_walk(const ReadsView& v, uint32_t n_col, GroupScratch& s, const uint32_t* lo, const uint32_t* hi,
                       const uint64_t* col_pos, const uint32_t* strag_off, uint32_t* strag_col, uint32_t* strag_idx,
                       cudaStream_t st);
void pileup_compact_masks(uint32_t n_col, const ColumnOutput& o, uint32_t* last_read, uint64_t* mask, uint32_t* nstrag,
                          uint32_t* strag_off, GroupScratch& s, cudaStream_t st, const uint32_t* redo = nullptr);
void pileup_compact_stragglers(uint32_t n_col, const ColumnOutput& o, const uint32_t* strag_off, uint32_t* strag_col,
                               uint32_t* strag_idx, cudaStream_t st);
uint32_t pileup_pack_blocks(uint64_t n_entries);
void pileup_pack_bases(uint64_t n_entries, const uint8_t* base, uint8_t* base4, uint32_t* block_special, uint32_t* block_off,
                       GroupScratch& s, uint32_t* scan_tmp, cudaStream_t st);
void pileup_pack_specials(uint64_t n_entries, const uint8_t* base, const uint32_t* block_off, uint32_t* special_entry,
                          uint8_t* special_base, cudaStream_t st);
void pileup_position_runs_scan(uint32_t n_col, const ColumnOutput& o, uint32_t* flag, uint32_t* incl, GroupScratch& s,
                               cudaStream_t st);
void pileup_position_runs_scatter(uint32_t n_col, const ColumnOutput& o, const uint32_t* flag, const uint32_t* incl,
                                  uint64_t* run_pos, uint32_t* run_first_col, cudaStream_t st);
void pileup_carry(const ReadsView& v, uint32_t g0, uint32_t g1, const int32_t* block_size, int64_t limit, GroupScratch& s,
                  CarryOut& out, cudaStream_t st);
void pileup_carry_copy(const ReadsView& v, uint32_t g0, uint32_t g1, const int32_t* block_size, uint32_t n_carry_out,
                       GroupScratch& s, CarryOut& out, cudaStream_t st);

}  // namespace biodb
