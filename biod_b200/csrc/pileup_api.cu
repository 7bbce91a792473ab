// Host side of the pileup: walks the file batch by batch (Pass), splits each batch into reference
// groups, runs the pileup kernels per group and carries reads that reach past the batch forward —
// the halo scheme of PileupChunkRange (bio/std/hts/bam/pileup.d:859-987): a batch emits columns up to
// (not including) the position of its last read, later columns wait for the next batch.
#include <algorithm>
#include <condition_variable>
#include <cstdlib>
#include <cstring>
#include <deque>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <memory>

#include "bai.h"
#include "maq.h"
#include "md_chain.h"
#include "runtime.h"
#include "scan.cuh"

using namespace biodb;

namespace {

struct CarryBufs {
  DevBuf a[9], data;     // pos,end,ref,rec_off,bin_mq_nl,flag_nc,l_seq,block_size,gidx
  uint32_t n = 0;
  uint64_t bytes = 0;
  CarryOut out() const {
    CarryOut o;
    o.pos = a[0].as<int32_t>(); o.end_pos = a[1].as<int32_t>(); o.ref_id = a[2].as<int32_t>();
    o.rec_off = a[3].as<uint64_t>(); o.bin_mq_nl = a[4].as<uint32_t>(); o.flag_nc = a[5].as<uint32_t>();
    o.l_seq = a[6].as<int32_t>(); o.block_size = a[7].as<int32_t>(); o.gidx = a[8].as<uint32_t>();
    o.data = data.as<uint8_t>();
    return o;
  }
};

// One set of output buffers: device side written by the pileup kernels, host side (pinned) filled by an
// asynchronous copy on the copy stream.  Two sets alternate so that batch k+1 is computed and copied while
// the caller still reads batch k.
struct OutSet {
  DevBuf d[21];     // col_pos, col_off, nstart, read_idx, base, qual, qoff, counts, last_read, live_mask, strag_off,
                    // strag_idx, strag_col, run_pos, run_first_col, base4, block_special, block_off, special_entry, special_base,
                    // reference_base
  PinBuf h[17];     // col_pos, col_off, nstart, read_idx, base|qual, qoff, counts, last_read, live_mask, run_pos,
                    // strag_idx, strag_col, run_first_col, base4, special_entry, special_base, reference_base
  // maq_mode: per column gt0, gt1, s0, s1, n_valid; calls: col, pos, gt, ref, qual
  DevBuf dm[10];
  PinBuf hm[10];
  uint32_t n_calls = 0;
  uint32_t n_strag = 0, n_runs = 0, n_special = 0;
  size_t col_cap = 0, ent_cap = 0;
  cudaEvent_t computed = nullptr, done = nullptr;
};

struct Desc {       // what the producer hands to biodb_pileup_next
  biodb_status st = BIODB_OK;
  biodb_error err{};
  biodb_column_batch cols{};
  int set = -1;
};

}  // namespace

struct biodb_pileup {
  biodb_reader* r = nullptr;
  biodb_pileup_params prm{};
  Pass pass;
  CarryBufs carry[2];
  int cur = 0;
  // continuation of a reference group across batches
  bool cont = false;          // the last batch ended inside a group that had live reads
  int32_t cont_ref = 0;
  int64_t cont_pos = 0;       // columns below this position are already emitted
  // shard mode (biodb_pileup_begin_shard)
  bool sharded = false;
  biodb_shard_info shard{};
  bool own_reached = false;          // the first own record has been met: index_bias is final
  uint64_t index_bias = 0;           // records of the halo: read_idx counts from the first record read
  uint32_t shard_index = 0, shard_count = 0, shard_last = 0;   // the pass covers shards [shard_index, shard_last)
  std::vector<int32_t> later_ref;    // (ref, pos) of the cuts behind this shard, i.e. of shards shard_index+1 ...
  std::vector<int64_t> later_pos;
  std::vector<uint64_t> reach;       // [shard_count] first own record reaching into each later shard (voffset), ~0 = none
  DevBuf d_later, d_reach;           // the keys / the per-batch minima on the device
  PinBuf h_reach;
  DevBuf pack_tmp;                   // scan scratch of the base packing (compact_reads)
  // reference bases from MD tags (use_md_tag; mdtag.cu, md_chain.h)
  std::unique_ptr<MdChain> md;       // which read's dna() serves which positions; its state runs across batches
  std::vector<MdSegment> md_segs;    // segments of the group being produced
  DevBuf maq_ent[2];                 // maq_mode: the entries as the caller sees them (base | strand, min(quality, mapq))
  DevBuf redo;                       // chunks of 32 columns the tile kernel left to the read-stationary one
  int tile_mode = 1;                 // BIODB_PILEUP_TILE: 0 = the read-stationary kernel alone, 2 = the column-stationary
                                     // kernel for every form, default 1 = for the forms that do not deliver read_idx
                                     // (compact_reads, MAQ), where it saves writing and re-reading 4 bytes per entry
  DevBuf md_len, md_dsegs, md_keep_buf[2][2];
  PinBuf md_h, md_hsegs;
  MdKeep md_keep{{~0ull, ~0ull}, {0, 0}, {nullptr, nullptr}};   // providers materialised by the previous batch
  int md_keep_set = 0;
  // maq_mode (maq.h)
  MaqParams maq;
  MaqDevTables maq_tab{nullptr, nullptr, nullptr};
  // region mode (biodb_pileup_begin_region): the reads are those of bam[ref][beg .. end) — the index's chunks read one
  // after the other, every batch reduced to the reads of the region (region.cu) before the pileup sees it
  bool region = false, region_done = false;
  std::vector<VoChunk> chunks;
  size_t chunk_i = 0;
  uint32_t reg_ref = 0, reg_beg = 0, reg_end = 0;
  uint64_t region_index = 0;         // reads of the region handed to the pileup so far (read_idx counts these)
  DevBuf d_sel[10], d_rscratch, d_rinfo;
  uint64_t sel_cap = 0, sel_cig_cap = 0;
  // single_ref state
  bool started = false, done = false;
  int32_t target_ref = -1;
  // current batch
  bool have_batch = false;
  std::vector<uint32_t> bounds;   // group boundaries of the current view, [0, ..., n]
  size_t gi = 0;
  uint32_t n_view = 0, n_carry_view = 0;
  uint64_t first_index = 0;
  // scratch
  DevBuf g[13], tmp[4], info, bnd, cs[3];
  size_t read_cap = 0, col_cap = 0;
  PinBuf h_small, h_bnd;
  uint64_t tot_cols = 0, tot_entries = 0;
  // pipelining: producer thread + two output sets
  OutSet sets[2];
  cudaStream_t copy_st = nullptr;
  std::thread worker;
  std::mutex mu;
  std::condition_variable cv;
  std::deque<Desc> ready;
  int free_sets = 2;         // sets the producer may still fill
  int next_set = 0;
  int held_set = -1;         // set whose pointers the caller currently holds
  bool stop = false, worker_started = false, worker_done = false;
  cudaEvent_t last_done = nullptr;

  ~biodb_pileup();
  void shutdown();
  void reset(const biodb_pileup_params* p);

  biodb_status fail(int status, const std::string& msg) { return pass.fail(status, 0, 0, msg); }
};

#define PL_TRY(expr)                                                                            \
  do {                                                                                          \
    cudaError_t _e = (expr);                                                                    \
    if (_e != cudaSuccess) return pl->fail(BIODB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

static ReadsView make_view(biodb_pileup* pl) {
  Pass& p = pl->pass;
  ReadsView v;
  RecordArrays a = p.arrays(0);
  v.pos = a.pos; v.end_pos = a.end_pos; v.ref_id = a.ref_id; v.rec_off = a.rec_off;
  v.bin_mq_nl = a.bin_mq_nl; v.flag_nc = a.flag_nc; v.l_seq = a.l_seq;
  v.carry_gidx = pl->carry[pl->cur].a[8].as<uint32_t>();
  v.carry_data = pl->carry[pl->cur].data.as<uint8_t>();
  v.u = p.d_u.as<uint8_t>();
  v.n_carry = pl->n_carry_view;
  v.n = pl->n_view;
  v.first_index = pl->first_index;
  return v;
}

static biodb_status ensure_read_scratch(biodb_pileup* pl, size_t n) {
  if (n + 8 <= pl->read_cap) return BIODB_OK;
  size_t cap = n + n / 4 + 1024;
  cudaStream_t st = pl->pass.st;
  static const size_t esz[12] = {4, 4, 4, 4, 4, 4, 4, 8, 4, 4, 4, 8};   // eend pm flag iid1 i.start i.end i.first i.cs ncol colbase cflag/cslot(2x) cbytes
  for (int k = 0; k < 12; ++k) {
    size_t b = cap * esz[k] * (k == 10 ? 2 : 1);
    PL_TRY(pl->g[k].ensure(b, st));
  }
  PL_TRY(pl->g[12].ensure(cap * 16, st));
  pl->read_cap = cap;
  return BIODB_OK;
}

static biodb_status ensure_tmp(biodb_pileup* pl, size_t elems) {
  cudaStream_t st = pl->pass.st;
  size_t t = scan_temp_elems(elems) + 8;
  PL_TRY(pl->tmp[0].ensure(t * 4, st));
  PL_TRY(pl->tmp[1].ensure(t * 4, st));
  PL_TRY(pl->tmp[2].ensure(t * 4, st));
  PL_TRY(pl->tmp[3].ensure(t * 8, st));
  return BIODB_OK;
}

static GroupScratch scratch(biodb_pileup* pl) {
  GroupScratch s;
  s.eend = pl->g[0].as<int32_t>(); s.pm = pl->g[1].as<int32_t>(); s.flag = pl->g[2].as<uint32_t>();
  s.rinfo = pl->g[12].as<uint4>();
  s.iid1 = pl->g[3].as<uint32_t>();
  s.islands.start = pl->g[4].as<int32_t>(); s.islands.end = pl->g[5].as<int32_t>();
  s.islands.first = pl->g[6].as<uint32_t>(); s.islands.cs = pl->g[7].as<int64_t>();
  s.ncol = pl->g[8].as<uint32_t>(); s.colbase = pl->g[9].as<uint32_t>();
  s.cflag = pl->g[10].as<uint32_t>(); s.cslot = pl->g[10].as<uint32_t>() + pl->read_cap;
  s.cbytes = pl->g[11].as<uint64_t>();
  s.tmp_i32 = pl->tmp[0].as<int32_t>(); s.tmp_u32 = pl->tmp[1].as<uint32_t>(); s.tmp_u32b = pl->tmp[2].as<uint32_t>();
  s.tmp_u64 = pl->tmp[3].as<uint64_t>();
  s.info = pl->info.as<int32_t>();
  s.clo = s.chi = 0;
  return s;
}

// Region mode: reduce the batch the pass has just produced to the reads of the region (BamReadFilter,
// randomaccessmanager.d:366-462), in place, so that the pileup kernels see nothing else.
static biodb_status region_reduce(biodb_pileup* pl, uint64_t front) {
  Pass& p = pl->pass;
  cudaStream_t st = p.st;
  const bool more_chunks = pl->chunk_i + 1 < pl->chunks.size();
  if (p.n) {
    if (p.n > 0xfffffff0ull) return pl->fail(BIODB_ERR_NOMEM, "too many records in one batch; lower blocks_per_batch");
    if (p.n + 8 > pl->sel_cap || p.n_cigar + 8 > pl->sel_cig_cap) {
      pl->sel_cap = std::max<uint64_t>(pl->sel_cap, p.n + p.n / 8 + 1024);
      pl->sel_cig_cap = std::max<uint64_t>(pl->sel_cig_cap, p.n_cigar + p.n_cigar / 8 + 1024);
      static const size_t esz[10] = {8, 4, 4, 4, 4, 4, 4, 4, 8, 4};
      for (int a = 0; a < 9; ++a) PL_TRY(pl->d_sel[a].ensure((size_t)(pl->sel_cap + 2) * esz[a], st));
      PL_TRY(pl->d_sel[9].ensure((size_t)(pl->sel_cig_cap + 2) * 4, st));
    }
    PL_TRY(pl->d_rscratch.ensure(region_scratch_elems(p.n) * 4, st));
    PL_TRY(pl->d_rinfo.ensure(64, st));
    RecordArrays in = p.arrays(front);
    RecordArrays out{pl->d_sel[0].as<uint64_t>(), pl->d_sel[1].as<int32_t>(), pl->d_sel[2].as<int32_t>(), pl->d_sel[3].as<int32_t>(),
                     pl->d_sel[4].as<int32_t>(), pl->d_sel[5].as<uint32_t>(), pl->d_sel[6].as<uint32_t>(), pl->d_sel[7].as<int32_t>(),
                     pl->d_sel[8].as<uint64_t>(), pl->d_sel[9].as<uint32_t>(), pl->sel_cap, pl->sel_cig_cap};
    p.stage_begin();
    PL_TRY(launch_region_filter(in, p.n, pl->reg_ref, pl->reg_beg, pl->reg_end, out, pl->d_rscratch.as<uint32_t>(),
                                pl->d_rinfo.as<uint32_t>(), st));
    p.stage_end(&p.stats.scan_ms);
    uint32_t* hi = pl->h_small.as<uint32_t>() + 32;
    PL_TRY(launch_copy_bytes(hi, pl->d_rinfo.p, 12, st));
    PL_TRY(cudaStreamSynchronize(st));
    const uint64_t n_keep = hi[1], n_cig = hi[2];
    if (hi[0] != 0xffffffffu) pl->region_done = true;       // a read beyond the region: nothing further can overlap it
    // the kept records take the place of the batch's records
    static const size_t esz[8] = {8, 4, 4, 4, 4, 4, 4, 4};
    void* dst[8] = {in.rec_off, in.block_size, in.ref_id, in.pos, in.end_pos, in.bin_mq_nl, in.flag_nc, in.l_seq};
    for (int a = 0; a < 8; ++a)
      if (n_keep) PL_TRY(launch_copy_bytes(dst[a], pl->d_sel[a].p, (size_t)n_keep * esz[a], st));
    PL_TRY(launch_copy_bytes(in.cigar_off, pl->d_sel[8].p, (size_t)(n_keep + 1) * 8, st));
    if (n_cig) PL_TRY(launch_copy_bytes(in.cigar, pl->d_sel[9].p, (size_t)n_cig * 4, st));
    p.n = n_keep;
    p.n_cigar = n_cig;
    pl->region_index += n_keep;
  }
  if (pl->region_done) {
    // the rest of this chunk and the chunks behind it are not read
    p.finished = true;
    p.final_slice = true;
    pl->chunk_i = pl->chunks.size();
  } else if (p.final_slice && more_chunks && !p.pending.status) {
    p.final_slice = false;                                  // the next chunk goes on with reads of the same reference
  }
  return BIODB_OK;
}

// Fetch the next batch of records and lay the carried reads in front of it.
static biodb_status load_batch(biodb_pileup* pl) {
  Pass& p = pl->pass;
  CarryBufs& c = pl->carry[pl->cur];
  uint64_t first = p.n_records_total;
  if (pl->region) {
    // a chunk's pass is over: go on with the next chunk (the pass then hands out that chunk's first batch)
    while (p.finished && !p.pending.status && !pl->region_done && pl->chunk_i + 1 < pl->chunks.size()) {
      const VoChunk& ch = pl->chunks[++pl->chunk_i];
      p.rewind(ch.beg >> 16, (uint32_t)(ch.beg & 0xFFFF));
      p.stop_coffset = ch.end >> 16;
      p.stop_uoffset = (uint32_t)(ch.end & 0xFFFF);
    }
    first = pl->region_index;
  }
  biodb_status s = p.next((uint32_t)p.r->opts.blocks_per_batch, c.n);
  if (s != BIODB_OK) return s;
  if (pl->region) {
    biodb_status rs = region_reduce(pl, c.n);
    if (rs != BIODB_OK) return rs;
  }
  if (pl->sharded && p.n) {
    // which of this batch's records are the shard's own (start at or after its first voffset)?
    uint64_t first_own = 0;                                  // index within the batch
    if (!pl->own_reached) {
      const uint64_t oc = pl->shard.first_voffset >> 16, ou = pl->shard.first_voffset & 0xFFFF;
      uint64_t x = ~0ull;                                    // slice offset of the first own byte
      for (const Seg& sg : p.segs) {
        if (sg.coffset > oc) { x = sg.ustart; break; }       // (only when the voffset was a block end)
        if (sg.coffset == oc && ou >= sg.within && ou < (uint64_t)sg.within + sg.len) { x = sg.ustart + (ou - sg.within); break; }
      }
      if (x == ~0ull) {
        first_own = p.n;                                     // the whole batch is halo
      } else {
        uint64_t* hh = (uint64_t*)pl->h_small.p;
        RecordArrays a = p.arrays(c.n);
        pileup_count_below(a.rec_off, (uint32_t)p.n, x, (uint64_t*)pl->info.p + 4, p.st);
        PL_TRY(launch_copy_bytes(hh + 8, (uint64_t*)pl->info.p + 4, 8, p.st));
        PL_TRY(cudaStreamSynchronize(p.st));
        first_own = hh[8];
        pl->own_reached = true;
      }
      pl->index_bias += first_own;
      pl->shard.n_halo_records = pl->index_bias;
    }
    // exact halos: the first own record that reaches across each later cut
    const uint32_t nk = (uint32_t)pl->later_ref.size();
    if (nk && first_own < p.n) {
      RecordArrays a = p.arrays(c.n);
      PL_TRY(cudaMemsetAsync(pl->d_reach.p, 0xff, (size_t)nk * 8, p.st));
      pileup_reach(a.ref_id, a.pos, a.end_pos, a.rec_off, (uint32_t)first_own, (uint32_t)p.n, pl->d_later.as<int32_t>(),
                   (const int64_t*)(pl->d_later.as<uint8_t>() + (((size_t)nk * 4 + 7) & ~(size_t)7)), nk,
                   pl->d_reach.as<unsigned long long>(), p.st);
      PL_TRY(launch_copy_bytes(pl->h_reach.p, pl->d_reach.p, (size_t)nk * 8, p.st));
      PL_TRY(cudaStreamSynchronize(p.st));
      const uint64_t* hr = pl->h_reach.as<uint64_t>();
      for (uint32_t k = 0; k < nk; ++k) {
        uint64_t& slot = pl->reach[pl->shard_last + k];
        if (hr[k] != ~0ull && slot == ~0ull) slot = p.voffset_of(hr[k]);   // batches come in file order: the first hit is the minimum
      }
    }
  }
  pl->first_index = first;      // shard mode: counted from the first halo record; the stitch subtracts n_halo_records
  pl->n_carry_view = c.n;
  pl->n_view = c.n + (uint32_t)p.n;
  cudaStream_t st = p.st;
  if (c.n) {
    RecordArrays a = p.arrays(0);
    CarryOut o = c.out();
    const size_t n = c.n;
    PL_TRY(launch_copy_bytes(a.pos, o.pos, n * 4, st));
    PL_TRY(launch_copy_bytes(a.end_pos, o.end_pos, n * 4, st));
    PL_TRY(launch_copy_bytes(a.ref_id, o.ref_id, n * 4, st));
    PL_TRY(launch_copy_bytes(a.rec_off, o.rec_off, n * 8, st));
    PL_TRY(launch_copy_bytes(a.bin_mq_nl, o.bin_mq_nl, n * 4, st));
    PL_TRY(launch_copy_bytes(a.flag_nc, o.flag_nc, n * 4, st));
    PL_TRY(launch_copy_bytes(a.l_seq, o.l_seq, n * 4, st));
    PL_TRY(launch_copy_bytes(a.block_size, o.block_size, n * 4, st));
  }
  biodb_status es = ensure_read_scratch(pl, pl->n_view);
  if (es != BIODB_OK) return es;
  es = ensure_tmp(pl, pl->n_view + 1);
  if (es != BIODB_OK) return es;
  // reference groups
  pl->bounds.clear();
  pl->bounds.push_back(0);
  if (pl->n_view) {
    const uint32_t cap = pl->n_view;
    PL_TRY(pl->bnd.ensure((size_t)(cap + 2) * 4, st));
    PL_TRY(pl->h_bnd.ensure((size_t)(cap + 2) * 4));
    ReadsView v = make_view(pl);
    uint32_t* d_b = pl->bnd.as<uint32_t>();
    pileup_find_groups(v, d_b + 1, d_b, cap, st);
    PL_TRY(launch_copy_bytes(pl->h_bnd.p, d_b, 4, st));
    PL_TRY(cudaStreamSynchronize(st));
    uint32_t nbnd = pl->h_bnd.as<uint32_t>()[0];
    if (nbnd) {
      PL_TRY(launch_copy_bytes(pl->h_bnd.as<uint32_t>() + 1, d_b + 1, (size_t)nbnd * 4, st));
      PL_TRY(cudaStreamSynchronize(st));
      uint32_t* hb = pl->h_bnd.as<uint32_t>() + 1;
      std::sort(hb, hb + nbnd);
      for (uint32_t k = 0; k < nbnd; ++k) pl->bounds.push_back(hb[k]);
    }
    pl->bounds.push_back(pl->n_view);
  }
  pl->gi = 0;
  pl->have_batch = true;
  return BIODB_OK;
}

biodb_pileup::~biodb_pileup() {
  shutdown();
  for (OutSet& o : sets) {
    if (o.computed) cudaEventDestroy(o.computed);
    if (o.done) cudaEventDestroy(o.done);
  }
  if (copy_st) cudaStreamDestroy(copy_st);
}

void biodb_pileup::shutdown() {
  if (worker_started) {
    {
      std::lock_guard<std::mutex> lk(mu);
      stop = true;
    }
    cv.notify_all();
    if (worker.joinable()) worker.join();
    worker_started = false;
  }
  if (pass.st) cudaStreamSynchronize(pass.st);
  if (copy_st) cudaStreamSynchronize(copy_st);
}

void biodb_pileup::reset(const biodb_pileup_params* p) {
  if (p) prm = *p;
  else { memset(&prm, 0, sizeof prm); prm.skip_zero_coverage = 1; prm.single_ref = 1; prm.end_at = ~0ull; }
  if (!prm.single_ref) { prm.start_from = 0; prm.end_at = ~0ull; }
  carry[0].n = carry[1].n = 0;
  cur = 0;
  cont = false;
  started = done = false;
  target_ref = -1;
  have_batch = false;
  bounds.clear();
  gi = 0;
  tot_cols = tot_entries = 0;
  ready.clear();
  free_sets = 2;
  next_set = 0;
  held_set = -1;
  stop = worker_done = false;
  last_done = nullptr;
  sharded = false;
  memset(&shard, 0, sizeof shard);
  own_reached = false;
  index_bias = 0;
  shard_index = shard_count = shard_last = 0;
  later_ref.clear();
  later_pos.clear();
  reach.clear();
  region = region_done = false;
  chunks.clear();
  chunk_i = 0;
  region_index = 0;
  md.reset(prm.use_md_tag ? new MdChain(prm.skip_zero_coverage != 0) : nullptr);
  maq = MaqParams();
  maq_tab = MaqDevTables{nullptr, nullptr, nullptr};
  {
    const char* e = getenv("BIODB_PILEUP_TILE");
    tile_mode = (e && e[0] >= '0' && e[0] <= '2') ? e[0] - '0' : 1;
  }
  md_segs.clear();
  md_keep.id[0] = md_keep.id[1] = ~0ull;
  md_keep_set = 0;
}

static biodb_status produce(biodb_pileup* pl, OutSet& os, biodb_column_batch* cols);

static void producer_main(biodb_pileup* pl) {
  cudaSetDevice(pl->r->device);
  while (true) {
    int set;
    {
      std::unique_lock<std::mutex> lk(pl->mu);
      pl->cv.wait(lk, [&] { return pl->stop || pl->free_sets > 0; });
      if (pl->stop) break;
      --pl->free_sets;
      set = pl->next_set;
      pl->next_set ^= 1;
    }
    Desc d;
    d.set = set;
    d.st = produce(pl, pl->sets[set], &d.cols);
    if (d.st != BIODB_OK) {
      cudaStreamSynchronize(pl->pass.st);
      pl->pass.collect_timing();
      if (d.st != BIODB_EOF) d.err = pl->r->err;
    }
    {
      std::lock_guard<std::mutex> lk(pl->mu);
      pl->ready.push_back(d);
      if (d.st != BIODB_OK) pl->worker_done = true;
    }
    pl->cv.notify_all();
    if (d.st != BIODB_OK) break;
  }
}

extern "C" {

biodb_status biodb_pileup_begin(biodb_reader* r, const biodb_pileup_params* p, biodb_pileup** out) {
  if (!r || !out) return BIODB_ERR_ARG;
  biodb_pileup* pl = nullptr;
  {
    std::lock_guard<std::mutex> lk(r->pool_mu);
    if (!r->pileup_pool.empty()) { pl = (biodb_pileup*)r->pileup_pool.back(); r->pileup_pool.pop_back(); }
  }
  if (pl) {
    pl->reset(p);
    pl->pass.rewind(r->reads_start_coffset, r->reads_start_uoffset);
    *out = pl;
    return BIODB_OK;
  }
  pl = new biodb_pileup;
  pl->r = r;
  pl->reset(p);
  biodb_status s = pl->pass.init(r, r->reads_start_coffset, r->reads_start_uoffset);
  if (s == BIODB_OK && (pl->info.ensure(64) != cudaSuccess || pl->h_small.ensure(256) != cudaSuccess)) s = BIODB_ERR_CUDA;
  if (s == BIODB_OK && cudaStreamCreateWithFlags(&pl->copy_st, cudaStreamNonBlocking) != cudaSuccess) s = BIODB_ERR_CUDA;
  for (OutSet& o : pl->sets)
    if (s == BIODB_OK && (cudaEventCreate(&o.computed) != cudaSuccess || cudaEventCreate(&o.done) != cudaSuccess)) s = BIODB_ERR_CUDA;
  if (s != BIODB_OK) { delete pl; return s; }
  *out = pl;
  return BIODB_OK;
}

biodb_status biodb_pileup_maq_params(biodb_pileup* pl, const biodb_maq_params* mp) {
  if (!pl || pl->worker_started) return BIODB_ERR_ARG;
  pl->maq = MaqParams();
  if (mp) {
    pl->maq.depcorr = mp->depcorr; pl->maq.eta = mp->eta; pl->maq.minimum_call_quality = mp->minimum_call_quality;
    pl->maq.minimum_base_quality = mp->minimum_base_quality;
  }
  if (pl->maq.minimum_base_quality < 0 || pl->maq.minimum_base_quality > 255) return BIODB_ERR_ARG;
  pl->maq_tab = MaqDevTables{nullptr, nullptr, nullptr};
  return BIODB_OK;
}

// The MAQ coefficient tables of (depcorr, eta) on the device: computed on the host the way ErrorModelCoefficients'
// constructor does (maq.d:80-120) once per reader and pair of values.
static biodb_status maq_tables(biodb_pileup* pl) {
  if (pl->maq_tab.fk) return BIODB_OK;
  biodb_reader* r = pl->r;
  std::lock_guard<std::mutex> lk(r->pool_mu);
  for (auto& c : r->maq_cache)
    if (c->depcorr == pl->maq.depcorr && c->eta == pl->maq.eta) {
      pl->maq_tab = MaqDevTables{c->fk.as<double>(), c->beta.as<double>(), c->lhet.as<double>()};
      return BIODB_OK;
    }
  // MaqSnpCaller keeps depcorr / eta as floats and hands them to ErrorModel(float, float) (maq.d:329-338)
  MaqTablesHost h((double)pl->maq.depcorr, (double)pl->maq.eta);
  std::unique_ptr<biodb_reader::MaqCache> c(new biodb_reader::MaqCache);
  c->depcorr = pl->maq.depcorr;
  c->eta = pl->maq.eta;
  if (c->fk.ensure(h.fk.size() * 8) != cudaSuccess || c->beta.ensure(h.beta.size() * 8) != cudaSuccess ||
      c->lhet.ensure(h.lhet.size() * 8) != cudaSuccess ||
      cudaMemcpy(c->fk.p, h.fk.data(), h.fk.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(c->beta.p, h.beta.data(), h.beta.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess ||
      cudaMemcpy(c->lhet.p, h.lhet.data(), h.lhet.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess)
    return BIODB_ERR_CUDA;
  pl->maq_tab = MaqDevTables{c->fk.as<double>(), c->beta.as<double>(), c->lhet.as<double>()};
  r->maq_cache.push_back(std::move(c));
  return BIODB_OK;
}

// makePileup(bam[ref][beg .. end), ...) (examples/read_bam_file.d:22-25): the pileup of the reads that overlap a region,
// fetched through the BAI index.  read_idx counts the reads of the region (the order biodb_reads_begin_region yields).
biodb_status biodb_pileup_begin_region(biodb_reader* r, const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end,
                                       const biodb_pileup_params* prm, biodb_pileup** out) {
  if (!r || !ix || !out || beg >= end) return BIODB_ERR_ARG;
  std::vector<VoChunk> c;
  if (biodb_index_region_chunks(ix, ref_id, beg, end, &c) != BIODB_OK) return BIODB_ERR_ARG;
  biodb_status s = biodb_pileup_begin(r, prm, out);
  if (s != BIODB_OK) return s;
  biodb_pileup* pl = *out;
  pl->region = true;
  pl->chunks.swap(c);
  pl->chunk_i = 0;
  pl->reg_ref = ref_id;
  pl->reg_beg = beg;
  pl->reg_end = end;
  if (pl->chunks.empty()) {
    pl->done = true;                                        // no chunk can hold a read of the region
  } else {
    const VoChunk& ch = pl->chunks[0];
    pl->pass.rewind(ch.beg >> 16, (uint32_t)(ch.beg & 0xFFFF));
    pl->pass.stop_coffset = ch.end >> 16;
    pl->pass.stop_uoffset = (uint32_t)(ch.end & 0xFFFF);
  }
  return BIODB_OK;
}

// The first record that starts in or after the BGZF block at `coffset`: its virtual offset, reference and position.
// In a file whose records straddle blocks the chain's entry into the block is searched for (Pass::entry_search).
// *vo = ~0 when no record follows.
static biodb_status peek_first_record(biodb_reader* r, uint64_t coffset, uint64_t* vo, int32_t* ref, int64_t* pos) {
  *vo = ~0ull; *ref = -1; *pos = INT64_MAX;
  Pass p;
  const bool at_start = coffset == r->reads_start_coffset;
  biodb_status s = p.init(r, coffset, at_start ? r->reads_start_uoffset : 0);
  if (s != BIODB_OK) return s;
  p.entry_search = !at_start;
  for (int tries = 0; tries < 64; ++tries) {
    s = p.next(4, 0);
    if (s == BIODB_EOF) return BIODB_OK;
    if (s != BIODB_OK) return s;
    if (p.n) break;
    if (p.finished) return BIODB_OK;
  }
  if (p.n == 0) return BIODB_OK;
  RecordArrays a = p.arrays(0);
  int32_t h[2];
  uint64_t off = 0;
  if (cudaMemcpy(&h[0], a.ref_id, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(&h[1], a.pos, 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
      cudaMemcpy(&off, a.rec_off, 8, cudaMemcpyDeviceToHost) != cudaSuccess)
    return p.fail(BIODB_ERR_CUDA, 0, 0, "peek failed");
  *vo = p.voffset_of(off);
  *ref = h[0];
  *pos = h[1];
  return BIODB_OK;
}

// biodb_shard_cuts: cut k = the first record in or after the block at compressed-byte fraction k/n of the data blocks
static biodb_status shard_cuts(biodb_reader* r, uint32_t n_shards, const biodb_reader::ShardCuts** out) {
  for (const auto& c : r->shard_cuts)
    if (c.n == n_shards) { *out = &c; return BIODB_OK; }
  biodb_status s = r->build_block_index();
  if (s != BIODB_OK) return s;
  const std::vector<uint64_t>& bi = r->block_index;
  const size_t nb = bi.size();
  const uint64_t end_vo = r->data_end_coffset << 16;
  biodb_reader::ShardCuts c;
  c.n = n_shards;
  for (uint32_t k = 0; k <= n_shards; ++k) {
    uint64_t vo = end_vo;
    int32_t ref = -1;
    int64_t pos = INT64_MAX;
    if (k < n_shards && nb) {
      size_t b = 0;
      if (k > 0) {
        const uint64_t a = bi[0], z = r->data_end_coffset;
        const uint64_t target = a + (uint64_t)((double)(z - a) * k / n_shards);
        b = (size_t)(std::lower_bound(bi.begin(), bi.end(), target) - bi.begin());
      }
      if (b < nb) {
        s = peek_first_record(r, bi[b], &vo, &ref, &pos);
        if (s != BIODB_OK) return s;
        if (vo == ~0ull) { vo = end_vo; ref = -1; pos = INT64_MAX; }
      }
    }
    if (!c.vo.empty() && vo < c.vo.back()) vo = c.vo.back();      // (cuts never run backwards)
    c.vo.push_back(vo);
    c.ref.push_back(ref);
    c.pos.push_back(pos);
  }
  r->shard_cuts.push_back(std::move(c));
  *out = &r->shard_cuts.back();
  return BIODB_OK;
}

biodb_status biodb_shard_cuts(biodb_reader* r, uint32_t n_shards, uint64_t* cut_voffset, int32_t* cut_ref, int64_t* cut_pos) {
  if (!r || n_shards == 0) return BIODB_ERR_ARG;
  const biodb_reader::ShardCuts* c = nullptr;
  biodb_status s = shard_cuts(r, n_shards, &c);
  if (s != BIODB_OK) return s;
  for (uint32_t k = 0; k <= n_shards; ++k) {
    if (cut_voffset) cut_voffset[k] = c->vo[k];
    if (cut_ref) cut_ref[k] = c->ref[k];
    if (cut_pos) cut_pos[k] = c->pos[k];
  }
  return BIODB_OK;
}

// common part of the shard / range passes: read [from_vo, to_vo), own records from own_vo on, columns in [lo, hi)
static biodb_status begin_clipped(biodb_reader* r, const biodb_pileup_params* prm, uint64_t from_vo, uint64_t own_vo, uint64_t to_vo,
                                  int32_t lo_ref, int64_t lo_pos, int32_t hi_ref, int64_t hi_pos, biodb_pileup** out) {
  biodb_pileup_params p2;
  if (prm) p2 = *prm; else { memset(&p2, 0, sizeof p2); p2.skip_zero_coverage = 1; }
  p2.single_ref = 0;                       // pileupColumns semantics
  p2.start_from = 0;
  p2.end_at = ~0ull;
  biodb_status s = biodb_pileup_begin(r, &p2, out);
  if (s != BIODB_OK) return s;
  biodb_pileup* pl = *out;
  pl->sharded = true;
  biodb_shard_info sh{};
  sh.first_voffset = own_vo;
  sh.end_voffset = to_vo;
  sh.halo_voffset = from_vo;
  sh.lo_ref = lo_ref; sh.lo_pos = lo_pos; sh.hi_ref = hi_ref; sh.hi_pos = hi_pos;
  pl->shard = sh;
  pl->own_reached = own_vo <= from_vo;
  pl->pass.rewind(from_vo >> 16, (uint32_t)(from_vo & 0xFFFF));
  if (to_vo != ~0ull) {
    pl->pass.stop_coffset = to_vo >> 16;
    pl->pass.stop_uoffset = (uint32_t)(to_vo & 0xFFFF);
  }
  if (to_vo != ~0ull && to_vo <= from_vo) pl->done = true;     // nothing to read
  return BIODB_OK;
}

biodb_status biodb_pileup_begin_range(biodb_reader* r, const biodb_pileup_params* prm, uint64_t from_voffset, uint64_t to_voffset,
                                      int32_t lo_ref, int64_t lo_pos, int32_t hi_ref, int64_t hi_pos, biodb_pileup** out) {
  if (!r || !out) return BIODB_ERR_ARG;
  return begin_clipped(r, prm, from_voffset, from_voffset, to_voffset, lo_ref, lo_pos, hi_ref, hi_pos, out);
}

// shards [shard, shard + count) of n_shards as ONE pass (count = 1: a single shard)
static biodb_status begin_shard_common(biodb_reader* r, const biodb_pileup_params* prm, uint32_t shard, uint32_t count,
                                       uint32_t n_shards, bool by_blocks, uint32_t halo_blocks, uint64_t halo_vo,
                                       biodb_pileup** out) {
  if (!r || !out || n_shards == 0 || count == 0 || shard >= n_shards || count > n_shards - shard) return BIODB_ERR_ARG;
  const uint32_t last = shard + count;                    // first shard behind the span
  const biodb_reader::ShardCuts* c = nullptr;
  biodb_status s = shard_cuts(r, n_shards, &c);
  if (s != BIODB_OK) return s;
  const uint64_t own = c->vo[shard], end = c->vo[last];
  uint64_t from = own;
  if (shard > 0 && by_blocks && halo_blocks) {
    // a guess: the halo starts halo_blocks BGZF blocks in front of the block the shard's first record starts in
    const std::vector<uint64_t>& bi = r->block_index;
    size_t b = (size_t)(std::upper_bound(bi.begin(), bi.end(), own >> 16) - bi.begin());   // blocks at or before own's
    b = b ? b - 1 : 0;
    const size_t hb = b > halo_blocks ? b - halo_blocks : 0;
    if (hb < bi.size() && bi[hb] < (own >> 16)) {
      int32_t rf; int64_t ps;
      s = peek_first_record(r, bi[hb], &from, &rf, &ps);
      if (s != BIODB_OK) return s;
      if (from > own) from = own;
    }
  } else if (shard > 0 && !by_blocks) {
    from = std::min(halo_vo, own);
  }
  // shard 0 also owns whatever precedes its first record's position; the last shard everything to the end
  const int32_t lo_ref = shard == 0 ? 0 : c->ref[shard];
  const int64_t lo_pos = shard == 0 ? INT64_MIN : c->pos[shard];
  s = begin_clipped(r, prm, from, own, end, lo_ref, lo_pos, c->ref[last], c->pos[last], out);
  if (s != BIODB_OK) return s;
  biodb_pileup* pl = *out;
  pl->shard_index = shard;
  pl->shard_count = n_shards;
  pl->reach.assign(n_shards, ~0ull);
  pl->shard_last = last;
  for (uint32_t t = last; t < n_shards; ++t) { pl->later_ref.push_back(c->ref[t]); pl->later_pos.push_back(c->pos[t]); }
  const size_t nk = pl->later_ref.size();
  if (nk) {
    const size_t poff = (nk * 4 + 7) & ~(size_t)7;
    std::vector<uint8_t> buf(poff + nk * 8);
    memcpy(buf.data(), pl->later_ref.data(), nk * 4);
    memcpy(buf.data() + poff, pl->later_pos.data(), nk * 8);
    if (pl->d_later.ensure(buf.size()) != cudaSuccess || pl->d_reach.ensure(nk * 8) != cudaSuccess ||
        pl->h_reach.ensure(nk * 8) != cudaSuccess ||
        cudaMemcpy(pl->d_later.p, buf.data(), buf.size(), cudaMemcpyHostToDevice) != cudaSuccess) {
      biodb_pileup_end(pl);
      return BIODB_ERR_CUDA;
    }
  }
  return BIODB_OK;
}

biodb_status biodb_pileup_begin_shard(biodb_reader* r, const biodb_pileup_params* prm, uint32_t shard, uint32_t n_shards,
                                      uint32_t halo_blocks, biodb_pileup** out) {
  return begin_shard_common(r, prm, shard, 1, n_shards, true, halo_blocks, 0, out);
}

biodb_status biodb_pileup_begin_shard_at(biodb_reader* r, const biodb_pileup_params* prm, uint32_t shard, uint32_t n_shards,
                                         uint64_t halo_voffset, biodb_pileup** out) {
  return begin_shard_common(r, prm, shard, 1, n_shards, false, 0, halo_voffset, out);
}

biodb_status biodb_pileup_begin_shard_span(biodb_reader* r, const biodb_pileup_params* prm, uint32_t first, uint32_t count,
                                           uint32_t n_shards, uint32_t halo_blocks, biodb_pileup** out) {
  return begin_shard_common(r, prm, first, count, n_shards, true, halo_blocks, 0, out);
}

biodb_status biodb_pileup_begin_shard_span_at(biodb_reader* r, const biodb_pileup_params* prm, uint32_t first, uint32_t count,
                                              uint32_t n_shards, uint64_t halo_voffset, biodb_pileup** out) {
  return begin_shard_common(r, prm, first, count, n_shards, false, 0, halo_voffset, out);
}

void biodb_pileup_shard_info(const biodb_pileup* pl, biodb_shard_info* out) {
  if (!pl || !out) return;
  *out = pl->shard;
  out->n_own_records = pl->pass.n_records_total - pl->index_bias;
}

void biodb_pileup_shard_reach(const biodb_pileup* pl, uint64_t* reach) {
  if (!pl || !reach) return;
  for (uint32_t t = 0; t < pl->shard_count; ++t) reach[t] = t < pl->reach.size() ? pl->reach[t] : ~0ull;
}

void biodb_pileup_end(biodb_pileup* pl) {
  if (!pl) return;
  pl->shutdown();
  biodb_reader* r = pl->r;
  {
    std::lock_guard<std::mutex> lk(r->pool_mu);
    if (r->pileup_pool.size() < 2) { r->pileup_pool.push_back(pl); return; }   // keep the buffers for the next pass
  }
  delete pl;
}
void biodb_pileup_destroy_pooled(void* p) { delete (biodb_pileup*)p; }
void biodb_pileup_stats(const biodb_pileup* pl, biodb_stats* out) {
  if (pl && out) *out = pl->pass.stats;
}
int32_t biodb_pileup_ref_id(const biodb_pileup* pl) { return pl ? pl->target_ref : -1; }
void biodb_pileup_totals(const biodb_pileup* pl, uint64_t* n_records, uint64_t* n_columns, uint64_t* n_entries) {
  if (!pl) return;
  if (n_records) *n_records = pl->region ? pl->region_index : pl->pass.n_records_total;
  if (n_columns) *n_columns = pl->tot_cols;
  if (n_entries) *n_entries = pl->tot_entries;
}

// Consumer side: hand out the next finished batch.  The producer thread works one batch ahead; the buffers of
// the batch returned here stay valid until the next call.
biodb_status biodb_pileup_next(biodb_pileup* pl, biodb_column_batch* cols) {
  if (!pl || !cols) return BIODB_ERR_ARG;
  memset(cols, 0, sizeof *cols);
  if (!pl->worker_started) {
    pl->worker_started = true;
    pl->worker = std::thread(producer_main, pl);
  }
  Desc d;
  {
    std::unique_lock<std::mutex> lk(pl->mu);
    if (pl->held_set >= 0) {            // the caller is done with the previous batch: its set may be refilled
      pl->held_set = -1;
      ++pl->free_sets;
      pl->cv.notify_all();
    }
    pl->cv.wait(lk, [&] { return !pl->ready.empty(); });
    d = pl->ready.front();
    if (d.st == BIODB_OK) pl->ready.pop_front();   // EOF / errors stay at the head: every later call repeats them
  }
  if (d.st == BIODB_OK) {
    OutSet& os = pl->sets[d.set];
    cudaEventSynchronize(os.done);
    pl->last_done = os.done;
    pl->held_set = d.set;
    *cols = d.cols;
    pl->tot_cols += d.cols.n_columns;
    pl->tot_entries += d.cols.n_entries;
    return BIODB_OK;
  }
  // end of the pass: the pass time runs from its first operation to the completion of the last copy
  if (pl->last_done && pl->pass.began) {
    float ms = 0;
    if (cudaEventElapsedTime(&ms, pl->pass.ev_begin, pl->last_done) == cudaSuccess && ms > pl->pass.stats.total_ms)
      pl->pass.stats.total_ms = ms;
  }
  if (d.st != BIODB_EOF) pl->r->err = d.err;
  return d.st;
}

}  // extern "C"

// Producer side: compute the next non-empty column batch into `os` and start its copy to the host.
static biodb_status produce(biodb_pileup* pl, OutSet& os, biodb_column_batch* cols) {
  memset(cols, 0, sizeof *cols);
  Pass& p = pl->pass;
  cudaStream_t st = p.st;
  const bool single = pl->prm.single_ref != 0;
  const int skip_zero = pl->prm.skip_zero_coverage != 0;
  while (true) {
    if (pl->done) return BIODB_EOF;
    if (!pl->have_batch || pl->gi + 1 >= pl->bounds.size()) {
      pl->have_batch = false;
      // (region mode: a finished pass is only the end of a chunk while further chunks wait)
      const bool more_chunks = pl->region && !pl->region_done && !p.pending.status && pl->chunk_i + 1 < pl->chunks.size();
      if (p.finished && !more_chunks && (pl->carry[pl->cur].n == 0 || p.pending.status != 0)) {
        biodb_status s = p.next(1, 0);          // raises the pending error, or EOF
        if (s == BIODB_EOF) pl->done = true;
        return s;
      }
      biodb_status s = load_batch(pl);
      if (s == BIODB_EOF) { pl->done = true; return s; }
      if (s != BIODB_OK) return s;
      if (pl->n_view == 0) continue;
    }
    // ---- one reference group of the current batch -------------------------------------------------
    const size_t gi = pl->gi++;
    const uint32_t g0 = pl->bounds[gi], g1 = pl->bounds[gi + 1];
    const bool last_group = gi + 2 == pl->bounds.size();
    // later batches may add reads to this group; with an error pending, stop where the reference's
    // sequential engine would have thrown (it needs the next read before it can finish the column)
    const bool trailing = last_group && (!p.final_slice || p.pending.status != 0);
    ReadsView v = make_view(pl);
    GroupScratch s = scratch(pl);
    int32_t* h = pl->h_small.as<int32_t>();
    // reference id of the group
    PL_TRY(launch_copy_bytes(h, v.ref_id + g0, 4, st));
    PL_TRY(cudaMemsetAsync(s.info, 0, 16, st));
    PL_TRY(cudaStreamSynchronize(st));
    const int32_t ref = h[0];
    const bool continues = pl->cont && gi == 0 && pl->cont_ref == ref;
    if (!continues && gi == 0) pl->cont = false;
    uint32_t drop_before = g0;
    if (single) {
      if (!pl->started) {
        // pileupInstance (pileup.d:482-493): drop leading reads that end before start_from; the first
        // read kept fixes the reference of the whole pileup
        uint32_t* d_first = (uint32_t*)(s.info + 2);
        pileup_first_kept(v, g0, g1, pl->prm.start_from, d_first, st);
        PL_TRY(launch_copy_bytes(h, d_first, 4, st));
        PL_TRY(cudaStreamSynchronize(st));
        uint32_t first = (uint32_t)h[0];
        if (first == 0xffffffffu) {
          // nothing kept in this group: carried reads (none can exist yet) and the group are skipped
          if (last_group) { pl->carry[pl->cur].n = 0; }
          continue;
        }
        pl->started = true;
        pl->target_ref = ref;
        drop_before = first;
      } else if (ref != pl->target_ref) {
        pl->done = true;                         // takeUntil!"a.ref_id != b" (pileup.d:494)
        return BIODB_EOF;
      }
    }
    int64_t clo = INT64_MIN, chi = INT64_MAX;
    if (continues) clo = pl->cont_pos;
    if (pl->sharded) {
      // column keys (ref, pos) of this shard: [lo, hi); unmapped (ref -1) sorts last
      auto key = [](int32_t rf) { return rf < 0 ? (int64_t)1 << 40 : (int64_t)rf; };
      const biodb_shard_info& sh = pl->shard;
      if (key(ref) < key(sh.lo_ref) || key(ref) > key(sh.hi_ref)) {
        if (last_group) { pl->carry[pl->cur].n = 0; pl->cont = false; }
        continue;                                    // a reference that belongs to a neighbouring shard
      }
      if (ref == sh.lo_ref) clo = std::max(clo, sh.lo_pos);
      if (ref == sh.hi_ref) chi = std::min(chi, sh.hi_pos);
    }
    if (single) {
      if (pl->prm.start_from > (uint64_t)INT64_MAX) { pl->done = true; return BIODB_EOF; }
      clo = std::max<int64_t>(clo, (int64_t)pl->prm.start_from);
      if (pl->prm.end_at <= (uint64_t)INT64_MAX) chi = (int64_t)pl->prm.end_at;
    }
    // ---- phase 1: liveness, running max, islands ----------------------------------------------------
    p.stage_begin();
    pileup_phase1(v, g0, g1, drop_before, skip_zero, clo, chi, s, st);
    p.stage_end(&p.stats.pileup_ms);
    // use_md_tag: length of dna(read) of the reads that are new in this batch (carried ones were fed to the chain by
    // the batch they came with), brought to the host with their positions by the sync below
    const bool use_md = pl->md != nullptr;
    const uint32_t md_a0 = std::max(g0, pl->n_carry_view);
    const size_t md_n = (use_md && md_a0 < g1) ? g1 - md_a0 : 0;
    if (md_n) {
      RecordArrays a = p.arrays(0);
      PL_TRY(pl->md_len.ensure(pl->read_cap * 4, st));
      PL_TRY(pl->md_h.ensure(md_n * 12 + 64));
      p.stage_begin();
      md_dna_lengths(v, a.block_size, s.eend, md_a0, g1, pl->md_len.as<int32_t>(), st);
      p.stage_end(&p.stats.pileup_ms);
      int32_t* hm = pl->md_h.as<int32_t>();
      PL_TRY(launch_copy_bytes(hm, v.pos + md_a0, md_n * 4, st));
      PL_TRY(launch_copy_bytes(hm + md_n, s.eend + md_a0, md_n * 4, st));
      PL_TRY(launch_copy_bytes(hm + 2 * md_n, pl->md_len.as<int32_t>() + md_a0, md_n * 4, st));
      p.stats.d2h_bytes += md_n * 12;
    }
    const uint32_t ng = g1 - g0;
    const uint32_t t_reads = (uint32_t)((ng + SCAN_TILE - 1) / SCAN_TILE);
    PL_TRY(launch_copy_bytes(h, s.tmp_u32 + t_reads, 4, st));           // n_islands
    PL_TRY(launch_copy_bytes(h + 1, s.info, 8, st));                    // status, last live + 1
    PL_TRY(cudaStreamSynchronize(st));
    const uint32_t n_islands = (uint32_t)h[0];
    const int32_t st1 = h[1];
    const uint32_t last_live1 = (uint32_t)h[2];
    if (st1 != 0)
      return pl->fail(st1, st1 == BIODB_ERR_CIGAR ? "Invalid read - CIGAR has no usable reference-consuming operation"
                                                  : "reads are not sorted by coordinate within the reference");
    CarryBufs& ccur = pl->carry[pl->cur];
    if (last_live1 == 0 || n_islands == 0) {
      // no live reads here.  A trailing group keeps any earlier continuation alive only if it is the same ref.
      if (last_group) ccur.n = 0;
      if (!continues) pl->cont = false;
      if (trailing && continues) pl->cont = true;
      continue;
    }
    int64_t E = 0;
    if (trailing) {
      PL_TRY(launch_copy_bytes(h, v.pos + (last_live1 - 1), 4, st));
      PL_TRY(cudaStreamSynchronize(st));
      E = h[0];
      chi = std::min(chi, E);
    }
    // use_md_tag: the chain over the new reads (host): segments "positions [first, first+count) read dna(read)[offset...]".
    // A group that later batches continue is drained up to E, the position its columns stop at; a complete one is
    // finished.  It needs nothing from the device beyond what the syncs above brought, and the device needs its segments
    // only for md_replay — so it runs on a thread of its own from here on, beside the kernels that build the columns
    // (and, in maq_mode, the likelihood kernel), and is joined where md_replay is launched.
    struct MdThread {
      std::thread t;
      void join() { if (t.joinable()) t.join(); }
      ~MdThread() { join(); }
    } md_thread;
    if (use_md) {
      pl->md_segs.clear();
      const int32_t* hm = pl->md_h.as<int32_t>();
      const uint64_t id0 = pl->first_index + (md_a0 - pl->n_carry_view);
      md_thread.t = std::thread([pl, hm, id0, ref, md_n, trailing, E]() {
        pl->md->admit_many(id0, ref, hm, hm + md_n, hm + 2 * md_n, md_n, &pl->md_segs);   // (end INT32_MIN = not a read of the pileup)
        if (trailing) pl->md->drain(E, &pl->md_segs);
        else pl->md->finish_reference(&pl->md_segs);
      });
    }
    s.clo = clo;
    s.chi = chi;
    // with skip_zero_coverage=false the column run of a continued group restarts exactly where the
    // previous batch stopped, even if no read covers that position (pileup.d:389-392)
    const bool force_lo = !skip_zero && continues;
    if (force_lo) {
      // island 0 of the group is the whole group: move its start down to the continuation point
      int32_t lo32 = (int32_t)std::max<int64_t>(clo, INT32_MIN);
      PL_TRY(launch_copy_bytes(h, s.islands.start, 4, st));
      PL_TRY(cudaStreamSynchronize(st));
      if (h[0] > lo32) {
        h[0] = lo32;
        PL_TRY(launch_copy_bytes(s.islands.start, h, 4, st));
      }
    }
    p.stage_begin();
    pileup_island_cols(n_islands, s, st);
    p.stage_end(&p.stats.pileup_ms);
    const uint32_t t_isl = (uint32_t)((n_islands + SCAN_TILE - 1) / SCAN_TILE);
    PL_TRY(launch_copy_bytes(h, s.tmp_u32b + t_isl, 4, st));
    PL_TRY(cudaStreamSynchronize(st));
    const uint32_t n_col = (uint32_t)h[0];
    if (n_col > 0x7ff00000u) return pl->fail(BIODB_ERR_NOMEM, "too many pileup columns in one batch; lower blocks_per_batch");
    // ---- phase 2: columns ---------------------------------------------------------------------------------
    uint64_t n_entries = 0;
    if (!n_col) md_thread.join();
    if (n_col) {
      if ((size_t)n_col + 8 > pl->col_cap) {
        size_t cap = (size_t)n_col + n_col / 4 + 1024;
        for (int k = 0; k < 3; ++k) PL_TRY(pl->cs[k].ensure(cap * 4, st));
        pl->col_cap = cap;
      }
      if ((size_t)n_col + 8 > os.col_cap) {
        size_t cap = (size_t)n_col + n_col / 4 + 1024;
        PL_TRY(os.d[0].ensure(cap * 8, st));
        PL_TRY(os.d[1].ensure(cap * 8, st));
        PL_TRY(os.d[2].ensure(cap * 4, st));
        if (!p.r->opts.device_output) {
          PL_TRY(os.h[0].ensure(cap * 8));
          PL_TRY(os.h[1].ensure(cap * 8));
          PL_TRY(os.h[2].ensure(cap * 4));
        }
        os.col_cap = cap;
      }
      biodb_status es = ensure_tmp(pl, std::max<size_t>(pl->n_view, n_col) + 2);
      if (es != BIODB_OK) return es;
      s = scratch(pl);
      s.clo = clo;
      s.chi = chi;
      ColumnScratch c{pl->cs[0].as<int32_t>(), os.d[2].as<uint32_t>(), pl->cs[1].as<uint32_t>(), pl->cs[2].as<uint32_t>()};
      ColumnOutput o{os.d[0].as<uint64_t>(), os.d[1].as<uint64_t>(), nullptr, nullptr, nullptr, nullptr, nullptr, -1};
      p.stage_begin();
      pileup_phase2(v, g0, g1, n_islands, n_col, s, c, o, st);
      p.stage_end(&p.stats.pileup_ms);
      PL_TRY(launch_copy_bytes(pl->h_small.p, o.col_off + n_col, 8, st));
      PL_TRY(cudaStreamSynchronize(st));
      n_entries = *pl->h_small.as<uint64_t>();
      const int maq_mode = pl->prm.maq_mode;
      const bool maq = maq_mode > 0;
      const bool counts_only = pl->prm.counts_only != 0 && !maq;
      const bool want_q = pl->prm.want_query_offset != 0 && !counts_only && !maq;
      const bool compact = pl->prm.compact_reads != 0 && !counts_only && !maq;
      if (compact && n_entries >= 0xfffffff0ull)      // special_entry[] indexes entries with 32 bits
        return pl->fail(BIODB_ERR_NOMEM, "too many column entries in one batch for the compact encoding; lower blocks_per_batch");
      if (counts_only) {
        PL_TRY(os.d[7].ensure(os.col_cap * 24, st));
        if (!p.r->opts.device_output) PL_TRY(os.h[6].ensure(os.col_cap * 24));
        o.counts = os.d[7].as<uint32_t>();
      }
      if (maq) {
        // the entries stay on the device, as the caller's view of them: base | strand, min(quality, mapping quality)
        biodb_status ms = maq_tables(pl);
        if (ms != BIODB_OK) return pl->fail(ms, "cannot set up the MAQ coefficient tables on the device");
        PL_TRY(pl->maq_ent[0].ensure((size_t)n_entries + 64, st));
        PL_TRY(pl->maq_ent[1].ensure((size_t)n_entries + 64, st));
        o.base = pl->maq_ent[0].as<uint8_t>();
        o.qual = pl->maq_ent[1].as<uint8_t>();
        o.maq_min_base_quality = pl->maq.minimum_base_quality;
      }
      if (!counts_only && !maq && n_entries + 8 > os.ent_cap) {
        size_t cap = (size_t)n_entries + n_entries / 8 + 4096;
        PL_TRY(os.d[3].ensure(cap * 4, st));
        PL_TRY(os.d[4].ensure(cap, st));
        PL_TRY(os.d[5].ensure(cap, st));
        if (!p.r->opts.device_output) {
          PL_TRY(os.h[3].ensure(cap * 4));
          PL_TRY(os.h[4].ensure(cap * 2));
        }
        os.ent_cap = cap;
      }
      if (want_q) {
        PL_TRY(os.d[6].ensure(os.ent_cap * 4, st));
        if (!p.r->opts.device_output) PL_TRY(os.h[5].ensure(os.ent_cap * 4));
      }
      if (!counts_only && !maq) {
        o.read_idx = os.d[3].as<uint32_t>();
        o.base = os.d[4].as<uint8_t>();
        o.qual = os.d[5].as<uint8_t>();
      }
      o.qoff = want_q ? os.d[6].as<uint32_t>() : nullptr;
      // entries: the column-stationary kernel for the plain / compact / MAQ forms (chunks deeper than its tile, and
      // the counts / query-offset forms, go to the read-stationary kernel)
      const bool tile = !counts_only && !want_q && (pl->tile_mode == 2 || (pl->tile_mode == 1 && (compact || maq)));
      uint32_t* redo = nullptr;
      if (tile) {
        PL_TRY(pl->redo.ensure((size_t)(pileup_tile_chunks(n_col) + 8) * 4, st));
        redo = pl->redo.as<uint32_t>();
        if (compact) {
          PL_TRY(os.d[8].ensure(os.col_cap * 4, st));
          PL_TRY(os.d[9].ensure(os.col_cap * 8, st));
        }
      }
      p.stage_begin();
      if (tile) {
        pileup_entries_tile(v, n_col, s, c, o, maq ? 2 : compact ? 1 : 0, os.d[8].as<uint32_t>(), os.d[9].as<uint64_t>(),
                            pl->cs[0].as<uint32_t>(), redo, st);
        pileup_entries(v, n_col, s, c, o, st, redo);
      } else {
        pileup_entries(v, n_col, s, c, o, st);
      }
      p.stage_end(&p.stats.pileup_ms);
      MaqColumns mq{nullptr, nullptr, nullptr, nullptr, nullptr};
      if (maq) {
        // genotype likelihoods of every column (maq.cu): they do not depend on the reference bases, so they go first
        PL_TRY(os.dm[0].ensure(os.col_cap, st));
        PL_TRY(os.dm[1].ensure(os.col_cap, st));
        PL_TRY(os.dm[2].ensure(os.col_cap * 4, st));
        PL_TRY(os.dm[3].ensure(os.col_cap * 4, st));
        PL_TRY(os.dm[4].ensure(os.col_cap * 2, st));
        mq = MaqColumns{os.dm[0].as<uint8_t>(), os.dm[1].as<uint8_t>(), os.dm[2].as<float>(), os.dm[3].as<float>(),
                        os.dm[4].as<uint16_t>()};
        p.stage_begin();
        maq_columns(o.col_off, o.base, o.qual, n_col, pl->maq_tab, mq, st);
        p.stage_end(&p.stats.pileup_ms);
      }
      md_thread.join();                        // the chain's segments are needed from here on
      if (use_md) {
        // reference_base: 'N' everywhere, then the providers' dna() replayed over the columns their segments cover
        static_assert(sizeof(MdSegment) == sizeof(MdSeg), "MdSegment is uploaded as MdSeg");
        const size_t nseg = pl->md_segs.size();
        PL_TRY(os.d[20].ensure(os.col_cap, st));
        if (!p.r->opts.device_output) PL_TRY(os.h[16].ensure(os.col_cap));
        PL_TRY(cudaMemsetAsync(os.d[20].p, 'N', n_col, st));
        if (nseg) {
          RecordArrays a = p.arrays(0);
          PL_TRY(pl->md_dsegs.ensure(nseg * sizeof(MdSeg), st));
          PL_TRY(pl->md_hsegs.ensure(nseg * sizeof(MdSeg)));
          memcpy(pl->md_hsegs.p, pl->md_segs.data(), nseg * sizeof(MdSeg));
          PL_TRY(launch_copy_bytes(pl->md_dsegs.p, pl->md_hsegs.p, nseg * sizeof(MdSeg), st));
          p.stats.h2d_bytes += nseg * sizeof(MdSeg);
          p.stage_begin();
          md_replay(v, a.block_size, pl->md_dsegs.as<MdSeg>(), (uint32_t)nseg, pl->md_keep, o.col_pos, n_col,
                    os.d[20].as<uint8_t>(), st);
          p.stage_end(&p.stats.pileup_ms);
        }
      }
      os.n_strag = 0;
      os.n_runs = 0;
      os.n_special = 0;
      os.n_calls = 0;
      if (maq) {
        // findSNPs' filter as a compaction
        uint32_t* flag = pl->cs[1].as<uint32_t>();          // the candidate windows (lo / hi) are dead after the entries kernel
        uint32_t* incl = pl->cs[2].as<uint32_t>();
        const uint8_t* refb = use_md ? os.d[20].as<uint8_t>() : nullptr;
        p.stage_begin();
        maq_call_flags(mq, refb, n_col, pl->maq.minimum_call_quality, flag, st);
        device_scan<true>(flag, incl, (uint64_t)n_col, s.tmp_u32b, OpAdd(), 0u, st);
        p.stage_end(&p.stats.pileup_ms);
        PL_TRY(launch_copy_bytes(pl->h_small.p, incl + (n_col - 1), 4, st));
        PL_TRY(cudaStreamSynchronize(st));
        os.n_calls = pl->h_small.as<uint32_t>()[0];
        static const size_t csz[5] = {4, 8, 1, 1, 4};
        for (int k = 0; k < 5; ++k) {
          PL_TRY(os.dm[5 + k].ensure((size_t)os.n_calls * csz[k] + 64, st));
          if (!p.r->opts.device_output) PL_TRY(os.hm[5 + k].ensure((size_t)os.n_calls * csz[k] + 64));
        }
        if (os.n_calls) {
          p.stage_begin();
          maq_call_gather(mq, refb, o.col_pos, n_col, flag, incl, os.dm[5].as<uint32_t>(), os.dm[6].as<uint64_t>(),
                          os.dm[7].as<uint8_t>(), os.dm[8].as<uint8_t>(), os.dm[9].as<float>(), st);
          p.stage_end(&p.stats.pileup_ms);
        }
        if (maq_mode >= 2 && !p.r->opts.device_output) {
          static const size_t msz[5] = {1, 1, 4, 4, 2};
          for (int k = 0; k < 5; ++k) PL_TRY(os.hm[k].ensure(os.col_cap * msz[k]));
        }
      }
      if (compact) {
        // sequential compact encoding (include/biod_b200.h): read lists as last read + window mask + stragglers
        // (compact_mask_kernel), positions as runs of consecutive positions (run_flag_kernel)
        PL_TRY(os.d[8].ensure(os.col_cap * 4, st));
        PL_TRY(os.d[9].ensure(os.col_cap * 8, st));
        PL_TRY(os.d[10].ensure((os.col_cap + 1) * 4, st));
        if (!p.r->opts.device_output) {
          PL_TRY(os.h[7].ensure(os.col_cap * 4));
          PL_TRY(os.h[8].ensure(os.col_cap * 8));
        }
        uint32_t* run_flag = pl->cs[1].as<uint32_t>();      // the candidate windows (lo / hi) are dead after the entries kernel
        uint32_t* run_incl = pl->cs[2].as<uint32_t>();
        p.stage_begin();
        pileup_compact_masks(n_col, o, os.d[8].as<uint32_t>(), os.d[9].as<uint64_t>(), pl->cs[0].as<uint32_t>(),
                             os.d[10].as<uint32_t>(), s, st, redo);
        // the stragglers of the columns that have any: walked out of the candidate reads now, while their windows (lo / hi)
        // are still there — the position-run scratch below reuses those arrays
        PL_TRY(launch_copy_bytes(pl->h_small.p, os.d[10].as<uint32_t>() + n_col, 4, st));
        PL_TRY(cudaStreamSynchronize(st));
        os.n_strag = pl->h_small.as<uint32_t>()[0];
        PL_TRY(os.d[11].ensure((size_t)os.n_strag * 4 + 64, st));
        PL_TRY(os.d[12].ensure((size_t)os.n_strag * 4 + 64, st));
        if (os.n_strag)
          pileup_strag_walk(v, n_col, s, c.lo, c.hi, o.col_pos, os.d[10].as<uint32_t>(), os.d[12].as<uint32_t>(),
                            os.d[11].as<uint32_t>(), st);
        pileup_position_runs_scan(n_col, o, run_flag, run_incl, s, st);
        // bases two per byte + the list of entries that are not a base ('-' in D / N, 0 past l_seq)
        const uint32_t npb = pileup_pack_blocks(n_entries);
        PL_TRY(os.d[15].ensure(os.ent_cap / 2 + 64, st));            // sized like the entry arrays (headroom included)
        PL_TRY(os.d[16].ensure((size_t)(pileup_pack_blocks(os.ent_cap) + 2) * 4, st));
        PL_TRY(os.d[17].ensure((size_t)(pileup_pack_blocks(os.ent_cap) + 2) * 4, st));
        PL_TRY(pl->pack_tmp.ensure((scan_temp_elems((uint64_t)pileup_pack_blocks(os.ent_cap) + 2) + 8) * 4, st));
        if (!p.r->opts.device_output) PL_TRY(os.h[13].ensure(os.ent_cap / 2 + 64));
        pileup_pack_bases(n_entries, o.base, os.d[15].as<uint8_t>(), os.d[16].as<uint32_t>(), os.d[17].as<uint32_t>(), s,
                          pl->pack_tmp.as<uint32_t>(), st);
        p.stage_end(&p.stats.pileup_ms);
        PL_TRY(launch_copy_bytes(pl->h_small.p, os.d[10].as<uint32_t>() + n_col, 4, st));
        PL_TRY(launch_copy_bytes(pl->h_small.as<uint32_t>() + 1, run_incl + (n_col - 1), 4, st));
        if (npb) PL_TRY(launch_copy_bytes(pl->h_small.as<uint32_t>() + 2, os.d[17].as<uint32_t>() + npb, 4, st));
        PL_TRY(cudaStreamSynchronize(st));
        os.n_runs = pl->h_small.as<uint32_t>()[1];
        os.n_special = npb ? pl->h_small.as<uint32_t>()[2] : 0;
        PL_TRY(os.d[18].ensure((size_t)os.n_special * 4 + 64, st));
        PL_TRY(os.d[19].ensure((size_t)os.n_special + 64, st));
        if (!p.r->opts.device_output) {
          PL_TRY(os.h[14].ensure((size_t)os.n_special * 4 + 64));
          PL_TRY(os.h[15].ensure((size_t)os.n_special + 64));
        }
        PL_TRY(os.d[13].ensure((size_t)os.n_runs * 8 + 64, st));
        PL_TRY(os.d[14].ensure((size_t)(os.n_runs + 1) * 4 + 64, st));
        if (!p.r->opts.device_output) {
          PL_TRY(os.h[9].ensure((size_t)os.n_runs * 8 + 64));
          PL_TRY(os.h[10].ensure((size_t)os.n_strag * 4 + 64));
          PL_TRY(os.h[11].ensure((size_t)os.n_strag * 4 + 64));
          PL_TRY(os.h[12].ensure((size_t)(os.n_runs + 1) * 4 + 64));
        }
        p.stage_begin();
        pileup_position_runs_scatter(n_col, o, run_flag, run_incl, os.d[13].as<uint64_t>(), os.d[14].as<uint32_t>(), st);
        if (os.n_special)
          pileup_pack_specials(n_entries, o.base, os.d[17].as<uint32_t>(), os.d[18].as<uint32_t>(), os.d[19].as<uint8_t>(), st);
        p.stage_end(&p.stats.pileup_ms);
      }
      PL_TRY(cudaEventRecord(os.computed, st));
      // results to the host on the copy stream: overlaps with the carry kernels and the next batch
      if (!p.r->opts.device_output) {
        cudaStream_t cs = pl->copy_st;
        PL_TRY(cudaStreamWaitEvent(cs, os.computed, 0));
        if (maq) {
          static const size_t csz[5] = {4, 8, 1, 1, 4}, msz[5] = {1, 1, 4, 4, 2};
          for (int k = 0; k < 5 && os.n_calls; ++k)
            PL_TRY(cudaMemcpyAsync(os.hm[5 + k].p, os.dm[5 + k].p, (size_t)os.n_calls * csz[k], cudaMemcpyDeviceToHost, cs));
          p.stats.d2h_bytes += (uint64_t)os.n_calls * 18;
          if (maq_mode >= 2) {
            PL_TRY(cudaMemcpyAsync(os.h[0].p, o.col_pos, (size_t)n_col * 8, cudaMemcpyDeviceToHost, cs));
            PL_TRY(cudaMemcpyAsync(os.h[1].p, o.col_off, (size_t)(n_col + 1) * 8, cudaMemcpyDeviceToHost, cs));
            PL_TRY(cudaMemcpyAsync(os.h[2].p, c.nstart, (size_t)n_col * 4, cudaMemcpyDeviceToHost, cs));
            for (int k = 0; k < 5; ++k)
              PL_TRY(cudaMemcpyAsync(os.hm[k].p, os.dm[k].p, (size_t)n_col * msz[k], cudaMemcpyDeviceToHost, cs));
            p.stats.d2h_bytes += (uint64_t)n_col * 32 + 8;
          }
        } else if (compact) {
          p.stats.d2h_bytes += (uint64_t)n_col * 16 + n_entries + (n_entries + 1) / 2 + n_entries * (want_q ? 4 : 0) +
                               (uint64_t)os.n_strag * 8 + (uint64_t)os.n_runs * 12 + 4 + (uint64_t)os.n_special * 5;
          PL_TRY(cudaMemcpyAsync(os.h[2].p, c.nstart, (size_t)n_col * 4, cudaMemcpyDeviceToHost, cs));
          PL_TRY(cudaMemcpyAsync(os.h[7].p, os.d[8].p, (size_t)n_col * 4, cudaMemcpyDeviceToHost, cs));
          PL_TRY(cudaMemcpyAsync(os.h[8].p, os.d[9].p, (size_t)n_col * 8, cudaMemcpyDeviceToHost, cs));
          PL_TRY(cudaMemcpyAsync(os.h[9].p, os.d[13].p, (size_t)os.n_runs * 8, cudaMemcpyDeviceToHost, cs));
          PL_TRY(cudaMemcpyAsync(os.h[12].p, os.d[14].p, (size_t)(os.n_runs + 1) * 4, cudaMemcpyDeviceToHost, cs));
          if (os.n_strag) {
            PL_TRY(cudaMemcpyAsync(os.h[10].p, os.d[11].p, (size_t)os.n_strag * 4, cudaMemcpyDeviceToHost, cs));
            PL_TRY(cudaMemcpyAsync(os.h[11].p, os.d[12].p, (size_t)os.n_strag * 4, cudaMemcpyDeviceToHost, cs));
          }
          PL_TRY(cudaMemcpyAsync(os.h[13].p, os.d[15].p, (size_t)(n_entries + 1) / 2, cudaMemcpyDeviceToHost, cs));
          if (os.n_special) {
            PL_TRY(cudaMemcpyAsync(os.h[14].p, os.d[18].p, (size_t)os.n_special * 4, cudaMemcpyDeviceToHost, cs));
            PL_TRY(cudaMemcpyAsync(os.h[15].p, os.d[19].p, (size_t)os.n_special, cudaMemcpyDeviceToHost, cs));
          }
          PL_TRY(cudaMemcpyAsync(os.h[4].as<uint8_t>() + n_entries, o.qual, (size_t)n_entries, cudaMemcpyDeviceToHost, cs));
        } else {
          p.stats.d2h_bytes += (uint64_t)n_col * 20 + 8 + (counts_only ? (uint64_t)n_col * 24 : n_entries * (6 + (want_q ? 4 : 0)));
          PL_TRY(cudaMemcpyAsync(os.h[0].p, o.col_pos, (size_t)n_col * 8, cudaMemcpyDeviceToHost, cs));
          PL_TRY(cudaMemcpyAsync(os.h[1].p, o.col_off, (size_t)(n_col + 1) * 8, cudaMemcpyDeviceToHost, cs));
          PL_TRY(cudaMemcpyAsync(os.h[2].p, c.nstart, (size_t)n_col * 4, cudaMemcpyDeviceToHost, cs));
          if (counts_only) {
            PL_TRY(cudaMemcpyAsync(os.h[6].p, o.counts, (size_t)n_col * 24, cudaMemcpyDeviceToHost, cs));
          } else {
            PL_TRY(cudaMemcpyAsync(os.h[3].p, o.read_idx, (size_t)n_entries * 4, cudaMemcpyDeviceToHost, cs));
            PL_TRY(cudaMemcpyAsync(os.h[4].p, o.base, (size_t)n_entries, cudaMemcpyDeviceToHost, cs));
            PL_TRY(cudaMemcpyAsync(os.h[4].as<uint8_t>() + n_entries, o.qual, (size_t)n_entries, cudaMemcpyDeviceToHost, cs));
          }
        }
        if (want_q) PL_TRY(cudaMemcpyAsync(os.h[5].p, o.qoff, (size_t)n_entries * 4, cudaMemcpyDeviceToHost, cs));
        if (use_md && !(maq && maq_mode < 2)) {
          PL_TRY(cudaMemcpyAsync(os.h[16].p, os.d[20].p, (size_t)n_col, cudaMemcpyDeviceToHost, cs));
          p.stats.d2h_bytes += n_col;
        }
        PL_TRY(cudaEventRecord(os.done, cs));
      } else {
        PL_TRY(cudaEventRecord(os.done, st));
      }
      PL_TRY(launch_copy_bytes(h + 8, s.info, 4, st));
      PL_TRY(cudaStreamSynchronize(st));
      if (h[8] != 0) return pl->fail(h[8], "Invalid read - query offset beyond the sequence while building a column");
    }
    if (use_md) {
      // providers whose dna() later batches may still ask for: write the strings out now, while the records are here
      MdKeep nk{{~0ull, ~0ull}, {0, 0}, {nullptr, nullptr}};
      if (trailing) {
        uint64_t ids[2];
        int64_t lens[2];
        const int np = pl->md->live_providers(ids, lens);
        const int set = pl->md_keep_set ^ 1;
        for (int k = 0; k < np; ++k) {
          nk.id[k] = ids[k];
          nk.len[k] = (uint32_t)std::min<int64_t>(std::max<int64_t>(lens[k], 0), 0x7fffffff);
          PL_TRY(pl->md_keep_buf[set][k].ensure((size_t)nk.len[k] + 16, st));
          nk.data[k] = pl->md_keep_buf[set][k].as<uint8_t>();
        }
        if (np) {
          RecordArrays a = p.arrays(0);
          md_keep(v, a.block_size, pl->md_keep, nk, st);
        }
        pl->md_keep_set = set;
      }
      pl->md_keep = nk;
    }
    // ---- carry reads that reach past this batch ---------------------------------------------------------------
    bool stop_after = false;
    if (single && pl->prm.end_at <= (uint64_t)INT64_MAX && trailing && E >= (int64_t)pl->prm.end_at) stop_after = true;
    if (last_group) {
      CarryBufs& nxt = pl->carry[pl->cur ^ 1];
      nxt.n = 0;
      nxt.bytes = 0;
      if (trailing && !stop_after) {
        RecordArrays a = p.arrays(0);
        CarryOut dummy{};
        p.stage_begin();
        pileup_carry(v, g0, g1, a.block_size, E, s, dummy, st);
        p.stage_end(&p.stats.pileup_ms);
        const uint32_t t_g = (uint32_t)((ng + SCAN_TILE - 1) / SCAN_TILE);
        PL_TRY(launch_copy_bytes(h, s.tmp_u32 + t_g, 4, st));
        PL_TRY(launch_copy_bytes(h + 2, s.tmp_u64 + t_g, 8, st));
        PL_TRY(cudaStreamSynchronize(st));
        const uint32_t nc = (uint32_t)h[0];
        const uint64_t nbytes = *(uint64_t*)(h + 2);
        if (nc) {
          static const size_t esz[9] = {4, 4, 4, 8, 4, 4, 4, 4, 4};
          for (int k = 0; k < 9; ++k) PL_TRY(nxt.a[k].ensure((size_t)(nc + 8) * esz[k], st));
          PL_TRY(nxt.data.ensure((size_t)nbytes + 256, st));
          CarryOut co = nxt.out();
          pileup_carry_copy(v, g0, g1, a.block_size, nc, s, co, st);
          PL_TRY(cudaStreamSynchronize(st));
        }
        nxt.n = nc;
        nxt.bytes = nbytes;
        pl->cont = true;
        pl->cont_ref = ref;
        pl->cont_pos = E;
      } else {
        pl->cont = false;
      }
      pl->cur ^= 1;
    }
    if (stop_after) pl->done = true;
    if (single && !trailing) pl->done = true;   // the group of the target reference is complete
    if (n_col == 0) {
      if (pl->done) return BIODB_EOF;
      continue;
    }
    p.mark_end();
    cols->n_columns = n_col;
    cols->n_entries = n_entries;
    cols->ref_id = ref;
    cols->last_of_pileup = pl->done ? 1 : 0;
    if (p.r->opts.device_output) {
      cols->position = os.d[0].as<uint64_t>();
      cols->col_off = os.d[1].as<uint64_t>();
      cols->n_starting_here = os.d[2].as<uint32_t>();
      cols->read_idx = os.d[3].as<uint32_t>();
      cols->base = os.d[4].as<uint8_t>();
      cols->qual = os.d[5].as<uint8_t>();
      cols->query_offset = pl->prm.want_query_offset ? os.d[6].as<uint32_t>() : nullptr;
      if (pl->prm.counts_only) {
        cols->read_idx = nullptr; cols->base = nullptr; cols->qual = nullptr; cols->query_offset = nullptr;
        cols->counts = os.d[7].as<uint32_t>();
      } else if (pl->prm.compact_reads) {
        cols->read_idx = nullptr; cols->position = nullptr; cols->col_off = nullptr; cols->base = nullptr;
        cols->base4 = os.d[15].as<uint8_t>();
        cols->n_special = os.n_special;
        cols->special_entry = os.d[18].as<uint32_t>();
        cols->special_base = os.d[19].as<uint8_t>();
        cols->last_read = os.d[8].as<uint32_t>();
        cols->live_mask = os.d[9].as<uint64_t>();
        cols->n_stragglers = os.n_strag;
        cols->strag_idx = os.d[11].as<uint32_t>();
        cols->strag_col = os.d[12].as<uint32_t>();
        cols->n_runs = os.n_runs;
        cols->run_pos = os.d[13].as<uint64_t>();
        cols->run_first_col = os.d[14].as<uint32_t>();
      }
      cols->reference_base = use_md ? os.d[20].as<uint8_t>() : nullptr;
      if (pl->prm.maq_mode > 0) {
        cols->read_idx = nullptr; cols->base = nullptr; cols->qual = nullptr; cols->query_offset = nullptr;
        cols->n_calls = os.n_calls;
        cols->call_col = os.dm[5].as<uint32_t>(); cols->call_pos = os.dm[6].as<uint64_t>(); cols->call_gt = os.dm[7].as<uint8_t>();
        cols->call_ref = os.dm[8].as<uint8_t>(); cols->call_qual = os.dm[9].as<float>();
        cols->maq_gt0 = os.dm[0].as<uint8_t>(); cols->maq_gt1 = os.dm[1].as<uint8_t>(); cols->maq_s0 = os.dm[2].as<float>();
        cols->maq_s1 = os.dm[3].as<float>(); cols->maq_n_valid = os.dm[4].as<uint16_t>();
      }
      return BIODB_OK;
    }
    cols->position = os.h[0].as<uint64_t>();
    cols->col_off = os.h[1].as<uint64_t>();
    cols->n_starting_here = os.h[2].as<uint32_t>();
    cols->read_idx = os.h[3].as<uint32_t>();
    cols->base = os.h[4].as<uint8_t>();
    cols->qual = os.h[4].as<uint8_t>() + n_entries;
    cols->query_offset = pl->prm.want_query_offset ? os.h[5].as<uint32_t>() : nullptr;
    if (pl->prm.counts_only) {
      cols->read_idx = nullptr; cols->base = nullptr; cols->qual = nullptr; cols->query_offset = nullptr;
      cols->counts = os.h[6].as<uint32_t>();
    } else if (pl->prm.compact_reads) {
      cols->read_idx = nullptr; cols->position = nullptr; cols->col_off = nullptr; cols->base = nullptr;
      cols->base4 = os.h[13].as<uint8_t>();
      cols->n_special = os.n_special;
      cols->special_entry = os.h[14].as<uint32_t>();
      cols->special_base = os.h[15].as<uint8_t>();
      cols->last_read = os.h[7].as<uint32_t>();
      cols->live_mask = os.h[8].as<uint64_t>();
      cols->n_stragglers = os.n_strag;
      cols->strag_idx = os.h[10].as<uint32_t>();
      cols->strag_col = os.h[11].as<uint32_t>();
      cols->n_runs = os.n_runs;
      cols->run_pos = os.h[9].as<uint64_t>();
      cols->run_first_col = os.h[12].as<uint32_t>();
    }
    cols->reference_base = use_md ? os.h[16].as<uint8_t>() : nullptr;
    if (pl->prm.maq_mode > 0) {
      cols->read_idx = nullptr; cols->base = nullptr; cols->qual = nullptr; cols->query_offset = nullptr;
      cols->n_calls = os.n_calls;
      cols->call_col = os.hm[5].as<uint32_t>(); cols->call_pos = os.hm[6].as<uint64_t>(); cols->call_gt = os.hm[7].as<uint8_t>();
      cols->call_ref = os.hm[8].as<uint8_t>(); cols->call_qual = os.hm[9].as<float>();
      if (pl->prm.maq_mode >= 2) {
        cols->maq_gt0 = os.hm[0].as<uint8_t>(); cols->maq_gt1 = os.hm[1].as<uint8_t>(); cols->maq_s0 = os.hm[2].as<float>();
        cols->maq_s1 = os.hm[3].as<float>(); cols->maq_n_valid = os.hm[4].as<uint16_t>();
      } else {
        cols->position = nullptr; cols->col_off = nullptr; cols->n_starting_here = nullptr;
        if (use_md) cols->reference_base = nullptr;       // (the calls carry their reference base)
      }
    }
    return BIODB_OK;
  }
}

