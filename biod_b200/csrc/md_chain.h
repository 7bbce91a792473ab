// Host side of the MD-tag reference bases (reference: PileupRangeUsingMdTag, pileup.d:522-654).
//
// BioD decides column by column which read's reconstructed reference (`dna(read)`, md/reconstruct.d:38-214) supplies
// PileupColumn.reference_base: the "chunk" of the current provider is consumed one base per column, and when it is used
// up the pending "next chunk provider" takes over.  Between two events — a read being admitted, a chunk running out —
// nothing happens but the cursor moving one base per column, so the whole thing is a sequential scan over the READS
// with O(1) state, not over the columns.  MdChain restates exactly that scan in position space: fed the live reads of
// a pileup in file order (position, end position, length of their dna() string), it emits segments
//     reference_base[P] = dna(read)[offset + (P - first)]   for P in [first, first + count)
// and every column position outside all segments has reference_base 'N'.  The device then only has to replay the
// providers' dna() strings into the column table ("N1 on the device").
//
// Nothing here touches the GPU; tests/test_md_chain.py checks it against a column-by-column restatement of the sweep.
#pragma once
#include <stdint.h>

#include <algorithm>
#include <vector>

namespace biodb {

struct MdSegment {
  int64_t first;      // first reference position served
  int64_t count;      // positions served
  uint64_t read;      // provider: the index the caller passed to admit()
  int64_t offset;     // index into dna(read) of the base for `first`
};

class MdChain {
 public:
  explicit MdChain(bool skip_zero_coverage) : skip_zero_(skip_zero_coverage) {}

  // Reads must come in file order; only reads the pileup keeps (basesCovered() > 0, inside the kept range) are given.
  void admit(uint64_t read, int32_t ref_id, int64_t pos, int64_t end, int64_t dna_len, std::vector<MdSegment>* out) {
    Rd r{read, pos, end, dna_len};
    if (!started_ || ref_id != ref_) {
      // a new reference: the sweep first runs to the end of the old one, then initNewReference() (pileup.d:598-611)
      if (started_) finish_island(out);
      started_ = true;
      ref_ = ref_id;
      begin_island(r, out);
      return;
    }
    if (pos == adm_pos_) {                    // admitted in the same column as the previous read
      if (settled_) {                         // (its step is already simulated and cannot end in a switch: see below)
        md_add(r, adm_had_zero_);
        max_end_ = std::max(max_end_, r.end);
      } else {
        pending_.push_back(r);
      }
      return;
    }
    flush_admissions(out);                    // columns up to and including the previous admission position are settled
    if (skip_zero_ && max_end_ < pos) {
      // no read covers the positions in front of this one: the sweep jumps (pileup.d:389-392)
      finish_island(out);
      begin_island(r, out);
      return;
    }
    adm_pos_ = pos;
    adm_init_ = false;
    adm_had_zero_ = max_end_ < pos;           // coverage of the previous column was 0 (only without skip_zero_coverage)
    // Fast path (nearly every read of a deep pileup): the chunk has bases left beyond this column, so neither the steps
    // up to it nor the step at it can end in a switch, whatever else is admitted at this position — the step can be
    // simulated right away instead of waiting for the next position to show up.
    if (have_chunk_ && ci_ + (pos - cur_) < chunk_.len) {
      ci_ += pos - cur_;
      cur_ = pos;
      md_add(r, adm_had_zero_);
      max_end_ = std::max(max_end_, r.end);
      settled_ = true;
      return;
    }
    settled_ = false;
    pending_.push_back(r);
  }

  // The live reads of one reference group of a batch, in file order: read k has index id0 + k, position pos[k], end
  // position end[k] (INT32_MIN = not a read of the pileup: skipped) and dna() length len[k].  Same result as admit()
  // read by read; runs of reads that fall inside the stretch the current chunk still serves — where nothing can happen
  // but the pending provider being replaced — go through a tight loop.
  void admit_many(uint64_t id0, int32_t ref_id, const int32_t* pos, const int32_t* end, const int32_t* len, size_t n,
                  std::vector<MdSegment>* out) {
    size_t k = 0;
    while (k < n) {
      if (end[k] == INT32_MIN) { ++k; continue; }
      if (started_ && !fin_ && ref_id == ref_ && curr_ref_ == ref_ && settled_ && pending_.empty() && have_chunk_ && ci_ < chunk_.len &&
          pos[k] >= cur_) {
        // positions up to last_safe are served by the chunk with bases to spare: no step there ends in a switch
        const int64_t last_safe = cur_ + (chunk_.len - ci_) - 1;
        int64_t at = cur_;                 // position of the last read taken
        bool hz = adm_had_zero_;
        const size_t k0 = k;
        // The stretch in one sweep: its reads (positions up to last_safe) and the largest end among them.  If the reads
        // admitted so far already reach past the stretch's last position and that position lies inside the chunk's
        // reference span, neither a jump of the sweep nor the rule of pileup.d:579 can come up in it, and all the
        // stretch does to the state is: the pending provider becomes the first read with the largest end, if that end
        // beats the chunk's and the pending provider's.
        size_t j = k;
        int32_t m = INT32_MIN;
        for (; j < n && pos[j] <= last_safe; ++j) m = end[j] > m ? end[j] : m;
        size_t jl = j;                     // one past the last read of the stretch that is a read of the pileup
        while (jl > k && end[jl - 1] == INT32_MIN) --jl;
        if (jl > k && max_end_ >= pos[jl - 1] && (uint32_t)pos[jl - 1] <= chunk_end_) {
          if ((uint32_t)m > chunk_end_) {
            if (!has_provider_ || m > (int32_t)provider_.end) {
              size_t a = k;
              while (end[a] != m) ++a;
              provider_ = Rd{id0 + a, pos[a], m, len[a]};
            }
            has_provider_ = true;
          }
          if (m > max_end_) max_end_ = m;
          if (pos[jl - 1] != at) { hz = false; at = pos[jl - 1]; }
          k = j;
        } else {
          for (; k < n; ++k) {
            const int32_t e = end[k];
            if (e == INT32_MIN) continue;
            const int64_t p = pos[k];
            if (p > last_safe) break;
            if (p != at) {
              if (skip_zero_ && max_end_ < p) break;                   // the sweep jumps: the general path handles it
              hz = max_end_ < p;
              at = p;
            }
            if ((uint32_t)p > chunk_end_) break;                       // (the rare rule of pileup.d:579: general path)
            if ((uint32_t)e > chunk_end_ && (!has_provider_ || e > (int32_t)provider_.end))
              provider_ = Rd{id0 + k, p, e, len[k]};
            if ((uint32_t)e > chunk_end_) has_provider_ = true;
            if (e > max_end_) max_end_ = e;
          }
        }
        if (at != cur_ || k != k0) {
          ci_ += at - cur_;
          cur_ = at;
          adm_pos_ = at;
          adm_init_ = false;
          adm_had_zero_ = hz;
          settled_ = true;
        }
        if (k != k0) continue;
      }
      admit(id0 + k, ref_id, pos[k], end[k], len[k], out);
      ++k;
    }
  }

  // End of the pileup (or of the batch range that will ever be simulated): settle what is pending.
  void finish(std::vector<MdSegment>* out) {
    if (!started_) return;
    finish_island(out);
    started_ = false;
  }

  // The reference of the reads admitted so far is complete (the batch pipeline knows that before the first read of the
  // next reference shows up): emit everything up to its last column now.  The admit() that follows finds it done.
  void finish_reference(std::vector<MdSegment>* out) {
    if (started_) finish_island(out);
  }

  // Streaming use: more reads of this reference will follow, none of them in front of `limit` — the position of the
  // last read admitted, which is where the batch pipeline stops emitting columns.  Emits the segments of every column
  // position below `limit`; afterwards no later segment reaches below it.
  void drain(int64_t limit, std::vector<MdSegment>* out) {
    if (!started_ || fin_) return;
    if (!pending_.empty() && !adm_init_ && limit == adm_pos_) run_to(limit, out);   // what flush_admissions() will do first
    if (seg_open_ && seg_first_ < limit) {
      const int64_t span = limit - seg_first_;
      const int64_t n = std::min<int64_t>(chunk_.len - seg_off_, span);
      if (n > 0) out->push_back(MdSegment{seg_first_, n, chunk_.id, seg_off_});
      seg_first_ = limit;
      seg_off_ += span;
      if (seg_off_ >= chunk_.len) seg_open_ = false;
    }
  }

  // The reads whose dna() may still be asked for by later segments: the current chunk's (if bases are left) and the
  // pending provider's.  Returns how many of ids[0..2) / lens[0..2) were filled.
  int live_providers(uint64_t* ids, int64_t* lens) const {
    int n = 0;
    if (have_chunk_ && !chunk_empty()) { ids[n] = chunk_.id; lens[n] = chunk_.len; ++n; }
    if (has_provider_ && !(n == 1 && ids[0] == provider_.id)) { ids[n] = provider_.id; lens[n] = provider_.len; ++n; }
    return n;
  }

 private:
  struct Rd { uint64_t id; int64_t pos, end, len; };

  // ---- state of PileupRangeUsingMdTag ------------------------------------------------------------------------
  bool skip_zero_;
  bool started_ = false;
  int32_t ref_ = -1;          // reference of the reads being admitted
  int32_t curr_ref_ = -1;     // _curr_ref_id
  // current chunk
  bool have_chunk_ = false;
  Rd chunk_{0, 0, 0, 0};
  int64_t ci_ = 0;            // index of the next base of the chunk to be consumed
  uint32_t chunk_end_ = 0;    // _chunk_end_position
  bool has_provider_ = false;
  Rd provider_{0, 0, 0, 0};
  // open segment of the current chunk: positions from seg_first_ on are served with offset seg_off_
  bool seg_open_ = false;
  int64_t seg_first_ = 0, seg_off_ = 0;
  // the sweep
  int64_t cur_ = 0;           // position of the last column whose step has been simulated
  int64_t max_end_ = 0;       // largest end position of the reads admitted so far on this reference
  // reads admitted at the same position, processed together when the next position (or the end) shows up
  std::vector<Rd> pending_;
  int64_t adm_pos_ = 0;
  bool adm_init_ = false, adm_had_zero_ = false;
  bool fin_ = false;          // finish_island() has run and no read has been admitted since
  bool settled_ = false;      // the step at adm_pos_ has been simulated already (fast path of admit())

  bool chunk_empty() const { return !have_chunk_ || ci_ >= chunk_.len; }

  void close_segment(int64_t end_excl, std::vector<MdSegment>* out) {
    if (!seg_open_) return;
    seg_open_ = false;
    // the chunk serves positions while it has bases left
    const int64_t avail = chunk_.len - seg_off_;
    const int64_t n = std::min<int64_t>(avail, end_excl - seg_first_);
    if (n > 0) out->push_back(MdSegment{seg_first_, n, chunk_.id, seg_off_});
  }

  // override add (pileup.d:563-596), after super.add(read)
  void md_add(const Rd& r, bool had_zero) {
    if (curr_ref_ != ref_) {
      curr_ref_ = ref_;
      has_provider_ = true;
      provider_ = r;
      return;
    }
    if ((uint32_t)r.pos > chunk_end_ && !had_zero) return;
    if ((uint32_t)r.end > chunk_end_) {
      if (!has_provider_) { has_provider_ = true; provider_ = r; }
      else if ((int32_t)r.end > (int32_t)provider_.end) provider_ = r;
    }
  }

  // _chunk = dna(provider); popFrontN(position - provider.position); reference_base = front; popFront  (:639-652, :601-607)
  // (initNewReference takes the chunk from its first base whatever the position: from_start)
  void switch_chunk(int64_t position, bool from_start, std::vector<MdSegment>* out) {
    close_segment(position, out);
    chunk_ = provider_;
    have_chunk_ = true;
    chunk_end_ = (uint32_t)provider_.end;
    has_provider_ = false;
    const int64_t off = from_start ? 0 : std::min<int64_t>(std::max<int64_t>(position - provider_.pos, 0), chunk_.len);
    if (off < chunk_.len) {
      seg_open_ = true;
      seg_first_ = position;
      seg_off_ = off;
      ci_ = off + 1;
    } else {
      ci_ = chunk_.len;          // nothing left: this column reads 'N'
    }
  }

  // Simulate the column steps at positions cur_+1 .. target-1 (all inside the current stretch of consecutive columns):
  // each consumes one base of the chunk; the one that uses the chunk up switches to the pending provider.
  void run_to(int64_t target, std::vector<MdSegment>* out) {
    if (has_provider_ && !chunk_empty()) {
      const int64_t pe = cur_ + (chunk_.len - ci_);       // the step that consumes the last base
      if (pe < target) {
        ci_ = chunk_.len;
        cur_ = pe;
        switch_chunk(pe, false, out);                     // overrides the base of that column
      }
    }
    if (target - 1 > cur_) {
      if (have_chunk_) ci_ = std::min<int64_t>(chunk_.len, ci_ + (target - 1 - cur_));
      cur_ = target - 1;
    }
  }

  // the column step at the admission position (popFront: consume, admit, maybe switch)
  void flush_admissions(std::vector<MdSegment>* out) {
    if (pending_.empty()) return;
    const int64_t p = adm_pos_;
    if (adm_init_) {
      // initNewReference: _prev_coverage = 0, reads admitted, then the provider's chunk from its first base
      for (const Rd& r : pending_) { md_add(r, true); max_end_ = std::max(max_end_, r.end); }
      if (has_provider_) {
        switch_chunk(p, true, out);
      } else if (have_chunk_ && !chunk_empty()) {
        // (not reachable with consistent input) the old chunk goes on: this column reads 'N', the next its next base
        seg_open_ = true;
        seg_first_ = p + 1;
        seg_off_ = ci_;
      }
      cur_ = p;
      pending_.clear();
      return;
    }
    run_to(p, out);
    // step to p: (A) one base of the chunk
    const bool was_empty = chunk_empty();
    if (!was_empty) ++ci_;
    cur_ = p;
    for (const Rd& r : pending_) { md_add(r, adm_had_zero_); max_end_ = std::max(max_end_, r.end); }
    if (chunk_empty() && has_provider_) switch_chunk(p, false, out);
    pending_.clear();
  }

  void begin_island(const Rd& r, std::vector<MdSegment>* out) {
    (void)out;
    adm_pos_ = r.pos;
    adm_init_ = true;
    adm_had_zero_ = true;
    fin_ = false;
    settled_ = false;
    pending_.clear();
    pending_.push_back(r);     // (max_end_ is updated when the admissions are flushed)
  }

  // the sweep runs on to the last covered position of the stretch; then (skip_zero_coverage or a new reference) it jumps
  void finish_island(std::vector<MdSegment>* out) {
    if (fin_) return;
    fin_ = true;
    flush_admissions(out);
    const int64_t end = max_end_;                 // columns exist up to max_end_ - 1
    run_to(end, out);
    // the step that leaves the stretch consumes one more base before the jump is noticed (popFront :616-622, then
    // initNewReference overrides the column)
    close_segment(end, out);
    if (have_chunk_ && !chunk_empty()) ++ci_;
    cur_ = end;
    max_end_ = 0;
  }
};

}  // namespace biodb
