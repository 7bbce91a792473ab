/*
 * biod_b200.h — C ABI of libbiod_b200.so: BGZF inflate -> BAM record/CIGAR decode -> pileup on one B200.
 *
 * This is the drop-in boundary for BioD's hot path.  BioD's only existing FFI is libz
 * (bio/core/utils/zlib.d:6-245, linked with -L-lz, Makefile:10); the D-side binding for this library
 * sits next to it (see INTEGRATION.md and bindings/d/biod_b200.d) and feeds BioD's own range types:
 *
 *   biodb_open / biodb_open_memory        <- BamReader.this(string) / this(Stream)   bam/reader.d:100-138
 *   biodb_header_text / biodb_n_refs ...  <- readSamHeader / readReferenceSequencesInfo  reader.d:579-597
 *   biodb_reads_begin / _next / _end      <- BamReader.reads -> BamReadRange          reader.d:228-231,
 *                                            readNext                                   readrange.d:118-173
 *                                            (BGZF part: fillBgzfBufferFromStream       bgzf/inputstream.d:54-199,
 *                                             decompressBgzfBlock                        bgzf/block.d:127-216)
 *   biodb_pileup_begin / _next / _end     <- makePileup / pileupColumns / PileupRange   bam/pileup.d:683,509,295
 *   biodb_dev_*                           <- the same three stages on caller-owned DEVICE buffers
 *                                            (what bench.py's HBM-resident `value` times)
 *
 * Conventions: plain pointers and sizes only; every call returns a biodb_status; no exceptions, no
 * callbacks into the host language.  One consumer thread per handle; any number of handles may be used
 * concurrently (MultiBamReader opens N readers, bam/multireader.d:218).  All host arrays handed out are
 * owned by the library (pinned host memory), stay valid until the next *_next call on the same iterator
 * or its *_end, and are WRITABLE (BamRead's constructor flips bit 0 of the read-name NUL, read.d:495,880).
 */
#ifndef BIOD_B200_H
#define BIOD_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum biodb_status {
  BIODB_OK = 0,
  BIODB_EOF = 1,            /* iterator exhausted (not an error) */
  BIODB_ERR_BGZF = -1,      /* -> BgzfException            bgzf/inputstream.d:41-43,63-64 */
  BIODB_ERR_ZLIB = -2,      /* -> ZlibException(errnum)    core/utils/zlib.d:247-274; errnum in biodb_error.zlib_errnum */
  BIODB_ERR_FORMAT = -3,    /* -> Exception: "Invalid file format: expected BAM\1" reader.d:113; ISIZE>65536 block.d:150 */
  BIODB_ERR_TRUNCATED = -4, /* -> ReadException from readExact on a cut record     readrange.d:169 */
  BIODB_ERR_IO = -5,        /* open/read failure */
  BIODB_ERR_CUDA = -6,      /* CUDA runtime failure, or no usable device: there is NO CPU fallback */
  BIODB_ERR_CIGAR = -7,     /* PileupRead.assertCigarIndexIsValid pileup.d:224-228 */
  BIODB_ERR_UNSORTED = -8,  /* pileup input not coordinate-sorted within a reference */
  BIODB_ERR_ARG = -9,
  BIODB_ERR_NOMEM = -10
} biodb_status;

typedef struct biodb_error {
  int32_t status;        /* biodb_status of the failed call */
  int32_t zlib_errnum;   /* Z_DATA_ERROR (-3) / Z_BUF_ERROR (-5) when status == BIODB_ERR_ZLIB */
  uint64_t file_offset;  /* compressed offset of the BGZF block at fault, when known */
  char message[256];     /* same text BioD puts in the exception */
} biodb_error;

typedef struct biodb_options {
  int32_t device;            /* CUDA ordinal; -1 = current device */
  int32_t blocks_per_batch;  /* BGZF blocks inflated per GPU batch; 0 = default: three full waves of the inflate
                                kernel on the device; -k = k full waves */
  int32_t verify_crc;        /* 1 = check each block's CRC32 on the device (debug builds of BioD assert it, block.d:187) */
  int32_t want_offsets;      /* 1 = fill start/end virtual offsets (withOffsets policy, readrange.d:51-66) */
  int32_t pin_input;         /* biodb_open_memory: 1 = cudaHostRegister the caller's whole buffer at open; 2 = page-lock on
                                demand, only the byte ranges that passes really read (a shard of the file) */
  int32_t resident_input;    /* 1 = copy the whole compressed file to HBM once at open; passes then read it from there */
  int32_t device_output;     /* 1 = biodb_pileup_next hands out DEVICE pointers (no device->host copy of the columns) */
  int32_t reserved[1];
} biodb_options;

typedef struct biodb_reader biodb_reader;
typedef struct biodb_reads biodb_reads;
typedef struct biodb_pileup biodb_pileup;

/* ---- lifetime -------------------------------------------------------------------------------------- */
const char* biodb_version(void);
void biodb_default_options(biodb_options* o);
/* Opens a BAM file (read into pinned host memory) and parses the header on the way: the header blocks
 * are inflated ON THE DEVICE.  Errors in later blocks surface when iteration reaches them
 * (test/unittests.d:139). */
biodb_status biodb_open(const char* path, const biodb_options* opts, biodb_reader** out);
/* Same over a caller-owned buffer (the MemoryStream case, bgzf/outputstream.d:233-242).  The buffer must
 * outlive the reader. */
biodb_status biodb_open_memory(const void* data, size_t len, const biodb_options* opts, biodb_reader** out);
void biodb_close(biodb_reader* r);
/* Last error of this reader / its iterators.  Never NULL. */
const biodb_error* biodb_last_error(const biodb_reader* r);
/* Error of a failed biodb_open* call (thread-local). */
const biodb_error* biodb_open_error(void);

/* ---- header (reader.d:150-215) ----------------------------------------------------------------------- */
biodb_status biodb_header_text(const biodb_reader* r, const char** text, size_t* len);
int32_t biodb_n_refs(const biodb_reader* r);
biodb_status biodb_ref_info(const biodb_reader* r, int32_t i, const char** name, int32_t* name_len, int32_t* length);
uint64_t biodb_reads_start_voffset(const biodb_reader* r);   /* reader.d:121-123 */
uint64_t biodb_file_size(const biodb_reader* r);
/* 1 if the input buffer of biodb_open_memory was page-locked (options.pin_input succeeded), else 0. */
int32_t biodb_input_is_pinned(const biodb_reader* r);

/* ---- records ------------------------------------------------------------------------------------------ */
typedef struct biodb_record_batch {
  uint64_t n;                  /* records in this batch */
  uint64_t first_index;        /* file-order index of record 0 of the batch */
  uint8_t* data;               /* raw uncompressed stream slice; record i = data[rec_off[i]+4 .. rec_off[i]+4+block_size[i]) */
  uint64_t data_len;
  const uint64_t* rec_off;     /* [n] offset of each record's 4-byte block_size prefix inside data */
  const int32_t* block_size;   /* [n] */
  const int32_t* ref_id;       /* [n]  read.d:85 */
  const int32_t* pos;          /* [n]  read.d:93 */
  const int32_t* end_pos;      /* [n]  position + basesCovered()  read.d:255-262,1380-1383 */
  const uint32_t* bin_mq_nl;   /* [n]  bin<<16 | mapq<<8 | l_read_name   read.d:952-962 */
  const uint32_t* flag_nc;     /* [n]  flag<<16 | n_cigar_op             read.d:963-968 */
  const int32_t* l_seq;        /* [n] */
  const uint64_t* cigar_off;   /* [n+1] into cigar */
  const uint32_t* cigar;       /* packed len<<4|op words  cigar.d:58-148 */
  const uint64_t* start_voffset; /* [n] or NULL unless options.want_offsets   readrange.d:38-48 */
  const uint64_t* end_voffset;   /* [n] or NULL */
} biodb_record_batch;

/* Every call starts an independent pass over the file from the first record, like BamReader.reads
 * (reader.d:228-231,553-569): any number may be alive and interleaved. */
biodb_status biodb_reads_begin(biodb_reader* r, biodb_reads** out);
/* BIODB_OK with a batch, BIODB_EOF at the end, or the error met at this point of the stream. */
biodb_status biodb_reads_next(biodb_reads* it, biodb_record_batch* batch);
void biodb_reads_end(biodb_reads* it);
/* Progress in [0,1] by compressed bytes consumed (readsWithProgress, reader.d:233-279). */
float biodb_reads_progress(const biodb_reads* it);

/* ---- BAI random access: bam["chr"][beg .. end) -------------------------------------- */
/* BaiFile (bio/std/hts/bam/baifile.d:85-169): parse the bytes of a .bai file.  Host only (needs no GPU).
 * BIODB_ERR_FORMAT ("Invalid file format: expected BAI\1") / BIODB_ERR_TRUNCATED, message in biodb_open_error(). */
typedef struct biodb_index biodb_index;
biodb_status biodb_index_open(const void* bai, size_t len, biodb_index** out);
void biodb_index_close(biodb_index* ix);
int32_t biodb_index_n_refs(const biodb_index* ix);
/* RandomAccessManager.getChunks (randomaccessmanager.d:222-244): the merged (beg, end) virtual-offset pairs that hold
 * every read overlapping [beg, end) of reference ref_id.  Writes at most cap pairs into out2 (may be NULL), returns
 * their number, -1 for an invalid reference index. */
int64_t biodb_index_chunks(const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end, uint64_t* out2, uint64_t cap);
/* The last entry of the last non-empty linear index among references [0, n_refs): where BamReader.unmappedReads starts
 * looking for the reads without a reference (reader.d:369-390).  Returns 1 and writes *out, or 0 if there is none. */
int32_t biodb_index_last_linear_offset(const biodb_index* ix, int32_t n_refs, uint64_t* out);
/* IndexBuilder / createIndex (bio/std/hts/bam/bai/indexing.d:56-366): builds the bytes of a .bai file from the reads of
 * a coordinate-sorted file, given in file order, batch by batch, as the arrays a biodb_record_batch holds (reader opened
 * with options.want_offsets).  Host only.  check_bins: verify every read's bin (indexing.d:236-246).
 * BIODB_ERR_UNSORTED ("BAM file is not coordinate-sorted ...") / BIODB_ERR_FORMAT ("Bin in read ... is set
 * incorrectly"), message in _error.  _finish: the file's bytes, valid until _end.  Bins are written in ascending order
 * (the reference's order is that of a D associative array). */
typedef struct biodb_index_builder biodb_index_builder;
biodb_status biodb_index_builder_begin(int32_t n_refs, int32_t check_bins, biodb_index_builder** out);
biodb_status biodb_index_builder_put(biodb_index_builder* b, uint64_t n, const int32_t* ref_id, const int32_t* pos,
                                     const int32_t* end_pos, const uint32_t* bin_mq_nl, const uint32_t* flag_nc,
                                     const uint64_t* start_voffset, const uint64_t* end_voffset);
biodb_status biodb_index_builder_finish(biodb_index_builder* b, const uint8_t** data, size_t* len);
const char* biodb_index_builder_error(const biodb_index_builder* b);
void biodb_index_builder_end(biodb_index_builder* b);
/* ReferenceSequence.opSlice(beg, end) = RandomAccessManager.getReads(BamRegion) (reference.d:76-81,
 * randomaccessmanager.d:300-305): an iterator over the reads of reference ref_id that overlap [beg, end) — the chunks
 * are inflated and scanned like any other stretch of the file, then filtered on the device (BamReadFilter, :366-462).
 * Batches come from biodb_reads_next as usual and hold only the reads of the region (first_index is 0; data is the
 * whole slice the reads lie in).  BIODB_ERR_ARG: beg >= end or invalid reference index. */
biodb_status biodb_reads_begin_region(biodb_reader* r, const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end,
                                      biodb_reads** out);
/* getReads(BamRegion[]) = BamReader.getReadsOverlapping(regions) (reader.d:361, randomaccessmanager.d:246-296,316-337) for
 * the n regions [begs[k], ends[k]) of ONE reference: they are sorted and overlapping ones joined, the chunks of all their
 * bins are read as one stream and filtered by the multi-region BamReadFilter (:366-462) on the device.  The caller runs
 * the groups of several references one after the other, in reference order, as the reference does.  Restatement-defined:
 * the reference indexes a bitset of 37449 entries with every bin id of the index, which fails on the pseudo-bin 37450 that
 * samtools and BioD's own IndexBuilder write; bins beyond the bitset are passed over here.  BIODB_ERR_ARG: a begin >= its
 * end, n == 0, or an invalid reference index.  biodb_index_regions_chunks: the chunk list of such a group (host only). */
biodb_status biodb_reads_begin_regions(biodb_reader* r, const biodb_index* ix, uint32_t ref_id, uint32_t n, const uint32_t* begs,
                                       const uint32_t* ends, biodb_reads** out);
int64_t biodb_index_regions_chunks(const biodb_index* ix, uint32_t ref_id, uint32_t n, const uint32_t* begs, const uint32_t* ends,
                                   uint64_t* out2, uint64_t cap);

/* getReadsBetween(from, to) (reader.d:350-356, randomaccessmanager.d:186-196): the records from virtual offset `from`
 * (which must point at the start of a record) up to virtual offset `to` — the end offset of some record, or
 * UINT64_MAX for "to the end of the file".  getReadAt(offset) (reader.d:336-339) is its first record; max_blocks > 0
 * bounds the BGZF blocks inflated per batch for such point reads (0 = options.blocks_per_batch).
 * Restatement-defined: a `to` inside a record is a BIODB_ERR_TRUNCATED after the whole records in front of it (BioD
 * stops silently). */
biodb_status biodb_reads_begin_between(biodb_reader* r, uint64_t from_voffset, uint64_t to_voffset, uint32_t max_blocks,
                                       biodb_reads** out);

/* ---- pileup -------------------------------------------------------------------------------------------- */
typedef struct biodb_pileup_params {
  int32_t single_ref;          /* 1 = makePileup (first reference only, pileup.d:490-494); 0 = pileupColumns */
  int32_t skip_zero_coverage;  /* pileup.d:389-392 */
  int32_t use_md_tag;          /* 1 = also reconstruct PileupColumn.reference_base from the reads' MD tags
                                  (PileupRangeUsingMdTag, pileup.d:522-654) */
  int32_t want_query_offset;   /* 1 = also return PileupRead.query_offset per entry (pileup.d:146-149) */
  uint64_t start_from;         /* pileup.d:482-489,497-504 (single_ref only) */
  uint64_t end_at;             /* pileup.d:505 (single_ref only); UINT64_MAX = none */
  int32_t counts_only;         /* 1 = per-column A,C,G,T,other,del counts instead of entries */
  int32_t compact_reads;       /* 1 = sequential compact encoding of the column table (position runs, read lists as
                                  last_read / live_mask / stragglers): see biodb_column_batch */
  int32_t maq_mode;            /* MAQ genotype likelihoods over the columns, computed on the device (SURVEY.md 8f row N3;
                                  MaqSnpCaller, bio/std/hts/snpcallers/maq.d:319-540) INSTEAD of delivering the entries:
                                  1 = findSNPs — only the calls that differ from the reference base and exceed
                                      minimum_call_quality leave the device (a few bytes per call);
                                  2 = also genotypeLikelihoodInfo / makeCall of every column (12 bytes per column) with
                                      position, col_off, n_starting_here.
                                  Set the caller's knobs with biodb_pileup_maq_params before the first _next.  findSNPs
                                  needs reference bases: combine with use_md_tag (without it every base is 'N'). */
  int32_t reserved[1];
} biodb_pileup_params;

/* MaqSnpCaller's knobs (maq.d:327-380) and their defaults */
typedef struct biodb_maq_params {
  float depcorr;               /* 0.17 */
  float eta;                   /* 0.03 */
  float minimum_call_quality;  /* 6.0 */
  int32_t minimum_base_quality;/* 13 */
} biodb_maq_params;

typedef struct biodb_column_batch {
  uint64_t n_columns;
  uint64_t n_entries;
  int32_t ref_id;              /* all columns of one batch lie on one reference (PileupColumn.ref_id, pileup.d:257-259) */
  int32_t last_of_pileup;      /* 1 on the final batch */
  const uint64_t* position;    /* [n_columns]   PileupColumn.position  pileup.d:262-264 */
  const uint64_t* col_off;     /* [n_columns+1] entries of column c = [col_off[c], col_off[c+1]); coverage = difference */
  const uint32_t* n_starting_here; /* [n_columns] reads_starting_here = last n entries  pileup.d:272-274 */
  const uint32_t* read_idx;    /* [n_entries] file-order record index, column-major, file order inside a column */
  const uint8_t* base;         /* [n_entries] current_base ('-' inside D/N)            pileup.d:115-122 */
  const uint8_t* qual;         /* [n_entries] current_base_quality (255 inside D/N)    pileup.d:127-134 */
  const uint32_t* query_offset;/* [n_entries] or NULL */
  const uint32_t* counts;      /* [n_columns*6] A,C,G,T,other,deletion — only with counts_only */
  /* compact_reads = 1: a sequential, lossless encoding of the same columns that moves 16 bytes per column + 1.5 bytes
   * per entry over PCIe instead of 20 + 6.  position, col_off and read_idx are NULL and
   *  - positions come as n_runs runs of consecutive positions: columns [run_first_col[r], run_first_col[r+1]) have
   *    positions run_pos[r], run_pos[r]+1, ...  (one run per stretch of non-zero coverage);
   *  - the reads of column c, in column (= file) order, are its stragglers — the entries k of strag_col[] / strag_idx[]
   *    with strag_col[k] == c (sorted by column; reads more than 63 records older than the column's last read) —
   *    followed by record index last_read[c] - d for every d = 63..0 with bit d of live_mask[c] set;
   *  - coverage(c) = stragglers(c) + popcount(live_mask[c]); the entries of column c in base4 / qual / query_offset
   *    start where those of column c-1 end (col_off is the running sum of the coverages);
   *  - base is NULL: entry e's base is "=ACMGRSVTWYHKDBN"[code], code = high nibble of base4[e/2] for even e, low
   *    nibble for odd e (BAM's SEQ packing, read.d:364-383) — except for the n_special entries listed, in ascending
   *    order, in special_entry[] (index e within the batch), whose base is special_base[]: '-' inside a deletion /
   *    reference skip (pileup.d:115-122), 0 for a base asked past l_seq. */
  const uint32_t* last_read;     /* [n_columns] */
  const uint64_t* live_mask;     /* [n_columns] */
  uint64_t n_stragglers;
  const uint32_t* strag_col;     /* [n_stragglers] */
  const uint32_t* strag_idx;     /* [n_stragglers] */
  uint64_t n_runs;
  const uint64_t* run_pos;       /* [n_runs] */
  const uint32_t* run_first_col; /* [n_runs+1] */
  const uint8_t* base4;          /* [(n_entries+1)/2] */
  uint64_t n_special;
  const uint32_t* special_entry; /* [n_special] */
  const uint8_t* special_base;   /* [n_special] */
  /* use_md_tag = 1 (any encoding): PileupColumn.reference_base per column (pileup.d:252-254,614-653), 'N' where no
   * read's MD tag supplies it; NULL otherwise (BioD's default column has 'N', pileup.d:239). */
  const uint8_t* reference_base; /* [n_columns] */
  /* maq_mode >= 1: the calls of findSNPs (maq.d:489-540) among this batch's columns, in column order.  Genotypes are
   * DiploidGenotype!Base5 codes, first allele * 5 + second with A C G T N = 0..4 (bio/core/genotype.d:33-37); a
   * heterozygote is stored as (later nucleotide)|(earlier nucleotide), as computeLikelihoods keys it (maq.d:218-228).
   * call_qual = score of the second best genotype - score of the best (maq.d:480-482). */
  uint64_t n_calls;
  const uint32_t* call_col;      /* [n_calls] column index within the batch */
  const uint64_t* call_pos;      /* [n_calls] */
  const uint8_t* call_gt;        /* [n_calls] */
  const uint8_t* call_ref;       /* [n_calls] reference base character */
  const float* call_qual;        /* [n_calls] */
  /* maq_mode == 2: per column the two best genotypes (255 = no base passed the filters) and their scores
   * (GenotypeLikelihoodInfo[0], [1], maq.d:252-310), and the number of bases that passed the filters. */
  const uint8_t* maq_gt0;        /* [n_columns] */
  const uint8_t* maq_gt1;
  const float* maq_s0;
  const float* maq_s1;
  const uint16_t* maq_n_valid;
} biodb_column_batch;

biodb_status biodb_pileup_begin(biodb_reader* r, const biodb_pileup_params* p, biodb_pileup** out);
/* Sharded pileup (pileupChunks semantics, pileup.d:859-1015).  The records of the file are cut into
 * n_shards consecutive ranges at equal compressed-byte fractions: cut k is the first record that starts in or after the
 * BGZF block at fraction k/n (found by inflating that block and — in files whose records straddle blocks — searching
 * for the record chain's entry, as the sequential reader does).  Shard s owns the records that start in
 * [cut s, cut s+1) and emits exactly the columns from the position of its first record up to (not including) the
 * position of the next shard's first record; it reads, besides its own records, a HALO of earlier records, from
 * `halo_voffset` on, because reads that start before the shard's first record can reach into its columns.
 * Concatenating the shards' batches in shard order gives the columns of an unsharded biodb_pileup_begin pass; read_idx
 * counts from the first record the shard reads (its first halo record): global index = read_idx - n_halo_records +
 * (records owned by all earlier shards), the base the stitch provides.  pileupColumns semantics; use_md_tag is allowed —
 * the chain of MD providers then starts at the halo's first read, exactly as BioD's own chunks start theirs at the first
 * read of makePileup(chain(prev_chunk, chunk), ...) (pileup.d:905-913).
 *
 * The halo is EXACT, in two steps (BioD approximates it by 2 x the median read length, pileup.d:941-985):
 *  1. _begin_shard starts the halo `halo_blocks` BGZF blocks in front of the shard, a guess;
 *  2. every shard reports, for every LATER shard t, the first of its own records that reaches into t's columns
 *     (biodb_pileup_shard_reach: same reference as cut t and end_position > cut t's position).  The minimum over the
 *     earlier shards is the exact start of shard t's halo; a shard whose guess began later than that is run again with
 *     _begin_shard_at(that voffset).  biod_b200.stitch.run_shards / exact_halos do this (one all-gather of n_shards
 *     offsets per rank across GPUs). */
typedef struct biodb_shard_info {
  uint64_t first_voffset, end_voffset;   /* own records: those that start in [first, end) */
  uint64_t halo_voffset;                 /* where reading starts (the start of a record, <= first_voffset) */
  int32_t lo_ref, hi_ref;                /* column keys (ref, pos): [lo, hi) ; ref -1 sorts last */
  int64_t lo_pos, hi_pos;
  uint64_t n_halo_records;               /* records read in front of the first own one (valid once it has been met) */
  uint64_t n_own_records;                /* records starting in the own range (valid at EOF) */
} biodb_shard_info;
/* The n_shards + 1 cuts: virtual offset, reference id and position of the record each cut falls on (cut n_shards = the
 * end of the data: ref -1, position INT64_MAX).  Arrays of n_shards + 1 elements; any may be NULL. */
biodb_status biodb_shard_cuts(biodb_reader* r, uint32_t n_shards, uint64_t* cut_voffset, int32_t* cut_ref, int64_t* cut_pos);
biodb_status biodb_pileup_begin_shard(biodb_reader* r, const biodb_pileup_params* p, uint32_t shard, uint32_t n_shards,
                                      uint32_t halo_blocks, biodb_pileup** out);
/* The same shard with its halo starting at the record at halo_voffset (<= the shard's first record). */
biodb_status biodb_pileup_begin_shard_at(biodb_reader* r, const biodb_pileup_params* p, uint32_t shard, uint32_t n_shards,
                                         uint64_t halo_voffset, biodb_pileup** out);
/* Shards [first, first + count) of n_shards as ONE pass: own records from cut `first` to cut `first + count`.  With
 * n_shards a multiple of the number of workers this hands workers unequal shares of one file (bench.py sizes the shares of
 * the end-to-end pass by each GPU's measured host-link speed) while every cut stays one of the n_shards + 1 cuts all workers
 * agree on.  _shard_reach then reports the later cuts t >= first + count. */
biodb_status biodb_pileup_begin_shard_span(biodb_reader* r, const biodb_pileup_params* p, uint32_t first, uint32_t count,
                                           uint32_t n_shards, uint32_t halo_blocks, biodb_pileup** out);
biodb_status biodb_pileup_begin_shard_span_at(biodb_reader* r, const biodb_pileup_params* p, uint32_t first, uint32_t count,
                                              uint32_t n_shards, uint64_t halo_voffset, biodb_pileup** out);
void biodb_pileup_shard_info(const biodb_pileup* pl, biodb_shard_info* out);
/* reach[t], t in (shard, n_shards): virtual offset of the first OWN record of this shard that reaches into the columns
 * of shard t, UINT64_MAX if none does; entries t <= shard are UINT64_MAX.  Valid at EOF.  `reach` has n_shards elements. */
void biodb_pileup_shard_reach(const biodb_pileup* pl, uint64_t* reach);
/* The pileup (pileupColumns semantics, use_md_tag allowed) of the records that start in [from_voffset, to_voffset) —
 * from_voffset the start of a record, to_voffset the start of a record or UINT64_MAX for the end of the file — clipped to
 * the columns with (ref, position) in [(lo_ref, lo_pos), (hi_ref, hi_pos)); ref -1 sorts last.  This is one element of
 * BioD's pileupChunks: makePileup(chain(prev_chunk, chunk), use_md_tag, beg, end) (pileup.d:905-913) with from_voffset
 * the first halo read and [lo, hi) the chunk's column interval.  read_idx counts from the record at from_voffset. */
biodb_status biodb_pileup_begin_range(biodb_reader* r, const biodb_pileup_params* p, uint64_t from_voffset, uint64_t to_voffset,
                                      int32_t lo_ref, int64_t lo_pos, int32_t hi_ref, int64_t hi_pos, biodb_pileup** out);
/* makePileup(bam[ref][beg .. end), use_md_tag, start_from, end_at, skip_zero_coverage) — "any range of reads is
 * acceptable" (examples/read_bam_file.d:22-25, transverse_multiple_bam_files.d:15): the pileup of the reads of
 * reference ref_id that overlap [beg, end), fetched through the BAI index (biodb_reads_begin_region's reads, reduced on
 * the device before the pileup kernels see the batch).  read_idx counts the reads of the region, in the order
 * biodb_reads_begin_region yields them.  BIODB_ERR_ARG: beg >= end or invalid reference index. */
biodb_status biodb_pileup_begin_region(biodb_reader* r, const biodb_index* ix, uint32_t ref_id, uint32_t beg, uint32_t end,
                                       const biodb_pileup_params* p, biodb_pileup** out);
/* The knobs of maq_mode (NULL = MaqSnpCaller's defaults).  Call before the first biodb_pileup_next. */
biodb_status biodb_pileup_maq_params(biodb_pileup* pl, const biodb_maq_params* mp);
biodb_status biodb_pileup_next(biodb_pileup* pl, biodb_column_batch* cols);
void biodb_pileup_end(biodb_pileup* pl);
/* Reference id of the pileup (AbstractPileup.ref_id, pileup.d:455-457); valid after the first _next. */
int32_t biodb_pileup_ref_id(const biodb_pileup* pl);
/* Totals over the whole pass so far (for benches and checks). */
void biodb_pileup_totals(const biodb_pileup* pl, uint64_t* n_records, uint64_t* n_columns, uint64_t* n_entries);

/* ---- BGZF compression (first part) -------------------------------------------------- */
/* bgzfCompress (bio/core/bgzf/compress.d:43-103) over a whole buffer, cut like BgzfOutputStream does
 * (bgzf/outputstream.d:50-223): one BGZF block per 0xFF00 bytes, compressed on the device, plus the 28-byte EOF block
 * when add_eof != 0 (close(), :218-221).  level: -1..9 as for zlib; 0 stores, every other value selects the one
 * effort this encoder has (greedy LZ77; dynamic or fixed Huffman codes, whichever is shorter).  The compressed bytes are valid DEFLATE but not zlib's —
 * the reference asks for the round trip only (outputstream.d:225-247).  device < 0: the current device.
 * BIODB_ERR_NOMEM when cap < what is needed (biodb_bgzf_compress_bound(len) always suffices); BIODB_ERR_CUDA without
 * a device (no CPU fallback). */
size_t biodb_bgzf_compress_bound(size_t len);
biodb_status biodb_bgzf_compress(int32_t device, const void* data, size_t len, int32_t level, int32_t add_eof, void* out,
                                 size_t cap, size_t* out_len);
/* BamWriter (bio/std/hts/bam/writer.d:67-300) over that compressor.  The writer collects the uncompressed stream on the
 * host exactly as BamWriter + BgzfOutputStream lay it out — "BAM\1", header text, reference table, a block boundary,
 * then the records, where a record that would not fit into the current block starts a new one (writer.d:259-267) and
 * a record longer than a block is cut every 0xFF00 bytes (outputstream.d:107-132) — and compresses the blocks on the
 * device: all at _finish, or as they accumulate when the caller drains the writer (_drain), in which case the writer
 * holds a few thousand blocks at a time, not the file.
 *  _header : writeSamHeader + writeReferenceSequenceInfo (names are NUL-terminated strings); once, before records
 *  _records: writeRecord for every record of the buffer (block_size prefix + body, back to back); the bin field is
 *            recalculated (read.d:1028-1030); BIODB_ERR_ARG "Read reference ID is out of range" (writer.d:245-246)
 *  _flush  : ends the current block
 *  _drain  : if at least min_blocks complete blocks have accumulated: compresses them and hands their BGZF bytes out
 *            (*data valid until the next call on this writer; *len = 0 when there was not enough to do) and releases
 *            their uncompressed bytes
 *  _finish : the rest of the file (the whole file if _drain was never used), EOF block included; valid until _end
 *  _layout : host-only view of the uncompressed bytes and the block starts chosen so far (for tests) */
typedef struct biodb_writer biodb_writer;
biodb_status biodb_writer_begin(int32_t device, int32_t level, biodb_writer** out);
biodb_status biodb_writer_header(biodb_writer* w, const char* text, size_t text_len, int32_t n_refs, const char* const* names,
                                 const int32_t* lengths);
biodb_status biodb_writer_records(biodb_writer* w, const uint8_t* records, size_t len);
biodb_status biodb_writer_flush(biodb_writer* w);
biodb_status biodb_writer_drain(biodb_writer* w, uint32_t min_blocks, const uint8_t** data, size_t* len);
biodb_status biodb_writer_finish(biodb_writer* w, const uint8_t** data, size_t* len);
biodb_status biodb_writer_layout(const biodb_writer* w, const uint8_t** data, size_t* len, const uint64_t** cuts, size_t* n_cuts);
/* The BAI index of the finished file: what BamWriter builds while it writes coordinate-sorted output (writer.d:139-195:
 * IndexBuilder with check_bins, every record with the virtual offsets it got).  After _finish; valid until _end.
 * BIODB_ERR_UNSORTED when the records were not in coordinate order. */
biodb_status biodb_writer_index(biodb_writer* w, const uint8_t** data, size_t* len);
/* Host-only test hook: use `data` (a BGZF stream with the writer's block layout) as the finished file. */
biodb_status biodb_writer_debug_set_output(biodb_writer* w, const uint8_t* data, size_t len);
const char* biodb_writer_error(const biodb_writer* w);
void biodb_writer_end(biodb_writer* w);
/* Host-only test hook: the device's DEFLATE encoder as plain host loops (csrc/deflate_enc.h: deflate_block_host — the
 * same windows, candidates, tokens and codes as the warp of deflate_warp_kernel, so the same bytes); raw DEFLATE of one
 * chunk of at most 65535 bytes.  Returns the size, 0 if cap is too small. */
int64_t biodb_debug_deflate_block(const uint8_t* in, uint32_t n, uint8_t* out, uint32_t cap, int32_t level);
/* Statistics of the last biodb_bgzf_compress / biodb_writer_finish on `device` (-1: the current one): out[0] =
 * microseconds the encoder kernel of the first slab (at most 2048 blocks) took, out[1] = CTAs of its persistent grid,
 * out[2] = calls so far, out[3] = 0. */
biodb_status biodb_debug_deflate_stats(int32_t device, uint64_t* out);

/* ---- measurement ---------------------------------------------------------------------------------------- */
typedef struct biodb_stats {
  double total_ms;        /* CUDA-event time from the first to the last operation of the pass, on its stream */
  double inflate_ms;      /* sum over launches of the inflate kernel (events around each launch) */
  double scan_ms;         /* record scanner kernels */
  double pileup_ms;       /* pileup kernels */
  uint64_t inflate_launches, kernel_launches;
  uint64_t h2d_bytes, d2h_bytes;
  uint64_t compressed_bytes, uncompressed_bytes;   /* algorithmic bytes moved by the inflate kernel */
  uint64_t n_blocks, n_records;
} biodb_stats;
void biodb_reads_stats(const biodb_reads* it, biodb_stats* out);
void biodb_pileup_stats(const biodb_pileup* pl, biodb_stats* out);

/* ---- device-resident stage API (all pointers are DEVICE pointers; stream is a cudaStream_t) ------------- */
/* Inflate n BGZF blocks.  Block i's raw-DEFLATE payload is comp[payload_off[i] .. +cdata_size[i]); its
 * ISIZE bytes are written at out + out_off[i].  `comp` must be readable 32 bytes past the last payload.
 * status[i] = 0, Z_DATA_ERROR (-3) or Z_BUF_ERROR (-5), matching inflate(Z_FINISH) of block.d:172.
 * crc (may be NULL): when given, the CRC32 of each block's output is written there. */
biodb_status biodb_dev_inflate(const uint8_t* comp, const uint64_t* payload_off, const uint32_t* cdata_size,
                               const uint64_t* out_off, const uint32_t* isize, uint32_t n_blocks, uint8_t* out,
                               int32_t* status, uint32_t* crc, void* stream);

/* Diagnostics of the lane-parallel inflate kernels since the last reset (synchronises the device):
 * out8[0] blocks they gave up on (redone by the warp-serial kernel: malformed or unusual streams), [1] super-chunks,
 * [2] decode rounds, [3] matches read back from L2, [4] matches, [5] DEFLATE blocks; of the blocks given up: [6] those
 * whose record stream outgrew its arena, [7] those whose sub-sequences never synchronised (codes of one length). */
biodb_status biodb_debug_inflate_counters(uint64_t* out8, int32_t reset);
/* Host-only building blocks of the MD-tag reference bases (row N1 of the plan, pileup.d:522-654), exported so that the
 * CPU test suite can check them against the oracle.  Neither needs a GPU.
 * biodb_debug_md_chain (csrc/md_chain.h): which read's dna() string supplies PileupColumn.reference_base at which
 * positions.  Input: the n reads a pileup keeps, in file order (reference id, position, end position, length of
 * dna(read)).  Output: segments (first position, count, read index, offset into that read's dna()) as int64 quadruples;
 * positions outside every segment read 'N'.  batch_reads > 0 drains the chain every batch_reads reads the way the batch
 * pipeline does and marks each drain with a quadruple (limit, -1, ref, 0).  Returns the number of quadruples (which
 * may exceed cap).
 * biodb_debug_md_dna (csrc/md_walk.h): dna(read) (md/reconstruct.d:38-214) of one raw record body (the block_size bytes
 * that follow the block_size field); returns its length and writes at most cap characters. */
int64_t biodb_debug_md_chain(const int32_t* ref_id, const int64_t* pos, const int64_t* end, const int64_t* dna_len, uint64_t n,
                             int32_t skip_zero_coverage, uint64_t batch_reads, int64_t* seg4, uint64_t cap);
int64_t biodb_debug_md_dna(const uint8_t* body, int64_t block_size, uint8_t* out, uint64_t cap);

typedef struct biodb_dev_records {
  uint64_t capacity;     /* in: room (records) in each array below */
  uint64_t cigar_capacity;
  uint64_t* rec_off;
  int32_t* block_size;
  int32_t* ref_id;
  int32_t* pos;
  int32_t* end_pos;
  uint32_t* bin_mq_nl;
  uint32_t* flag_nc;
  int32_t* l_seq;
  uint64_t* cigar_off;   /* [capacity+1] */
  uint32_t* cigar;
} biodb_dev_records;

/* Walk the records of an uncompressed stream slice `u[0..u_len)` that starts at a record boundary.
 * block_uoff[n_blocks+1] are the BGZF block boundaries inside u (used to walk blocks in parallel).
 * On return (after the stream is synchronised) result[0] = number of complete records,
 * result[1] = offset of the first byte not consumed (tail of a cut record), result[2] = cigar words,
 * result[3] = 0 or a biodb_status (malformed record).  `final_slice` selects end-of-file semantics
 * (readrange.d:139-150). */
biodb_status biodb_dev_scan_records(const uint8_t* u, uint64_t u_len, const uint64_t* block_uoff, uint32_t n_blocks,
                                    int32_t final_slice, biodb_dev_records* out, uint64_t* result /* device, [4] */,
                                    void* workspace, size_t workspace_bytes, void* stream);
size_t biodb_dev_scan_workspace_bytes(uint32_t n_blocks);

#ifdef __cplusplus
}
#endif
#endif /* BIOD_B200_H */
