/*
    D binding of libbiod_b200.so (include/biod_b200.h) for BioD.

    Place next to bio/core/utils/zlib.d (BioD's only other FFI, bio/core/utils/zlib.d:6-245) and link with
    -L-lbiod_b200 in addition to -L-lz (Makefile:10).  Written against BioD v0.2.4; it could not be compiled
    in the build image (no D toolchain), see INTEGRATION.md.

    GpuBamReader implements IBamSamReader (bio/std/hts/bam/abstractreader.d:34-49) so every consumer of the
    range API keeps working; BamRead stays BioD's own struct — a view over raw record bytes
    (bio/std/hts/bam/read.d:865) — built over slices of the library's writable record buffer.
*/
module bio.std.hts.bam.gpu;

import bio.std.hts.bam.read : BamRead;
import bio.std.hts.bam.readrange : BamReadBlock;
import bio.std.hts.bam.abstractreader : IBamSamReader;
import bio.std.hts.bam.referenceinfo : ReferenceSequenceInfo;
import bio.std.hts.sam.header : SamHeader;
import bio.core.bgzf.inputstream : BgzfException;
import bio.core.bgzf.virtualoffset : VirtualOffset;
import bio.core.utils.zlib : ZlibException;
import contrib.undead.stream : ReadException;
import std.parallelism : TaskPool, taskPool;
import std.string : toStringz;
import std.conv : to;
import std.exception : enforce;
import std.range : InputRange, inputRangeObject;

extern (C) nothrow @nogc {
    enum : int {
        BIODB_OK = 0, BIODB_EOF = 1, BIODB_ERR_BGZF = -1, BIODB_ERR_ZLIB = -2, BIODB_ERR_FORMAT = -3,
        BIODB_ERR_TRUNCATED = -4, BIODB_ERR_IO = -5, BIODB_ERR_CUDA = -6, BIODB_ERR_CIGAR = -7,
        BIODB_ERR_UNSORTED = -8, BIODB_ERR_ARG = -9, BIODB_ERR_NOMEM = -10
    }
    struct biodb_error { int status; int zlib_errnum; ulong file_offset; char[256] message; }
    struct biodb_options {
        int device = -1; int blocks_per_batch; int verify_crc; int want_offsets; int pin_input;
        int resident_input; int device_output; int[1] reserved;
    }
    struct biodb_reader; struct biodb_reads; struct biodb_pileup;
    struct biodb_record_batch {
        ulong n; ulong first_index; ubyte* data; ulong data_len;
        const(ulong)* rec_off; const(int)* block_size; const(int)* ref_id; const(int)* pos; const(int)* end_pos;
        const(uint)* bin_mq_nl; const(uint)* flag_nc; const(int)* l_seq; const(ulong)* cigar_off; const(uint)* cigar;
        const(ulong)* start_voffset; const(ulong)* end_voffset;
    }
    struct biodb_pileup_params {
        int single_ref; int skip_zero_coverage; int use_md_tag; int want_query_offset;
        ulong start_from; ulong end_at; int counts_only; int compact_reads; int maq_mode; int[1] reserved;
    }
    struct biodb_maq_params { float depcorr = 0.17f; float eta = 0.03f; float minimum_call_quality = 6.0f; int minimum_base_quality = 13; }
    struct biodb_shard_info {
        ulong first_voffset, end_voffset, halo_voffset; int lo_ref, hi_ref; long lo_pos, hi_pos;
        ulong n_halo_records, n_own_records;
    }
    struct biodb_column_batch {
        ulong n_columns; ulong n_entries; int ref_id; int last_of_pileup;
        const(ulong)* position; const(ulong)* col_off; const(uint)* n_starting_here;
        const(uint)* read_idx; const(ubyte)* base; const(ubyte)* qual; const(uint)* query_offset; const(uint)* counts;
        // compact_reads: position, col_off and read_idx are null; see include/biod_b200.h for the sequential encoding
        const(uint)* last_read; const(ulong)* live_mask; ulong n_stragglers; const(uint)* strag_col; const(uint)* strag_idx;
        ulong n_runs; const(ulong)* run_pos; const(uint)* run_first_col;
        const(ubyte)* base4; ulong n_special; const(uint)* special_entry; const(ubyte)* special_base;
        const(ubyte)* reference_base;   // use_md_tag: one per column (pileup.d:252-254), else null
        // maq_mode: the calls of MaqSnpCaller.findSNPs among the batch's columns; mode 2: the two best genotypes per column
        ulong n_calls; const(uint)* call_col; const(ulong)* call_pos; const(ubyte)* call_gt; const(ubyte)* call_ref;
        const(float)* call_qual;
        const(ubyte)* maq_gt0; const(ubyte)* maq_gt1; const(float)* maq_s0; const(float)* maq_s1; const(ushort)* maq_n_valid;
    }
    void biodb_default_options(biodb_options*);
    int biodb_open(const(char)* path, const(biodb_options)*, biodb_reader**);
    int biodb_open_memory(const(void)* data, size_t len, const(biodb_options)*, biodb_reader**);
    void biodb_close(biodb_reader*);
    const(biodb_error)* biodb_last_error(const(biodb_reader)*);
    const(biodb_error)* biodb_open_error();
    int biodb_header_text(const(biodb_reader)*, const(char)** text, size_t* len);
    int biodb_n_refs(const(biodb_reader)*);
    int biodb_ref_info(const(biodb_reader)*, int i, const(char)** name, int* name_len, int* length);
    ulong biodb_reads_start_voffset(const(biodb_reader)*);
    int biodb_reads_begin(biodb_reader*, biodb_reads**);
    int biodb_reads_next(biodb_reads*, biodb_record_batch*);
    void biodb_reads_end(biodb_reads*);
    float biodb_reads_progress(const(biodb_reads)*);
    int biodb_pileup_begin(biodb_reader*, const(biodb_pileup_params)*, biodb_pileup**);
    int biodb_pileup_next(biodb_pileup*, biodb_column_batch*);
    int biodb_pileup_maq_params(biodb_pileup*, const(biodb_maq_params)*);
    // shards / chunks (pileupChunks semantics, pileup.d:859-1015) with exact halos: see include/biod_b200.h
    int biodb_shard_cuts(biodb_reader*, uint n_shards, ulong* cut_voffset, int* cut_ref, long* cut_pos);
    int biodb_pileup_begin_shard(biodb_reader*, const(biodb_pileup_params)*, uint shard, uint n_shards, uint halo_blocks, biodb_pileup**);
    int biodb_pileup_begin_shard_at(biodb_reader*, const(biodb_pileup_params)*, uint shard, uint n_shards, ulong halo_voffset,
                                    biodb_pileup**);
    int biodb_pileup_begin_shard_span(biodb_reader*, const(biodb_pileup_params)*, uint first, uint count, uint n_shards,
                                      uint halo_blocks, biodb_pileup**);
    int biodb_pileup_begin_shard_span_at(biodb_reader*, const(biodb_pileup_params)*, uint first, uint count, uint n_shards,
                                         ulong halo_voffset, biodb_pileup**);
    void biodb_pileup_shard_info(const(biodb_pileup)*, biodb_shard_info*);
    void biodb_pileup_shard_reach(const(biodb_pileup)*, ulong* reach);
    int biodb_pileup_begin_range(biodb_reader*, const(biodb_pileup_params)*, ulong from_voffset, ulong to_voffset,
                                 int lo_ref, long lo_pos, int hi_ref, long hi_pos, biodb_pileup**);
    void biodb_pileup_end(biodb_pileup*);
    int biodb_pileup_ref_id(const(biodb_pileup)*);
    // BAI random access (row N2)
    struct biodb_index;
    int biodb_index_open(const(void)* bai, size_t len, biodb_index**);
    void biodb_index_close(biodb_index*);
    int biodb_index_n_refs(const(biodb_index)*);
    long biodb_index_chunks(const(biodb_index)*, uint ref_id, uint beg, uint end, ulong* out2, ulong cap);
    int biodb_reads_begin_region(biodb_reader*, const(biodb_index)*, uint ref_id, uint beg, uint end, biodb_reads**);
    int biodb_reads_begin_regions(biodb_reader*, const(biodb_index)*, uint ref_id, uint n, const(uint)* begs, const(uint)* ends,
                                  biodb_reads**);
    int biodb_reads_begin_between(biodb_reader*, ulong from_voffset, ulong to_voffset, uint max_blocks, biodb_reads**);
    int biodb_pileup_begin_region(biodb_reader*, const(biodb_index)*, uint ref_id, uint beg, uint end,
                                  const(biodb_pileup_params)*, biodb_pileup**);
    // BGZF compression (row N4, first part)
    struct biodb_writer;
    int biodb_writer_begin(int device, int level, biodb_writer**);
    int biodb_writer_header(biodb_writer*, const(char)* text, size_t text_len, int n_refs, const(char*)* names, const(int)* lengths);
    int biodb_writer_records(biodb_writer*, const(ubyte)* records, size_t len);
    int biodb_writer_flush(biodb_writer*);
    int biodb_writer_drain(biodb_writer*, uint min_blocks, const(ubyte)** data, size_t* len);
    int biodb_writer_finish(biodb_writer*, const(ubyte)** data, size_t* len);
    const(char)* biodb_writer_error(const(biodb_writer)*);
    void biodb_writer_end(biodb_writer*);
    size_t biodb_bgzf_compress_bound(size_t len);
    int biodb_bgzf_compress(int device, const(void)* data, size_t len, int level, int add_eof, void* out_, size_t cap,
                            size_t* out_len);
}

/// The BGZF stream of `data` (bgzf/outputstream.d:50-223 over a complete buffer): blocks of 0xFF00 bytes compressed on
/// the GPU and framed like bgzfCompress frames them (bgzf/compress.d:43-103), plus the EOF block when `eof`.
ubyte[] gpuBgzfCompress(const(ubyte)[] data, int level = -1, bool eof = true) {
    auto buf = new ubyte[biodb_bgzf_compress_bound(data.length)];
    size_t n;
    auto st = biodb_bgzf_compress(-1, data.ptr, data.length, level, eof, buf.ptr, buf.length, &n);
    enforce(st == BIODB_OK, "biodb_bgzf_compress failed");
    return buf[0 .. n];
}

/// Maps a status + error record to the exception classes BioD's tests pin (test/unittests.d:132-142).
private void raise(const(biodb_error)* e) {
    auto msg = to!string(e.message.ptr);
    switch (e.status) {
        case BIODB_ERR_BGZF:      throw new BgzfException(msg);                 // bgzf/inputstream.d:41-43
        case BIODB_ERR_ZLIB:      throw new ZlibException(e.zlib_errnum);       // core/utils/zlib.d:247-274
        case BIODB_ERR_TRUNCATED: throw new ReadException(msg);                 // readrange.d:169
        default:                  throw new Exception(msg);                     // reader.d:113, block.d:150
    }
}

/// Input range of BamRead over GPU-decoded batches: `empty/front/popFront` like BamReadRange (readrange.d:82-188).
struct GpuBamReadRange(bool withOffsets = false) {
    private biodb_reader* _h;
    private biodb_reads* _it;
    private biodb_record_batch _b;
    private size_t _i;
    private bool _empty;
    private BamRead _current;
    private IBamSamReader _reader;

    this(biodb_reader* h, IBamSamReader reader) {
        _h = h; _reader = reader;
        if (biodb_reads_begin(h, &_it) != BIODB_OK) raise(biodb_last_error(h));
        fetch();
    }
    /// the reads of reference ref_id overlapping [beg, end): RandomAccessManager.getReads (randomaccessmanager.d:300-305)
    this(biodb_reader* h, IBamSamReader reader, const(biodb_index)* ix, uint ref_id, uint beg, uint end) {
        _h = h; _reader = reader;
        if (biodb_reads_begin_region(h, ix, ref_id, beg, end, &_it) != BIODB_OK) raise(biodb_last_error(h));
        fetch();
    }
    /// the reads of reference ref_id overlapping any of the regions [begs[k], ends[k]) — one reference group of
    /// getReads(BamRegion[]) (randomaccessmanager.d:286-296); getReadsOverlapping chains the groups (:316-337)
    this(biodb_reader* h, IBamSamReader reader, const(biodb_index)* ix, uint ref_id, const(uint)[] begs, const(uint)[] ends) {
        _h = h; _reader = reader;
        if (biodb_reads_begin_regions(h, ix, ref_id, cast(uint)begs.length, begs.ptr, ends.ptr, &_it) != BIODB_OK)
            raise(biodb_last_error(h));
        fetch();
    }
    ~this() { if (_it !is null) { biodb_reads_end(_it); _it = null; } }
    @disable this(this);

    bool empty() @property const { return _empty; }
    static if (withOffsets) {
        BamReadBlock front() @property {
            return BamReadBlock(VirtualOffset(_b.start_voffset[_i]), VirtualOffset(_b.end_voffset[_i]), _current);
        }
    } else {
        ref BamRead front() @property { return _current; }
    }
    void popFront() { ++_i; if (_i >= _b.n) fetch(); else load(); }

    private void fetch() {
        auto st = biodb_reads_next(_it, &_b);
        if (st == BIODB_EOF) { _empty = true; return; }
        if (st != BIODB_OK) raise(biodb_last_error(_h));
        _i = 0;
        load();
    }
    private void load() {
        auto o = cast(size_t)_b.rec_off[_i] + 4;
        // the slice is writable library-owned pinned memory, valid until the next batch: the same rule as
        // BioD's own slab (readrange.d:176-184); BamRead.dup copies (read.d:585-592)
        _current = BamRead(_b.data[o .. o + _b.block_size[_i]]);               // read.d:482-504
        _current.associateWithReader(_reader);                                  // read.d:841-843
    }
}

/// `bam["chr1"]` (bam/reference.d:37-160): slicing gives the reads that overlap [start, end), fetched through the BAI
/// index — chunks inflated and scanned on the GPU, BamReadFilter (randomaccessmanager.d:366-462) as a device-side filter.
struct GpuReferenceSequence {
    private GpuBamReader _bam;
    private int _ref_id;
    private ReferenceSequenceInfo _info;
    string name() @property const { return _info.name; }
    int length() @property const { return _info.length; }
    int id() @property const { return _ref_id; }
    auto opSlice(uint start, uint end) {
        enforce(start < end, "start must be less than end");                      // reference.d:77
        enforce(_ref_id >= 0, "invalid reference id");
        return GpuBamReadRange!true(_bam.handle, _bam, _bam.index, cast(uint)_ref_id, start, end);
    }
    auto opSlice() { return opSlice(0, length); }
    /// the same region as a value makePileup accepts (a D range cannot be told apart from any other range of reads)
    GpuRegion region(uint start, uint end) { enforce(start < end, "start must be less than end"); return GpuRegion(_bam, cast(uint)_ref_id, start, end); }
    VirtualOffset startVirtualOffset() { auto r = opSlice(); enforce(!r.empty); return r.front.start_virtual_offset; }
    int firstPosition() { auto r = opSlice(); return r.empty ? -1 : r.front.read.position; }
}

/// Drop-in for BamReader on the streaming path (bam/reader.d:80-598).
class GpuBamReader : IBamSamReader {
    private biodb_reader* _h;
    private string _filename;
    private SamHeader _header;
    private string _headertext;
    private ReferenceSequenceInfo[] _refs;

    this(string filename, TaskPool pool = taskPool) {     // the TaskPool is accepted and ignored (reader.d:100-101)
        _filename = filename;
        biodb_options o; biodb_default_options(&o);
        o.want_offsets = 1;
        if (biodb_open(filename.toStringz, &o, &_h) != BIODB_OK) raise(biodb_open_error());
        const(char)* t; size_t n;
        biodb_header_text(_h, &t, &n);
        _headertext = t[0 .. n].idup;
        foreach (i; 0 .. biodb_n_refs(_h)) {
            const(char)* nm; int nl, len;
            biodb_ref_info(_h, i, &nm, &nl, &len);
            _refs ~= ReferenceSequenceInfo(nm[0 .. nl].idup, len);
        }
    }
    ~this() { if (_ix !is null) biodb_index_close(_ix); if (_h !is null) biodb_close(_h); }

    SamHeader header() @property { if (_header is null) _header = new SamHeader(_headertext); return _header; }
    const(ReferenceSequenceInfo)[] reference_sequences() @property const nothrow { return _refs; }
    string filename() @property const { return _filename; }
    auto reads(alias IteratePolicy = void)() @property { return GpuBamReadRange!false(_h, this); }
    auto readsWithOffsets() @property { return GpuBamReadRange!true(_h, this); }
    InputRange!BamRead allReads() @property { return inputRangeObject(reads()); }
    void assumeSequentialProcessing() {}   // batches already reuse their buffer (reader.d:324)
    package biodb_reader* handle() { return _h; }

    // random access (reader.d:424-447): the index is looked for next to the file, as BaiFile does (baifile.d:95-113)
    private biodb_index* _ix;
    package const(biodb_index)* index() {
        if (_ix is null) {
            import std.file : exists, read;
            auto first = _filename ~ ".bai";
            enforce(exists(first), "BAM index file (.bai) must be provided");     // randomaccessmanager.d:202-204
            auto bytes = cast(ubyte[])read(first);
            if (biodb_index_open(bytes.ptr, bytes.length, &_ix) != BIODB_OK) raise(biodb_open_error());
        }
        return _ix;
    }
    GpuReferenceSequence opIndex(string ref_name) {                               // reader.d:424-429
        foreach (i, r; _refs) if (r.name == ref_name) return GpuReferenceSequence(this, cast(int)i, r);
        throw new Exception("Reference with name " ~ ref_name ~ " does not exist");
    }
    GpuReferenceSequence reference(int ref_id) { return GpuReferenceSequence(this, ref_id, _refs[ref_id]); }   // reader.d:435-440
}

/// Pileup column over a GPU column batch; same surface as PileupColumn (pileup.d:236-290) for the fields
/// MaqSnpCaller reads (snpcallers/maq.d:396-410,468-483).
struct GpuPileupColumn {
    private const(biodb_column_batch)* _b;
    private size_t _c;
    // where this column sits in the per-entry arrays and in the straggler list; filled by GpuPileup, which walks the
    // columns in order (the compact encoding of a batch is sequential: see include/biod_b200.h)
    private ulong _pos;
    private size_t _off, _cov, _sk, _ns, _sp;
    ulong position() @property const { return _pos; }
    int ref_id() @property const { return _b.ref_id; }
    size_t coverage() @property const { return _cov; }
    /// pileup.d:252-254; 'N' (PileupColumn's default, pileup.d:239) unless the pileup was made with use_md_tag
    char reference_base() @property const { return _b.reference_base is null ? 'N' : cast(char)_b.reference_base[_c]; }
    /// Record indices of the reads of the column, in file order.
    const(uint)[] reads() @property const {
        if (_b.read_idx !is null) return _b.read_idx[_off .. _off + _cov];
        auto r = new uint[_cov];
        size_t k = 0;
        foreach (i; _sk .. _sk + _ns) r[k++] = _b.strag_idx[i];
        foreach_reverse (d; 0 .. 64) if ((_b.live_mask[_c] >> d) & 1) r[k++] = _b.last_read[_c] - cast(uint)d;
        return r;
    }
    const(uint)[] reads_starting_here() @property const {
        auto r = reads;
        return r[$ - _b.n_starting_here[_c] .. $];
    }
    /// Bases of the column's reads (pileup.d:277-279).  The compact encoding packs them two per byte and lists the
    /// entries that are not a base ('-' in a deletion, 0 past l_seq) apart; `_sp` is the cursor into that list.
    const(char)[] bases() @property const {
        if (_b.base !is null) return cast(const(char)[])_b.base[_off .. _off + _cov];
        auto r = new char[_cov];
        size_t sp = _sp;
        foreach (k; 0 .. _cov) {
            auto e = _off + k;
            if (sp < _b.n_special && _b.special_entry[sp] == e) { r[k] = cast(char)_b.special_base[sp++]; continue; }
            auto b = _b.base4[e >> 1];
            r[k] = "=ACMGRSVTWYHKDBN"[(e & 1) ? (b & 15) : (b >> 4)];
        }
        return r;
    }
    const(ubyte)[] base_qualities() @property const { return _b.qual[_off .. _off + _cov]; }
}

/// Input range of columns; `makePileup(GpuBamReader, ...)` mirrors pileup.d:683-694, `pileupColumns` pileup.d:509-519.
struct GpuPileup {
    private biodb_reader* _h;
    private biodb_pileup* _p;
    private biodb_column_batch _b;
    private size_t _c;
    private bool _empty;
    private GpuPileupColumn _col;      // cursors of the current column
    private size_t _run;               // current run of consecutive positions (compact encoding)
    this(biodb_reader* h, bool single_ref, bool use_md_tag, ulong start_from, ulong end_at, bool skip_zero_coverage,
         bool compact = true) {
        _h = h;
        biodb_pileup_params prm;
        prm.single_ref = single_ref; prm.skip_zero_coverage = skip_zero_coverage; prm.use_md_tag = use_md_tag;
        prm.start_from = start_from; prm.end_at = end_at; prm.compact_reads = compact;
        if (biodb_pileup_begin(h, &prm, &_p) != BIODB_OK) raise(biodb_last_error(h));
        fetch();
    }
    /// the pileup of the reads of a region: makePileup(bam["chr2"][150 .. 160], ...) (examples/read_bam_file.d:21-25)
    this(biodb_reader* h, const(biodb_index)* ix, uint ref_id, uint beg, uint end, bool use_md_tag, ulong start_from,
         ulong end_at, bool skip_zero_coverage, bool compact = true) {
        _h = h;
        biodb_pileup_params prm;
        prm.single_ref = true; prm.skip_zero_coverage = skip_zero_coverage; prm.use_md_tag = use_md_tag;
        prm.start_from = start_from; prm.end_at = end_at; prm.compact_reads = compact;
        if (biodb_pileup_begin_region(h, ix, ref_id, beg, end, &prm, &_p) != BIODB_OK) raise(biodb_last_error(h));
        fetch();
    }
    /// one element of pileupChunks: the records that start in [from_voffset, to_voffset), columns [beg, end) of ref_id
    /// (makePileup(chain(prev_chunk, chunk), use_md_tag, beg, end), pileup.d:905-913)
    this(biodb_reader* h, ulong from_voffset, ulong to_voffset, int ref_id, ulong beg, ulong end, bool use_md_tag) {
        _h = h;
        biodb_pileup_params prm;
        prm.skip_zero_coverage = true; prm.use_md_tag = use_md_tag; prm.compact_reads = true;
        if (biodb_pileup_begin_range(h, &prm, from_voffset, to_voffset, ref_id, cast(long)beg, ref_id, cast(long)end, &_p) != BIODB_OK)
            raise(biodb_last_error(h));
        _start = beg; _end = end; _ref = ref_id;
        fetch();
    }
    private ulong _start, _end; private int _ref = -1;
    /// AbstractPileup.start_position / end_position (pileup.d:440-452) of a chunk
    ulong start_position() @property const { return _start; }
    ulong end_position() @property const { return _end; }
    ~this() { if (_p !is null) { biodb_pileup_end(_p); _p = null; } }
    @disable this(this);
    bool empty() @property const { return _empty; }
    GpuPileupColumn front() @property { return _col; }
    void popFront() {
        if (++_c >= _b.n_columns) { fetch(); return; }
        _col._off += _col._cov;
        _col._sk += _col._ns;
        while (_col._sp < _b.n_special && _b.special_entry[_col._sp] < _col._off) ++_col._sp;
        place();
    }
    // cursors of column _c, given those of column _c - 1
    private void place() {
        _col._c = _c;
        if (_b.read_idx !is null) {                       // explicit table
            _col._pos = _b.position[_c];
            _col._off = cast(size_t)_b.col_off[_c];
            _col._cov = cast(size_t)(_b.col_off[_c + 1] - _b.col_off[_c]);
            return;
        }
        while (_c >= _b.run_first_col[_run + 1]) ++_run;
        _col._pos = _b.run_pos[_run] + (_c - _b.run_first_col[_run]);
        size_t ns = 0;
        while (_col._sk + ns < _b.n_stragglers && _b.strag_col[_col._sk + ns] == _c) ++ns;
        _col._ns = ns;
        import core.bitop : popcnt;
        _col._cov = ns + popcnt(_b.live_mask[_c]);
    }
    int ref_id() @property { return biodb_pileup_ref_id(_p); }
    private void fetch() {
        auto st = biodb_pileup_next(_p, &_b);
        if (st == BIODB_EOF) { _empty = true; return; }
        if (st != BIODB_OK) raise(biodb_last_error(_h));
        _c = 0;
        _run = 0;
        _col = GpuPileupColumn(&_b, 0);
        place();
    }
}

auto makePileup(GpuBamReader bam, bool use_md_tag = false, ulong start_from = 0, ulong end_at = ulong.max,
                bool skip_zero_coverage = true) {
    return GpuPileup(bam.handle, true, use_md_tag, start_from, end_at, skip_zero_coverage);
}
/// makePileup over a region's reads: `makePileup(bam["chr2"].region(150, 160), false, 155, 158)`
struct GpuRegion { GpuBamReader bam; uint ref_id, start, end; }
auto makePileup(GpuRegion r, bool use_md_tag = false, ulong start_from = 0, ulong end_at = ulong.max,
                bool skip_zero_coverage = true) {
    return GpuPileup(r.bam.handle, r.bam.index, r.ref_id, r.start, r.end, use_md_tag, start_from, end_at, skip_zero_coverage);
}
auto pileupColumns(GpuBamReader bam, bool use_md_tag = false, bool skip_zero_coverage = true) {
    return GpuPileup(bam.handle, false, use_md_tag, 0, ulong.max, skip_zero_coverage);
}

/// pileupChunks(reads, use_md_tag, block_size, start_from, end_at) (pileup.d:1011-1015): non-overlapping consecutive
/// pileups that can be processed in parallel — here each one is a range pileup on the GPU, independent of the others.
/// The cuts follow chunksConsumingLessThan (splitter.d:66-90) and PileupChunkRange (pileup.d:876-940) over the record
/// tables of one reads pass (reader opened with want_offsets); the halo of a chunk is EXACT (the first earlier read of
/// the reference that reaches beyond the chunk's first column) where BioD keeps 2 x the median read length.
/// biod_b200/bam.py: chunk_plan is the same arithmetic, checked against a line-by-line simulation of the D code.
struct GpuChunkPlan { size_t first, last, halo; int ref_id; ulong start_position, end_position; }

GpuChunkPlan[] chunkPlan(const(int)[] ref_id, const(int)[] pos, const(int)[] end_pos, const(int)[] rec_size,
                         size_t block_size = 16_384_000, ulong start_from = 0, ulong end_at = ulong.max) {
    import std.algorithm : max, min;
    GpuChunkPlan[] plan;
    const n = ref_id.length;
    size_t[2][] chunks;
    for (size_t i = 0; i < n;) {
        size_t total = 4 + rec_size[i], j = i + 1;
        while (total <= block_size && j < n && ref_id[j] == ref_id[i]) { total += 4 + rec_size[j]; ++j; }
        chunks ~= [i, j];
        i = j;
    }
    bool started = false, have_prev = false;
    size_t prev_last;
    foreach (k, ch; chunks) {
        const i = ch[0], j = ch[1];
        const rid = ref_id[i];
        if (rid < 0) continue;
        long right_end = long.min;
        foreach (r; i .. j) right_end = max(right_end, end_pos[r]);
        long beg;
        size_t halo = i;
        if (!started) {
            beg = pos[i];
            if (cast(ulong)beg >= end_at) break;
            if (right_end <= cast(long)start_from) continue;
            started = true;
        } else if (have_prev && ref_id[prev_last - 1] == rid) {
            beg = pos[prev_last - 1];
            size_t r0 = i;
            while (r0 > 0 && ref_id[r0 - 1] == rid) --r0;
            foreach (r; r0 .. i) if (end_pos[r] > beg) { halo = r; break; }
        } else {
            beg = pos[i];
        }
        long end = pos[j - 1];
        if (k + 1 >= chunks.length || ref_id[chunks[k + 1][0]] != ref_id[j - 1]) end = right_end;
        plan ~= GpuChunkPlan(i, j, halo, rid, max(cast(ulong)beg, start_from), min(cast(ulong)end, end_at));
        have_prev = true;
        prev_last = j;
    }
    return plan;
}

/// MaqSnpCaller.findSNPs over the GPU pileup (maq.d:489-540): the likelihoods are computed on the device, only the
/// calls come back.  Genotype codes are DiploidGenotype!Base5's (first * 5 + second).
struct GpuSnpCall { int ref_id; ulong position; char reference_base; ubyte genotype; float quality; }

GpuSnpCall[] findSNPs(GpuBamReader bam, biodb_maq_params mp = biodb_maq_params.init, bool single_ref = true) {
    biodb_pileup_params prm;
    prm.single_ref = single_ref; prm.skip_zero_coverage = true; prm.use_md_tag = true; prm.end_at = ulong.max; prm.maq_mode = 1;
    biodb_pileup* p;
    if (biodb_pileup_begin(bam.handle, &prm, &p) != BIODB_OK) raise(biodb_last_error(bam.handle));
    scope (exit) biodb_pileup_end(p);
    enforce(biodb_pileup_maq_params(p, &mp) == BIODB_OK, "invalid MAQ parameters");
    GpuSnpCall[] calls;
    biodb_column_batch b;
    while (true) {
        const st = biodb_pileup_next(p, &b);
        if (st == BIODB_EOF) break;
        if (st != BIODB_OK) raise(biodb_last_error(bam.handle));
        foreach (k; 0 .. cast(size_t)b.n_calls)
            calls ~= GpuSnpCall(b.ref_id, b.call_pos[k], cast(char)b.call_ref[k], b.call_gt[k], b.call_qual[k]);
    }
    return calls;
}
