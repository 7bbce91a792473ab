"""Host-side mirror of BioD's reader / pileup API over the C ABI (include/biod_b200.h).

The reference is D (bio/std/hts/bam/{reader,read,pileup}.d); no D toolchain exists in this image, so this
module plays the role the D binding (bindings/d/biod_b200.d) plays for real BioD users: same names,
argument meaning and error classes, every byte of work done by the CUDA library.

    bam = BamReader("file.bam")                    # bam/reader.d:127-138
    for read in bam.reads(): ...                   # reader.d:228-231  -> BamRead views (read.d:81)
    for column in makePileup(bam, start_from=..):  # pileup.d:683-694  -> PileupColumn (pileup.d:236-290)
    for column in pileupColumns(bam): ...          # pileup.d:509-519
"""
import ctypes as C
import os

import numpy as np

from . import _capi as capi


class BgzfException(Exception):
    """bio/core/bgzf/inputstream.d:41-43"""


class ZlibException(Exception):
    """bio/core/utils/zlib.d:247-274"""

    def __init__(self, msg, errnum):
        super().__init__(msg)
        self.errnum = errnum


class BamFormatException(Exception):
    """plain Exception of reader.d:113 / block.d:150"""


class ReadException(Exception):
    """contrib/undead/stream.d ReadException raised by readExact (readrange.d:169)"""


class PileupException(Exception):
    pass


class CudaUnavailable(RuntimeError):
    pass


def _raise(err):
    st, msg = err.status, err.message.decode("latin-1")
    if st == capi.ERR_BGZF:
        raise BgzfException(msg)
    if st == capi.ERR_ZLIB:
        raise ZlibException(msg, err.zlib_errnum)
    if st == capi.ERR_FORMAT:
        raise BamFormatException(msg)
    if st == capi.ERR_TRUNCATED:
        raise ReadException(msg)
    if st == capi.ERR_CUDA:
        raise CudaUnavailable(msg)
    if st in (capi.ERR_CIGAR, capi.ERR_UNSORTED):
        raise PileupException(msg)
    raise RuntimeError(f"biod_b200 status {st}: {msg}")


def _np(ptr, n, dtype):
    n = int(n)
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = n * np.dtype(dtype).itemsize
    addr = C.addressof(ptr.contents)
    return np.frombuffer((C.c_uint8 * nbytes).from_address(addr), dtype=dtype)


CIGAR_CHARS = "MIDNSHP=X???????"          # cigar.d:109-111
BASE_CHARS = np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)   # bio/core/base.d:85


class ReferenceSequenceInfo:
    """bam/referenceinfo.d:34-62"""

    def __init__(self, name, length):
        self.name, self.length = name, length


class RecordBatch:
    """One batch of decoded records: SoA tables (numpy views of library-owned pinned memory, valid until the
    next batch is fetched) plus the raw uncompressed slice the BamRead views point into."""

    FIELDS = {"rec_off": np.uint64, "block_size": np.int32, "ref_id": np.int32, "pos": np.int32, "end_pos": np.int32,
              "bin_mq_nl": np.uint32, "flag_nc": np.uint32, "l_seq": np.int32}

    def __init__(self, b, copy):
        n = int(b.n)
        self.n = n
        self.first_index = int(b.first_index)
        g = (lambda a: a.copy()) if copy else (lambda a: a)
        self.data = g(_np(b.data, b.data_len, np.uint8))
        for f, dt in self.FIELDS.items():
            setattr(self, f, g(_np(getattr(b, f), n, dt)))
        self.cigar_off = g(_np(b.cigar_off, n + 1, np.uint64))
        self.cigar = g(_np(b.cigar, int(self.cigar_off[n]) if n else 0, np.uint32))
        self.start_voffset = g(_np(b.start_voffset, n, np.uint64)) if b.start_voffset else None
        self.end_voffset = g(_np(b.end_voffset, n, np.uint64)) if b.end_voffset else None

    # derived columns (read.d:952-968)
    @property
    def flag(self):
        return (self.flag_nc >> 16).astype(np.uint16)

    @property
    def n_cigar(self):
        return (self.flag_nc & 0xFFFF).astype(np.uint16)

    @property
    def mapq(self):
        return ((self.bin_mq_nl >> 8) & 0xFF).astype(np.uint8)

    @property
    def l_read_name(self):
        return (self.bin_mq_nl & 0xFF).astype(np.uint8)

    @property
    def bin(self):
        return (self.bin_mq_nl >> 16).astype(np.uint16)

    def read(self, i):
        return BamRead(self, i)


class BamRead:
    """Zero-copy view over one raw BAM record (bam/read.d:81-1031, layout :907-1003)."""

    __slots__ = ("_b", "_i", "raw")

    def __init__(self, batch, i):
        self._b, self._i = batch, i
        o = int(batch.rec_off[i]) + 4
        self.raw = batch.data[o:o + int(batch.block_size[i])]

    ref_id = property(lambda s: int(s._b.ref_id[s._i]))
    position = property(lambda s: int(s._b.pos[s._i]))
    end_position = property(lambda s: int(s._b.end_pos[s._i]))
    flag = property(lambda s: int(s._b.flag_nc[s._i]) >> 16)
    mapping_quality = property(lambda s: (int(s._b.bin_mq_nl[s._i]) >> 8) & 0xFF)
    sequence_length = property(lambda s: int(s._b.l_seq[s._i]))
    index = property(lambda s: s._b.first_index + s._i)
    is_unmapped = property(lambda s: bool(s.flag & 4))
    # BamReadBlock (readrange.d:38-48); None unless the reader was opened with want_offsets
    start_virtual_offset = property(lambda s: None if s._b.start_voffset is None else int(s._b.start_voffset[s._i]))
    end_virtual_offset = property(lambda s: None if s._b.end_voffset is None else int(s._b.end_voffset[s._i]))
    is_reverse_strand = property(lambda s: bool(s.flag & 16))

    def basesCovered(self):
        return 0 if self.is_unmapped else self.end_position - self.position

    @property
    def name(self):
        ln = int(self._b.bin_mq_nl[self._i]) & 0xFF
        return self.raw[32:32 + ln - 1].tobytes().decode("latin-1")

    @property
    def cigar(self):
        a, b = int(self._b.cigar_off[self._i]), int(self._b.cigar_off[self._i + 1])
        return [(int(x) >> 4, CIGAR_CHARS[int(x) & 15]) for x in self._b.cigar[a:b]]

    def cigarString(self):
        c = self.cigar
        return "".join(f"{l}{o}" for l, o in c) if c else "*"

    def _seq_off(self):
        return 32 + (int(self._b.bin_mq_nl[self._i]) & 0xFF) + 4 * (int(self._b.flag_nc[self._i]) & 0xFFFF)

    @property
    def sequence(self):
        n, o = self.sequence_length, self._seq_off()
        packed = self.raw[o:o + (n + 1) // 2]
        codes = np.empty(2 * len(packed), dtype=np.uint8)
        codes[0::2] = packed >> 4
        codes[1::2] = packed & 15
        return BASE_CHARS[codes[:n]].tobytes().decode()

    @property
    def base_qualities(self):
        n = self.sequence_length
        o = self._seq_off() + (n + 1) // 2
        return self.raw[o:o + n]

    @property
    def tags_raw(self):
        n = self.sequence_length
        return self.raw[self._seq_off() + (n + 1) // 2 + n:].tobytes()


class BaiFile:
    """bam/baifile.d:85-169: a parsed BAI index.  `source` is the bytes of the .bai file, its path, or the path of the
    BAM file (then `<bam>.bai` and `<bam without extension>.bai` are tried, baifile.d:95-113).  Host only."""

    def __init__(self, source):
        self._L = L = capi.lib()
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            data = bytes(source)
        else:
            path = str(source)
            if not path.endswith(".bai"):
                first, second = path + ".bai", path[:path.rfind(".") + 1] + "bai"
                if os.path.exists(first):
                    path = first
                elif os.path.exists(second):
                    path = second
                else:
                    raise Exception(f"searched for {first} or {second}, found neither")
            with open(path, "rb") as f:
                data = f.read()
        self._buf = np.frombuffer(data, dtype=np.uint8)
        h = C.c_void_p()
        st = L.biodb_index_open(self._buf.ctypes.data, self._buf.size, C.byref(h))
        if st != capi.OK:
            _raise(L.biodb_open_error().contents)
        self._h = h
        self.n_refs = int(L.biodb_index_n_refs(h))

    def close(self):
        if getattr(self, "_h", None):
            self._L.biodb_index_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def chunks(self, ref_id, beg, end):
        """RandomAccessManager.getChunks (randomaccessmanager.d:222-244): [(beg voffset, end voffset), ...]."""
        n = int(self._L.biodb_index_chunks(self._h, ref_id, beg, end, None, 0))
        if n < 0:
            raise Exception("Invalid reference sequence index")
        out = np.zeros(2 * max(n, 1), dtype=np.uint64)
        self._L.biodb_index_chunks(self._h, ref_id, beg, end, out.ctypes.data, n)
        return [(int(out[2 * k]), int(out[2 * k + 1])) for k in range(n)]


class RegionReads:
    """The range `bam["chr"][start .. end)` (reference.d:76-81): iterating it yields the reads of the region; handing it
    to makePileup / pileupColumns gives the pileup of exactly these reads (examples/read_bam_file.d:22-25)."""

    def __init__(self, reader, ref_id, start, end):
        if not start < end:
            raise Exception("start must be less than end")                       # reference.d:77
        self.reader, self.ref_id, self.start, self.end = reader, ref_id, start, end

    def __iter__(self):
        return self.reader.region_reads(self.ref_id, self.start, self.end)


class ReferenceSequence:
    """bam/reference.d:37-160: `reader["chr1"]`; slicing it gives the reads that overlap [start, end)."""

    def __init__(self, reader, ref_id, info):
        self._reader, self.id, self.name, self.length = reader, ref_id, info.name, info.length

    def __getitem__(self, sl):
        start = 0 if sl.start is None else sl.start
        end = self.length if sl.stop is None else sl.stop
        return RegionReads(self._reader, self.id, start, end)

    def reads(self):
        return self[0:self.length]

    def startVirtualOffset(self):
        """reference.d:96-102 (the EOF fallback for references without reads is not mirrored)."""
        for r in self.reads():
            return r.start_virtual_offset
        return None

    def firstPosition(self):
        for r in self.reads():
            return r.position
        return -1


class BamReader:
    """bam/reader.d:80-598 — the subset on the hot path."""

    def __init__(self, source, blocks_per_batch=0, want_offsets=False, device=-1, task_pool=None, verify_crc=False,
                 index=None):
        # task_pool is accepted and ignored (reader.d:100-101): the device is the pool
        self._L = L = capi.lib()
        o = capi.Options()
        L.biodb_default_options(C.byref(o))
        o.device = device
        if blocks_per_batch:
            o.blocks_per_batch = blocks_per_batch
        o.want_offsets = int(want_offsets)
        o.verify_crc = int(verify_crc)
        h = C.c_void_p()
        if isinstance(source, (bytes, bytearray, memoryview, np.ndarray)):
            self._buf = np.frombuffer(source, dtype=np.uint8)
            self.filename = None
            st = L.biodb_open_memory(self._buf.ctypes.data, self._buf.size, C.byref(o), C.byref(h))
        else:
            self.filename = str(source)
            st = L.biodb_open(self.filename.encode(), C.byref(o), C.byref(h))
        if st != capi.OK:
            _raise(L.biodb_open_error().contents)
        self._h = h
        t, n = C.c_char_p(), C.c_size_t()
        L.biodb_header_text(h, C.byref(t), C.byref(n))
        self.header_text = C.string_at(t, n.value).decode("latin-1") if n.value else ""
        self.reference_sequences = []
        for i in range(L.biodb_n_refs(h)):
            nm, nl, ln = C.c_char_p(), C.c_int32(), C.c_int32()
            L.biodb_ref_info(h, i, C.byref(nm), C.byref(nl), C.byref(ln))
            self.reference_sequences.append(ReferenceSequenceInfo(C.string_at(nm, nl.value).decode("latin-1"), ln.value))
        self.reads_start_voffset = int(L.biodb_reads_start_voffset(h))
        # random access (reader.d:424-447): an index given explicitly, or looked for next to the file on first use
        self._index = index if (index is None or isinstance(index, BaiFile)) else BaiFile(index)

    def _bai(self):
        if self._index is None:
            if self.filename is None:
                raise Exception("BAM index file (.bai) must be provided")      # randomaccessmanager.d:202-204
            self._index = BaiFile(self.filename)
        return self._index

    def __getitem__(self, ref_name):
        """reader.d:424-429"""
        for i, r in enumerate(self.reference_sequences):
            if r.name == ref_name:
                return ReferenceSequence(self, i, r)
        raise Exception(f"Reference with name {ref_name} does not exist")

    def reference(self, ref_id):
        """reader.d:435-440"""
        return ReferenceSequence(self, ref_id, self.reference_sequences[ref_id])

    def region_batches(self, ref_id, start, end, copy=False):
        """Batches holding the reads of reference ref_id that overlap [start, end) (randomaccessmanager.d:300-305)."""
        if not start < end:
            raise Exception("start must be less than end")                       # reference.d:77
        L = self._L
        it = C.c_void_p()
        st = L.biodb_reads_begin_region(self._h, self._bai()._h, ref_id, start, end, C.byref(it))
        if st == capi.ERR_ARG:
            raise Exception(L.biodb_last_error(self._h).contents.message.decode("latin-1"))
        if st != capi.OK:
            self._err()
        try:
            while True:
                b = capi.RecordBatch()
                st = L.biodb_reads_next(it, C.byref(b))
                if st == capi.EOF:
                    return
                if st != capi.OK:
                    self._err()
                yield RecordBatch(b, copy)
        finally:
            L.biodb_reads_end(it)

    def regions_batches(self, regions, copy=False):
        """Batches of getReadsOverlapping(regions) (reader.d:361, randomaccessmanager.d:316-337): regions = BamRegion-like
        triples (ref_id, start, end) in any order; they are grouped by reference, and the groups are read one after the
        other in reference order, each through biodb_reads_begin_regions."""
        by_ref = {}
        for ref_id, start, end in regions:
            if not start < end:
                raise Exception("start must be less than end")
            by_ref.setdefault(int(ref_id), []).append((int(start), int(end)))
        L = self._L
        for ref_id in sorted(by_ref):
            begs = np.array([a for a, _ in by_ref[ref_id]], dtype=np.uint32)
            ends = np.array([b for _, b in by_ref[ref_id]], dtype=np.uint32)
            it = C.c_void_p()
            st = L.biodb_reads_begin_regions(self._h, self._bai()._h, ref_id, len(begs), begs.ctypes.data, ends.ctypes.data, C.byref(it))
            if st == capi.ERR_ARG:
                raise Exception(L.biodb_last_error(self._h).contents.message.decode("latin-1"))
            if st != capi.OK:
                self._err()
            try:
                while True:
                    b = capi.RecordBatch()
                    st = L.biodb_reads_next(it, C.byref(b))
                    if st == capi.EOF:
                        break
                    if st != capi.OK:
                        self._err()
                    yield RecordBatch(b, copy)
            finally:
                L.biodb_reads_end(it)

    def getReadsOverlapping(self, regions):
        """reader.d:361: the reads that overlap any of the regions [(ref_id, start, end), ...], every read once, in file
        order within a reference, references in ascending order."""
        for batch in self.regions_batches(regions, copy=True):
            for i in range(batch.n):
                yield BamRead(batch, i)

    def getReadsBetween(self, from_voffset, to_voffset=None, max_blocks=0):
        """reader.d:350-356: the reads from one virtual offset (the start of a record) to another (the end of one;
        None = the end of the file)."""
        L = self._L
        it = C.c_void_p()
        to = 2**64 - 1 if to_voffset is None else int(to_voffset)
        if L.biodb_reads_begin_between(self._h, int(from_voffset), to, max_blocks, C.byref(it)) != capi.OK:
            self._err()
        try:
            while True:
                b = capi.RecordBatch()
                st = L.biodb_reads_next(it, C.byref(b))
                if st == capi.EOF:
                    return
                if st != capi.OK:
                    self._err()
                batch = RecordBatch(b, True)
                for i in range(batch.n):
                    yield BamRead(batch, i)
        finally:
            L.biodb_reads_end(it)

    def getReadAt(self, voffset):
        """reader.d:336-339: the read that starts at a virtual offset."""
        for r in self.getReadsBetween(voffset, None, max_blocks=2):
            return r
        raise ReadException("not enough data in stream")

    BGZF_EOF = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])

    def eofVirtualOffset(self):
        """reader.d:177-179, randomaccessmanager.d:113-137: the start of the EOF block if the file ends with one, else the
        physical end of the file."""
        if self.filename is None:
            size, tail = int(self._buf.size), bytes(self._buf[-28:])
        else:
            size = os.path.getsize(self.filename)
            with open(self.filename, "rb") as f:
                f.seek(max(0, size - 28))
                tail = f.read()
        return ((size - 28) << 16) if tail == self.BGZF_EOF else (size << 16)

    def unmappedReads(self):
        """reader.d:369-390: the reads at the end of a coordinate-sorted, indexed file whose reference id is -1."""
        bai = self._bai()
        start = self.eofVirtualOffset()
        try:
            if self.getReadAt(self.reads_start_voffset).ref_id == -1:
                start = self.reads_start_voffset
        except ReadException:
            pass                                                                # no reads at all
        off = C.c_uint64()
        if self._L.biodb_index_last_linear_offset(bai._h, len(self.reference_sequences), C.byref(off)):
            start = int(off.value)
        started = False
        for r in self.getReadsBetween(start, None):
            if started or r.ref_id == -1:
                started = True
                yield r

    def region_reads(self, ref_id, start, end):
        for batch in self.region_batches(ref_id, start, end, copy=True):
            for i in range(batch.n):
                yield BamRead(batch, i)

    def close(self):
        if getattr(self, "_h", None):
            self._L.biodb_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _err(self):
        _raise(self._L.biodb_last_error(self._h).contents)

    def hasReference(self, name):
        return any(r.name == name for r in self.reference_sequences)

    def read_batches(self, copy=False):
        """Batches of decoded records; every call starts a fresh, independent pass (reader.d:228-231,553-569)."""
        L = self._L
        it = C.c_void_p()
        if L.biodb_reads_begin(self._h, C.byref(it)) != capi.OK:
            self._err()
        try:
            while True:
                b = capi.RecordBatch()
                st = L.biodb_reads_next(it, C.byref(b))
                if st == capi.EOF:
                    return
                if st != capi.OK:
                    self._err()
                yield RecordBatch(b, copy)
        finally:
            L.biodb_reads_end(it)

    def reads(self):
        for batch in self.read_batches(copy=True):
            for i in range(batch.n):
                yield BamRead(batch, i)

    def column_batches(self, single_ref, use_md_tag=False, start_from=0, end_at=2**64 - 1, skip_zero_coverage=True,
                       want_query_offset=False, copy=False, shard=None, halo_blocks=8, halo_voffset=None, shard_info=None,
                       counts_only=False, compact_reads=False, region=None, record_range=None, maq=None, maq_mode=0):
        """shard=(index, n_shards) runs one shard of a sharded pileup (pileupChunks semantics, pileup.d:859-1015) — or
        shard=(first, n_shards, count) the shards [first, first + count) as one pass: its halo
        starts halo_blocks BGZF blocks in front of it (a guess) or at the record at halo_voffset; shard_info receives
        the biodb_shard_info fields plus "reach" (see biod_b200.stitch.exact_halos).
        region=(ref_id, start, end) piles up the reads of bam[ref][start .. end) only (read_idx then counts those).
        record_range=(from_voffset, to_voffset, lo_ref, lo_pos, hi_ref, hi_pos): the records that start in
        [from, to), columns clipped to the keys [lo, hi) (biodb_pileup_begin_range)."""
        L = self._L
        p = capi.PileupParams()
        p.single_ref, p.skip_zero_coverage, p.use_md_tag = int(single_ref), int(skip_zero_coverage), int(use_md_tag)
        p.want_query_offset, p.start_from, p.end_at = int(want_query_offset), start_from, end_at
        p.counts_only = int(counts_only)
        p.compact_reads = int(compact_reads)
        p.maq_mode = int(maq_mode)
        pl = C.c_void_p()
        if region is not None:
            st = L.biodb_pileup_begin_region(self._h, self._bai()._h, region[0], region[1], region[2], C.byref(p), C.byref(pl))
        elif record_range is not None:
            st = L.biodb_pileup_begin_range(self._h, C.byref(p), *[int(x) for x in record_range], C.byref(pl))
        elif shard is not None and halo_voffset is not None:
            st = L.biodb_pileup_begin_shard_span_at(self._h, C.byref(p), shard[0], shard[2] if len(shard) > 2 else 1, shard[1],
                                                    int(halo_voffset), C.byref(pl))
        elif shard is not None:
            st = L.biodb_pileup_begin_shard_span(self._h, C.byref(p), shard[0], shard[2] if len(shard) > 2 else 1, shard[1],
                                                 halo_blocks, C.byref(pl))
        else:
            st = L.biodb_pileup_begin(self._h, C.byref(p), C.byref(pl))
        if st == capi.ERR_ARG:
            raise ValueError("biodb_pileup_begin: invalid arguments")
        if st != capi.OK:
            self._err()
        if maq_mode and maq is not None:
            mp = capi.MaqParams(maq.depcorr, maq.eta, maq.minimum_call_quality, maq.minimum_base_quality)
            if L.biodb_pileup_maq_params(pl, C.byref(mp)) != capi.OK:
                L.biodb_pileup_end(pl)
                raise ValueError("biodb_pileup_maq_params: invalid arguments")
        try:
            while True:
                cb = capi.ColumnBatch()
                st = L.biodb_pileup_next(pl, C.byref(cb))
                if st == capi.EOF:
                    if shard_info is not None:
                        si = capi.ShardInfo()
                        L.biodb_pileup_shard_info(pl, C.byref(si))
                        shard_info.update({f: getattr(si, f) for f, _ in si._fields_})
                        if shard is not None:
                            reach = (C.c_uint64 * shard[1])()
                            L.biodb_pileup_shard_reach(pl, reach)
                            shard_info["reach"] = [int(x) for x in reach]
                    return
                if st != capi.OK:
                    self._err()
                yield ColumnBatch(cb, copy)
        finally:
            L.biodb_pileup_end(pl)

    def shard_cuts(self, n_shards):
        """(voffset, ref_id, position) of the n_shards + 1 cuts of a sharded pileup (biodb_shard_cuts)."""
        vo = (C.c_uint64 * (n_shards + 1))()
        rf = (C.c_int32 * (n_shards + 1))()
        ps = (C.c_int64 * (n_shards + 1))()
        if self._L.biodb_shard_cuts(self._h, n_shards, vo, rf, ps) != capi.OK:
            self._err()
        return [(int(vo[k]), int(rf[k]), int(ps[k])) for k in range(n_shards + 1)]

    def sharded_column_batches(self, n_shards, halo_blocks=8, shard_infos=None, **kw):
        """All shards of a sharded pileup, one after the other on this GPU, with EXACT halos: shard t's halo starts at
        the first record of an earlier shard that reaches into its columns (the minimum of the earlier shards'
        "reach" reports — known by the time shard t runs, so nothing is run twice here; across GPUs the shards run
        at once from a guess and biod_b200.stitch.exact_halos says which to run again).  Yields (shard, batch)."""
        need = [2**64 - 1] * n_shards
        cuts = self.shard_cuts(n_shards)
        spans = kw.pop("spans", None) or [1] * n_shards        # spans=[2, 1, 4]: shards 0-1, 2, 3-6 as three passes
        assert sum(spans) == n_shards and all(c > 0 for c in spans)
        s = 0
        for count in spans:
            info = {}
            hv = min(need[s], cuts[s][0])
            for b in self.column_batches(False, shard=(s, n_shards, count), halo_voffset=hv, shard_info=info, **kw):
                yield s, b
            for t in range(s + count, n_shards):
                need[t] = min(need[t], info["reach"][t])
            if shard_infos is not None:
                shard_infos.append(info)
            s += count


def _popcount64(x):
    x = x.astype(np.uint64)
    c = np.zeros(len(x), dtype=np.int64)
    for k in range(8):
        c += _POP8[((x >> np.uint64(8 * k)) & np.uint64(0xFF)).astype(np.int64)]
    return c


_POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int64)


def expand_compact_columns(n_columns, last, mask, strag_col, strag_idx, run_pos, run_first_col):
    """The explicit column table (position, col_off, read_idx) from the sequential compact encoding
    (include/biod_b200.h, biodb_column_batch with compact_reads)."""
    nc = int(n_columns)
    position = np.zeros(nc, dtype=np.uint64)
    for r in range(len(run_pos)):
        a, b = int(run_first_col[r]), int(run_first_col[r + 1])
        position[a:b] = np.uint64(run_pos[r]) + np.arange(b - a, dtype=np.uint64)
    ns = np.bincount(strag_col.astype(np.int64), minlength=nc).astype(np.int64) if nc else np.zeros(0, dtype=np.int64)
    cov = ns + _popcount64(mask)
    col_off = np.concatenate([[0], np.cumsum(cov)]).astype(np.uint64)
    soff = np.concatenate([[0], np.cumsum(ns)]).astype(np.uint32)
    read_idx = expand_compact_reads(col_off, last, mask, soff, strag_idx, int(col_off[-1]))
    return position, col_off, read_idx


def expand_compact_reads(col_off, last, mask, soff, sidx, n_entries):
    """read_idx of every entry from (last read, window mask, straggler offsets / indices) per column."""
    nc = len(last)
    out = np.zeros(n_entries, dtype=np.uint32)
    if nc == 0:
        return out
    cov = np.diff(col_off).astype(np.int64)
    ns = np.diff(soff).astype(np.int64)
    # bits of every mask, d = 63..0 (column order = ascending record index = descending d)
    bits = ((mask[:, None] >> np.arange(63, -1, -1, dtype=np.uint64)[None, :]) & np.uint64(1)).astype(bool)
    nb = bits.sum(axis=1)
    assert np.array_equal(nb + ns, cov), "compact read lists do not add up to the coverage"
    d = np.broadcast_to(np.arange(63, -1, -1, dtype=np.int64)[None, :], bits.shape)[bits]
    col = np.broadcast_to(np.arange(nc, dtype=np.int64)[:, None], bits.shape)[bits]
    win = (last[col].astype(np.int64) - d).astype(np.uint32)
    # place: stragglers first, then the window reads
    start = col_off[:-1].astype(np.int64)
    k_in_col = np.arange(len(col), dtype=np.int64) - np.repeat(np.cumsum(nb) - nb, nb)
    out[start[col] + ns[col] + k_in_col] = win
    if len(sidx):
        scol = np.repeat(np.arange(nc, dtype=np.int64), ns)
        k = np.arange(len(sidx), dtype=np.int64) - np.repeat(soff[:-1].astype(np.int64), ns)
        out[start[scol] + k] = sidx
    return out


class ColumnBatch:
    def __init__(self, cb, copy):
        g = (lambda a: a.copy()) if copy else (lambda a: a)
        nc, ne = int(cb.n_columns), int(cb.n_entries)
        self.n_columns, self.n_entries, self.ref_id = nc, ne, int(cb.ref_id)
        self.compact = None
        # maq_mode: the calls of findSNPs, and (mode 2) the two best genotypes of every column
        ncall = int(cb.n_calls)
        self.calls = None
        if cb.call_pos or cb.maq_gt0:
            self.calls = dict(col=g(_np(cb.call_col, ncall, np.uint32)), pos=g(_np(cb.call_pos, ncall, np.uint64)),
                              gt=g(_np(cb.call_gt, ncall, np.uint8)), ref=g(_np(cb.call_ref, ncall, np.uint8)),
                              qual=g(_np(cb.call_qual, ncall, np.float32)))
        self.maq = None
        if cb.maq_gt0:
            self.maq = dict(gt0=g(_np(cb.maq_gt0, nc, np.uint8)), gt1=g(_np(cb.maq_gt1, nc, np.uint8)),
                            s0=g(_np(cb.maq_s0, nc, np.float32)), s1=g(_np(cb.maq_s1, nc, np.float32)),
                            n_valid=g(_np(cb.maq_n_valid, nc, np.uint16)))
        if self.calls is not None and not cb.position:
            # maq_mode 1: nothing but the calls leaves the device
            self.position = self.col_off = self.n_starting_here = self.read_idx = self.base = self.qual = None
            self.query_offset = self.counts = self.reference_base = None
            return
        self.n_starting_here = g(_np(cb.n_starting_here, nc, np.uint32))
        if cb.last_read:
            # compact_reads: sequential encoding -> the same explicit table
            last = _np(cb.last_read, nc, np.uint32).copy()
            mask = _np(cb.live_mask, nc, np.uint64).copy()
            nst, nr = int(cb.n_stragglers), int(cb.n_runs)
            scol = _np(cb.strag_col, nst, np.uint32).copy()
            sidx = _np(cb.strag_idx, nst, np.uint32).copy()
            rpos = _np(cb.run_pos, nr, np.uint64).copy()
            rfirst = _np(cb.run_first_col, nr + 1, np.uint32).copy()
            self.compact = (last, mask, scol, sidx, rpos, rfirst)
            self.position, self.col_off, self.read_idx = expand_compact_columns(nc, last, mask, scol, sidx, rpos, rfirst)
            assert int(self.col_off[-1]) == ne
        else:
            self.position = g(_np(cb.position, nc, np.uint64))
            self.col_off = g(_np(cb.col_off, nc + 1, np.uint64))
            self.read_idx = g(_np(cb.read_idx, ne, np.uint32)) if cb.read_idx else np.zeros(0, dtype=np.uint32)
        if cb.base4 and not cb.base:
            # compact_reads: bases two per byte + the few entries that are not a base
            b4 = _np(cb.base4, (ne + 1) // 2, np.uint8)
            codes = np.empty(2 * len(b4), dtype=np.uint8)
            codes[0::2] = b4 >> 4
            codes[1::2] = b4 & 15
            self.base = BASE_CHARS[codes[:ne]]
            nsp = int(cb.n_special)
            if nsp:
                self.base[_np(cb.special_entry, nsp, np.uint32).astype(np.int64)] = _np(cb.special_base, nsp, np.uint8)
        else:
            self.base = g(_np(cb.base, ne, np.uint8)) if cb.base else None
        self.qual = g(_np(cb.qual, ne, np.uint8)) if cb.qual else None
        self.query_offset = g(_np(cb.query_offset, ne, np.uint32)) if cb.query_offset else None
        self.counts = g(_np(cb.counts, nc * 6, np.uint32)).reshape(nc, 6) if cb.counts else None
        # use_md_tag: PileupColumn.reference_base per column (pileup.d:252-254)
        self.reference_base = g(_np(cb.reference_base, nc, np.uint8)) if cb.reference_base else None


class PileupColumn:
    """bam/pileup.d:236-290"""

    __slots__ = ("_b", "_c")

    def __init__(self, batch, c):
        self._b, self._c = batch, c

    def _rng(self):
        return int(self._b.col_off[self._c]), int(self._b.col_off[self._c + 1])

    position = property(lambda s: int(s._b.position[s._c]))
    ref_id = property(lambda s: s._b.ref_id)

    @property
    def reference_base(self):
        """pileup.d:252-254; 'N' unless the pileup was made with use_md_tag (PileupColumn's default, pileup.d:239)."""
        rb = self._b.reference_base
        return "N" if rb is None else chr(int(rb[self._c]))

    @property
    def coverage(self):
        a, b = self._rng()
        return b - a

    @property
    def reads(self):
        a, b = self._rng()
        return self._b.read_idx[a:b]

    @property
    def reads_starting_here(self):
        a, b = self._rng()
        return self._b.read_idx[b - int(self._b.n_starting_here[self._c]):b]

    @property
    def bases(self):
        a, b = self._rng()
        return self._b.base[a:b].tobytes().decode()

    @property
    def base_qualities(self):
        a, b = self._rng()
        return self._b.qual[a:b]


def _columns(reader, single_ref, **kw):
    if isinstance(reader, RegionReads):                      # makePileup(bam["chr"][a .. b), ...)
        kw["region"] = (reader.ref_id, reader.start, reader.end)
        reader = reader.reader
    for batch in reader.column_batches(single_ref, copy=True, **kw):
        for c in range(batch.n_columns):
            yield PileupColumn(batch, c)


def makePileup(reader, use_md_tag=False, start_from=0, end_at=2**64 - 1, skip_zero_coverage=True):
    """bam/pileup.d:683-694"""
    return _columns(reader, True, use_md_tag=use_md_tag, start_from=start_from, end_at=end_at,
                    skip_zero_coverage=skip_zero_coverage)


def pileupColumns(reader, use_md_tag=False, skip_zero_coverage=True):
    """bam/pileup.d:509-519"""
    return _columns(reader, False, use_md_tag=use_md_tag, skip_zero_coverage=skip_zero_coverage)


# ---- MAQ SNP caller (bio/std/hts/snpcallers/maq.d) ---------------------------------------------------------------------
_BASE5 = "ACGTN"


class DiploidCall5:
    """bio/core/call.d:29-92 (Call!(DiploidGenotype, Base5)): what MaqSnpCaller.makeCall / findSNPs hand out."""

    __slots__ = ("sample", "chromosome", "position", "reference_base", "genotype_code", "quality")

    def __init__(self, sample, chromosome, position, reference_base, genotype_code, quality):
        self.sample, self.chromosome, self.position = sample, chromosome, int(position)
        self.reference_base, self.genotype_code, self.quality = reference_base, int(genotype_code), float(quality)

    base1 = property(lambda s: _BASE5[s.genotype_code // 5])             # genotype.d:53-60
    base2 = property(lambda s: _BASE5[s.genotype_code % 5])
    is_heterozygous = property(lambda s: s.base1 != s.base2)
    is_homozygous = property(lambda s: s.base1 == s.base2)
    genotype = property(lambda s: s.base1 + "|" + s.base2)              # genotype.d:73-75

    @property
    def is_variant(self):                                                # call.d:88-90
        r = self.reference_base.upper()
        r = r if r in "ACGT" else "N"
        return (self.base1, self.base2) != (r, r)


class MaqSnpCaller:
    """bio/std/hts/snpcallers/maq.d:319-540.  The likelihoods are computed on the GPU over the columns of the pileup it
    builds (maq_mode of biodb_pileup_params): only calls — or, for genotypeLikelihoods, 12 bytes per column — come back."""

    def __init__(self, depcorr=0.17, eta=0.03, minimum_call_quality=6.0, minimum_base_quality=13):
        self.depcorr, self.eta = depcorr, eta
        self.minimum_call_quality, self.minimum_base_quality = minimum_call_quality, minimum_base_quality

    def findSNPs(self, reads, reference="", sample="", single_ref=True, **kw):
        """maq.d:489-540: the calls that differ from the reference base (from the reads' MD tags) and whose quality exceeds
        minimum_call_quality.  `reads`: a BamReader or bam["chr"][a:b]; single_ref=True is makePileup(reads, true)."""
        region = None
        if isinstance(reads, RegionReads):
            region, reads = (reads.ref_id, reads.start, reads.end), reads.reader
        for b in reads.column_batches(single_ref, use_md_tag=True, region=region, maq=self, maq_mode=1, copy=True, **kw):
            chrom = reference or (reads.reference_sequences[b.ref_id].name if 0 <= b.ref_id < len(reads.reference_sequences) else "")
            c = b.calls
            for k in range(len(c["pos"])):
                yield DiploidCall5(sample, chrom, c["pos"][k], chr(int(c["ref"][k])), c["gt"][k], c["qual"][k])

    def genotypeLikelihoods(self, reads, single_ref=True, use_md_tag=True, **kw):
        """genotypeLikelihoodInfo / makeCall of every column (maq.d:388-486), batch by batch: yields ColumnBatch objects
        whose .maq holds gt0 / gt1 / s0 / s1 / n_valid per column (the call's quality is s1 - s0) next to position,
        col_off, n_starting_here and reference_base."""
        region = None
        if isinstance(reads, RegionReads):
            region, reads = (reads.ref_id, reads.start, reads.end), reads.reader
        yield from reads.column_batches(single_ref, use_md_tag=use_md_tag, region=region, maq=self, maq_mode=2, copy=True, **kw)


# ---- pileupChunks (bam/pileup.d:859-1015, bam/splitter.d:66-101) ---------------------------------------------------
def chunk_plan(ref_id, pos, end_pos, rec_size, block_size=16_384_000, start_from=0, end_at=2**64 - 1):
    """Where BioD's pileupChunks cuts a range of reads, from the per-read tables alone (host arithmetic, no GPU):
    chunksConsumingLessThan (splitter.d:72-90: a chunk takes reads while the bytes taken so far are <= block_size and the
    reference stays that of its first read; size_in_bytes = 4 + block_size of the record, read.d:609-611) and
    PileupChunkRange's constructor / front / popFront (pileup.d:876-940).  Returns a list of dicts
    (first, last, halo, ref_id, start_position, end_position): the chunk's reads [first, last), the first read of its
    halo, and its column interval.  The halo is EXACT — the first earlier read of the reference that reaches beyond the
    interval's start — where BioD keeps the previous chunk's reads from 2 x the median read length before it
    (pileup.d:941-985, an approximation that loses longer reads)."""
    n = len(ref_id)
    size = rec_size.astype(np.int64) + 4
    cum = np.concatenate([[0], np.cumsum(size)])
    chunks = []
    i = 0
    while i < n:
        # getNextChunk: the first read always; then reads while total_size <= block_size (checked before each is added)
        j = int(np.searchsorted(cum, cum[i] + block_size, side="right"))     # first j with cum[j] - cum[i] > block_size
        j = max(i + 1, min(j, n))
        other = np.nonzero(ref_id[i + 1:j] != ref_id[i])[0]
        if other.size:
            j = i + 1 + int(other[0])
        chunks.append((i, j))
        i = j
    # exact halos: per reference run, the running maximum of the end positions (non-decreasing, so "the first read
    # that reaches beyond x" is a binary search)
    run_start = np.zeros(n, dtype=np.int64)
    pm = np.zeros(n, dtype=np.int64)
    if n:
        edges = np.concatenate([[0], np.flatnonzero(np.diff(ref_id)) + 1, [n]])
        for a, b in zip(edges[:-1], edges[1:]):
            run_start[a:b] = a
            pm[a:b] = np.maximum.accumulate(end_pos[a:b].astype(np.int64))
    plan = []
    started = False
    prev = None                                    # (first, last) of the previous chunk handed out
    for k, (i, j) in enumerate(chunks):
        rid = int(ref_id[i])
        if rid < 0:
            continue
        right_end = int(end_pos[i:j].max())
        if not started:
            beg = int(pos[i])
            if beg >= end_at:
                break
            if right_end <= start_from:
                continue
            started = True
            halo = i
        else:
            if prev is not None and int(ref_id[prev[1] - 1]) == rid:
                beg = int(pos[prev[1] - 1])
            else:
                beg = int(pos[i])
                prev = None
            halo = i
            if prev is not None:
                # the first earlier read of this reference that covers a position beyond beg
                r0 = int(run_start[i])
                halo = min(i, r0 + int(np.searchsorted(pm[r0:i], beg, side="right")))
        # front: up to the position of the chunk's last read, or — when the next chunk of the range (whatever its
        # reference id) is on another reference, or there is none — to the right end of the chunk's reads
        nxt = chunks[k + 1] if k + 1 < len(chunks) else None
        end = int(pos[j - 1])
        if nxt is None or int(ref_id[nxt[0]]) != int(ref_id[j - 1]):
            end = right_end
        plan.append(dict(first=i, last=j, halo=halo, ref_id=rid, start_position=max(beg, start_from), end_position=min(end, end_at)))
        prev = (i, j)
    return plan


class PileupChunk:
    """One element of pileupChunks: the pileup of a chunk's reads and its halo over [start_position, end_position)
    (makePileup(chain(prev_chunk, chunk), use_md_tag, beg, end), pileup.d:905-913).  Chunks are independent of one
    another: any number may be iterated at the same time, on one GPU or several."""

    def __init__(self, reader, plan, from_vo, to_vo, use_md_tag):
        self._reader, self._from, self._to, self._md = reader, from_vo, to_vo, use_md_tag
        self.ref_id, self.start_position, self.end_position = plan["ref_id"], plan["start_position"], plan["end_position"]
        self.first_read_index = plan["halo"]          # read_idx of the columns counts from this read of the file
        self.reads = (plan["first"], plan["last"])

    def column_batches(self, **kw):
        if self.start_position >= self.end_position:
            return iter(())
        return self._reader.column_batches(False, use_md_tag=self._md,
                                           record_range=(self._from, self._to, self.ref_id, self.start_position, self.ref_id,
                                                         self.end_position), **kw)

    def __iter__(self):
        for batch in self.column_batches(copy=True):
            for c in range(batch.n_columns):
                yield PileupColumn(batch, c)


def pileupChunks(reader, use_md_tag=False, block_size=16_384_000, start_from=0, end_at=2**64 - 1):
    """bam/pileup.d:1011-1015: non-overlapping consecutive pileups that can be processed in parallel.  One pass over the
    records (GPU: inflate + record scan, field tables only) finds the cuts; every chunk is then a range pileup of its
    own (biodb_pileup_begin_range).  The reader must have been opened with want_offsets=True."""
    ref, pos, end, size, vo = [], [], [], [], []
    for b in reader.read_batches():
        if b.start_voffset is None:
            raise Exception("pileupChunks needs virtual offsets: open the reader with want_offsets=True")
        ref.append(b.ref_id.copy()); pos.append(b.pos.copy()); end.append(b.end_pos.copy())
        size.append(b.block_size.copy()); vo.append(b.start_voffset.copy())
    cat = lambda v, dt: np.concatenate(v) if v else np.zeros(0, dtype=dt)  # noqa: E731
    ref, pos, end, size, vo = cat(ref, np.int32), cat(pos, np.int32), cat(end, np.int32), cat(size, np.int32), cat(vo, np.uint64)
    for pl in chunk_plan(ref, pos, end, size, block_size, start_from, end_at):
        to_vo = int(vo[pl["last"]]) if pl["last"] < len(vo) else 2**64 - 1
        yield PileupChunk(reader, pl, int(vo[pl["halo"]]), to_vo, use_md_tag)


# ---- BGZF compression (bgzf/compress.d, bgzf/outputstream.d) --------------------------------------------------------
def bgzf_compress(data, level=-1, eof=True, device=-1):
    """The BGZF stream of `data`: one block per 0xFF00 bytes (outputstream.d:50-223), each compressed on the GPU the way
    bgzfCompress frames it (compress.d:43-103), plus the EOF block of close() when eof is true."""
    L = capi.lib()
    if not -1 <= level <= 9:
        raise ValueError("level must be within -1 .. 9")                         # compress.d:46-48
    buf = np.frombuffer(bytes(data), dtype=np.uint8)
    cap = int(L.biodb_bgzf_compress_bound(buf.size))
    out = np.empty(cap, dtype=np.uint8)
    n = C.c_size_t()
    st = L.biodb_bgzf_compress(device, buf.ctypes.data if buf.size else None, buf.size, level, int(eof), out.ctypes.data, cap,
                               C.byref(n))
    if st == capi.ERR_CUDA:
        raise CudaUnavailable("biodb_bgzf_compress needs a CUDA device: biod_b200 has no CPU fallback")
    if st != capi.OK:
        raise RuntimeError(f"biodb_bgzf_compress: status {st}")
    return out[:n.value].tobytes()


class BgzfOutputStream:
    """bgzf/outputstream.d:50-223 over a file-like sink: bytes written are cut into BGZF blocks and compressed on the
    GPU at flush() / close(); close() appends the EOF block (:218-221)."""

    def __init__(self, sink, compression_level=-1):
        self._sink, self._level, self._buf, self._open = sink, compression_level, bytearray(), True

    def write(self, data):
        assert self._open, "stream is closed"
        self._buf += data
        return len(data)

    writeExact = write

    def flush(self):
        if self._buf:
            self._sink.write(bgzf_compress(self._buf, self._level, eof=False))
            self._buf = bytearray()

    def addEofBlock(self):
        self._sink.write(bytes(BamReader.BGZF_EOF))

    def close(self):
        if self._open:
            self.flush()
            self.addEofBlock()
            self._open = False


class BamWriter:
    """bam/writer.d:67-300: collects header and records the way BamWriter lays them out and compresses the BGZF blocks on
    the GPU — in slabs of `stream_blocks` blocks as they accumulate, the rest at finish().  `sink` is a file-like object
    or a path."""

    def __init__(self, sink, compression_level=-1, task_pool=None, device=-1, stream_blocks=2048):
        """stream_blocks: complete BGZF blocks are compressed and written to the sink whenever that many have accumulated
        (0: everything at finish()), so the writer holds ~stream_blocks x 64 KiB, not the file."""
        self._L = L = capi.lib()
        self._stream_blocks = int(stream_blocks)
        if not -1 <= compression_level <= 9:
            raise ValueError("compression level must be within -1 .. 9")
        self._sink = open(sink, "wb") if isinstance(sink, (str, os.PathLike)) else sink
        self._own = self._sink is not sink
        h = C.c_void_p()
        if L.biodb_writer_begin(device, compression_level, C.byref(h)) != capi.OK:
            raise RuntimeError("biodb_writer_begin failed")
        self._h = h

    def _check(self, st):
        if st == capi.ERR_CUDA:
            raise CudaUnavailable("BamWriter needs a CUDA device: biod_b200 has no CPU fallback")
        if st != capi.OK:
            raise Exception(self._L.biodb_writer_error(self._h).decode("latin-1") or f"biod_b200 status {st}")

    def writeSamHeader(self, header_text):
        self._text = header_text if isinstance(header_text, bytes) else header_text.encode("latin-1")

    def writeReferenceSequenceInfo(self, reference_sequences):
        """writer.d:151-181; reference_sequences: objects with .name and .length (or (name, length) pairs)."""
        refs = [(r.name, r.length) if hasattr(r, "name") else r for r in reference_sequences]
        names = (C.c_char_p * max(1, len(refs)))(*[n.encode("latin-1") for n, _ in refs])
        lens = (C.c_int32 * max(1, len(refs)))(*[int(l) for _, l in refs])
        text = getattr(self, "_text", b"")
        self._check(self._L.biodb_writer_header(self._h, text, len(text), len(refs), names, lens))

    def writeRecord(self, read, prefixed=False):
        """writer.d:244-268; `read` is a BamRead, or the raw bytes of one record BODY (what follows the block_size field) —
        the block_size prefix is always prepended here.  Bytes that already carry their prefix are passed with
        prefixed=True (or through writeRecords); nothing is guessed from the bytes themselves."""
        raw = read.raw.tobytes() if hasattr(read, "raw") else bytes(read)
        if prefixed and not hasattr(read, "raw"):
            if len(raw) < 4 or int.from_bytes(raw[:4], "little", signed=True) != len(raw) - 4:
                raise Exception("malformed record: the block_size prefix does not match the length of the bytes given")
        else:
            raw = len(raw).to_bytes(4, "little") + raw
        self.writeRecords(raw)

    def writeRecords(self, records):
        """Any number of records, back to back, each with its block_size prefix."""
        buf = np.frombuffer(bytes(records), dtype=np.uint8)
        self._check(self._L.biodb_writer_records(self._h, buf.ctypes.data if buf.size else None, buf.size))
        if self._stream_blocks > 0:
            self._drain(self._stream_blocks)

    def _drain(self, min_blocks):
        d, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.biodb_writer_drain(self._h, min_blocks, C.byref(d), C.byref(n)))
        if n.value:
            self._sink.write(C.string_at(d, n.value))

    def flush(self):
        self._check(self._L.biodb_writer_flush(self._h))

    def layout(self):
        """(uncompressed bytes, block starts) chosen so far — host only."""
        d, n, c, k = C.c_void_p(), C.c_size_t(), C.c_void_p(), C.c_size_t()
        self._check(self._L.biodb_writer_layout(self._h, C.byref(d), C.byref(n), C.byref(c), C.byref(k)))
        data = C.string_at(d, n.value) if n.value else b""
        cuts = list(np.frombuffer(C.string_at(c, 8 * k.value), dtype=np.uint64).astype(np.int64))
        return data, cuts

    def index(self):
        """The BAI index of the finished file (call between finish_keep() and close(); writer.d:139-195)."""
        d, n = C.c_void_p(), C.c_size_t()
        self._check(self._L.biodb_writer_index(self._h, C.byref(d), C.byref(n)))
        return C.string_at(d, n.value)

    def finish(self, want_index=None):
        """writer.d:276-280.  Like BamWriter, creates `<file>.bai` next to a coordinate-sorted `.bam` written to a path
        (writer.d:139-146,171-175); want_index=True returns the index bytes whatever the sink is."""
        d, n = C.c_void_p(), C.c_size_t()
        try:
            self._check(self._L.biodb_writer_finish(self._h, C.byref(d), C.byref(n)))
            self._sink.write(C.string_at(d, n.value))
            path = self._sink.name if self._own else None
            auto = bool(path) and str(path).endswith(".bam") and b"SO:coordinate" in getattr(self, "_text", b"") and \
                not getattr(self, "_no_index", False)
            # (unsorted records under an SO:coordinate header, or a wrong bin, make index() raise — the reference fails at
            #  writeRecord time; either way the file and the native handle are released below)
            bai = self.index() if (want_index or (auto and want_index is None)) else None
            if auto and bai is not None:
                with open(str(path) + ".bai", "wb") as f:
                    f.write(bai)
        finally:
            if self._own:
                self._sink.close()
            self.close()
        return bai

    def disableAutoIndexCreation(self):
        """writer.d:108-110"""
        self._no_index = True

    def close(self):
        if getattr(self, "_h", None):
            self._L.biodb_writer_end(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IndexBuilder:
    """bam/bai/indexing.d:56-351 over the C ABI (host only): put() batches of reads in file order, finish() returns the
    bytes of the .bai file."""

    def __init__(self, n_refs, check_bins=False):
        self._L = L = capi.lib()
        h = C.c_void_p()
        if L.biodb_index_builder_begin(n_refs, int(check_bins), C.byref(h)) != capi.OK:
            raise RuntimeError("biodb_index_builder_begin failed")
        self._h = h

    def put(self, batch):
        """batch: a RecordBatch of a reader opened with want_offsets (or any object with the same arrays)."""
        if batch.start_voffset is None:
            raise Exception("the index needs virtual offsets: open the reader with want_offsets=True")
        arrs = [np.ascontiguousarray(getattr(batch, f), dtype=dt) for f, dt in
                (("ref_id", np.int32), ("pos", np.int32), ("end_pos", np.int32), ("bin_mq_nl", np.uint32),
                 ("flag_nc", np.uint32), ("start_voffset", np.uint64), ("end_voffset", np.uint64))]
        st = self._L.biodb_index_builder_put(self._h, len(arrs[0]), *[a.ctypes.data for a in arrs])
        if st != capi.OK:
            raise Exception(self._L.biodb_index_builder_error(self._h).decode("latin-1"))

    def finish(self):
        d, n = C.c_void_p(), C.c_size_t()
        st = self._L.biodb_index_builder_finish(self._h, C.byref(d), C.byref(n))
        if st != capi.OK:
            raise Exception(self._L.biodb_index_builder_error(self._h).decode("latin-1"))
        return C.string_at(d, n.value)

    def close(self):
        if getattr(self, "_h", None):
            self._L.biodb_index_builder_end(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def createIndex(reader, check_bins=False):
    """bam/bai/indexing.d:356-366: the BAI index of a coordinate-sorted file, as bytes.  The reads and their virtual
    offsets come from a GPU pass over the file (the reader must have been opened with want_offsets=True)."""
    b = IndexBuilder(len(reader.reference_sequences), check_bins)
    try:
        for batch in reader.read_batches(copy=True):
            b.put(batch)
        return b.finish()
    finally:
        b.close()
