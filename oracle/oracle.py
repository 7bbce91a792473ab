"""ctypes loader for the CPU oracle (TEST INFRASTRUCTURE ONLY — see oracle.cpp).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this module.  The product (biod_b200/) never does.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

OK, ERR_BGZF, ERR_ZLIB, ERR_FORMAT, ERR_TRUNC, ERR_CIGAR, ERR_UNSORTED = 0, -1, -2, -3, -4, -7, -8


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    src = os.path.join(_HERE, "oracle.cpp")
    if force or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"], stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        L = C.CDLL(build())
        u8p, u64 = C.POINTER(C.c_uint8), C.c_uint64
        L.orc_open.restype = C.c_void_p
        L.orc_open.argtypes = [C.c_void_p, u64]
        L.orc_close.argtypes = [C.c_void_p]
        for f in ("orc_open_status", "orc_status", "orc_zlib_errnum", "orc_n_refs", "orc_decode"):
            getattr(L, f).restype = C.c_int
            getattr(L, f).argtypes = [C.c_void_p]
        for f in ("orc_err_offset", "orc_reads_start_voffset", "orc_reads_start_uoffset", "orc_n_blocks",
                  "orc_n_records", "orc_n_cigar_total"):
            getattr(L, f).restype = u64
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_errmsg.restype = C.c_char_p
        L.orc_errmsg.argtypes = [C.c_void_p]
        L.orc_header_text.restype = C.c_void_p
        L.orc_header_text.argtypes = [C.c_void_p, C.POINTER(u64)]
        L.orc_ref_name.restype = C.c_char_p
        L.orc_ref_name.argtypes = [C.c_void_p, C.c_int]
        L.orc_ref_len.restype = C.c_int
        L.orc_ref_len.argtypes = [C.c_void_p, C.c_int]
        L.orc_block_field.restype = u64
        L.orc_block_field.argtypes = [C.c_void_p, u64, C.c_int]
        L.orc_udata.restype = C.c_void_p
        L.orc_udata.argtypes = [C.c_void_p, C.POINTER(u64)]
        for name in list(_REC_ARRAYS) + ["cigar_off", "cigar"]:
            fn = getattr(L, "orc_" + name)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p]
        L.orc_pileup_run.restype = C.c_void_p
        L.orc_pileup_run.argtypes = [C.c_void_p, C.c_int, u64, u64, C.c_int]
        L.orc_pileup_run_md.restype = C.c_void_p
        L.orc_pileup_run_md.argtypes = [C.c_void_p, C.c_int, u64, u64, C.c_int]
        L.orc_dna_of_read.restype = u64
        L.orc_dna_of_read.argtypes = [C.c_void_p, u64, C.c_char_p, u64]
        L.orc_pileup_run_list.restype = C.c_void_p
        L.orc_pileup_run_list.argtypes = [C.c_void_p, C.c_void_p, u64, C.c_int, u64, u64, C.c_int, C.c_int]
        L.orc_pileup_run_range.restype = C.c_void_p
        L.orc_pileup_run_range.argtypes = [C.c_void_p, C.c_int, u64, u64, C.c_int, C.c_int64, C.c_int64]
        L.orc_pileup_free.argtypes = [C.c_void_p]
        L.orc_pileup_status.restype = C.c_int
        L.orc_pileup_status.argtypes = [C.c_void_p]
        L.orc_pileup_errmsg.restype = C.c_char_p
        L.orc_pileup_errmsg.argtypes = [C.c_void_p]
        L.orc_pileup_ref_id.restype = C.c_int
        L.orc_pileup_ref_id.argtypes = [C.c_void_p]
        for f in ("orc_pileup_n_columns", "orc_pileup_n_entries"):
            getattr(L, f).restype = u64
            getattr(L, f).argtypes = [C.c_void_p]
        for name in list(_PILEUP_ARRAYS) + ["col_off", "ref_base"]:
            fn = getattr(L, "orc_pileup_" + name)
            fn.restype = C.c_void_p
            fn.argtypes = [C.c_void_p]
        L.orc_bai_open.restype = C.c_void_p
        L.orc_bai_open.argtypes = [C.c_void_p, u64]
        L.orc_bai_close.argtypes = [C.c_void_p]
        L.orc_bai_status.restype = C.c_int
        L.orc_bai_status.argtypes = [C.c_void_p]
        L.orc_bai_errmsg.restype = C.c_char_p
        L.orc_bai_errmsg.argtypes = [C.c_void_p]
        for f in ("orc_bai_n_bins", "orc_bai_n_intervals"):
            getattr(L, f).restype = u64
            getattr(L, f).argtypes = [C.c_void_p, u64]
        L.orc_bai_n_refs.restype = u64
        L.orc_bai_n_refs.argtypes = [C.c_void_p]
        L.orc_region_chunks.restype = C.c_int64
        L.orc_region_chunks.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, u64]
        L.orc_region_reads.restype = C.c_int64
        L.orc_region_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p,
                                       C.c_void_p, u64]
        L.orc_reads_between.restype = C.c_int64
        L.orc_reads_between.argtypes = [C.c_void_p, u64, u64, C.c_void_p, u64]
        L.orc_maq_new.restype = C.c_void_p
        L.orc_maq_new.argtypes = [C.c_double, C.c_double]
        L.orc_maq_free.argtypes = [C.c_void_p]
        for f in ("orc_maq_fk", "orc_maq_beta", "orc_maq_lhet"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_maq_compute.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, u64, C.c_void_p]
        L.orc_maq_run.restype = C.c_void_p
        L.orc_maq_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        L.orc_maq_result_free.argtypes = [C.c_void_p]
        for f in ("orc_maq_n_valid", "orc_maq_gt0", "orc_maq_gt1", "orc_maq_s0", "orc_maq_s1", "orc_maq_scores"):
            getattr(L, f).restype = C.c_void_p
            getattr(L, f).argtypes = [C.c_void_p]
        L.orc_cpu_baseline.restype = C.c_int
        L.orc_cpu_baseline.argtypes = [C.c_void_p, u64, C.c_int, C.c_int] + [C.POINTER(C.c_double)] * 3 + \
            [C.POINTER(u64)] * 4
        _LIB = L
    return _LIB


_REC_ARRAYS = {
    "rec_off": np.uint64, "block_size": np.int32, "ref_id": np.int32, "pos": np.int32, "end_pos": np.int32,
    "l_seq": np.int32, "next_ref": np.int32, "next_pos": np.int32, "tlen": np.int32, "bin": np.uint16,
    "flag": np.uint16, "n_cigar": np.uint16, "mapq": np.uint8, "l_read_name": np.uint8,
    "start_vo": np.uint64, "end_vo": np.uint64,
}
_PILEUP_ARRAYS = {
    "col_ref": np.int32, "col_pos": np.uint64, "n_start": np.uint32, "read_idx": np.uint32, "qoff": np.uint32,
    "op_index": np.uint32, "op_offset": np.uint32, "base": np.uint8, "qual": np.uint8,
}


def _arr(ptr, n, dtype):
    if n == 0 or not ptr:
        return np.zeros(0, dtype=dtype)
    nbytes = int(n) * np.dtype(dtype).itemsize
    buf = (C.c_uint8 * nbytes).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


class OracleError(Exception):
    def __init__(self, status, msg, zerr=0, offset=0):
        super().__init__(f"oracle status {status}: {msg}")
        self.status, self.msg, self.zerr, self.offset = status, msg, zerr, offset


class Pileup:
    """Result of makePileup / pileupColumns restated on the CPU."""

    def __init__(self, L, h, keep=False):
        self._L, self._h = L, (h if keep else None)
        self.status = L.orc_pileup_status(h)
        self.msg = L.orc_pileup_errmsg(h).decode()
        self.ref_id = L.orc_pileup_ref_id(h)
        nc, ne = L.orc_pileup_n_columns(h), L.orc_pileup_n_entries(h)
        self.n_columns, self.n_entries = int(nc), int(ne)
        for name, dt in _PILEUP_ARRAYS.items():
            n = nc if name in ("col_ref", "col_pos", "n_start") else ne
            setattr(self, name, _arr(getattr(L, "orc_pileup_" + name)(h), n, dt))
        self.col_off = _arr(L.orc_pileup_col_off(h), nc + 1, np.uint64)
        self.ref_base = _arr(L.orc_pileup_ref_base(h), nc, np.uint8)     # PileupColumn.reference_base ('N' unless use_md_tag)
        if not keep:
            L.orc_pileup_free(h)

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_pileup_free(self._h)
            self._h = None

    def bases(self, c):
        a, b = int(self.col_off[c]), int(self.col_off[c + 1])
        return self.base[a:b].tobytes().decode()


class Bam:
    """Restated BamReader: open parses the header; decode() walks every record."""

    def __init__(self, data: bytes):
        self._L = L = lib()
        self._data = np.frombuffer(data, dtype=np.uint8)  # keep alive (oracle does not copy)
        self._h = L.orc_open(self._data.ctypes.data, len(data))
        st = L.orc_open_status(self._h)
        if st:
            e = self._error()
            self.close()
            raise e
        n = C.c_uint64()
        p = L.orc_header_text(self._h, C.byref(n))
        self.header_text = C.string_at(p, n.value).decode("latin-1") if n.value else ""
        self.ref_names = [L.orc_ref_name(self._h, i).decode() for i in range(L.orc_n_refs(self._h))]
        self.ref_lens = [L.orc_ref_len(self._h, i) for i in range(L.orc_n_refs(self._h))]
        self.reads_start_voffset = int(L.orc_reads_start_voffset(self._h))
        self.reads_start_uoffset = int(L.orc_reads_start_uoffset(self._h))
        self._decoded = False

    def _error(self):
        L = self._L
        return OracleError(L.orc_status(self._h), L.orc_errmsg(self._h).decode(), L.orc_zlib_errnum(self._h),
                           int(L.orc_err_offset(self._h)))

    def close(self):
        if self._h:
            self._L.orc_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def decode(self, raise_on_error=True):
        L = self._L
        st = L.orc_decode(self._h)
        self.status = st
        self.error = self._error() if st else None
        n = int(L.orc_n_records(self._h))
        self.n_records = n
        for name, dt in _REC_ARRAYS.items():
            setattr(self, name, _arr(getattr(L, "orc_" + name)(self._h), n, dt))
        self.cigar_off = _arr(L.orc_cigar_off(self._h), n + 1, np.uint64)
        self.cigar = _arr(L.orc_cigar(self._h), int(L.orc_n_cigar_total(self._h)), np.uint32)
        ln = C.c_uint64()
        p = L.orc_udata(self._h, C.byref(ln))
        self.udata = _arr(p, ln.value, np.uint8)
        nb = int(L.orc_n_blocks(self._h))
        self.n_blocks = nb
        self.blocks = np.array([[L.orc_block_field(self._h, i, f) for f in range(7)] for i in range(nb)],
                               dtype=np.uint64).reshape(nb, 7)
        self._decoded = True
        if st and raise_on_error:
            raise self.error
        return self

    # -- record views (restating bam/read.d accessors) --------------------
    def record_bytes(self, i):
        o = int(self.rec_off[i]) + 4
        return self.udata[o:o + int(self.block_size[i])]

    def name(self, i):
        r = self.record_bytes(i)
        return r[32:32 + int(self.l_read_name[i]) - 1].tobytes().decode("latin-1")

    def cigar_ops(self, i):
        a, b = int(self.cigar_off[i]), int(self.cigar_off[i + 1])
        return [(int(x) >> 4, "MIDNSHP=X???????"[int(x) & 15]) for x in self.cigar[a:b]]

    def cigar_string(self, i):
        ops = self.cigar_ops(i)
        return "".join(f"{l}{o}" for l, o in ops) if ops else "*"

    def sequence(self, i):
        r = self.record_bytes(i)
        o = 32 + int(self.l_read_name[i]) + 4 * int(self.n_cigar[i])
        n = int(self.l_seq[i])
        packed = r[o:o + (n + 1) // 2]
        codes = np.empty(2 * len(packed), dtype=np.uint8)
        codes[0::2] = packed >> 4
        codes[1::2] = packed & 15
        return np.frombuffer(b"=ACMGRSVTWYHKDBN", dtype=np.uint8)[codes[:n]].tobytes().decode()

    def qualities(self, i):
        r = self.record_bytes(i)
        n = int(self.l_seq[i])
        o = 32 + int(self.l_read_name[i]) + 4 * int(self.n_cigar[i]) + (n + 1) // 2
        return r[o:o + n]

    def tags_raw(self, i):
        r = self.record_bytes(i)
        n = int(self.l_seq[i])
        o = 32 + int(self.l_read_name[i]) + 4 * int(self.n_cigar[i]) + (n + 1) // 2 + n
        return r[o:].tobytes()

    # -- pileup ------------------------------------------------------------
    def make_pileup(self, start_from=0, end_at=2**64 - 1, skip_zero_coverage=True, use_md_tag=False, keep=False):
        run = self._L.orc_pileup_run_md if use_md_tag else self._L.orc_pileup_run
        return Pileup(self._L, run(self._h, 1, start_from, end_at, int(skip_zero_coverage)), keep)

    def pileup_columns(self, skip_zero_coverage=True, use_md_tag=False, keep=False):
        run = self._L.orc_pileup_run_md if use_md_tag else self._L.orc_pileup_run
        return Pileup(self._L, run(self._h, 0, 0, 2**64 - 1, int(skip_zero_coverage)), keep)

    def maq(self, pileup, tables, min_base_quality=13):
        """MaqSnpCaller.genotypeLikelihoodInfo over every column of `pileup` (made with keep=True); see Maq."""
        assert pileup._h, "make the pileup with keep=True"
        L = self._L
        r = L.orc_maq_run(self._h, pileup._h, tables._h, int(min_base_quality))
        nc = pileup.n_columns
        out = dict(n_valid=_arr(L.orc_maq_n_valid(r), nc, np.uint32), gt0=_arr(L.orc_maq_gt0(r), nc, np.uint8),
                   gt1=_arr(L.orc_maq_gt1(r), nc, np.uint8), s0=_arr(L.orc_maq_s0(r), nc, np.float32),
                   s1=_arr(L.orc_maq_s1(r), nc, np.float32), scores=_arr(L.orc_maq_scores(r), nc * 25, np.float32).reshape(nc, 25))
        L.orc_maq_result_free(r)
        return out

    def make_pileup_of(self, indices, start_from=0, end_at=2**64 - 1, skip_zero_coverage=True, use_md_tag=False,
                       single_ref=True):
        """makePileup over the reads with the given record indices (any range of reads: examples/read_bam_file.d:22-25)."""
        idx = np.ascontiguousarray(indices, dtype=np.int64)
        return Pileup(self._L, self._L.orc_pileup_run_list(self._h, idx.ctypes.data, len(idx), int(single_ref), start_from,
                                                           end_at, int(skip_zero_coverage), int(use_md_tag)))

    def dna(self, i):
        """dna(read) of record i: reference bases over its aligned and deleted positions (md/reconstruct.d:38-214)."""
        buf = C.create_string_buffer(1 << 16)
        n = int(self._L.orc_dna_of_read(self._h, i, buf, len(buf)))
        return buf.raw[:min(n, len(buf))].decode("latin1")

    def make_pileup_range(self, rec_begin, rec_end, start_from=0, end_at=2**64 - 1, skip_zero_coverage=True,
                          single_ref=True):
        return Pileup(self._L, self._L.orc_pileup_run_range(self._h, int(single_ref), start_from, end_at,
                                                            int(skip_zero_coverage), rec_begin, rec_end))


class Maq:
    """ErrorModelCoefficients of the MAQ model (bio/std/hts/snpcallers/maq.d:66-132) and computeLikelihoods (:138-248)."""

    def __init__(self, depcorr=0.17, eta=0.03):
        # MaqSnpCaller keeps depcorr / eta as floats and hands them to ErrorModel(float, float) (maq.d:329-338,346-347)
        self._L = lib()
        self._h = self._L.orc_maq_new(float(np.float32(depcorr)), float(np.float32(eta)))
        self.fk = _arr(self._L.orc_maq_fk(self._h), 256, np.float64)
        self.beta = _arr(self._L.orc_maq_beta(self._h), 256 * 256 * 64, np.float64)
        self.lhet = _arr(self._L.orc_maq_lhet(self._h), 256 * 256, np.float64)

    def compute(self, bases, quals, reverse):
        """computeLikelihoods over explicit read bases: 25 scores by DiploidGenotype!Base5 code (FLT_MIN = absent)."""
        b = np.frombuffer(bases if isinstance(bases, bytes) else bytes(bases), dtype=np.uint8).copy()
        q = np.ascontiguousarray(quals, dtype=np.uint8)
        r = np.ascontiguousarray(reverse, dtype=np.uint8)
        out = np.zeros(25, dtype=np.float32)
        self._L.orc_maq_compute(self._h, b.ctypes.data, q.ctypes.data, r.ctypes.data, len(b), out.ctypes.data)
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_maq_free(self._h)
            self._h = None


class Bai:
    """BaiFile (bam/baifile.d:85-169) + RandomAccessManager.getChunks (randomaccessmanager.d:222-244)."""

    def __init__(self, data: bytes):
        self._L = lib()
        self._buf = np.frombuffer(bytes(data), dtype=np.uint8)
        self._h = self._L.orc_bai_open(self._buf.ctypes.data, len(self._buf))
        st = self._L.orc_bai_status(self._h)
        if st:
            raise OracleError(st, self._L.orc_bai_errmsg(self._h).decode())
        self.n_refs = int(self._L.orc_bai_n_refs(self._h))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.orc_bai_close(self._h)
            self._h = None

    def chunks(self, ref_id, beg, end):
        out = np.zeros(2 * 4096, dtype=np.uint64)
        n = int(self._L.orc_region_chunks(self._h, ref_id, beg, end, out.ctypes.data, len(out) // 2))
        if n < 0:
            raise OracleError(-9, "Invalid reference sequence index")
        assert n <= len(out) // 2
        return [(int(out[2 * k]), int(out[2 * k + 1])) for k in range(n)]


def region_reads(bam, bai, ref_id, beg, end):
    """bam[ref][beg .. end) (reference.d:76-81): (record indices, start voffsets, end voffsets)."""
    cap = max(16, bam.n_records + 16)
    idx = np.zeros(cap, dtype=np.int64)
    sv = np.zeros(cap, dtype=np.uint64)
    ev = np.zeros(cap, dtype=np.uint64)
    n = int(bam._L.orc_region_reads(bam._h, bai._h, ref_id, beg, end, idx.ctypes.data, sv.ctypes.data, ev.ctypes.data, cap))
    if n < 0:
        raise OracleError(n, "region read failed")
    return idx[:n].copy(), sv[:n].copy(), ev[:n].copy()


def group_chunks(bai, ref_id, regions):
    """getGroupChunks (randomaccessmanager.d:246-296) for sorted, non-overlapping regions [(start, end), ...] of one
    reference: [(beg voffset, end voffset), ...]."""
    L = bai._L if hasattr(bai, "_L") else lib()
    begs = np.array([a for a, _ in regions], dtype=np.uint32)
    ends = np.array([b for _, b in regions], dtype=np.uint32)
    L.orc_group_chunks.restype = C.c_int64
    L.orc_group_chunks.argtypes = [C.c_void_p, C.c_uint32, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    cap = 1 << 16
    out = np.zeros(2 * cap, dtype=np.uint64)
    n = int(L.orc_group_chunks(bai._h, ref_id, len(begs), begs.ctypes.data, ends.ctypes.data, out.ctypes.data, cap))
    if n < 0:
        raise OracleError(n, "group chunks failed")
    return [(int(out[2 * k]), int(out[2 * k + 1])) for k in range(n)]


def regions_reads(bam, bai, regions):
    """getReads(BamRegion[]) (randomaccessmanager.d:316-337): regions = [(ref_id, start, end), ...] in any order.
    Returns (record indices, start voffsets, end voffsets)."""
    cap = max(16, 2 * bam.n_records + 16)
    idx = np.zeros(cap, dtype=np.int64)
    sv = np.zeros(cap, dtype=np.uint64)
    ev = np.zeros(cap, dtype=np.uint64)
    rg = np.ascontiguousarray(np.array(regions, dtype=np.uint32).reshape(-1, 3))
    bam._L.orc_regions_reads.restype = C.c_int64
    bam._L.orc_regions_reads.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
    n = int(bam._L.orc_regions_reads(bam._h, bai._h, rg.ctypes.data, len(rg), idx.ctypes.data, sv.ctypes.data, ev.ctypes.data, cap))
    if n < 0:
        raise OracleError(n, "multi-region read failed")
    return idx[:n].copy(), sv[:n].copy(), ev[:n].copy()


def reads_between(bam, from_vo, to_vo):
    """getReadsBetween (randomaccessmanager.d:186-196): record indices."""
    cap = max(16, bam.n_records + 16)
    idx = np.zeros(cap, dtype=np.int64)
    n = int(bam._L.orc_reads_between(bam._h, from_vo, to_vo, idx.ctypes.data, cap))
    if n < 0:
        raise OracleError(n, "reads between failed")
    return idx[:n].copy()


def cpu_baseline(data, threads, do_pileup=True):
    """Time the restated CPU path on `data` (bytes-like BAM).  Returns a dict."""
    L = lib()
    a = np.frombuffer(data, dtype=np.uint8)
    ti, td, tp = C.c_double(), C.c_double(), C.c_double()
    nr, nc, ne, cs = C.c_uint64(), C.c_uint64(), C.c_uint64(), C.c_uint64()
    st = L.orc_cpu_baseline(a.ctypes.data, len(a), threads, int(do_pileup), C.byref(ti), C.byref(td), C.byref(tp),
                            C.byref(nr), C.byref(nc), C.byref(ne), C.byref(cs))
    if st:
        raise OracleError(st, "cpu baseline failed")
    L.orc_cpu_baseline_lazy_seconds.restype = C.c_double
    return dict(t_inflate=ti.value, t_decode=td.value, t_pileup=tp.value, t_pileup_lazy=L.orc_cpu_baseline_lazy_seconds(),
                n_records=nr.value, n_columns=nc.value, n_entries=ne.value, checksum=cs.value)
