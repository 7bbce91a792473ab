"""BAI index on the host (row N2; csrc/bai.h through biodb_index_*): parse + getChunks against the oracle's
restatement (randomaccessmanager.d:222-244) on the reference's .bai files and on an index built for a synthetic BAM.
No GPU involved."""
import numpy as np
import pytest

from baiutil import build_bai
from conftest import fixture_bytes
from oracle import oracle as orc


def product_chunks(bai_bytes, ref, beg, end):
    from biod_b200 import BaiFile
    ix = BaiFile(bai_bytes)
    try:
        return ix.n_refs, ix.chunks(ref, beg, end)
    finally:
        ix.close()


@pytest.mark.parametrize("name", ["bins.bam.bai", "ex1_header.bam.bai", "tags.bam.bai"])
def test_chunks_match_oracle(name):
    from biod_b200 import BaiFile
    raw = fixture_bytes(name)
    o = orc.Bai(raw)
    ix = BaiFile(raw)
    assert ix.n_refs == o.n_refs
    rng = np.random.default_rng(3)
    regions = [(0, 1), (0, 2**31 - 1), (0, 2**32 - 1), (16383, 16385), (2**29 - 1, 2**29 + 5), (1400, 1500), (50_000, 50_100)]
    regions += [tuple(sorted(int(x) for x in rng.integers(0, 1 << int(rng.integers(4, 31)), 2))) for _ in range(300)]
    some = 0
    for r in range(o.n_refs):
        for beg, end in regions:
            if beg < end:
                want = o.chunks(r, beg, end)
                assert ix.chunks(r, beg, end) == want, (name, r, beg, end)
                some += len(want)
    assert some > 0
    with pytest.raises(Exception):
        ix.chunks(o.n_refs, 0, 10)


def test_index_of_a_synthetic_bam():
    from test_md_chain import random_pileup
    data = random_pileup(np.random.default_rng(9), 3000, refs=3, block_size=2500)
    b = orc.Bam(data).decode()
    raw = build_bai(b)
    o = orc.Bai(raw)
    rng = np.random.default_rng(4)
    for r in range(3):
        for _ in range(40):
            beg, end = sorted(int(x) for x in rng.integers(0, 30000, 2))
            if beg < end:
                assert product_chunks(raw, r, beg, end)[1] == o.chunks(r, beg, end)
                # the index is a valid one: the oracle's region read through it equals the naive filter
                got = [int(i) for i in orc.region_reads(b, o, r, beg, end)[0]]
                assert got == [i for i in range(b.n_records) if b.ref_id[i] == r and b.pos[i] < end and b.end_pos[i] > beg
                               and (b.pos[i] > beg or b.end_pos[i] > beg)]


def test_parse_errors():
    from biod_b200 import BaiFile, BamFormatException, ReadException
    good = fixture_bytes("bins.bam.bai")
    with pytest.raises(BamFormatException):
        BaiFile(b"BAM\1" + good[4:])
    with pytest.raises(ReadException):
        BaiFile(good[:len(good) // 2])
    with pytest.raises(ReadException):
        BaiFile(good[:6])
    # negative counts are a corrupt index, not an empty table (n_ref, then the first reference's n_bin)
    import struct
    for off in (4, 8):
        with pytest.raises(BamFormatException, match="negative"):
            BaiFile(good[:off] + struct.pack("<i", -3) + good[off + 4:])


def test_last_linear_offset():
    # reader.d:380-383 (what unmappedReads starts from): the last entry of the last non-empty linear index
    import ctypes as C
    import struct
    from biod_b200 import BaiFile, _capi
    L = _capi.lib()
    for name in ("bins.bam.bai", "ex1_header.bam.bai", "tags.bam.bai"):
        raw = fixture_bytes(name)
        # a plain walk of the file format (SAM spec 5.2)
        o, n_ref = 8, struct.unpack_from("<i", raw, 4)[0]
        last = []
        for _ in range(n_ref):
            n_bin = struct.unpack_from("<i", raw, o)[0]
            o += 4
            for _ in range(n_bin):
                n_chunk = struct.unpack_from("<i", raw, o + 4)[0]
                o += 8 + 16 * n_chunk
            n_intv = struct.unpack_from("<i", raw, o)[0]
            o += 4
            last.append(struct.unpack_from("<Q", raw, o + 8 * (n_intv - 1))[0] if n_intv else None)
            o += 8 * n_intv
        ix = BaiFile(raw)
        for n in range(n_ref + 2):
            want = next((v for v in reversed(last[:n]) if v is not None), None)
            out = C.c_uint64()
            got = L.biodb_index_last_linear_offset(ix._h, n, C.byref(out))
            assert (int(out.value) if got else None) == want, (name, n)
