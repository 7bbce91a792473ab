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


def merged_groups(regions):
    """regions [(ref, start, end), ...] -> {ref: [(start, end), ...]} sorted, overlapping ones joined (algo.d:95-162)."""
    by = {}
    for r, a, b in sorted(regions):
        g = by.setdefault(r, [])
        if g and g[-1][1] >= a:
            g[-1][1] = max(g[-1][1], b)
        else:
            g.append([a, b])
    return {r: [tuple(x) for x in g] for r, g in by.items()}


def random_regions(rng, o, n_max=6, span=None):
    out = []
    for _ in range(int(rng.integers(1, n_max + 1))):
        r = int(rng.integers(0, len(o.ref_names)))
        a = int(rng.integers(0, max(1, o.ref_lens[r])))
        b = a + 1 + int(rng.integers(0, span or max(1, o.ref_lens[r])))
        out.append((r, a, b))
    return out


@pytest.mark.parametrize("name", ["bins.bam", "ex1_header.bam", "tags.bam"])
def test_multi_region_reads_of_the_oracle_and_group_chunks(name):
    """getReads(BamRegion[]) (randomaccessmanager.d:246-296,316-337; no reference test exercises it): the oracle's
    restatement returns, per reference, the sorted union of what the single-region reads of the merged regions return —
    every read once — and the product's chunk lists (biodb_index_regions_chunks) are the oracle's."""
    import ctypes as C
    from biod_b200 import BaiFile
    o = orc.Bam(fixture_bytes(name)).decode()
    raw = fixture_bytes(name + ".bai")
    bai = orc.Bai(raw)
    ix = BaiFile(raw)
    rng = np.random.default_rng(41)
    for trial in range(60):
        regions = random_regions(rng, o, span=3000 if trial % 2 else None)
        got = [int(i) for i in orc.regions_reads(o, bai, regions)[0]]
        want = []
        for r, group in sorted(merged_groups(regions).items()):
            one = set()
            for a, b in group:
                one |= set(int(i) for i in orc.region_reads(o, bai, r, a, b)[0])
            want += sorted(one)
            # the product's chunks for this group
            begs = np.array([a for a, _ in group], dtype=np.uint32)
            ends = np.array([b for _, b in group], dtype=np.uint32)
            n = int(ix._L.biodb_index_regions_chunks(ix._h, r, len(begs), begs.ctypes.data, ends.ctypes.data, None, 0))
            out = np.zeros(2 * max(n, 1), dtype=np.uint64)
            ix._L.biodb_index_regions_chunks(ix._h, r, len(begs), begs.ctypes.data, ends.ctypes.data, out.ctypes.data, n)
            assert [(int(out[2 * k]), int(out[2 * k + 1])) for k in range(n)] == orc.group_chunks(bai, r, group), (name, regions)
        assert got == want, (name, regions)
    ix.close()
