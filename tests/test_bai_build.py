"""IndexBuilder on the host (row N4, last part; csrc/bai_build.h through biodb_index_builder_*) against a plain-Python
restatement of bam/bai/indexing.d:56-351 (tests/baiutil.build_bai_biod): the same bytes; and the index it builds is a
working one — the oracle's region reads through it equal the naive filter.  No reference test pins createIndex (and
the .bai files in the reference's test data come from another tool: they differ from what indexing.d produces, e.g.
in the linear-index entry of a placed unmapped read), so the parity of this function is "restated, unpinned".
No GPU involved: the reads and their virtual offsets come from the oracle here."""
import numpy as np
import pytest

from baiutil import build_bai_biod
from conftest import fixture_bytes
from oracle import oracle as orc


class OracleBatch:
    """The arrays of a RecordBatch, taken from the oracle's record table."""

    def __init__(self, o, a, z):
        self.ref_id, self.pos, self.end_pos = o.ref_id[a:z], o.pos[a:z], o.end_pos[a:z]
        self.bin_mq_nl = (o.bin[a:z].astype(np.uint32) << 16) | (o.mapq[a:z].astype(np.uint32) << 8) | o.l_read_name[a:z]
        self.flag_nc = (o.flag[a:z].astype(np.uint32) << 16) | o.n_cigar[a:z]
        self.start_voffset, self.end_voffset = o.start_vo[a:z], o.end_vo[a:z]


def product_index(o, step=1000, check_bins=False):
    from biod_b200 import IndexBuilder
    b = IndexBuilder(len(o.ref_names), check_bins)
    for a in range(0, o.n_records, step):
        b.put(OracleBatch(o, a, min(o.n_records, a + step)))
    raw = b.finish()
    b.close()
    return raw


def naive(o, r, beg, end):
    return [i for i in range(o.n_records) if o.ref_id[i] == r and o.pos[i] < end and o.end_pos[i] > beg]


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "tags.bam", "mg1655_chunk.bam", "illu_20_chunk.bam"])
def test_built_index_equals_the_restatement_and_works(name):
    o = orc.Bam(fixture_bytes(name)).decode()
    raw = product_index(o, step=777)
    assert raw == build_bai_biod(o)
    assert raw == product_index(o, step=1)                       # batch boundaries do not matter
    bai = orc.Bai(raw)
    assert bai.n_refs == len(o.ref_names)
    rng = np.random.default_rng(13)
    for r in range(len(o.ref_names)):
        ln = o.ref_lens[r]
        for beg, end in [(0, ln), (0, 1)] + [tuple(sorted(int(x) for x in rng.integers(0, ln, 2))) for _ in range(15)]:
            if beg < end:
                got = [int(i) for i in orc.region_reads(o, bai, r, beg, end)[0]]
                want = [i for i in naive(o, r, beg, end) if o.pos[i] > beg or o.end_pos[i] > beg]
                assert got == want, (name, r, beg, end)


def test_synthetic_multi_reference_index():
    from test_md_chain import random_pileup
    from bamutil import bam_record, make_bam
    data = random_pileup(np.random.default_rng(19), 3000, refs=3, block_size=2000)
    o = orc.Bam(data).decode()
    assert product_index(o, check_bins=True) == build_bai_biod(o, check_bins=True)
    # references without reads in front of, between and behind the ones with reads; reads without a reference at the end
    recs = [bam_record("a", "ACGT", "4M", 10, ref_id=1), bam_record("b", "ACGT", "4M", 20000, ref_id=1),
            bam_record("c", "ACGT", "4M", 5, ref_id=3), bam_record("u", "ACGT", "", -1, ref_id=-1, flag=4)]
    o = orc.Bam(make_bam([(f"c{i}", 100000) for i in range(5)], recs)).decode()
    raw = product_index(o)
    assert raw == build_bai_biod(o)
    bai = orc.Bai(raw)
    assert [int(i) for i in orc.region_reads(o, bai, 1, 0, 100000)[0]] == [0, 1]
    assert [int(i) for i in orc.region_reads(o, bai, 3, 0, 100000)[0]] == [2]
    assert len(orc.region_reads(o, bai, 0, 0, 100000)[0]) == 0 and len(orc.region_reads(o, bai, 4, 0, 100000)[0]) == 0
    assert raw[-8:] == (1).to_bytes(8, "little")                 # n_no_coor
    # an empty file
    o = orc.Bam(make_bam([("c0", 1000)], [])).decode()
    assert product_index(o) == build_bai_biod(o) == b"BAI\1" + (1).to_bytes(4, "little") + bytes(8) + bytes(8)


def test_builder_errors():
    from bamutil import bam_record, make_bam
    from biod_b200 import IndexBuilder
    recs = [bam_record("a", "ACGT", "4M", 100), bam_record("b", "ACGT", "4M", 50)]
    o = orc.Bam(make_bam([("c0", 1000)], recs)).decode()
    with pytest.raises(Exception, match="not coordinate-sorted"):
        product_index(o)
    # a wrong bin is only an error when asked to check
    good = bam_record("a", "ACGT", "4M", 100)
    bad = bytearray(bam_record("b", "ACGT", "4M", 20000))
    bad[4 + 10:4 + 12] = (4681).to_bytes(2, "little")
    o = orc.Bam(make_bam([("c0", 100000)], [good, bytes(bad)])).decode()
    product_index(o)
    with pytest.raises(Exception, match="is set incorrectly"):
        product_index(o, check_bins=True)
    with pytest.raises(Exception, match="want_offsets"):
        class NoOffsets:
            start_voffset = None
        IndexBuilder(1).put(NoOffsets())


def _parse_bai(raw):
    import struct
    assert raw[:4] == b"BAI\1"
    n, p, refs = struct.unpack_from("<i", raw, 4)[0], 8, []
    for _ in range(n):
        nb = struct.unpack_from("<i", raw, p)[0]
        p += 4
        bins = {}
        for _ in range(nb):
            bid, nc = struct.unpack_from("<Ii", raw, p)
            p += 8
            bins[bid] = [struct.unpack_from("<QQ", raw, p + 16 * k) for k in range(nc)]
            p += 16 * nc
        nl = struct.unpack_from("<i", raw, p)[0]
        p += 4 + 8 * nl
        refs.append(bins)
    return refs


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "tags.bam"])
def test_bins_and_chunks_agree_with_the_fixture_indexes_of_another_tool(name):
    """An outside check (not a pin of indexing.d, which no reference test exercises): the reference's test data carries
    .bai files written by samtools.  Bin by bin, the chunk lists the product's builder writes are the ones samtools
    wrote, once three habits of that tool are set aside — it tells the END of the file (past the EOF block) as the end
    offset of the file's last record where BioD's stream tells the EOF block's start; one of the files has empty chunks
    (begin == end); older files lack the pseudo-bin 37450.  (The linear indexes differ in the entry of the first window,
    as the module docstring says, and are not compared.)"""
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    mine = _parse_bai(product_index(o, step=500))
    theirs = _parse_bai(fixture_bytes(name + ".bai"))
    eof_vo = (len(data) - 28) << 16
    assert len(mine) == len(theirs)
    for a, b in zip(mine, theirs):
        a = {k: v for k, v in a.items() if k != 37450}
        b = {k: [(x, min(y, eof_vo)) for x, y in v if x != y] for k, v in b.items() if k != 37450}
        assert a == {k: v for k, v in b.items() if v}
