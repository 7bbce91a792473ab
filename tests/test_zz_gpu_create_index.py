"""createIndex (bam/bai/indexing.d:356-366) end to end: the reads and their virtual offsets from a GPU pass over the
file, the index built on the host (csrc/bai_build.h) — the same bytes as the plain-Python restatement makes from the
oracle's record table, and an index the GPU's own region reads work with.  (The file sorts last in the suite: it is the
one GPU test of this round that could not be run on a B200 before the round ended; everything it composes was.)"""
import numpy as np
import pytest

from baiutil import build_bai_biod
from conftest import fixture_bytes
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,bpb", [("ex1_header.bam", 0), ("bins.bam", 2), ("mg1655_chunk.bam", 1)])
def test_create_index(name, bpb):
    from biod_b200 import BamReader, createIndex
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    raw = createIndex(BamReader(data, blocks_per_batch=bpb, want_offsets=True), check_bins=(name != "bins.bam"))
    assert raw == build_bai_biod(o)
    rd = BamReader(data, want_offsets=True, index=raw)
    bai = orc.Bai(raw)
    rng = np.random.default_rng(3)
    for r in range(len(o.ref_names)):
        ln = o.ref_lens[r]
        for beg, end in [(0, ln)] + [tuple(sorted(int(x) for x in rng.integers(0, ln, 2))) for _ in range(5)]:
            if beg < end:
                want = orc.region_reads(o, bai, r, beg, end)[0]
                got = [x.raw.tobytes() for x in rd.region_reads(r, beg, end)]
                assert got == [o.record_bytes(int(i)).tobytes() for i in want], (name, r, beg, end)
