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


def test_writer_creates_the_index(tmp_path):
    # writer.d:139-146,171-175: a coordinate-sorted file written to a path ending in .bam gets its .bai beside it
    import struct
    from biod_b200 import BamReader, BamWriter
    o = orc.Bam(fixture_bytes("ex1_header.bam")).decode()
    assert "SO:coordinate" in o.header_text
    path = tmp_path / "out.bam"
    w = BamWriter(str(path))
    w.writeSamHeader(o.header_text)
    w.writeReferenceSequenceInfo(list(zip(o.ref_names, o.ref_lens)))
    w.writeRecords(b"".join(struct.pack("<i", int(o.block_size[i])) + o.record_bytes(i).tobytes() for i in range(o.n_records)))
    bai = w.finish()
    written = path.read_bytes()
    o2 = orc.Bam(written).decode()
    assert o2.n_records == o.n_records
    assert bai == build_bai_biod(o2, check_bins=True) == (tmp_path / "out.bam.bai").read_bytes()
    # and the reader finds it next to the file (baifile.d:95-113)
    rd = BamReader(str(path))
    chr2 = o.ref_names.index("chr2")
    got = [x.raw.tobytes() for x in rd["chr2"][150:160]]
    want = [i for i in range(o2.n_records) if o2.ref_id[i] == chr2 and o2.pos[i] < 160 and o2.end_pos[i] > 150]
    assert got == [o2.record_bytes(i).tobytes() for i in want]
