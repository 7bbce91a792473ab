"""GPU side of BamWriter (row N4; biodb_writer_finish: the blocks the writer laid out, compressed on the device): the
file must read back — through zlib block by block, through the oracle, and through this library's own reader — as the
header and the reads that were written (test/unittests.d:286-305), with exactly the block layout the host half chose."""
import io
import struct

import numpy as np
import pytest

from conftest import fixture_bytes
from oracle import oracle as orc
from test_gpu_x_deflate import bgzf_read

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,level", [("ex1_header.bam", -1), ("tags.bam", 0), ("bins.bam", 1)])
def test_written_bam_reads_back(name, level):
    from biod_b200 import BamReader, BamWriter
    o = orc.Bam(fixture_bytes(name)).decode()
    sink = io.BytesIO()
    w = BamWriter(sink, compression_level=level)
    w.writeSamHeader(o.header_text)
    w.writeReferenceSequenceInfo(list(zip(o.ref_names, o.ref_lens)))
    w.writeRecords(b"".join(struct.pack("<i", int(o.block_size[i])) + o.record_bytes(i).tobytes() for i in range(o.n_records)))
    data, cuts = w.layout()
    w.finish()
    stream = sink.getvalue()
    back, sizes = bgzf_read(stream)
    assert back == data
    cuts = cuts + ([len(data)] if cuts[-1] != len(data) else [])
    assert sizes == [b - a for a, b in zip(cuts[:-1], cuts[1:])] + [0]          # the layout's blocks, then the EOF block
    o2 = orc.Bam(stream).decode()
    assert o2.n_records == o.n_records and o2.header_text == o.header_text and o2.ref_names == o.ref_names
    rd = BamReader(stream)
    assert rd.header_text == o.header_text and [r.name for r in rd.reference_sequences] == o.ref_names
    raws = []
    for b in rd.read_batches(copy=True):
        for i in range(b.n):
            p = int(b.rec_off[i]) + 4
            raws.append(b.data[p:p + int(b.block_size[i])].tobytes())
    assert raws == [o2.record_bytes(i).tobytes() for i in range(o2.n_records)]
    for i in range(0, o.n_records, 97):                                          # unchanged but for the recalculated bin
        a, c = o.record_bytes(i).tobytes(), raws[i]
        assert a[:10] == c[:10] and a[12:] == c[12:]


def test_streaming_writer_writes_the_same_file_and_index():
    """biodb_writer_drain: blocks compressed and handed to the sink as they accumulate (stream_blocks = 3 here) give the
    same file and the same .bai as compressing everything at finish() — every block is compressed on its own."""
    from biod_b200 import BamWriter
    from tools import bamgen
    data = bamgen.generate(20000, 2, True, -1, bamgen.SEED_BASE + 5).tobytes()
    o = orc.Bam(data).decode()
    recs = [struct.pack("<i", int(o.block_size[i])) + o.record_bytes(i).tobytes() for i in range(o.n_records)]
    files, indexes, writes = [], [], []
    for stream_blocks in (0, 3):
        class Sink(io.BytesIO):
            n_writes = 0

            def write(self, b):
                Sink.n_writes += 1
                return super().write(b)
        sink = Sink()
        w = BamWriter(sink, stream_blocks=stream_blocks)
        w.writeSamHeader(o.header_text)
        w.writeReferenceSequenceInfo(list(zip(o.ref_names, o.ref_lens)))
        for a in range(0, len(recs), 500):
            w.writeRecords(b"".join(recs[a:a + 500]))
        indexes.append(w.finish(want_index=True))
        files.append(sink.getvalue())
        writes.append(Sink.n_writes)
    assert files[0] == files[1] and indexes[0] == indexes[1] and len(indexes[0]) > 100
    assert writes[0] == 1 and writes[1] > 5
    o2 = orc.Bam(files[1]).decode()
    assert o2.n_records == o.n_records
    assert [o2.record_bytes(i).tobytes()[12:] for i in range(0, o.n_records, 53)] == \
        [o.record_bytes(i).tobytes()[12:] for i in range(0, o.n_records, 53)]
