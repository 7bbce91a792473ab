"""BamWriter's host half (row N4; csrc/deflate.cu: biodb_writer_*): the uncompressed stream and the BGZF block
boundaries it chooses — magic, header, reference table, a boundary, then records where one that would not fit starts a
new block (bam/writer.d:139-181,244-268; bgzf/outputstream.d:107-161) — against the layout of the test suite's own
writer (tests/bamutil.make_bam, which follows the same lines) and read back through the oracle.  No GPU: the layout
is inspected before compression (biodb_writer_layout)."""
import struct
import zlib

import numpy as np
import pytest

from bamutil import bam_record, bgzf_block, make_bam, BGZF_EOF
from conftest import fixture_bytes
from oracle import oracle as orc


def blocks_of(stream):
    """Uncompressed payload of every BGZF block of a stream."""
    out, p = [], 0
    while p < len(stream):
        bsize = struct.unpack_from("<H", stream, p + 16)[0] + 1
        out.append(zlib.decompress(stream[p + 18:p + bsize - 8], -15))
        p += bsize
    return out


def layout_stream(w):
    """The writer's layout as a BGZF file (blocks compressed by zlib here: the device is not involved)."""
    data, cuts = w.layout()
    cuts = cuts + ([len(data)] if cuts[-1] != len(data) else [])
    parts = [data[a:b] for a, b in zip(cuts[:-1], cuts[1:])]
    return parts, b"".join(bgzf_block(p) for p in parts) + BGZF_EOF


def write_like(o, **kw):
    from biod_b200 import BamWriter
    import io
    w = BamWriter(io.BytesIO(), **kw)
    w.writeSamHeader(o.header_text)
    w.writeReferenceSequenceInfo(list(zip(o.ref_names, o.ref_lens)))
    return w


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "tags.bam", "mg1655_chunk.bam"])
def test_rewritten_fixture_reads_back(name):
    o = orc.Bam(fixture_bytes(name)).decode()
    w = write_like(o)
    recs = [struct.pack("<i", int(o.block_size[i])) + o.record_bytes(i).tobytes() for i in range(o.n_records)]
    w.writeRecords(b"".join(recs[:o.n_records // 2]))
    for r in recs[o.n_records // 2:]:
        w.writeRecord(r, prefixed=True)
    parts, stream = layout_stream(w)
    o2 = orc.Bam(stream).decode()
    assert o2.header_text == o.header_text and o2.ref_names == o.ref_names and o2.ref_lens == o.ref_lens
    assert o2.n_records == o.n_records
    # records come back unchanged, except for the bin, which BamWriter recalculates (read.d:1028-1030)
    for f in ("ref_id", "pos", "end_pos", "l_seq", "flag", "n_cigar", "mapq", "l_read_name", "next_ref", "next_pos", "tlen"):
        assert np.array_equal(getattr(o2, f), getattr(o, f)), f
    from bamutil import reg2bin
    want_bin = [reg2bin(int(p), int(e)) if p >= 0 else 4680 for p, e in zip(o.pos, o.end_pos)]
    mapped = o.pos >= 0
    assert np.array_equal(o2.bin[mapped], np.array(want_bin, dtype=np.uint16)[mapped])
    for i in range(o.n_records):
        a, b = o.record_bytes(i).tobytes(), o2.record_bytes(i).tobytes()
        assert a[:10] == b[:10] and a[12:] == b[12:], i
    # the same block layout as the reference's rule gives: header block(s) first, then no record crosses a block
    assert parts[0].startswith(b"BAM\1")
    assert all(len(p) <= 0xFF00 for p in parts)
    assert np.array_equal(o2.start_vo & 0xFFFF < 0x10000, np.ones(o.n_records, dtype=bool))
    first_block_of_records = int(o2.start_vo[0]) & 0xFFFF
    assert first_block_of_records == 0                         # writeReferenceSequenceInfo flushes the header block
    # no record straddles: every record ends in the block it starts in
    assert all((int(s) >> 16) == (int(e) >> 16) or (int(e) & 0xFFFF) == 0 for s, e in zip(o2.start_vo, o2.end_vo))


def test_layout_equals_the_suite_writer():
    rng = np.random.default_rng(17)
    recs = []
    pos = 0
    for k in range(4000):
        pos += int(rng.integers(0, 20))
        L = int(rng.integers(1, 400))
        seq = "".join("ACGT"[x] for x in rng.integers(0, 4, L))
        recs.append(bam_record(f"r{k}", seq, f"{L}M", pos, ref_id=k % 2))
    refs = [("c0", 1000000), ("c1", 500000)]
    want = blocks_of(make_bam(refs, recs))[:-1]                  # (without the EOF block)
    o = orc.Bam(make_bam(refs, recs)).decode()
    w = write_like(o)
    w.writeRecords(b"".join(recs))
    parts, stream = layout_stream(w)
    assert [len(p) for p in parts] == [len(p) for p in want]
    assert b"".join(parts) == b"".join(want)
    assert orc.Bam(stream).decode().n_records == len(recs)


def test_record_longer_than_a_block():
    # writer.d:259-267 + outputstream.d:107-132: the record starts a block of its own, the stream cuts it every 0xFF00
    # bytes, and — the writer's own count now exceeding a block — the next record starts a new block again
    small = [bam_record(f"s{k}", "ACGT" * 10, "40M", 100 + k) for k in range(6)]
    huge = bam_record("huge", "A" * 90000, "90000M", 5000)
    refs = [("c0", 1000000)]
    o = orc.Bam(make_bam(refs, small)).decode()
    w = write_like(o)
    w.writeRecords(b"".join(small[:3]) + huge + b"".join(small[3:]))
    parts, stream = layout_stream(w)
    n3, n_rest = sum(len(r) for r in small[:3]), sum(len(r) for r in small[3:])
    assert [len(p) for p in parts[1:]] == [n3, 0xFF00, 0xFF00, len(huge) - 2 * 0xFF00, n_rest]
    o2 = orc.Bam(stream).decode()
    assert o2.n_records == 7 and o2.name(3) == "huge" and int(o2.l_seq[3]) == 90000
    # a block that is filled exactly is emitted at once (outputstream.d:108: >=)
    fill = bam_record("f", "", "", 7, tags=b"XXZ" + b"y" * (0xFF00 - 4 - 32 - 2 - 4) + b"\0")
    assert len(fill) == 0xFF00
    w = write_like(o)
    w.writeRecords(fill + small[0])
    parts, _ = layout_stream(w)
    assert [len(p) for p in parts[1:]] == [0xFF00, len(small[0])]


def test_writer_argument_errors():
    from biod_b200 import BamWriter
    import io
    w = BamWriter(io.BytesIO())
    w.writeSamHeader("@HD\tVN:1.6\n")
    w.writeReferenceSequenceInfo([("c0", 1000)])
    with pytest.raises(Exception, match="Read reference ID is out of range"):
        w.writeRecord(bam_record("x", "ACGT", "4M", 5, ref_id=1), prefixed=True)
    w.writeRecord(bam_record("u", "ACGT", "", -1, ref_id=-1, flag=4), prefixed=True)      # unmapped reads are fine
    with pytest.raises(Exception):
        w.writeRecords(b"\x50\x00\x00\x00abc")                            # truncated
    with pytest.raises(ValueError):
        BamWriter(io.BytesIO(), compression_level=12)


def test_index_of_the_written_file():
    """biodb_writer_index (writer.d:139-195): the index BamWriter builds while writing == IndexBuilder over the reads
    and virtual offsets the finished file really has.  The file is compressed by zlib here (debug hook), so no GPU."""
    import ctypes as C
    from baiutil import build_bai_biod
    from biod_b200 import _capi
    L = _capi.lib()
    for name in ("ex1_header.bam", "bins.bam", "mg1655_chunk.bam"):
        o = orc.Bam(fixture_bytes(name)).decode()
        w = write_like(o)
        w.writeRecords(b"".join(struct.pack("<i", int(o.block_size[i])) + o.record_bytes(i).tobytes()
                                for i in range(o.n_records)))
        parts, stream = layout_stream(w)
        buf = np.frombuffer(stream, dtype=np.uint8)
        assert L.biodb_writer_debug_set_output(w._h, buf.ctypes.data, buf.size) == 0
        bai = w.index()
        o2 = orc.Bam(stream).decode()
        assert bai == build_bai_biod(o2, check_bins=True), name
    # unsorted records: the writer writes them, the index refuses
    w = write_like(o)
    w.writeRecord(bam_record("a", "ACGT", "4M", 100), prefixed=True)
    w.writeRecord(bam_record("b", "ACGT", "4M", 50), prefixed=True)
    parts, stream = layout_stream(w)
    buf = np.frombuffer(stream, dtype=np.uint8)
    assert L.biodb_writer_debug_set_output(w._h, buf.ctypes.data, buf.size) == 0
    with pytest.raises(Exception, match="not coordinate-sorted"):
        w.index()


def test_write_record_never_guesses_the_prefix():
    """A record body whose ref_id happens to equal len(body) - 4 (plausible with hundreds of contigs) used to be taken
    for bytes that already carry their block_size prefix; writeRecord(bytes) now always means a body."""
    from biod_b200 import BamWriter
    import io
    rec = bam_record("q", "ACGTACGT", "8M", 7, ref_id=0)
    body = bytearray(rec[4:])
    L = len(body)
    body[0:4] = struct.pack("<i", L - 4)                    # ref_id == len - 4
    w = BamWriter(io.BytesIO())
    w.writeSamHeader("@HD\tVN:1.6\n")
    w.writeReferenceSequenceInfo([("c%d" % i, 1000) for i in range(L)])
    w.writeRecord(bytes(body))
    data, cuts = w.layout()
    assert data[cuts[-1]:][:8] == struct.pack("<ii", L, L - 4) and data.endswith(bytes(body)[12:])   # (the bin is recalculated)
    with pytest.raises(Exception, match="block_size prefix"):
        w.writeRecord(bytes(body)[:-1], prefixed=True)      # said to be prefixed, but the prefix does not fit
