"""GPU parity of BAI random access: reader["chr"][beg:end] — the index's chunks inflated and
scanned on the device, then BamReadFilter (randomaccessmanager.d:366-462) as a device-side filter + compaction —
against the oracle's restatement: the same reads, in the same order, with the same raw bytes, fields, CIGAR words and
virtual offsets.  Regions of test/unittests.d:145-185 on bins.bam, the other indexed fixtures, and a synthetic file
whose chunks end in the middle of BGZF blocks and span several batches."""
import numpy as np
import pytest

from baiutil import build_bai
from conftest import fixture_bytes
from oracle import oracle as orc
from test_oracle_golden import BINS_REGIONS

pytestmark = pytest.mark.gpu


def gpu_region(rd, ref, beg, end):
    raws, sv, ev, fields, cig = [], [], [], [], []
    for b in rd.region_batches(ref, beg, end, copy=True):
        assert b.n > 0
        for i in range(b.n):
            o = int(b.rec_off[i]) + 4
            raws.append(b.data[o:o + int(b.block_size[i])].tobytes())
            fields.append((int(b.ref_id[i]), int(b.pos[i]), int(b.end_pos[i]), int(b.bin_mq_nl[i]), int(b.flag_nc[i]), int(b.l_seq[i])))
            cig.append(b.cigar[int(b.cigar_off[i]):int(b.cigar_off[i + 1])].tolist())
        sv += b.start_voffset.tolist()
        ev += b.end_voffset.tolist()
    return raws, sv, ev, fields, cig


def check_region(rd, o, bai, ref, beg, end):
    idx, wsv, wev = orc.region_reads(o, bai, ref, beg, end)
    raws, sv, ev, fields, cig = gpu_region(rd, ref, beg, end)
    assert len(raws) == len(idx), (ref, beg, end, len(raws), len(idx))
    for k, i in enumerate(idx):
        i = int(i)
        assert raws[k] == o.record_bytes(i).tobytes(), (ref, beg, end, k)
        assert fields[k] == (int(o.ref_id[i]), int(o.pos[i]), int(o.end_pos[i]),
                             (int(o.bin[i]) << 16) | (int(o.mapq[i]) << 8) | int(o.l_read_name[i]),
                             (int(o.flag[i]) << 16) | int(o.n_cigar[i]), int(o.l_seq[i]))
        assert cig[k] == o.cigar[int(o.cigar_off[i]):int(o.cigar_off[i + 1])].tolist()
    assert sv == wsv.tolist(), (ref, beg, end)
    assert ev == wev.tolist(), (ref, beg, end)
    return len(idx)


@pytest.mark.parametrize("bpb", [0, 1])
def test_bins_bam_regions(bpb):
    # test/unittests.d:145-185
    from biod_b200 import BamReader
    data = fixture_bytes("bins.bam")
    o = orc.Bam(data).decode()
    bai = orc.Bai(fixture_bytes("bins.bam.bai"))
    rd = BamReader(data, blocks_per_batch=bpb, want_offsets=True, index=fixture_bytes("bins.bam.bai"))
    large = o.ref_names.index("large")
    total = sum(check_region(rd, o, bai, large, beg, end) for beg, end in BINS_REGIONS)
    assert total > 0
    for name in o.ref_names:
        r = o.ref_names.index(name)
        assert check_region(rd, o, bai, r, 0, o.ref_lens[r]) > 0
        # test/unittests.d:193-205
        first = next(iter(rd[name][0:o.ref_lens[r]]))
        assert first.name == f"{name}:r1:0..1:len1:bin4681:hexbin0x1249"
        assert rd[name].firstPosition() == 0


@pytest.mark.parametrize("name", ["ex1_header.bam", "tags.bam"])
def test_other_indexed_fixtures(name):
    from biod_b200 import BamReader
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    raw = fixture_bytes(name + ".bai")
    bai = orc.Bai(raw)
    rd = BamReader(data, blocks_per_batch=2, want_offsets=True, index=raw)
    rng = np.random.default_rng(8)
    for r in range(len(o.ref_names)):
        ln = o.ref_lens[r]
        for beg, end in [(0, ln), (0, 1), (ln - 1, ln)] + [tuple(sorted(int(x) for x in rng.integers(0, ln, 2))) for _ in range(12)]:
            if beg < end:
                check_region(rd, o, bai, r, beg, end)


def test_synthetic_chunks_across_batches():
    from biod_b200 import BamReader
    from test_md_chain import random_pileup
    data = random_pileup(np.random.default_rng(21), 4000, refs=3, block_size=1800)
    o = orc.Bam(data).decode()
    raw = build_bai(o)
    bai = orc.Bai(raw)
    rng = np.random.default_rng(6)
    total = 0
    for bpb in (0, 1, 3):
        rd = BamReader(data, blocks_per_batch=bpb, want_offsets=True, index=raw)
        for r in range(3):
            for beg, end in [(0, 100000), (0, 1), (5000, 5001)] + [tuple(sorted(int(x) for x in rng.integers(0, 12000, 2))) for _ in range(10)]:
                if beg < end:
                    total += check_region(rd, o, bai, r, beg, end)
    assert total > 0


def test_region_argument_errors():
    from biod_b200 import BamReader
    data = fixture_bytes("bins.bam")
    rd = BamReader(data, index=fixture_bytes("bins.bam.bai"))
    with pytest.raises(Exception, match="start must be less than end"):
        list(rd["large"][10:10])
    with pytest.raises(Exception, match="does not exist"):
        rd["nope"]
    with pytest.raises(Exception, match="Invalid reference sequence index"):
        list(rd.region_reads(99, 0, 10))
    with pytest.raises(Exception, match="must be provided"):
        list(BamReader(data)["large"][0:10])
    # a sequential pass after region reads on the same reader is unaffected
    assert sum(b.n for b in rd.read_batches()) == orc.Bam(data).decode().n_records


def test_reads_between_and_read_at():
    # getReadsBetween (reader.d:350-356) and getReadAt (reader.d:336-339, test/unittests.d:193-205)
    from biod_b200 import BamReader
    data = fixture_bytes("ex1_header.bam")
    o = orc.Bam(data).decode()
    n = o.n_records
    for bpb in (0, 2):
        rd = BamReader(data, blocks_per_batch=bpb, want_offsets=True)
        for a, z in [(0, n), (0, 1), (5, 6), (100, 1500), (n - 3, n), (700, 701), (1234, 3000)]:
            want = [int(i) for i in orc.reads_between(o, int(o.start_vo[a]), int(o.end_vo[z - 1]))]
            assert want == list(range(a, z))
            got = list(rd.getReadsBetween(int(o.start_vo[a]), int(o.end_vo[z - 1])))
            assert [r.raw.tobytes() for r in got] == [o.record_bytes(i).tobytes() for i in want], (a, z)
            assert [r.start_virtual_offset for r in got] == o.start_vo[a:z].tolist()
        assert len(list(rd.getReadsBetween(int(o.start_vo[n - 10])))) == 10          # to the end of the file
        assert list(rd.getReadsBetween(int(o.start_vo[5]), int(o.start_vo[5]))) == []
        for i in (0, 1, 777, n - 1):
            r = rd.getReadAt(int(o.start_vo[i]))
            assert r.raw.tobytes() == o.record_bytes(i).tobytes() and r.start_virtual_offset == int(o.start_vo[i])
    data = fixture_bytes("bins.bam")
    rd = BamReader(data, want_offsets=True, index=fixture_bytes("bins.bam.bai"))
    for name in ("tiny", "small", "large"):
        assert rd.getReadAt(rd[name].startVirtualOffset()).name == f"{name}:r1:0..1:len1:bin4681:hexbin0x1249"


def check_region_pileup(rd, o, bai, ref, beg, end, **kw):
    """makePileup(bam[ref][beg .. end), ...) on the GPU against the oracle's pileup of the same reads."""
    from gpu_util import assert_pileup_equal
    idx = orc.region_reads(o, bai, ref, beg, end)[0]
    okw = dict(start_from=kw.get("start_from", 0), end_at=kw.get("end_at", 2**64 - 1),
               skip_zero_coverage=kw.get("skip_zero_coverage", True), use_md_tag=kw.get("use_md_tag", False))
    want = o.make_pileup_of(idx, **okw)
    assert want.status == 0
    g = dict(col_pos=[], col_ref=[], cov=[], n_start=[], read_idx=[], base=[], qual=[], qoff=[], ref_base=[])
    for b in rd.column_batches(True, want_query_offset=True, copy=True, region=(ref, beg, end), **kw):
        g["col_pos"].append(b.position)
        g["col_ref"].append(np.full(b.n_columns, b.ref_id, dtype=np.int32))
        g["cov"].append(np.diff(b.col_off).astype(np.uint64))
        g["n_start"].append(b.n_starting_here)
        g["read_idx"].append(b.read_idx)
        g["base"].append(b.base)
        g["qual"].append(b.qual)
        g["qoff"].append(b.query_offset)
        if b.reference_base is not None:
            g["ref_base"].append(b.reference_base)
    dts = dict(col_pos=np.uint64, col_ref=np.int32, cov=np.uint64, n_start=np.uint32, read_idx=np.uint32, base=np.uint8,
               qual=np.uint8, qoff=np.uint32, ref_base=np.uint8)
    g = {k: (np.concatenate(v) if v else np.zeros(0, dtype=dts[k])) for k, v in g.items()}
    g["col_off"] = np.concatenate([[0], np.cumsum(g["cov"])]).astype(np.uint64)
    # read_idx counts the reads of the region: map it to the record index of the whole file the oracle reports
    g["read_idx"] = idx[g["read_idx"].astype(np.int64)].astype(np.uint32) if len(g["read_idx"]) else g["read_idx"]
    assert_pileup_equal(g, want)
    if kw.get("use_md_tag"):
        assert g["ref_base"].tobytes() == np.asarray(want.ref_base, dtype=np.uint8).tobytes()
    return want.n_columns


def test_region_pileup_example():
    # examples/read_bam_file.d:21-25: makePileup(bam["chr2"][150 .. 160], false, 155, 158)
    from biod_b200 import BamReader, makePileup
    data = fixture_bytes("ex1_header.bam")
    raw = fixture_bytes("ex1_header.bam.bai")
    o = orc.Bam(data).decode()
    bai = orc.Bai(raw)
    rd = BamReader(data, index=raw)
    chr2 = o.ref_names.index("chr2")
    assert check_region_pileup(rd, o, bai, chr2, 150, 160, start_from=155, end_at=158) == 3
    cols = list(makePileup(rd["chr2"][150:160], start_from=155, end_at=158))
    assert [c.position for c in cols] == [155, 156, 157] and [c.coverage for c in cols] == [11, 10, 8]
    rng = np.random.default_rng(12)
    for bpb in (0, 1):
        rd = BamReader(data, blocks_per_batch=bpb, index=raw)
        for r in range(2):
            ln = o.ref_lens[r]
            for beg, end in [(0, ln), (100, 101)] + [tuple(sorted(int(x) for x in rng.integers(0, ln, 2))) for _ in range(6)]:
                if beg < end:
                    check_region_pileup(rd, o, bai, r, beg, end)
                    check_region_pileup(rd, o, bai, r, beg, end, start_from=beg, end_at=end, skip_zero_coverage=False)


def test_region_pileup_across_chunks_and_batches():
    from biod_b200 import BamReader
    from test_md_chain import random_pileup
    data = random_pileup(np.random.default_rng(33), 3000, refs=2, block_size=1500)
    o = orc.Bam(data).decode()
    raw = build_bai(o)
    bai = orc.Bai(raw)
    rng = np.random.default_rng(2)
    total = 0
    for bpb in (0, 1, 2):
        rd = BamReader(data, blocks_per_batch=bpb, index=raw)
        for r in range(2):
            for beg, end in [(0, 100000), (3000, 3001)] + [tuple(sorted(int(x) for x in rng.integers(0, 14000, 2))) for _ in range(5)]:
                if beg < end:
                    total += check_region_pileup(rd, o, bai, r, beg, end)
                    total += check_region_pileup(rd, o, bai, r, beg, end, use_md_tag=True, skip_zero_coverage=False)
    assert total > 0


def test_unmapped_reads_and_eof_offset():
    # reader.d:369-390 (unmappedReads), :177-179 (eofVirtualOffset)
    from biod_b200 import BamReader
    data = fixture_bytes("bins.bam")
    o = orc.Bam(data).decode()
    rd = BamReader(data, want_offsets=True, index=fixture_bytes("bins.bam.bai"))
    assert rd.eofVirtualOffset() == (len(data) - 28) << 16 == int(o.end_vo[-1])
    want = [i for i in range(o.n_records) if o.ref_id[i] == -1]
    assert want and want == list(range(want[0], o.n_records))        # they sit at the end of the file
    got = list(rd.unmappedReads())
    assert [r.raw.tobytes() for r in got] == [o.record_bytes(i).tobytes() for i in want]
    assert [r.start_virtual_offset for r in got] == o.start_vo[want].tolist()
    # a file without unmapped reads
    data = fixture_bytes("ex1_header.bam")
    rd = BamReader(data, want_offsets=True, index=fixture_bytes("ex1_header.bam.bai"))
    assert list(rd.unmappedReads()) == []


@pytest.mark.parametrize("name,bpb", [("bins.bam", 0), ("bins.bam", 1), ("ex1_header.bam", 0), ("tags.bam", 0)])
def test_reads_overlapping_several_regions(name, bpb):
    """getReadsOverlapping(BamRegion[]) (reader.d:361, randomaccessmanager.d:316-337): regions in any order, overlapping
    ones among them, several references — the same reads as the oracle's restatement, every read once, raw bytes and
    virtual offsets included."""
    from biod_b200 import BamReader
    from test_bai_index import random_regions
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    raw_bai = fixture_bytes(name + ".bai")
    bai = orc.Bai(raw_bai)
    rd = BamReader(data, blocks_per_batch=bpb, want_offsets=True, index=raw_bai)
    rng = np.random.default_rng(43)
    some = 0
    for trial in range(25):
        regions = random_regions(rng, o, span=3000 if trial % 2 else None)
        idx, wsv, wev = orc.regions_reads(o, bai, regions)
        raws, sv, ev = [], [], []
        for b in rd.regions_batches(regions, copy=True):
            for i in range(b.n):
                p = int(b.rec_off[i]) + 4
                raws.append(b.data[p:p + int(b.block_size[i])].tobytes())
            sv += b.start_voffset.tolist()
            ev += b.end_voffset.tolist()
        assert len(raws) == len(idx), (name, regions, len(raws), len(idx))
        assert raws == [o.record_bytes(int(i)).tobytes() for i in idx], (name, regions)
        assert sv == wsv.tolist() and ev == wev.tolist(), (name, regions)
        some += len(idx)
    assert some > 0
    got = [r.raw.tobytes() for r in rd.getReadsOverlapping([(0, 0, 50), (0, 20, 80)])]
    assert got == [o.record_bytes(int(i)).tobytes() for i in orc.regions_reads(o, bai, [(0, 0, 50), (0, 20, 80)])[0]]
    with pytest.raises(Exception):
        list(rd.getReadsOverlapping([(0, 10, 10)]))
