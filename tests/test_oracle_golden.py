"""Pins the CPU oracle against the reference's OWN golden vectors (no GPU needed).

Every assertion below restates one made by BioD's tests; the citation is next to it.
"""
import numpy as np
import pytest

from conftest import fixture_bytes
from bamutil import bam_record, make_bam, tag_z, tags_to_sam
from oracle import oracle as orc


@pytest.fixture(scope="module")
def ex1():
    return orc.Bam(fixture_bytes("ex1_header.bam")).decode()


def test_header_fields(ex1):
    # test/unittests.d:70-84
    assert ex1.ref_names == ["chr1", "chr2"]
    assert ex1.ref_lens == [1575, 1584]
    assert ex1.header_text.startswith("@HD")


def test_first_records(ex1):
    # test/unittests.d:88-103
    assert ex1.sequence(0) == "CTCAAGGTTGTTGCAAGGGGGTCTATGTGAACAAA"
    assert bytes(ex1.qualities(0) + 33).decode() == "<<<7<<<;<<<<<<<<8;;<7;4<;<;;;;;94<;"
    assert ex1.ref_names[ex1.ref_id[0]] == "chr1"
    assert ex1.name(0) == "EAS56_57:6:190:289:82"
    assert ex1.flag[0] == 69
    assert ex1.pos[0] == 99
    assert ex1.mapq[0] == 0
    assert ex1.cigar_string(2) == "35M"


def sam_line(b, i):
    """BamRead.toSam restated for the fields the SAM golden holds (bam/read.d:613-700)."""
    rn = "*" if b.ref_id[i] < 0 else b.ref_names[b.ref_id[i]]
    if b.next_ref[i] < 0:
        mrn = "*"
    elif b.next_ref[i] == b.ref_id[i]:
        mrn = "="
    else:
        mrn = b.ref_names[b.next_ref[i]]
    seq = b.sequence(i) or "*"
    q = b.qualities(i)
    qs = "*" if len(q) == 0 or q[0] == 255 else bytes(q + 33).decode("latin-1")
    f = [b.name(i), str(b.flag[i]), rn, str(b.pos[i] + 1), str(b.mapq[i]), b.cigar_string(i), mrn,
         str(b.next_pos[i] + 1), str(b.tlen[i]), seq, qs]
    return "\t".join(f + tags_to_sam(b.tags_raw(i)))


def test_third_record_sam_text(ex1):
    # test/unittests.d:103
    assert sam_line(ex1, 2) == ("EAS51_64:3:190:727:308\t99\tchr1\t103\t99\t35M\t=\t263\t195\t"
                                "GGTGCAGAGCCGAGTCACGGGGTTGCCAGCACAGG\t<<<<<<<<<<<<<<<<<<<<<<<<<<<::<<<844\t"
                                "MF:i:18\tAq:i:73\tNM:i:0\tUQ:i:0\tH0:i:1\tH1:i:0")


def test_all_records_equal_sam_golden(ex1):
    # test/unittests.d:303 (3270 records) and :309-312 (SAM file == BAM file record for record)
    assert ex1.n_records == 3270
    sam = [l for l in fixture_bytes("ex1_header.sam").decode().split("\n") if l and not l.startswith("@")]
    assert len(sam) == 3270
    for i, line in enumerate(sam):
        assert sam_line(ex1, i) == line, i


@pytest.mark.parametrize("name,cls", [
    ("duplicated_block_size.bam", orc.ERR_BGZF), ("no_block_size.bam", orc.ERR_BGZF),
    ("wrong_extra_gzip_length.bam", orc.ERR_BGZF), ("wrong_bc_subfield_length.bam", orc.ERR_BGZF),
    ("corrupted_zlib_archive.bam", orc.ERR_ZLIB)])
def test_corrupted_files_raise_the_pinned_classes(name, cls):
    # test/unittests.d:132-142 — either the constructor or the iteration throws
    with pytest.raises(orc.OracleError) as ei:
        orc.Bam(fixture_bytes(name)).decode()
    assert ei.value.status == cls
    if cls == orc.ERR_BGZF:
        assert ei.value.msg.startswith("Error reading BGZF block starting from offset ")


def test_error_messages_and_locations():
    with pytest.raises(orc.OracleError) as ei:
        orc.Bam(fixture_bytes("wrong_bc_subfield_length.bam")).decode()
    assert ei.value.offset == 36489 and "wrong BC subfield length: 5; expected 2" in ei.value.msg
    with pytest.raises(orc.OracleError) as ei:
        orc.Bam(fixture_bytes("duplicated_block_size.bam"))
    assert "duplicate field with block size" in ei.value.msg
    with pytest.raises(orc.OracleError) as ei:
        orc.Bam(fixture_bytes("no_block_size.bam"))
    assert "block size was not found in any subfield" in ei.value.msg
    with pytest.raises(orc.OracleError) as ei:
        orc.Bam(fixture_bytes("corrupted_zlib_archive.bam")).decode()
    assert ei.value.zerr == -3  # Z_DATA_ERROR


def test_lazy_error_keeps_earlier_records():
    # the fault is in block 3: records of blocks 1-2 are delivered before the throw (unittests.d:139)
    b = orc.Bam(fixture_bytes("wrong_bc_subfield_length.bam")).decode(raise_on_error=False)
    assert b.status == orc.ERR_BGZF and b.n_records > 0


@pytest.mark.parametrize("name,n_blocks,n_bytes,n_reads", [
    ("ex1_header.bam", 9, 456679, 3270), ("illu_20_chunk.bam", 3, 32310, 29), ("bins.bam", 28, 1670543, 16458),
    ("tags.bam", 3, 31612, 417), ("b7_295_chunk.bam", 22, None, 449), ("mg1655_chunk.bam", 26, None, 556),
    ("ion_20_chunk.bam", 13, None, 476), ("long_header.bam", 3, 75979, 0)])
def test_valid_fixtures_decode(name, n_blocks, n_bytes, n_reads):
    b = orc.Bam(fixture_bytes(name)).decode()
    assert b.n_records == n_reads
    assert b.n_blocks == n_blocks - 1  # SURVEY counts include the 28-byte EOF block
    if n_bytes is not None:
        assert len(b.udata) == n_bytes
    # virtual offsets are consistent: each record starts where the previous one ended
    if n_reads:
        assert b.start_vo[0] == b.reads_start_voffset
        assert np.array_equal(b.start_vo[1:], b.end_vo[:-1])


def test_makepileup_stops_at_first_reference(ex1):
    # test/unittests.d:320-330
    p = ex1.make_pileup()
    assert p.status == 0 and p.n_columns == 1470
    assert set(p.col_ref.tolist()) == {0}
    assert not (ex1.flag[p.read_idx] & 4).any()
    assert (ex1.ref_id[p.read_idx] == 0).all()


def test_pileupcolumns_counts_and_first_columns(ex1):
    # test/unittests.d:334-367
    p = ex1.pileup_columns()
    assert p.status == 0
    assert int((p.col_ref == 0).sum()) == 1470 and int((p.col_ref == 1).sum()) == 1567
    assert (np.diff(p.col_ref) >= 0).all()
    c0 = 0
    c1 = int(np.argmax(p.col_ref == 1))
    assert p.col_pos[c0] == 99 and ex1.name(int(p.read_idx[p.col_off[c0]])) == "EAS56_57:6:190:289:82"
    assert p.col_pos[c1] == 0 and ex1.name(int(p.read_idx[p.col_off[c1]])) == "B7_591:8:4:841:340"
    col_of_entry = np.repeat(np.arange(p.n_columns), np.diff(p.col_off).astype(np.int64))
    assert (ex1.ref_id[p.read_idx] == p.col_ref[col_of_entry]).all()
    assert not (ex1.flag[p.read_idx] & 4).any()


# ---- in-module vectors of bam/pileup.d:699-857 -----------------------------------------------
SEQS = ["ATTATGGACATTGTTTCCGTTATCATCATCATCATCATCATCATCATTATCATC",
        "GACATTGTTTCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATC",
        "ATTGTTTCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATCACC",
        "TGTTTCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATCACCAC",
        "TCCGTTATCATCATCATCATCATCATCATCATCATCATCATCATCACCACCACC",
        "GTTATCATCATCATCATCATCATCATCATCATCATCATCATCATCGTCACCCTG",
        "TCATCATCATCATAATCATCATCATCATCATCATCATCGTCACCCTGTGTTGAG",
        "TCATCATCATCGTCACCCTGTGTTGAGGACAGAAGTAATTTCCCTTTCTTGGCT",
        "TCATCATCATCATCACCACCACCACCCTGTGTTGAGGACAGAAGTAATATCCCT",
        "CACCACCACCCTGTGTTGAGGACAGAAGTAATTTCCCTTTCTTGGCTGGTCACC"]
CIGARS = ["54M", "54M", "50M3I1M", "54M", "54M", "54M", "2S52M", "16M15D38M", "13M3I38M", "54M"]
POSITIONS = [758, 764, 767, 769, 773, 776, 785, 795, 804, 817]
MDS = ["47C6", "54", "51", "50T3", "46T7", "45A0C7", "11C24A0C14", "11A3T0^CATCATCATCACCAC38", "15T29T5", "2T45T5"]


def pileup_vector_bam():
    recs = [bam_record(f"r{i}", SEQS[i], CIGARS[i], POSITIONS[i], tags=tag_z("MD", MDS[i])) for i in range(10)]
    return make_bam([("20", 63025520)], recs)


def test_pileup_unit_vector():
    # bam/pileup.d:776-825
    b = orc.Bam(pileup_vector_bam()).decode()
    sub = b.make_pileup(796, 849, False)
    full = b.make_pileup(0, 2**64 - 1, False)
    assert sub.status == 0 and full.status == 0
    assert sub.col_pos[0] == 796
    assert sub.col_pos[-1] == 848  # half-open: the range is empty once position >= end_at (pileup.d:505)
    k = int(np.argmax(full.col_pos == 796))
    n = sub.n_columns
    assert np.array_equal(np.diff(sub.col_off), np.diff(full.col_off)[k:k + n])  # :789
    col = {int(p): c for c, p in enumerate(full.col_pos)}
    assert full.bases(col[796]) == "CCCCCCAC"        # :793
    assert full.bases(col[805]) == "TCCCCCCCC"       # :797
    assert full.bases(col[806]) == "AAAAAAAGA"       # :801
    assert full.bases(col[821]) == "AAGG-AA"         # :815
    assert full.bases(col[826]) == "CCCCCC"          # :819
    assert full.bases(col[849]) == "TAT"             # :823

    def entry(pos, k_from_end):
        c = col[pos]
        return int(full.col_off[c + 1]) - k_from_end

    # :805 cigar_after.front.type == 'D' at 810 for reads[coverage-2]
    e = entry(810, 2)
    r = int(full.read_idx[e])
    assert b.cigar_ops(r)[int(full.op_index[e]) + 1][1] == "D"
    # :809 cigar_before.back.type == 'I' at 817
    e = entry(817, 2)
    r = int(full.read_idx[e])
    assert b.cigar_ops(r)[int(full.op_index[e]) - 1][1] == "I"
    # :813 cigar_operation.type == 'D' at 821 for reads[coverage-3]
    e = entry(821, 3)
    r = int(full.read_idx[e])
    assert b.cigar_ops(r)[int(full.op_index[e])][1] == "D"
    assert full.qual[e] == 255 and full.base[e] == ord("-")


def test_pileup_zero_coverage_vector():
    # bam/pileup.d:830-856: with skip_zero_coverage=false one column per reference position of dna(reads)
    seqs = ["CCCACATAGAAAGCTTGCTGTTTCTCTGTGGGAAGTTTTAACTTAGGTCAGCTT",
            "TAGAAAGCTTGCTGTTTCTCTGTGGGAAGTTTTAACTTAGGTTAGCTTCATCTA",
            "TTTTTCTTTCTTTCTTTGAAGAAGGCAGATTCCTGGTCCTGCCACTCAAATTTT",
            "TTTCTTTCTTTCTTTGAAGAAGGCAGATTCCTGGTCCTGCCACTCAAATTTTCA"]
    pos = [979, 985, 1046, 1048]
    recs = [bam_record(f"r{i + 1}", seqs[i], "54M", pos[i]) for i in range(4)]
    b = orc.Bam(make_bam([("20", 63025520)], recs)).decode()
    p = b.make_pileup(0, 2**64 - 1, False)
    assert p.n_columns == (1048 + 54) - 979
    assert np.array_equal(p.col_pos, np.arange(979, 1048 + 54, dtype=np.uint64))
    cov = np.diff(p.col_off)
    assert (cov[985 + 54 - 979:1046 - 979] == 0).all() and cov[0] == 1
    q = b.make_pileup(0, 2**64 - 1, True)
    assert q.n_columns == int((cov > 0).sum())


def test_make_pileup_example_invariant():
    # examples/make_pileup.d:20-30: joiner(columns.reads_starting_here) == bam.reads
    b = orc.Bam(fixture_bytes("illu_20_chunk.bam")).decode()
    assert b.n_records == 29
    p = b.make_pileup()
    starting = []
    for c in range(p.n_columns):
        hi = int(p.col_off[c + 1])
        starting += p.read_idx[hi - int(p.n_start[c]):hi].tolist()
    assert starting == list(range(29))
    # ... and column.reads.equalRange(column.position) is the same set (make_pileup.d:20-24)
    for c in range(p.n_columns):
        lo, hi = int(p.col_off[c]), int(p.col_off[c + 1])
        here = [int(r) for r in p.read_idx[lo:hi] if b.pos[r] == p.col_pos[c]]
        assert here == p.read_idx[hi - int(p.n_start[c]):hi].tolist()


def test_bins_bam_all_cigar_ops():
    # bins.bam is the only fixture with N, = and X operations
    b = orc.Bam(fixture_bytes("bins.bam")).decode()
    ops = set((b.cigar & 15).tolist())
    assert {0, 1, 2, 3, 7, 8} <= ops
    p = b.pileup_columns()
    assert p.status == 0 and p.n_entries > 0


# ---- reference bases from MD tags (oracle restatement of pileup.d:522-654) --------------
def _py_reference_of(seq, cigar, md):
    """Independent, plain reconstruction of the reference bases a well-formed read covers: {ref offset: base}."""
    import re
    ops = [(int(n), o) for n, o in re.findall(r"(\d+)([MIDNSHP=X])", cigar)]
    aligned, q = [], 0                    # query bases of M/=/X in order
    for n, o in ops:
        if o in "M=X":
            aligned += list(seq[q:q + n])
        if o in "MIS=X":
            q += n
    toks = re.findall(r"\d+|\^[A-Z]+|[A-Z]", md)
    ref, k = [], 0
    for t in toks:
        if t[0].isdigit():
            ref += aligned[k:k + int(t)]
            k += int(t)
        elif t[0] == "^":
            ref += list(t[1:])
        else:
            ref.append(t)
            k += 1
    # lay the string over reference offsets: M/=/X and D consume it, N consumes reference without bases
    out, off, i = {}, 0, 0
    for n, o in ops:
        if o in "M=XD":
            for _ in range(n):
                out[off] = ref[i]
                off += 1
                i += 1
        elif o == "N":
            off += n
    assert i == len(ref)
    return out


def _expected_reference(positions, seqs, cigars, mds):
    exp = {}
    for p, s_, c, m in zip(positions, seqs, cigars, mds):
        for off, b in _py_reference_of(s_, c, m).items():
            assert exp.setdefault(p + off, b) == b, "reads disagree on the reference"
    return exp


def test_reference_bases_unit_vector():
    # bam/pileup.d:776-786: column.reference_base == dna(reads)[column.position - first_read_position]
    b = orc.Bam(pileup_vector_bam()).decode()
    exp = _expected_reference(POSITIONS, SEQS, CIGARS, MDS)
    for args in ((796, 849, False), (0, 2**64 - 1, False), (0, 2**64 - 1, True)):
        p = b.make_pileup(*args, use_md_tag=True)
        assert p.status == 0 and p.n_columns > 0
        got = "".join(chr(x) for x in p.ref_base)
        want = "".join(exp.get(int(pos), "N") for pos in p.col_pos)
        assert got == want
    # dna(read) itself, read by read (md/reconstruct.d:216-260 checks the same function on other vectors)
    for i in range(10):
        d = _py_reference_of(SEQS[i], CIGARS[i], MDS[i])
        assert b.dna(i) == "".join(d[k] for k in sorted(d))
    # without the flag every column keeps PileupColumn's default (pileup.d:240)
    assert set(b.make_pileup(796, 849, False).ref_base.tolist()) == {ord("N")}


def test_reference_bases_zero_coverage_vector():
    # bam/pileup.d:830-856: equal(dna(reads), map!(c => c.reference_base)(makePileup(reads, true, 0, ulong.max, false)))
    seqs = ["CCCACATAGAAAGCTTGCTGTTTCTCTGTGGGAAGTTTTAACTTAGGTCAGCTT",
            "TAGAAAGCTTGCTGTTTCTCTGTGGGAAGTTTTAACTTAGGTTAGCTTCATCTA",
            "TTTTTCTTTCTTTCTTTGAAGAAGGCAGATTCCTGGTCCTGCCACTCAAATTTT",
            "TTTCTTTCTTTCTTTGAAGAAGGCAGATTCCTGGTCCTGCCACTCAAATTTTCA"]
    pos = [979, 985, 1046, 1048]
    mds = ["54", "42C7C3", "54", "54"]
    recs = [bam_record(f"r{i + 1}", seqs[i], "54M", pos[i], tags=tag_z("MD", mds[i])) for i in range(4)]
    b = orc.Bam(make_bam([("20", 63025520)], recs)).decode()
    exp = _expected_reference(pos, seqs, ["54M"] * 4, mds)
    p = b.make_pileup(0, 2**64 - 1, False, use_md_tag=True)
    got = "".join(chr(x) for x in p.ref_base)
    want = "".join(exp.get(q, "N") for q in range(979, 1048 + 54))     # dna(reads): 'N' where no read covers (reconstruct.d:262-270)
    assert got == want
    q = b.make_pileup(0, 2**64 - 1, True, use_md_tag=True)
    assert "".join(chr(x) for x in q.ref_base) == "".join(exp[int(x)] for x in q.col_pos)


def test_md_operations_and_missing_tags():
    # md/parse.d:148-182 vectors go through dna(read): matches, mismatches, deletions, zero matches are skipped
    recs = [bam_record("a", "ACGTACGTAC", "10M", 10, tags=tag_z("MD", "3A0C5")),
            bam_record("b", "ACGTACGTAC", "4M2D6M", 10, tags=tag_z("MD", "4^GG6")),
            bam_record("c", "ACGTACGTAC", "2S4M1I3M", 12, tags=tag_z("XX", "zz") + tag_z("MD", "0T6")),
            bam_record("d", "ACGTACGTAC", "10M", 14)]                      # no MD tag: empty dna
    b = orc.Bam(make_bam([("r", 1000)], recs)).decode()
    assert b.dna(0) == "ACG" + "A" + "C" + "CGTAC"        # 3 matches, mismatch A, (0 skipped), mismatch C, 5 matches
    assert b.dna(1) == "ACGT" + "GG" + "ACGTAC"           # the deleted bases come from the tag
    assert b.dna(2) == "T" + "TAC" + "TAC"                # soft clip and insertion are not reference positions
    assert b.dna(3) == ""
    p = b.pileup_columns(use_md_tag=True)
    assert p.status == 0 and len(p.ref_base) == p.n_columns


# ---- BAI random access --------------------------------------------------------------------
BINS_REGIONS = [(1400, 1500), (10, 123), (135, 1236), (1350, 3612), (643, 1732), (267, 1463), (0, 30), (1363, 1612),
                (361, 1231), (322, 612), (912, 938), (0, 3000), (0, 100), (0, 1000), (0, 1900), (1, 279)] + \
               [(i, i + 100) for i in range(50_000, 1_000_000, 50_000)]


def naive_region(b, ref_id, beg, end):
    # test/unittests.d:151-156: ref matches, position < end, position + basesCovered() > beg
    return [i for i in range(b.n_records)
            if b.ref_id[i] == ref_id and b.pos[i] < end and b.end_pos[i] > beg]


def test_bai_region_reads_equal_the_naive_filter():
    # test/unittests.d:145-185: bf["large"][beg .. end] == filter over all reads, for every listed interval
    data = fixture_bytes("bins.bam")
    b = orc.Bam(data).decode()
    bai = orc.Bai(fixture_bytes("bins.bam.bai"))
    assert bai.n_refs == len(b.ref_names)
    large = b.ref_names.index("large")
    for beg, end in BINS_REGIONS:
        idx, sv, ev = orc.region_reads(b, bai, large, beg, end)
        assert list(idx) == naive_region(b, large, beg, end), (beg, end)
        assert np.array_equal(sv, b.start_vo[idx])
        # end offsets: as in the sequential walk, except that after the last record of a chunk the stream has already
        # moved on to the next chunk (inputstream.d:516-524), whose start is what virtualTell reports
        begs = {a for a, _ in bai.chunks(large, beg, end)}
        assert all(int(e) == int(w) or int(e) in begs for e, w in zip(ev, b.end_vo[idx]))
    # every reference, whole and in parts
    for name in b.ref_names:
        r = b.ref_names.index(name)
        ln = b.ref_lens[r]
        for beg, end in [(0, ln), (0, 1), (ln // 2, ln // 2 + 1), (ln - 1, ln), (ln // 3, 2 * ln // 3)]:
            if beg < end:
                assert list(orc.region_reads(b, bai, r, beg, end)[0]) == naive_region(b, r, beg, end), (name, beg, end)


def test_bai_first_reads_of_references():
    # test/unittests.d:193-205: getReadAt(bf[name].startVirtualOffset()).name
    data = fixture_bytes("bins.bam")
    b = orc.Bam(data).decode()
    bai = orc.Bai(fixture_bytes("bins.bam.bai"))
    for name in ("tiny", "small", "large"):
        r = b.ref_names.index(name)
        idx, sv, _ = orc.region_reads(b, bai, r, 0, b.ref_lens[r])
        assert b.name(int(idx[0])) == f"{name}:r1:0..1:len1:bin4681:hexbin0x1249"
        # getReadAt: the first record of the stream that starts at that virtual offset
        assert int(orc.reads_between(b, int(sv[0]), int(b.end_vo[int(idx[0])]))[0]) == int(idx[0])


@pytest.mark.parametrize("name", ["ex1_header.bam", "tags.bam"])
def test_bai_other_fixtures(name):
    data = fixture_bytes(name)
    b = orc.Bam(data).decode()
    bai = orc.Bai(fixture_bytes(name + ".bai"))
    rng = np.random.default_rng(5)
    for r in range(len(b.ref_names)):
        ln = b.ref_lens[r]
        regions = [(0, ln)] + [tuple(sorted(int(x) for x in rng.integers(0, ln, 2))) for _ in range(25)]
        for beg, end in regions:
            if beg < end:
                # (no reference test pins these files; placed-but-unmapped reads need not be in the index's chunks)
                got = [int(i) for i in orc.region_reads(b, bai, r, beg, end)[0]]
                naive = naive_region(b, r, beg, end)
                assert got == [i for i in naive if i in set(got)], (name, r, beg, end)
                assert all(b.flag[i] & 4 for i in set(naive) - set(got)), (name, r, beg, end)


def test_reads_between_virtual_offsets():
    # getReadsBetween (randomaccessmanager.d:186-196): records from `from` on whose end offset is not beyond `to`
    data = fixture_bytes("ex1_header.bam")
    b = orc.Bam(data).decode()
    n = b.n_records
    for a, z in [(0, n), (0, 1), (5, 6), (100, 1500), (n - 3, n), (700, 701), (1234, 3000)]:
        got = orc.reads_between(b, int(b.start_vo[a]), int(b.end_vo[z - 1]))
        assert list(got) == list(range(a, z)), (a, z)


def test_bai_parse_errors():
    good = fixture_bytes("bins.bam.bai")
    with pytest.raises(orc.OracleError):
        orc.Bai(b"BAM\1" + good[4:])
    with pytest.raises(orc.OracleError):
        orc.Bai(good[:len(good) // 2])
    bai = orc.Bai(good)
    with pytest.raises(orc.OracleError):
        bai.chunks(bai.n_refs, 0, 10)          # "Invalid reference sequence index"
