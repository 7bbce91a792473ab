"""GPU parity: the CUDA path (through the C ABI) against the CPU oracle, bit for bit."""
import numpy as np
import pytest

from conftest import fixture_bytes
from bamutil import bam_record, make_bam, tag_z
from gpu_util import REC_FIELDS, assert_pileup_equal, gpu_pileup, gpu_records
from oracle import oracle as orc
from test_oracle_golden import CIGARS, MDS, POSITIONS, SEQS, pileup_vector_bam

pytestmark = pytest.mark.gpu

VALID = ["ex1_header.bam", "illu_20_chunk.bam", "bins.bam", "tags.bam", "b7_295_chunk.bam", "mg1655_chunk.bam",
         "ion_20_chunk.bam", "long_header.bam"]


def check_records(data, bpb):
    o = orc.Bam(data).decode()
    rd, g, raws, err = gpu_records(data, blocks_per_batch=bpb)
    assert err is None, err
    assert rd.header_text == o.header_text
    assert [r.name for r in rd.reference_sequences] == o.ref_names
    assert [r.length for r in rd.reference_sequences] == o.ref_lens
    assert rd.reads_start_voffset == o.reads_start_voffset
    assert len(raws) == o.n_records
    for f in REC_FIELDS:
        exp = getattr(o, f) if hasattr(o, f) else None
        if f == "bin_mq_nl":
            exp = (o.bin.astype(np.uint32) << 16) | (o.mapq.astype(np.uint32) << 8) | o.l_read_name
        if f == "flag_nc":
            exp = (o.flag.astype(np.uint32) << 16) | o.n_cigar
        assert np.array_equal(g[f], exp), f
    assert np.array_equal(g["cigar"], o.cigar)
    assert np.array_equal(g["n_cigar_rec"], np.diff(o.cigar_off))
    assert np.array_equal(g["start_voffset"], o.start_vo)
    assert np.array_equal(g["end_voffset"], o.end_vo)
    for i, raw in enumerate(raws):
        assert raw == o.record_bytes(i).tobytes(), i
    return o


@pytest.mark.parametrize("name", VALID)
@pytest.mark.parametrize("bpb", [0, 3, 1])
def test_records_match_oracle(name, bpb):
    check_records(fixture_bytes(name), bpb)


@pytest.mark.parametrize("name", VALID[:-1])
@pytest.mark.parametrize("bpb", [0, 2])
def test_pileup_columns_match_oracle(name, bpb):
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    assert_pileup_equal(gpu_pileup(data, False, bpb), o.pileup_columns())
    assert_pileup_equal(gpu_pileup(data, True, bpb), o.make_pileup())


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam"])
def test_pileup_with_zero_coverage_columns(name):
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    assert_pileup_equal(gpu_pileup(data, False, 2, skip_zero_coverage=False), o.pileup_columns(False))
    assert_pileup_equal(gpu_pileup(data, True, 3, skip_zero_coverage=False), o.make_pileup(0, 2**64 - 1, False))


def test_ex1_pins_on_gpu():
    # test/unittests.d:303, :334-367 through the CUDA path itself
    data = fixture_bytes("ex1_header.bam")
    g = gpu_pileup(data, False)
    assert int((g["col_ref"] == 0).sum()) == 1470 and int((g["col_ref"] == 1).sum()) == 1567
    assert g["col_pos"][0] == 99
    rd, rec, raws, err = gpu_records(data)
    assert err is None and len(raws) == 3270


def test_pileup_unit_vector_on_gpu():
    # bam/pileup.d:776-825
    data = pileup_vector_bam()
    o = orc.Bam(data).decode()
    for bpb in (0, 1):
        assert_pileup_equal(gpu_pileup(data, True, bpb, start_from=796, end_at=849, skip_zero_coverage=False),
                            o.make_pileup(796, 849, False))
        full = gpu_pileup(data, True, bpb, skip_zero_coverage=False)
        assert_pileup_equal(full, o.make_pileup(0, 2**64 - 1, False))
    col = {int(p): c for c, p in enumerate(full["col_pos"])}

    def bases(p):
        a, b = int(full["col_off"][col[p]]), int(full["col_off"][col[p] + 1])
        return full["base"][a:b].tobytes().decode()

    assert bases(796) == "CCCCCCAC" and bases(805) == "TCCCCCCCC" and bases(806) == "AAAAAAAGA"
    assert bases(821) == "AAGG-AA" and bases(826) == "CCCCCC" and bases(849) == "TAT"


@pytest.mark.parametrize("start,end", [(0, 2**64 - 1), (500, 900), (1000, 1001), (99, 100), (5000, 6000), (0, 99)])
@pytest.mark.parametrize("skip", [True, False])
def test_make_pileup_ranges(start, end, skip):
    data = fixture_bytes("ex1_header.bam")
    o = orc.Bam(data).decode()
    for bpb in (0, 2):
        assert_pileup_equal(gpu_pileup(data, True, bpb, start_from=start, end_at=end, skip_zero_coverage=skip),
                            o.make_pileup(start, end, skip))


@pytest.mark.parametrize("name,cls", [
    ("duplicated_block_size.bam", "BgzfException"), ("no_block_size.bam", "BgzfException"),
    ("wrong_extra_gzip_length.bam", "BgzfException"), ("wrong_bc_subfield_length.bam", "BgzfException"),
    ("corrupted_zlib_archive.bam", "ZlibException")])
def test_corrupted_files_raise_the_pinned_classes(name, cls):
    # test/unittests.d:132-142
    data = fixture_bytes(name)
    with pytest.raises(orc.OracleError) as oe:
        orc.Bam(data).decode()
    for bpb in (2, 0, 5):
        try:
            rd, g, raws, err = gpu_records(data, blocks_per_batch=bpb)
        except Exception as e:  # noqa: BLE001  raised by the constructor
            err, raws = e, []
        assert err is not None and type(err).__name__ == cls
        try:
            n_before = orc.Bam(data).decode(raise_on_error=False).n_records
        except orc.OracleError:
            n_before = 0
        assert len(raws) == n_before, (bpb, len(raws), n_before)
    if cls == "BgzfException":
        assert str(err) == oe.value.msg
    else:
        assert err.errnum == oe.value.zerr == -3
    # records ahead of the fault are still delivered, exactly as many as the oracle sees
    ob = None
    try:
        ob = orc.Bam(data).decode(raise_on_error=False)
    except orc.OracleError:
        pass
    if ob is not None:
        assert len(raws) == ob.n_records


def synthetic(n=3000, seed=7, refs=(("chrA", 100000), ("chrB", 50000)), indel=True):
    rng = np.random.default_rng(seed)
    recs, k = [], 0
    for rid, (_, ln) in enumerate(refs):
        pos = 0
        for _ in range(n // len(refs)):
            pos += int(rng.geometric(0.2)) - 1
            L = int(rng.integers(30, 151))
            seq = "".join("ACGTN"[i] for i in rng.choice(5, L, p=[.24, .24, .24, .24, .04]))
            qual = bytes(rng.integers(2, 42, L).tolist())
            cig = f"{L}M"
            r = rng.random()
            if indel and r < 0.5 and L > 40:
                a = int(rng.integers(5, L - 20))
                kind = rng.integers(0, 6)
                if kind == 0:
                    x = int(rng.integers(1, 10)); cig = f"{a}M{x}I{L - a - x}M"
                elif kind == 1:
                    x = int(rng.integers(1, 10)); cig = f"{a}M{x}D{L - a}M"
                elif kind == 2:
                    x = int(rng.integers(50, 400)); cig = f"{a}M{x}N{L - a}M"
                elif kind == 3:
                    x = int(rng.integers(1, 15)); cig = f"{x}S{L - x}M"
                elif kind == 4:
                    x = int(rng.integers(1, 15)); cig = f"3H{L - x}M{x}S"
                else:
                    cig = f"{a}={1}X{L - a - 1}M2P"
            flag = 0
            if rng.random() < 0.03:
                flag, cig = 4, "*" if rng.random() < 0.5 else cig
            if cig == "*":
                cig = ""
            recs.append(bam_record(f"q{k}", seq, cig, pos, ref_id=rid, flag=flag, qual=qual,
                                   mapq=int(rng.integers(0, 61)), tags=tag_z("XX", "y" * int(rng.integers(0, 30)))))
            k += 1
    return list(refs), recs


@pytest.mark.parametrize("straddle", [False, True])
@pytest.mark.parametrize("level", [0, 1, 6, 9])
def test_synthetic_mixed_cigar(straddle, level):
    refs, recs = synthetic()
    data = make_bam(refs, recs, level=level, straddle=straddle, block_size=0xFF00 if not straddle else 4000)
    o = check_records(data, 5)
    for skip in (True, False):
        assert_pileup_equal(gpu_pileup(data, False, 3, skip_zero_coverage=skip), o.pileup_columns(skip))
    assert_pileup_equal(gpu_pileup(data, True, 2, start_from=1234, end_at=40000), o.make_pileup(1234, 40000))


def test_empty_and_headers_only():
    data = make_bam([("c", 10)], [])
    rd, g, raws, err = gpu_records(data)
    assert err is None and raws == []
    assert gpu_pileup(data, False)["col_pos"].size == 0
    # a file that is only an EOF block is not a BAM (reader.d:113)
    from bamutil import BGZF_EOF
    with pytest.raises(Exception) as ei:
        gpu_records(BGZF_EOF)
    assert type(ei.value).__name__ == "BamFormatException"


def test_stream_stops_at_first_empty_block():
    # inputstream.d:393-394: an ISIZE==0 block in the middle silently ends the read range
    from bamutil import BGZF_EOF, bam_header, bgzf_block
    refs, recs = synthetic(200, indel=False)
    hdr = bam_header("@HD\tVN:1.6\n", refs)
    data = bgzf_block(hdr) + bgzf_block(b"".join(recs[:100])) + BGZF_EOF + bgzf_block(b"".join(recs[100:])) + BGZF_EOF
    o = orc.Bam(data).decode()
    assert o.n_records == 100
    rd, g, raws, err = gpu_records(data)
    assert err is None and len(raws) == 100


def test_truncated_record_raises_read_exception():
    from bamutil import bam_header, bgzf_block, BGZF_EOF
    refs, recs = synthetic(50, indel=False)
    hdr = bam_header("@HD\tVN:1.6\n", refs)
    body = b"".join(recs)
    data = bgzf_block(hdr) + bgzf_block(body[:-20]) + BGZF_EOF
    o = orc.Bam(data).decode(raise_on_error=False)
    assert o.status == orc.ERR_TRUNC
    rd, g, raws, err = gpu_records(data)
    assert type(err).__name__ == "ReadException" and len(raws) == o.n_records
    # fewer than 4 stray bytes end the range silently (readrange.d:139-150)
    data = bgzf_block(hdr) + bgzf_block(body + b"\x01\x02") + BGZF_EOF
    rd, g, raws, err = gpu_records(data)
    assert err is None and len(raws) == len(recs)


def test_independent_interleaved_passes():
    # every bam.reads call is a fresh pass (reader.d:228-231); examples/make_pileup.d iterates four times
    from biod_b200 import BamReader
    data = fixture_bytes("ex1_header.bam")
    rd = BamReader(data, blocks_per_batch=2)
    a, b = rd.reads(), rd.reads()
    names_a, names_b = [], []
    for _ in range(1000):
        names_a.append(next(a).name)
    for _ in range(500):
        names_b.append(next(b).name)
    assert names_a[:500] == names_b
    assert sum(1 for _ in rd.reads()) == 3270


def test_make_pileup_example_invariant_on_gpu():
    # examples/make_pileup.d:20-30
    from biod_b200 import BamReader, makePileup
    data = fixture_bytes("illu_20_chunk.bam")
    bam = BamReader(data)
    starting = []
    for column in makePileup(bam, True):
        starting += column.reads_starting_here.tolist()
    assert starting == list(range(29))


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "b7_295_chunk.bam", "ion_20_chunk.bam", "mg1655_chunk.bam"])
@pytest.mark.parametrize("n_shards,halo", [(2, 8), (3, 2), (5, 8), (4, 0)])
def test_sharded_pileup_equals_unsharded(name, n_shards, halo):
    """SURVEY.md §8e / pileup.d:859-1015: shards + EXACT halos; concatenated shard outputs == one sequential pass.
    bins.bam holds reads spanning megabases and the *_chunk files pile hundreds of reads on one spot, so a guessed halo
    of a few blocks is too short there: the shards' reach reports must say so and the re-run from the exact offset must
    be right — no halo size is ever chosen by hand.  mg1655_chunk.bam is htsjdk-written: its records straddle BGZF
    blocks, so every cut has to find the record chain's entry first."""
    from gpu_util import gpu_pileup_sharded
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    g = gpu_pileup_sharded(data, n_shards, halo_blocks=halo, blocks_per_batch=3)
    assert sum(i["n_own_records"] for i in g["shards"]) == o.n_records
    assert_pileup_equal(g, o.pileup_columns())
    if halo == 0 and name != "mg1655_chunk.bam":
        # without any halo every shard whose first column is covered by an earlier read must have been run again
        cuts = [(i["lo_ref"], i["lo_pos"]) for i in g["shards"]]
        covered = [t for t in range(1, n_shards) if g["shards"][t]["n_own_records"] and
                   np.any((o.ref_id == cuts[t][0]) & (o.pos < cuts[t][1]) & (o.end_pos > cuts[t][1]) & (o.end_pos > o.pos))]
        assert set(covered) <= set(g["redone"])


@pytest.mark.parametrize("skip", [True, False])
def test_sharded_pileup_synthetic(skip):
    from gpu_util import gpu_pileup_sharded
    from tools import bamgen
    data = bamgen.generate(60000, 3, True, level=1, threads=4).tobytes()
    o = orc.Bam(data).decode()
    for n_shards in (2, 4, 7):
        g = gpu_pileup_sharded(data, n_shards, halo_blocks=4, blocks_per_batch=16, skip_zero_coverage=skip)
        assert_pileup_equal(g, o.pileup_columns(skip))


def test_sharded_pileup_in_spans_of_shards():
    """biodb_pileup_begin_shard_span: 7 shards run as three passes of 2, 1 and 4 shards (unequal shares of one file for
    workers of unequal speed) — the same columns as the sequential pass, halos exact."""
    from gpu_util import gpu_pileup_sharded
    from tools import bamgen
    data = bamgen.generate(40000, 2, True, level=1, threads=4).tobytes()
    o = orc.Bam(data).decode()
    for spans in ([2, 1, 4], [7], [1, 6], [3, 3, 1]):
        g = gpu_pileup_sharded(data, 7, halo_blocks=1, blocks_per_batch=9, spans=spans)
        assert len(g["shards"]) == len(spans) and sum(i["n_own_records"] for i in g["shards"]) == o.n_records
        assert_pileup_equal(g, o.pileup_columns())
    from biod_b200 import BamReader
    rd = BamReader(data, blocks_per_batch=9)
    n_col = sum(b.n_columns for _, b in rd.sharded_column_batches(7, spans=[2, 1, 4]))
    assert n_col == o.pileup_columns().n_columns
    with pytest.raises(ValueError):
        list(rd.column_batches(False, shard=(5, 7, 3)))


def test_sharded_pileup_of_a_straddling_file():
    """htsjdk layout at scale: records cut across BGZF blocks, mixed CIGARs with long N-skips, 6 shards, halo guessed
    at one block — cuts enter through the plausibility search, exact halos repair the guess."""
    from gpu_util import gpu_pileup_sharded
    from tools import bamgen
    data = bamgen.generate(50000, 2, True, level=1, threads=4, straddle=True).tobytes()
    o = orc.Bam(data).decode()
    g = gpu_pileup_sharded(data, 6, halo_blocks=1, blocks_per_batch=7)
    assert sum(i["n_own_records"] for i in g["shards"]) == o.n_records
    assert_pileup_equal(g, o.pileup_columns())
    # the sequential helper of the mirror (exact halos known up front, nothing run twice) gives the same columns
    from biod_b200 import BamReader
    rd = BamReader(data, blocks_per_batch=7)
    n_col = n_ent = 0
    for _, b in rd.sharded_column_batches(6):
        n_col += b.n_columns
        n_ent += b.n_entries
    p = o.pileup_columns()
    assert (n_col, n_ent) == (p.n_columns, p.n_entries)


def _oracle_shard_with_md(o, info, first_index, skip=True):
    """What BioD computes for one chunk (pileup.d:905-913): makePileup over the halo reads + the chunk's reads with
    use_md_tag, clipped to the chunk's column interval."""
    lo, hi = (info["lo_ref"], info["lo_pos"]), (info["hi_ref"], info["hi_pos"])
    a = first_index - info["n_halo_records"]
    b = first_index + info["n_own_records"]
    p = o.make_pileup_of(np.arange(a, b), 0, 2**64 - 1, skip, use_md_tag=True, single_ref=False)
    key = lambda r, x: (2**40 if r < 0 else int(r), int(x))  # noqa: E731
    keep = np.array([key(*lo) <= key(r, x) < key(*hi) for r, x in zip(p.col_ref, p.col_pos.astype(np.int64))], dtype=bool)
    return p, keep


@pytest.mark.parametrize("name", ["illu_20_chunk.bam", "ex1_header.bam", "mg1655_chunk.bam"])
def test_sharded_pileup_with_md_tags(name):
    """use_md_tag in shards: every shard's reference bases are those of BioD's own chunk — the provider chain starts
    at the first halo read (makePileup(chain(prev_chunk, chunk), use_md_tag, beg, end))."""
    from gpu_util import gpu_pileup_sharded
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    n_shards = 3
    g = gpu_pileup_sharded(data, n_shards, halo_blocks=1, blocks_per_batch=2, use_md_tag=True)
    assert_pileup_equal(g, o.pileup_columns())
    first, want = 0, []
    for info in g["shards"]:
        p, keep = _oracle_shard_with_md(o, info, first)
        want.append(p.ref_base[keep])
        first += info["n_own_records"]
    want = np.concatenate(want)
    assert len(g["ref_base"]) == len(want) == o.pileup_columns().n_columns
    assert g["ref_base"].tobytes() == want.tobytes()
    if name == "illu_20_chunk.bam":
        assert set(want.tobytes()) - {ord("N")}


def test_crc_verification():
    # block.d:187: debug builds of BioD assert the CRC32 of every inflated block; verify_crc does it on the device
    import struct
    from biod_b200 import BamReader
    data = bytearray(fixture_bytes("ex1_header.bam"))
    rd = BamReader(bytes(data), verify_crc=True, blocks_per_batch=3)
    assert sum(1 for _ in rd.reads()) == 3270
    # corrupt the CRC field of the third block: inflate still succeeds, the check must fail when that block is reached
    o = orc.Bam(bytes(data)).decode()
    b = o.blocks[2]
    crc_at = int(b[6]) + int(b[2])
    data[crc_at] ^= 0x5A
    assert sum(1 for _ in BamReader(bytes(data), blocks_per_batch=3).reads()) == 3270     # not checked by default
    rd = BamReader(bytes(data), verify_crc=True, blocks_per_batch=2)
    got = 0
    with pytest.raises(Exception) as ei:
        for _ in rd.reads():
            got += 1
    assert type(ei.value).__name__ == "ZlibException" and "CRC32" in str(ei.value)
    assert 0 < got < 3270


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "ion_20_chunk.bam"])
def test_counts_only_mode(name):
    # fused consumer view: per-column A,C,G,T,other,deletion counts == histogram of the oracle's bases
    from biod_b200 import BamReader
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    p = o.pileup_columns()
    col = np.repeat(np.arange(p.n_columns), np.diff(p.col_off).astype(np.int64))
    cat = np.full(len(p.base), 4)
    for k, ch in enumerate(b"ACGT"):
        cat[p.base == ch] = k
    cat[p.base == ord("-")] = 5
    exp = np.zeros((p.n_columns, 6), dtype=np.uint32)
    np.add.at(exp, (col, cat), 1)
    rd = BamReader(data, blocks_per_batch=3)
    got, pos, cov = [], [], []
    for b in rd.column_batches(False, counts_only=True, copy=True):
        assert b.read_idx.size == 0
        got.append(b.counts)
        pos.append(b.position)
        cov.append(np.diff(b.col_off))
    assert np.array_equal(np.concatenate(pos), p.col_pos)
    assert np.array_equal(np.concatenate(got), exp)
    assert np.array_equal(np.concatenate(cov), np.diff(p.col_off))


def test_straddling_layout_matches_oracle():
    # htsjdk-style files fill every BGZF block regardless of record boundaries: the fused chain walk has to find
    # its entry point in every block (plausibility search) and the resolve kernel has to confirm it
    from tools import bamgen
    data = bamgen.generate(40000, 2, True, level=1, threads=4, straddle=True).tobytes()
    o = check_records(data, 7)
    assert_pileup_equal(gpu_pileup(data, False, 5), o.pileup_columns())


@pytest.mark.parametrize("mixed", [False, True])
def test_large_synthetic_properties(mixed):
    """Size-independent properties of a 2 M-read, 2-contig file (the oracle comparison at this size is
    tests/test_gpu_configs.py; these are the checks that also hold at sizes no oracle reaches): every live read starts in
    exactly one column, coverage sums to the sum of reference spans, columns are sorted, entries stay in file
    order, sharded == unsharded, counts == histogram of bases."""
    from biod_b200 import BamReader
    from tools import bamgen
    n = 2_000_000
    data = bamgen.generate(n, 2, mixed, level=1)
    rd = BamReader(data)
    pos, end = [], []
    for b in rd.read_batches():
        pos.append(b.pos.copy())
        end.append(b.end_pos.copy())
    pos, end = np.concatenate(pos), np.concatenate(end)
    assert pos.size == n
    spans = (end - pos).astype(np.int64)
    tot_cols = tot_ent = tot_start = 0
    last_key = (-1, -1)
    chk = 0
    for b in rd.column_batches(False):
        cov = np.diff(b.col_off).astype(np.int64)
        assert (cov > 0).all()
        assert (np.diff(b.position.astype(np.int64)) > 0).all()
        assert (b.ref_id, int(b.position[0])) > last_key
        last_key = (b.ref_id, int(b.position[-1]))
        tot_cols += b.n_columns
        tot_ent += b.n_entries
        tot_start += int(b.n_starting_here.sum())
        # file order inside columns: read_idx strictly increasing within each column
        d = np.diff(b.read_idx.astype(np.int64))
        ends = (b.col_off[1:-1] - b.col_off[0]).astype(np.int64) - 1      # boundaries between columns
        mask = np.ones(d.size, dtype=bool)
        mask[ends] = False
        assert (d[mask] > 0).all()
        chk += int(b.base.astype(np.uint64).sum() + 3 * b.qual.astype(np.uint64).sum())
    assert tot_start == n
    assert tot_ent == int(spans.sum())
    # sharded run: same totals and checksum
    s_cols = s_ent = s_chk = 0
    for _, b in rd.sharded_column_batches(3):
        s_cols += b.n_columns
        s_ent += b.n_entries
        s_chk += int(b.base.astype(np.uint64).sum() + 3 * b.qual.astype(np.uint64).sum())
    assert (s_cols, s_ent, s_chk) == (tot_cols, tot_ent, chk)
    # counts mode: per-column counts sum to the coverage
    c_tot = 0
    for b in rd.column_batches(False, counts_only=True):
        assert np.array_equal(b.counts.sum(1), np.diff(b.col_off).astype(np.uint32))
        c_tot += int(b.counts.sum())
    assert c_tot == tot_ent


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "ion_20_chunk.bam", "b7_295_chunk.bam"])
@pytest.mark.parametrize("bpb", [0, 2])
def test_compact_read_lists_match_oracle(name, bpb):
    """compact_reads (last read + window mask + stragglers per column) carries exactly the explicit read_idx lists."""
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    assert_pileup_equal(gpu_pileup(data, False, bpb, compact_reads=True), o.pileup_columns())
    assert_pileup_equal(gpu_pileup(data, True, bpb, compact_reads=True, skip_zero_coverage=False),
                        o.make_pileup(0, 2**64 - 1, False))


def test_compact_read_lists_with_stragglers():
    """Long N-skips keep a read alive across hundreds of later reads: those come back as stragglers."""
    from biod_b200 import BamReader
    refs, recs = synthetic(n=6000, seed=11, refs=(("chrA", 100000),))
    data = make_bam(refs, recs, level=6)
    o = orc.Bam(data).decode()
    for bpb in (0, 3):
        assert_pileup_equal(gpu_pileup(data, False, bpb, compact_reads=True), o.pileup_columns())
    n_strag = 0
    for b in BamReader(data).column_batches(False, compact_reads=True, copy=True):
        assert b.compact is not None
        n_strag += len(b.compact[3])
        assert len(b.compact[4]) >= 1 and b.compact[5][-1] == b.n_columns
    assert n_strag > 0


def test_sharded_pileup_with_compact_columns():
    """Shards deliver the same columns in the compact encoding (what bench.py's e2e leg asks for at N > 1)."""
    from gpu_util import gpu_pileup_sharded
    from tools import bamgen
    data = bamgen.generate(50000, 2, True, level=6, threads=4).tobytes()
    o = orc.Bam(data).decode()
    g = gpu_pileup_sharded(data, 3, halo_blocks=4, blocks_per_batch=8, compact_reads=True)
    assert_pileup_equal(g, o.pileup_columns())


def test_streamed_file_equals_in_memory(tmp_path):
    """biodb_open(path) streams the file (header window + two pinned slabs per pass, inputstream.d:467-478) instead of
    holding it: the same records, offsets and columns as over the in-memory buffer — on a fixture, and on a file several
    times the size of the header window whose batches cross the window many times."""
    from biod_b200 import BamReader
    from tools import bamgen
    for name, data, bpb in (("ex1_header.bam", fixture_bytes("ex1_header.bam"), 2),
                            ("synthetic", bamgen.generate(300_000, 2, True, level=1, threads=4).tobytes(), 97)):
        path = tmp_path / (name + ".bam")
        path.write_bytes(data)
        a = BamReader(str(path), blocks_per_batch=bpb, want_offsets=True)
        b = BamReader(data, blocks_per_batch=bpb, want_offsets=True)
        assert a.header_text == b.header_text and a.reads_start_voffset == b.reads_start_voffset
        n = 0
        for x, y in zip(a.read_batches(copy=True), b.read_batches(copy=True)):
            assert x.n == y.n and np.array_equal(x.pos, y.pos) and np.array_equal(x.end_pos, y.end_pos)
            assert np.array_equal(x.start_voffset, y.start_voffset) and np.array_equal(x.cigar, y.cigar)
            assert x.data[:int(x.rec_off[-1])].tobytes() == y.data[:int(y.rec_off[-1])].tobytes()
            n += x.n
        assert n == (3270 if name == "ex1_header.bam" else 300_000)
        ca = [(c.n_columns, c.n_entries, int(c.position[0]), int(c.base.astype(np.uint64).sum())) for c in a.column_batches(False)]
        cb = [(c.n_columns, c.n_entries, int(c.position[0]), int(c.base.astype(np.uint64).sum())) for c in b.column_batches(False)]
        assert ca == cb and ca
        # shards read the file through the same window
        sa = [(s, c.n_columns, c.n_entries) for s, c in a.sharded_column_batches(3)]
        assert sum(x[1] for x in sa) == sum(x[0] for x in ca)
    # a truncated file surfaces as it does in memory: the records before the cut, then the error
    data = fixture_bytes("ex1_header.bam")
    path = tmp_path / "cut.bam"
    path.write_bytes(data[:len(data) // 2])
    got = 0
    with pytest.raises(Exception):
        for _ in BamReader(str(path), blocks_per_batch=2).reads():
            got += 1
    assert 0 < got < 3270
