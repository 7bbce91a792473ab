"""pileupChunks (bam/pileup.d:859-1015, bam/splitter.d:66-101).
CPU: biod_b200.chunk_plan — the cuts and column intervals — against a line-by-line simulation of ReadRangeSplitter and
PileupChunkRange over the oracle's records.  GPU: every chunk's columns against the oracle's makePileup over the same
reads (halo + chunk) and interval, with and without MD tags; the chunks together against the sequential pileup."""
import numpy as np
import pytest

from conftest import fixture_bytes
from oracle import oracle as orc


def simulate_biod(o, block_size, start_from=0, end_at=2**64 - 1):
    """ReadRangeSplitter.getNextChunk (splitter.d:66-90) + PileupChunkRange (pileup.d:876-940) with lists and loops."""
    n = o.n_records
    chunks, i = [], 0
    while i < n:                                              # getNextChunk
        first = i
        total = 4 + int(o.block_size[i])
        i += 1
        while total <= block_size and i < n:
            if o.ref_id[i] != o.ref_id[first]:
                break
            total += 4 + int(o.block_size[i])
            i += 1
        chunks.append(list(range(first, i)))
    out = []
    k = 0
    cur = None
    while True:                                               # constructor (pileup.d:876-901)
        if k >= len(chunks):
            return out
        cur = chunks[k]
        k += 1
        if o.ref_id[cur[0]] < 0:
            continue
        beg = int(o.pos[cur[0]])
        if beg >= end_at:
            return out
        right_end = max(int(o.end_pos[r]) for r in cur)
        if right_end > start_from:
            break
    while True:
        end_pos = int(o.pos[cur[-1]])                         # front (pileup.d:905-913)
        if k >= len(chunks) or o.ref_id[chunks[k][0]] != o.ref_id[cur[-1]]:
            end_pos = right_end
        out.append((cur[0], cur[-1] + 1, int(o.ref_id[cur[0]]), max(beg, start_from), min(end_pos, end_at)))
        prev = cur                                            # popFront (pileup.d:915-940)
        while True:
            if k >= len(chunks):
                return out
            cur = chunks[k]
            k += 1
            if o.ref_id[cur[0]] >= 0:
                break
        right_end = max(int(o.end_pos[r]) for r in cur)
        if prev and o.ref_id[prev[-1]] == o.ref_id[cur[0]]:
            beg = int(o.pos[prev[-1]])
        else:
            beg = int(o.pos[cur[0]])


def plan_of(o, block_size, start_from=0, end_at=2**64 - 1):
    from biod_b200 import chunk_plan
    return chunk_plan(o.ref_id, o.pos, o.end_pos, o.block_size, block_size, start_from, end_at)


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "tags.bam", "illu_20_chunk.bam"])
@pytest.mark.parametrize("block_size", [300, 70_000, 16_384_000])
def test_plan_matches_the_simulation(name, block_size):
    o = orc.Bam(fixture_bytes(name)).decode()
    for start_from, end_at in ((0, 2**64 - 1), (300, 1200), (5000, 100)):
        want = simulate_biod(o, block_size, start_from, end_at)
        got = [(p["first"], p["last"], p["ref_id"], p["start_position"], p["end_position"]) for p in plan_of(o, block_size, start_from, end_at)]
        assert got == want
        for p in plan_of(o, block_size, start_from, end_at):
            # the halo is exact: it is the first earlier read of the reference that reaches beyond the interval's start,
            # and no read in front of it does
            h, i = p["halo"], p["first"]
            assert h <= i and all(o.ref_id[r] == p["ref_id"] for r in range(h, i))
            if h < i:
                assert o.end_pos[h] > p["start_position"] or o.end_pos[h] > o.pos[i - 1]


def chunk_tables(chunk):
    pos, cov, nstart, ridx, base, qual, refb = [], [], [], [], [], [], []
    for b in chunk.column_batches(copy=True, want_query_offset=False):
        pos.append(b.position)
        cov.append(np.diff(b.col_off).astype(np.uint64))
        nstart.append(b.n_starting_here)
        ridx.append(b.read_idx)
        base.append(b.base)
        qual.append(b.qual)
        if b.reference_base is not None:
            refb.append(b.reference_base)
    cat = lambda v, dt: np.concatenate(v) if v else np.zeros(0, dtype=dt)  # noqa: E731
    return dict(pos=cat(pos, np.uint64), cov=cat(cov, np.uint64), nstart=cat(nstart, np.uint32), ridx=cat(ridx, np.uint32),
                base=cat(base, np.uint8), qual=cat(qual, np.uint8), refb=cat(refb, np.uint8))


@pytest.mark.gpu
@pytest.mark.parametrize("name,block_size", [("ex1_header.bam", 20_000), ("bins.bam", 20_000), ("mg1655_chunk.bam", 30_000),
                                              ("illu_20_chunk.bam", 1500)])
@pytest.mark.parametrize("use_md", [False, True])
def test_chunks_match_the_oracle(name, block_size, use_md):
    from biod_b200 import BamReader, pileupChunks
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    rd = BamReader(data, want_offsets=True, blocks_per_batch=3)
    plan = plan_of(o, block_size)
    chunks = list(pileupChunks(rd, use_md, block_size))
    assert len(chunks) == len(plan) and len(plan) >= 2
    all_pos, all_ref = [], []
    for ch, pl in zip(chunks, plan):
        assert (ch.ref_id, ch.start_position, ch.end_position, ch.first_read_index) == \
            (pl["ref_id"], pl["start_position"], pl["end_position"], pl["halo"])
        want = o.make_pileup_of(np.arange(pl["halo"], pl["last"]), pl["start_position"], pl["end_position"], True,
                                use_md_tag=use_md, single_ref=True)
        g = chunk_tables(ch)
        assert np.array_equal(g["pos"], want.col_pos)
        assert np.array_equal(g["cov"], np.diff(want.col_off))
        assert np.array_equal(g["nstart"], want.n_start)
        assert np.array_equal(g["ridx"].astype(np.int64) + ch.first_read_index, want.read_idx)   # read_idx counts from the first halo read
        assert np.array_equal(g["base"], want.base) and np.array_equal(g["qual"], want.qual)
        if use_md:
            assert g["refb"].tobytes() == want.ref_base.tobytes()
        all_pos.append(g["pos"])
        all_ref.append(np.full(len(g["pos"]), ch.ref_id))
    # consecutive and non-overlapping: together the chunks are the columns of the sequential pileup — up to where the
    # LAST chunk of a reference ends, which is the right end of that chunk's own reads (pileup.d:905-907): a longer read
    # of an earlier chunk can reach further (mg1655_chunk.bam, illu_20_chunk.bam), and BioD's chunks drop those columns
    # too (each chunk was compared with the oracle's pileup over exactly BioD's reads and interval above)
    seq = o.pileup_columns()
    got_pos, got_ref = np.concatenate(all_pos), np.concatenate(all_ref)
    n_cut = 0
    for r in np.unique(seq.col_ref):
        g, w = got_pos[got_ref == r], seq.col_pos[seq.col_ref == r]
        assert len(g) <= len(w) and np.array_equal(g, w[:len(g)])
        n_cut += len(w) - len(g)
    assert set(np.unique(got_ref)) == set(np.unique(seq.col_ref))
    if name in ("ex1_header.bam", "bins.bam"):
        assert n_cut == 0
    # iterating a chunk yields PileupColumn objects, like BioD's range of pileups
    first = next(ch for ch in chunks if ch.start_position < ch.end_position)
    col = next(iter(first))
    assert col.position == int(seq.col_pos[0]) and col.coverage == int(seq.col_off[1])


@pytest.mark.gpu
def test_chunks_of_a_sub_range():
    from biod_b200 import BamReader, pileupChunks
    data = fixture_bytes("ex1_header.bam")
    o = orc.Bam(data).decode()
    rd = BamReader(data, want_offsets=True)
    # start_from / end_at clip every chunk (pileup.d:905-913), so every reference yields its columns [400, 900)
    got = {}
    for ch in pileupChunks(rd, False, 10_000, 400, 900):
        got.setdefault(ch.ref_id, []).append(chunk_tables(ch)["pos"])
    assert sorted(got) == [0, 1]
    for ref in (0, 1):
        idx = np.nonzero(o.ref_id == ref)[0]
        want = o.make_pileup_of(idx, 400, 900, True, use_md_tag=False, single_ref=True)
        assert np.array_equal(np.concatenate(got[ref]), want.col_pos)
    assert np.array_equal(np.concatenate(got[0]), o.make_pileup(400, 900, True).col_pos)
