"""GPU parity on the workloads BASELINE.json names, generated exactly as bench.py generates them (tools/bamgen, seed
0xB10D + config index): the CUDA path through the C ABI against the oracle, every column array.
  configs[0]  examples/make_pileup.d on a 10 k-read sorted BAM, 1 ref, 150 bp — whole, with MD tags, plus the example's
              own invariants (examples/make_pileup.d:16-30)
  configs[1]  100 M reads, 1 contig, CIGAR 150M        — the first 2 M reads of the bench file
  configs[2]  100 M reads, mixed CIGAR (M/I/D/S/N)      — the first 2 M reads of the bench file
  configs[3]  1 G reads, 24 contigs, mixed CIGAR        — a 24-contig prefix, cut into 8 shards as 8 GPUs cut it
The 2 M-read comparisons stream: GPU batches are checked against slices of the oracle's arrays as they arrive."""
import numpy as np
import pytest

from gpu_util import assert_pileup_equal, gpu_pileup, gpu_pileup_sharded
from oracle import oracle as orc

pytestmark = pytest.mark.gpu


def bench_file(config, n_reads):
    import bench
    from tools import bamgen
    c = bench.CONFIGS[config]
    return bamgen.generate(n_reads, c["refs"], bool(c["mixed"]), -1, bamgen.SEED_BASE + config)


def test_config0_make_pileup_example():
    from biod_b200 import BamReader, makePileup
    from tools import bamgen
    data = bamgen.generate(10_000, 1, False, -1, bamgen.SEED_BASE + 1).tobytes()
    o = orc.Bam(data).decode()
    assert o.n_records == 10_000 and len(o.ref_names) == 1
    # makePileup(bam.reads, true) — every column, every entry, every reference base
    want = o.make_pileup(0, 2**64 - 1, True, use_md_tag=True)
    for bpb in (0, 3):
        g = gpu_pileup(data, True, bpb, use_md_tag=True)
        assert_pileup_equal(g, want)
        assert g["ref_base"].tobytes() == want.ref_base.tobytes()
    assert set(want.ref_base.tobytes()) <= set(b"ACGT"), "every position of this file is covered by a read with an MD tag"
    # the example's own checks (examples/make_pileup.d:20-30): the reads that start in the columns, joined, are the reads
    # of the file; every column's reads are sorted by coordinate
    bam = BamReader(data)
    starting, n_cols = [], 0
    for column in makePileup(bam, True):
        starting.append(column.reads_starting_here)
        n_cols += 1
        assert np.all(np.diff(column.reads.astype(np.int64)) > 0)
    assert np.array_equal(np.concatenate(starting), np.arange(10_000))
    assert n_cols == want.n_columns


def stream_compare(data, want, **kw):
    """Column batches of the GPU pass against the oracle's tables, batch by batch."""
    from biod_b200 import BamReader
    rd = BamReader(data)
    c0 = e0 = 0
    for b in rd.column_batches(False, want_query_offset=True, **kw):
        nc, ne = b.n_columns, b.n_entries
        assert np.array_equal(b.position, want.col_pos[c0:c0 + nc])
        assert (want.col_ref[c0:c0 + nc] == b.ref_id).all()
        assert np.array_equal(np.diff(b.col_off), np.diff(want.col_off[c0:c0 + nc + 1]))
        assert np.array_equal(b.n_starting_here, want.n_start[c0:c0 + nc])
        assert np.array_equal(b.read_idx, want.read_idx[e0:e0 + ne])
        assert np.array_equal(b.base, want.base[e0:e0 + ne])
        assert np.array_equal(b.qual, want.qual[e0:e0 + ne])
        assert np.array_equal(b.query_offset, want.qoff[e0:e0 + ne])
        c0 += nc
        e0 += ne
    assert (c0, e0) == (want.n_columns, want.n_entries)


@pytest.mark.parametrize("config", [2, 3])
def test_first_2m_reads_of_the_bench_files(config):
    """configs[1] / configs[2] as bench.py times them, first 2 M reads: ~10 M columns, ~300 M entries, bit for bit."""
    data = bench_file(config, 2_000_000)
    o = orc.Bam(data).decode()
    assert o.n_records == 2_000_000
    want = o.pileup_columns()
    assert want.status == 0
    stream_compare(data, want)
    del want
    # the compact encoding the e2e leg delivers expands to the same columns (a smaller prefix: the expansion is Python)
    small = bench_file(config, 100_000)
    assert_pileup_equal(gpu_pileup(small, False, 0, compact_reads=True), orc.Bam(small).decode().pileup_columns())


def test_config3_prefix_in_8_shards():
    """24 contigs, mixed CIGAR, cut into 8 shards the way bench.py --gpus 8 cuts the file: cuts fall inside contigs and
    next to contig changes; halos are exact (reads with 1000-base N-skips reach past a guess of one block)."""
    data = bench_file(4, 600_000)
    o = orc.Bam(data).decode()
    assert len(o.ref_names) == 24 and len(set(o.ref_id.tolist())) == 24
    want = o.pileup_columns()
    g = gpu_pileup_sharded(data, 8, halo_blocks=1)
    assert sum(i["n_own_records"] for i in g["shards"]) == o.n_records
    assert_pileup_equal(g, want)
    refs_cut = {i["lo_ref"] for i in g["shards"][1:]}
    assert len(refs_cut) >= 4, "the cuts should fall on several different contigs"
