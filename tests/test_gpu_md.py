"""GPU parity of the MD-tag reference bases (makePileup / pileupColumns with use_md_tag = true,
PileupRangeUsingMdTag, bam/pileup.d:522-654): PileupColumn.reference_base of every column from the CUDA path, through
the C ABI, against the oracle's column-by-column restatement — on the reference's own pileup vectors, on its fixture
files and on random pileups with consistent and with broken MD tags, cut into batches of one to three BGZF blocks so
that the chain of providers, the drained segments and the kept dna() strings cross batch boundaries."""
import numpy as np
import pytest

from conftest import fixture_bytes
from gpu_util import assert_pileup_equal, gpu_pileup
from oracle import oracle as orc
from test_md_chain import random_pileup
from test_oracle_golden import pileup_vector_bam

pytestmark = pytest.mark.gpu


def text(a):
    return np.asarray(a, dtype=np.uint8).tobytes().decode("latin1")


def check(data, o, single_ref, bpb, skip=True, start_from=0, end_at=2**64 - 1, **kw):
    """The whole column table and its reference bases, GPU against oracle."""
    if single_ref:
        want = o.make_pileup(start_from, end_at, skip, use_md_tag=True)
        g = gpu_pileup(data, True, bpb, use_md_tag=True, skip_zero_coverage=skip, start_from=start_from, end_at=end_at, **kw)
    else:
        want = o.pileup_columns(skip, use_md_tag=True)
        g = gpu_pileup(data, False, bpb, use_md_tag=True, skip_zero_coverage=skip, **kw)
    assert want.status == 0
    assert_pileup_equal(g, want)
    assert len(g["ref_base"]) == want.n_columns
    got, exp = text(g["ref_base"]), text(want.ref_base)
    if got != exp:
        k = next(i for i in range(len(exp)) if got[i] != exp[i])
        raise AssertionError(f"reference_base differs at column {k} (ref {int(want.col_ref[k])}, position "
                             f"{int(want.col_pos[k])}): got {got[max(0, k - 20):k + 20]!r}, expected {exp[max(0, k - 20):k + 20]!r}")
    return exp


def test_reference_vectors_on_gpu():
    # bam/pileup.d:776-786 and :830-856
    data = pileup_vector_bam()
    o = orc.Bam(data).decode()
    for bpb in (0, 1):
        for skip in (True, False):
            full = check(data, o, True, bpb, skip)
            assert set(full) - {"N"}, "the vectors carry MD tags: some reference bases must be known"
            check(data, o, True, bpb, skip, 796, 849)
            check(data, o, False, bpb, skip)


@pytest.mark.parametrize("name", ["illu_20_chunk.bam", "ex1_header.bam", "tags.bam", "bins.bam", "mg1655_chunk.bam"])
def test_fixture_reference_bases(name):
    data = fixture_bytes(name)
    o = orc.Bam(data).decode()
    for bpb in (0, 1):
        check(data, o, False, bpb)
        check(data, o, True, bpb, False)
    if name == "illu_20_chunk.bam":                  # every read of this file carries MD:Z
        assert set(check(data, o, False, 2)) - {"N"}


@pytest.mark.parametrize("seed", range(4))
@pytest.mark.parametrize("consistent", [True, False])
def test_random_pileups_across_batches(seed, consistent):
    rng = np.random.default_rng(500 + seed)
    data = random_pileup(rng, 900, refs=1 + seed % 3, consistent=consistent, gap_p=0.03 if seed % 2 else 0.0,
                         block_size=700 + 300 * seed)
    o = orc.Bam(data).decode()
    for bpb in (0, 1, 3):
        for skip in (True, False):
            check(data, o, False, bpb, skip)
        check(data, o, True, bpb, True, start_from=int(rng.integers(50, 600)), end_at=int(rng.integers(900, 2500)))
        check(data, o, True, bpb, False)


def test_reference_bases_with_other_encodings():
    rng = np.random.default_rng(77)
    data = random_pileup(rng, 600, refs=2, block_size=900)
    o = orc.Bam(data).decode()
    exp = check(data, o, False, 2, compact_reads=True)
    # counts_only: no entries, but the same columns and reference bases
    from biod_b200 import BamReader
    rd = BamReader(data, blocks_per_batch=2)
    got = "".join(text(b.reference_base) for b in rd.column_batches(False, use_md_tag=True, counts_only=True, copy=True))
    assert got == exp
    # without the flag there is no array, and PileupColumn.reference_base is BioD's default 'N'
    from biod_b200 import pileupColumns
    cols = list(pileupColumns(BamReader(data)))
    assert cols and all(c.reference_base == "N" for c in cols[:50])
    cols = list(pileupColumns(BamReader(data, blocks_per_batch=1), use_md_tag=True))
    assert "".join(c.reference_base for c in cols) == exp
