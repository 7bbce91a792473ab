"""GPU parity of the MAQ genotype likelihoods (MaqSnpCaller, bio/std/hts/snpcallers/maq.d:319-540):
the fused device consumer of the pileup columns against the oracle's restatement (pinned in tests/test_oracle_maq.py).
Floating point: scores and call qualities within 2e-6 relative / 1e-4 absolute — the reference computes
`tmp1 - C * lhet` in x87 `real`, the kernel in double; genotypes bit-exact except where two scores are equal to that
tolerance (then either order of the tie is accepted)."""
import math

import numpy as np
import pytest

from bamutil import bam_record, make_bam, tag_z
from conftest import fixture_bytes
from oracle import oracle as orc

pytestmark = pytest.mark.gpu
REL, ABS = 2e-6, 1e-4
GT = lambda a, b: "ACGTN".index(a) * 5 + "ACGTN".index(b)  # noqa: E731


def gpu_maq(data, caller, single_ref=False, bpb=0, use_md=True):
    from biod_b200 import BamReader
    rd = BamReader(data, blocks_per_batch=bpb)
    out = {k: [] for k in ("pos", "ref", "gt0", "gt1", "s0", "s1", "n_valid", "refb")}
    calls = []
    for b in caller.genotypeLikelihoods(rd, single_ref=single_ref, use_md_tag=use_md):
        out["pos"].append(b.position)
        out["ref"].append(np.full(b.n_columns, b.ref_id))
        for k in ("gt0", "gt1", "s0", "s1", "n_valid"):
            out[k].append(b.maq[k])
        if b.reference_base is not None:
            out["refb"].append(b.reference_base)
        c = b.calls
        calls += [(b.ref_id, int(c["pos"][k]), int(c["gt"][k]), chr(int(c["ref"][k])), float(c["qual"][k])) for k in range(len(c["pos"]))]
    return {k: (np.concatenate(v) if v else np.zeros(0)) for k, v in out.items()}, calls


def check_against_oracle(data, caller, tables, bpb=0):
    o = orc.Bam(data).decode()
    p = o.pileup_columns(True, use_md_tag=True, keep=True)
    w = o.maq(p, tables, caller.minimum_base_quality)
    g, calls = gpu_maq(data, caller, False, bpb)
    assert np.array_equal(g["pos"], p.col_pos) and np.array_equal(g["ref"], p.col_ref)
    assert np.array_equal(g["n_valid"], np.minimum(w["n_valid"], 0xffff))
    assert g["refb"].tobytes() == p.ref_base.tobytes()
    assert np.allclose(g["s0"], w["s0"], rtol=REL, atol=ABS) and np.allclose(g["s1"], w["s1"], rtol=REL, atol=ABS)
    sc = w["scores"]
    for name, col in (("gt0", "s0"), ("gt1", "s1")):
        diff = np.flatnonzero(g[name] != w[name])
        for c in diff:                                    # only ties (to the tolerance) may come out in another order
            assert g[name][c] != 255 and math.isclose(sc[c, g[name][c]], w[col][c], rel_tol=REL, abs_tol=ABS), (name, c)
        assert len(diff) <= max(2, len(g[name]) // 1000)
    # findSNPs: reference base known, call differs from it, quality above the threshold
    code = {ord("A"): 0, ord("C"): 1, ord("G"): 2, ord("T"): 3}
    want = []
    for c in range(p.n_columns):
        if w["gt0"][c] == 255:
            continue
        r = code.get(int(p.ref_base[c]) & 0x5f, 4)
        q = float(w["s1"][c] - w["s0"][c])
        if w["gt0"][c] != r * 6 and q > caller.minimum_call_quality:
            want.append((int(p.col_ref[c]), int(p.col_pos[c]), int(w["gt0"][c]), chr(int(p.ref_base[c])), q))
    border = {(r, x) for r, x, _, _, q in want if abs(q - caller.minimum_call_quality) < 1e-3}
    border |= {(r, x) for r, x, _, _, q in calls if abs(q - caller.minimum_call_quality) < 1e-3}
    a = [t for t in calls if (t[0], t[1]) not in border]
    b = [t for t in want if (t[0], t[1]) not in border]
    assert [(t[0], t[1], t[3]) for t in a] == [(t[0], t[1], t[3]) for t in b]
    assert all(math.isclose(x[4], y[4], rel_tol=1e-5, abs_tol=1e-3) for x, y in zip(a, b))
    assert sum(1 for x, y in zip(a, b) if x[2] != y[2]) <= max(1, len(a) // 500)
    return len(want), p.n_columns


def test_hand_computed_columns_on_gpu():
    """One A of quality 30 covering one position: AA scores 0, the tie of C|A, G|A, T|A at 10 log10(2) resolves to C|A
    (the smallest genotype code), the call's quality is 3.0103 (tests/test_oracle_maq.py has the derivation)."""
    from biod_b200 import MaqSnpCaller
    refs = [("c", 1000)]
    recs = [bam_record("r1", "A", "1M", 10, qual=bytes([30]), tags=tag_z("MD", "1")),
            bam_record("r2", "AA", "2M", 20, qual=bytes([30, 30]), tags=tag_z("MD", "2")),
            bam_record("r3", "AC", "2M", 20, qual=bytes([20, 20]), tags=tag_z("MD", "2"))]
    data = make_bam(refs, recs)
    g, calls = gpu_maq(data, MaqSnpCaller(minimum_base_quality=13), True)
    assert g["pos"].tolist() == [10, 20, 21]
    assert (g["gt0"][0], g["gt1"][0]) == (GT("A", "A"), GT("C", "A"))
    assert g["s0"][0] == 0.0 and g["s1"][0] == pytest.approx(10 * math.log10(2), rel=1e-6)
    # position 20: A (q30) and A (q20), both forward: C|A = -C * lhet(2, 0) = 6.0206 behind A|A = 0
    assert (g["gt0"][1], g["gt1"][1]) == (GT("A", "A"), GT("C", "A"))
    assert g["s1"][1] == pytest.approx(20 * math.log10(2), rel=1e-6)
    assert g["refb"].tobytes() == b"AAA" or g["refb"].tobytes() == b"AAC"
    assert g["n_valid"].tolist() == [1, 2, 2]


@pytest.mark.parametrize("name", ["ex1_header.bam", "illu_20_chunk.bam", "ion_20_chunk.bam", "tags.bam"])
def test_fixture_likelihoods_and_calls(name):
    from biod_b200 import MaqSnpCaller
    tables = orc.Maq()
    data = fixture_bytes(name)
    for bpb in (0, 2):
        check_against_oracle(data, MaqSnpCaller(), tables, bpb)


def test_synthetic_with_substitutions_and_other_knobs():
    """configs[1]-style reads (0.5 % substitutions against the MD tags' reference, 30x): real SNP-like calls appear;
    other depcorr / eta / thresholds go through the same tables."""
    from biod_b200 import MaqSnpCaller
    from tools import bamgen
    data = bamgen.generate(40_000, 2, True, -1, bamgen.SEED_BASE + 3).tobytes()
    n_calls, n_cols = check_against_oracle(data, MaqSnpCaller(), orc.Maq(), 0)
    assert n_cols > 150_000
    caller = MaqSnpCaller(depcorr=0.3, eta=0.05, minimum_call_quality=20.0, minimum_base_quality=20)
    n2, _ = check_against_oracle(data, caller, orc.Maq(0.3, 0.05), 3)
    assert n2 <= n_calls or n_calls == 0


def test_find_snps_mirror():
    """MaqSnpCaller.findSNPs(bam.reads) as BioD's users call it (maq.d:321-326): DiploidCall5 objects."""
    from biod_b200 import BamReader, MaqSnpCaller
    data = fixture_bytes("ex1_header.bam")
    rd = BamReader(data)
    calls = list(MaqSnpCaller().findSNPs(rd, sample="s1"))
    assert calls and all(c.is_variant and c.quality > 6.0 and c.chromosome == "chr1" and c.sample == "s1" for c in calls)
    assert all(calls[i].position < calls[i + 1].position for i in range(len(calls) - 1))
    both = list(MaqSnpCaller().findSNPs(rd, single_ref=False))
    assert {c.chromosome for c in both} == {"chr1", "chr2"} and both[:len(calls)][0].position == calls[0].position
