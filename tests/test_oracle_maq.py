"""Pins of the oracle's MAQ restatement (bio/std/hts/snpcallers/maq.d:66-310,388-486).
The reference has no test for the caller, so the oracle is pinned on values computed BY HAND from the formulas of the
model (closed forms for one and two reads) and on an independent plain-double restatement over random columns; the
genotype order — GenotypeLikelihoodInfo walks the TinyMap in genotype-code order and insertion-sorts by score, equal
scores keeping code order (maq.d:258-277, tinymap.d:139-158) — is pinned on a column where three genotypes tie."""
import math

import numpy as np
import pytest

from conftest import fixture_bytes
from oracle import oracle as orc

C10 = 10.0 / math.log(10.0)
FLT_MIN = np.float32(1.1754943508222875e-38)
GT = lambda a, b: "ACGTN".index(a) * 5 + "ACGTN".index(b)  # noqa: E731  DiploidGenotype!Base5 code of a|b


@pytest.fixture(scope="module")
def maq():
    return orc.Maq()            # MaqSnpCaller's defaults: depcorr 0.17, eta 0.03 (as floats, maq.d:329-331)


def test_tables_closed_forms(maq):
    dep, eta = float(np.float32(0.17)), float(np.float32(0.03))
    for n in (0, 1, 2, 17, 255):
        assert maq.fk[n] == pytest.approx((1 - dep) ** n * (1 - eta) + eta, rel=1e-12)          # maq.d:85-87
    for n, k in ((1, 0), (2, 1), (10, 3), (30, 15), (255, 100)):
        want = math.log(math.comb(n, k)) - n * math.log(2)                                       # :100-101
        assert maq.lhet[n << 8 | k] == pytest.approx(want, rel=1e-9, abs=1e-9)
    # beta(q, n, k) = -10/ln10 * log(P(X >= k+1) / P(X >= k)), X ~ Binomial(n, e), e = 10^(-q/10)  (:105-119)
    for q in (4, 13, 30, 41, 63):
        e = 10.0 ** (-q / 10.0)
        assert maq.beta[q << 16 | 1 << 8 | 0] == pytest.approx(q, rel=1e-9)                      # P(X>=1)/P(X>=0) = e
        assert maq.beta[q << 16 | 2 << 8 | 0] == pytest.approx(-C10 * math.log(e * (2 - e)), rel=1e-9)
        assert maq.beta[q << 16 | 2 << 8 | 1] == pytest.approx(-C10 * math.log(e / (2 - e)), rel=1e-9)
        tail = lambda n, k: sum(math.comb(n, j) * e ** j * (1 - e) ** (n - j) for j in range(k, n + 1))  # noqa: E731
        for n, k in ((5, 2), (30, 0), (30, 7)):
            assert maq.beta[q << 16 | n << 8 | k] == pytest.approx(-C10 * math.log(tail(n, k + 1) / tail(n, k)), rel=1e-6)


def test_hand_computed_single_read(maq):
    """One A of quality 30 on the forward strand.  fsum[A] = fk(0) = 1, bsum[A] = beta(30, 1, 0) = 30.
    AA: nothing else was seen -> 0.  CC / GG / TT: bsum[A] = 30.  X|A: the other two nucleotides were not seen ->
    -C * lhet(1, 0) = 10 log10(2) = 3.0103.  Heterozygotes without A: 30 - C * lhet(0, 0) = 30."""
    s = maq.compute(b"A", [30], [0])
    assert s[GT("A", "A")] == 0.0
    for x in "CGT":
        assert s[GT(x, x)] == pytest.approx(30.0, rel=1e-6)
        assert s[GT(x, "A")] == pytest.approx(10 * math.log10(2), rel=1e-6)
        assert s[GT("A", x)] == FLT_MIN                           # symmetric=false: only (later, earlier) pairs are set
    for a, b in (("G", "C"), ("T", "C"), ("T", "G")):
        assert s[GT(a, b)] == pytest.approx(30.0, rel=1e-6)
    assert (s.reshape(5, 5)[4] == FLT_MIN).all() and (s.reshape(5, 5)[:, 4] == FLT_MIN).all()   # nothing with N


def test_hand_computed_two_reads(maq):
    """A (q30) and A (q20), both forward: the q30 base is taken first (w = c = 0), the q20 one second (w = c = 1):
    bsum[A] = fk(0) * beta(30, 2, 0) + fk(1) * beta(20, 2, 1) = 26.99187 + 0.8351 * 22.98853 = 46.18959."""
    fk1 = float(np.float32(0.83)) * float(np.float32(0.97)) + float(np.float32(0.03))
    e30, e20 = 1e-3, 1e-2
    want = -C10 * math.log(e30 * (2 - e30)) + fk1 * (-C10 * math.log(e20 / (2 - e20)))
    assert want == pytest.approx(46.18959, abs=2e-5)
    for order in ((30, 20), (20, 30)):                           # the order in the column does not matter here
        s = maq.compute(b"AA", list(order), [0, 0])
        assert s[GT("C", "C")] == pytest.approx(want, rel=1e-6)
        assert s[GT("A", "A")] == 0.0
        assert s[GT("C", "A")] == pytest.approx(2 * 10 * math.log10(2), rel=1e-6)     # -C * lhet(2, 0) = -C * log(1/4)
    # opposite strands: both bases are the first of their strand (w = 0 twice) but c still counts the base
    s = maq.compute(b"AA", [30, 20], [0, 1])
    assert s[GT("C", "C")] == pytest.approx(-C10 * math.log(e30 * (2 - e30)) + 1.0 * (-C10 * math.log(e20 / (2 - e20))), rel=1e-6)


def test_equal_qualities_keep_column_order(maq):
    """The pinned tie-break of the sort (see oracle.cpp): A fwd then A rev, both q30, after an A fwd of q35.  Stable
    ascending order, walked from the back: q35 (w_f 0, c 0), then the LATER q30 base = rev (w_r 0, c 1), then fwd (w_f 1, c 2)."""
    b = maq.beta
    fk = maq.fk
    n = 3
    want = fk[0] * b[35 << 16 | n << 8 | 0] + fk[0] * b[30 << 16 | n << 8 | 1] + fk[1] * b[30 << 16 | n << 8 | 2]
    other = fk[0] * b[35 << 16 | n << 8 | 0] + fk[1] * b[30 << 16 | n << 8 | 1] + fk[0] * b[30 << 16 | n << 8 | 2]
    assert abs(want - other) > 0.5                                 # the order is not a rounding matter
    s = maq.compute(b"AAA", [35, 30, 30], [0, 0, 1])
    assert s[GT("C", "C")] == pytest.approx(want, rel=1e-6)


def plain_restatement(maq, bases, quals, rev):
    """computeLikelihoods in straight double precision, written from maq.d:138-248 independently of oracle.cpp."""
    n = min(len(bases), 255)
    idx = sorted(range(n), key=lambda i: quals[i])                # stable
    w, c, fs, bs = {}, {}, {}, {}
    for i in reversed(idx):
        q = min(max(int(quals[i]), 4), 63)
        bb = chr(bases[i])
        key = (bb, int(rev[i]))
        f = maq.fk[w.get(key, 0)]
        fs[bb] = fs.get(bb, 0.0) + f
        bs[bb] = bs.get(bb, 0.0) + f * maq.beta[q << 16 | n << 8 | c.get(bb, 0)]
        c[bb] = c.get(bb, 0) + 1
        w[key] = w.get(key, 0) + 1
    out = {}
    nuc = "ACGT"
    for i, b1 in enumerate(nuc):
        rest = [x for x in nuc if x != b1]
        out[GT(b1, b1)] = sum(bs.get(x, 0.0) for x in rest) if sum(c.get(x, 0) for x in rest) > 0 else 0.0
        for b2 in nuc[i + 1:]:
            rest = [x for x in nuc if x not in (b1, b2)]
            lh = maq.lhet[(c.get(b1, 0) + c.get(b2, 0)) << 8 | c.get(b2, 0)]
            t1 = sum(bs.get(x, 0.0) for x in rest) if sum(c.get(x, 0) for x in rest) > 0 else 0.0
            out[GT(b2, b1)] = t1 - C10 * lh
    return {g: max(v, 0.0) for g, v in out.items()}


def test_random_columns_against_a_plain_restatement(maq):
    rng = np.random.default_rng(5)
    for trial in range(300):
        n = int(rng.integers(1, 60)) if trial % 10 else int(rng.integers(200, 300))
        bases = rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), n, p=[0.45, 0.3, 0.1, 0.1, 0.05]).astype(np.uint8)
        quals = rng.integers(0, 70, n).astype(np.uint8)
        rev = rng.integers(0, 2, n).astype(np.uint8)
        got = maq.compute(bases.tobytes(), quals, rev)
        want = plain_restatement(maq, bases, quals, rev)
        for g in range(25):
            if g in want:
                assert got[g] == pytest.approx(want[g], rel=2e-6, abs=2e-5), (trial, g)
            else:
                assert got[g] == FLT_MIN


def test_genotype_order_and_calls_on_a_fixture(maq):
    """GenotypeLikelihoodInfo / makeCall over ex1_header.bam: ten genotypes per column, sorted by score with equal scores
    in genotype-code order; the call's quality is the gap between the two best (maq.d:480-482)."""
    o = orc.Bam(fixture_bytes("ex1_header.bam")).decode()
    p = o.pileup_columns(True, use_md_tag=True, keep=True)
    r = o.maq(p, maq)
    assert len(r["gt0"]) == p.n_columns == 3037
    cov = np.diff(p.col_off)
    assert (r["n_valid"] <= cov).all() and r["n_valid"].max() > 20
    has = r["n_valid"] > 0
    assert (r["gt0"][has] != 255).all() and (r["gt0"][~has] == 255).all()
    sc = r["scores"]
    present = sc != FLT_MIN
    assert (present[has].sum(1) == 10).all() and (sc[present] >= 0).all()
    for c in np.flatnonzero(has)[::37]:
        order = sorted(np.flatnonzero(present[c]), key=lambda g: (sc[c, g], g))
        assert (r["gt0"][c], r["gt1"][c]) == (order[0], order[1])
        assert r["s0"][c] == sc[c, order[0]] and r["s1"][c] == sc[c, order[1]]
    # a column where a single kind of base was seen: its homozygote scores 0 and the three X|base heterozygotes tie;
    # the second best is the one with the smallest code
    c = next(c for c in np.flatnonzero(has) if (sc[c][present[c]] == 0).sum() == 1 and
             len({chr(x) for x, q in zip(p.base[int(p.col_off[c]):int(p.col_off[c + 1])], p.qual[int(p.col_off[c]):int(p.col_off[c + 1])]) if q >= 13 and x != ord("-")}) == 1)
    g0 = int(r["gt0"][c])
    assert g0 // 5 == g0 % 5
    ties = [g for g in np.flatnonzero(present[c]) if sc[c, g] == r["s1"][c]]
    assert len(ties) >= 2 and int(r["gt1"][c]) == min(ties)
