"""MdChain (biod_b200/csrc/md_chain.h, C ABI hook biodb_debug_md_chain — host code, no GPU): the read-by-read scan that
says which read's dna() supplies PileupColumn.reference_base where, checked against the oracle's column-by-column
restatement of PileupRangeUsingMdTag (pileup.d:522-654) on the reference's own vectors and on random pileups."""
import ctypes as C

import numpy as np
import pytest
from bamutil import bam_record, make_bam, tag_z

from oracle import oracle as orc


def chain_reference(b, skip_zero, single_ref, start_from=0):
    """reference_base per (ref, position) from MdChain + the oracle's dna(read) strings."""
    from biod_b200 import _capi
    L = _capi.lib()
    keep = [i for i in range(b.n_records) if b.end_pos[i] - b.pos[i] > 0]      # pileup.d:481
    if single_ref:
        k = 0
        while k < len(keep) and b.end_pos[keep[k]] < start_from:                # pileup.d:482-489
            k += 1
        keep = keep[k:]
        if keep:
            keep = [i for i in keep if b.ref_id[i] == b.ref_id[keep[0]]][:len(keep)]
            first_ref = b.ref_id[keep[0]]
            cut = next((j for j, i in enumerate(keep) if b.ref_id[i] != first_ref), len(keep))
            keep = keep[:cut]
    dna = [b.dna(i) for i in keep]
    n = len(keep)
    ref = np.array([b.ref_id[i] for i in keep], dtype=np.int32)
    pos = np.array([b.pos[i] for i in keep], dtype=np.int64)
    end = np.array([b.end_pos[i] for i in keep], dtype=np.int64)
    ln = np.array([len(d) for d in dna], dtype=np.int64)
    seg = np.zeros(4 * (4 * n + 8), dtype=np.int64)
    m = L.biodb_debug_md_chain(ref.ctypes.data, pos.ctypes.data, end.ctypes.data, ln.ctypes.data, n, int(skip_zero),
                               seg.ctypes.data, len(seg) // 4)
    assert 0 <= m <= len(seg) // 4
    out = {}
    for k in range(m):
        first, count, r, off = (int(x) for x in seg[4 * k:4 * k + 4])
        assert count > 0 and off >= 0 and off + count <= len(dna[r])
        for q in range(count):
            key = (int(ref[r]), first + q)
            assert key not in out, "segments overlap"
            out[key] = dna[r][off + q]
    return out


def check(data, skip_zero, single_ref, start_from=0, end_at=2**64 - 1):
    b = orc.Bam(data).decode()
    p = (b.make_pileup(start_from, end_at, skip_zero, use_md_tag=True) if single_ref
         else b.pileup_columns(skip_zero, use_md_tag=True))
    assert p.status == 0
    got = chain_reference(b, skip_zero, single_ref, start_from)
    want = "".join(chr(x) for x in p.ref_base)
    mine = "".join(got.get((int(r), int(q)), "N") for r, q in zip(p.col_ref, p.col_pos))
    assert mine == want
    return p.n_columns


def test_reference_vectors():
    from test_oracle_golden import pileup_vector_bam
    for skip in (True, False):
        assert check(pileup_vector_bam(), skip, True) > 0
        assert check(pileup_vector_bam(), skip, True, 796, 849) > 0
        assert check(pileup_vector_bam(), skip, False) > 0


def random_pileup(rng, n_reads, refs=2, consistent=True, gap_p=0.02, dup_p=0.15):
    """Reads with M / I / D / S / N operations and MD tags written against a random reference."""
    recs = []
    for rid in range(refs):
        genome = "".join("ACGT"[k] for k in rng.integers(0, 4, 6000))
        pos = int(rng.integers(0, 30))
        for k in range(n_reads // refs):
            if rng.random() > dup_p:
                pos += int(rng.integers(1, 9))
            if rng.random() < gap_p:
                pos += int(rng.integers(40, 200))
            L = int(rng.integers(8, 60))
            kind = int(rng.integers(0, 6))
            a = int(rng.integers(2, L - 2))
            if kind == 0:
                ops = [(L, "M")]
            elif kind == 1:
                ops = [(a, "M"), (int(rng.integers(1, 4)), "I"), (L - a, "M")]
            elif kind == 2:
                ops = [(a, "M"), (int(rng.integers(1, 6)), "D"), (L - a, "M")]
            elif kind == 3:
                ops = [(int(rng.integers(1, 5)), "S"), (L, "M")]
            elif kind == 4:
                ops = [(a, "M"), (int(rng.integers(5, 40)), "N"), (L - a, "M")]
            else:
                ops = [(a, "="), (1, "X"), (L - a, "M"), (2, "S")]
            seq, md, run, g = [], [], 0, pos
            for n, o in ops:
                if o in "M=X":
                    for _ in range(n):
                        base = genome[g]
                        if rng.random() < 0.06:                       # a mismatch
                            alt = "ACGT"[(("ACGT".index(base)) + int(rng.integers(1, 4))) % 4]
                            seq.append(alt)
                            md.append(str(run) + base)
                            run = 0
                        else:
                            seq.append(base)
                            run += 1
                        g += 1
                elif o in "IS":
                    seq += ["ACGT"[x] for x in rng.integers(0, 4, n)]
                elif o == "D":
                    md.append(str(run) + "^" + genome[g:g + n])
                    run = 0
                    g += n
                elif o == "N":
                    g += n
            md.append(str(run))
            mds = "".join(md)
            tags = tag_z("XA", "q") + tag_z("MD", mds)
            if not consistent:
                r = rng.random()
                if r < 0.15:
                    tags = tag_z("XA", "q")                           # no MD tag at all
                elif r < 0.3:
                    tags = tag_z("MD", mds[:max(1, len(mds) // 2)])   # cut short
                elif r < 0.4:
                    tags = tag_z("MD", mds + "A7")                    # longer than the read
            cig = "".join(f"{n}{o}" for n, o in ops)
            recs.append(bam_record(f"r{rid}_{k}", "".join(seq), cig, pos, ref_id=rid, tags=tags))
    return make_bam([(f"c{i}", 100000) for i in range(refs)], recs)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("consistent", [True, False])
def test_random_pileups(seed, consistent):
    rng = np.random.default_rng(100 + seed)
    data = random_pileup(rng, 400, refs=1 + seed % 3, consistent=consistent, gap_p=0.03 if seed % 2 else 0.0)
    for skip in (True, False):
        assert check(data, skip, False) > 0
        assert check(data, skip, True) > 0
        assert check(data, skip, True, start_from=int(rng.integers(50, 600))) > 0
