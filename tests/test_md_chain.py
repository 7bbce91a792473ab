"""Host halves of the MD-tag reference bases (row N1), checked on the CPU against the oracle:
MdChain (biod_b200/csrc/md_chain.h, C ABI hook biodb_debug_md_chain): the read-by-read scan that says which read's dna()
supplies PileupColumn.reference_base where, against the oracle's column-by-column restatement of PileupRangeUsingMdTag
(pileup.d:522-654) on the reference's own vectors and on random pileups — in one go and drained batch by batch the way
the pipeline does;
DnaWalk (biod_b200/csrc/md_walk.h, hook biodb_debug_md_dna): the allocation-free generator of dna(read) that the CUDA
kernels run per thread, compiled for the host here, against the oracle's string-based dna_of_read on the fixtures, on
random reads and on malformed MD strings / tag areas."""
import ctypes as C
import os

import numpy as np
import pytest
from bamutil import bam_record, make_bam, tag_z

from oracle import oracle as orc


def chain_reference(b, skip_zero, single_ref, start_from=0, batch_reads=0):
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
    seg = np.zeros(4 * (8 * n + 8), dtype=np.int64)
    m = L.biodb_debug_md_chain(ref.ctypes.data, pos.ctypes.data, end.ctypes.data, ln.ctypes.data, n, int(skip_zero),
                               batch_reads, seg.ctypes.data, len(seg) // 4)
    assert 0 <= m <= len(seg) // 4
    out = {}
    settled = {}                     # per reference: columns below this position were handed out by an earlier drain
    for k in range(m):
        first, count, r, off = (int(x) for x in seg[4 * k:4 * k + 4])
        if count == -1:              # a drain marker (limit, -1, ref, 0)
            assert first >= settled.get(r, -1 << 62)
            settled[r] = first
            continue
        assert count > 0 and off >= 0 and off + count <= len(dna[r])
        assert first >= settled.get(int(ref[r]), -1 << 62), "a segment reaches below an earlier drain"
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
    want = "".join(chr(x) for x in p.ref_base)
    for batch_reads in (0, 1, 7, 64, 10**9):
        got = chain_reference(b, skip_zero, single_ref, start_from, batch_reads)
        mine = "".join(got.get((int(r), int(q)), "N") for r, q in zip(p.col_ref, p.col_pos))
        assert mine == want, batch_reads
    return p.n_columns


def test_reference_vectors():
    from test_oracle_golden import pileup_vector_bam
    for skip in (True, False):
        assert check(pileup_vector_bam(), skip, True) > 0
        assert check(pileup_vector_bam(), skip, True, 796, 849) > 0
        assert check(pileup_vector_bam(), skip, False) > 0


def random_pileup(rng, n_reads, refs=2, consistent=True, gap_p=0.02, dup_p=0.15, **bam_kw):
    """Reads with M / I / D / S / N operations and MD tags written against a random reference."""
    recs = []
    for rid in range(refs):
        genome = "".join("ACGT"[k] for k in rng.integers(0, 4, 6000 + 12 * n_reads))
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
    return make_bam([(f"c{i}", 100000) for i in range(refs)], recs, **bam_kw)


@pytest.mark.parametrize("seed", range(6))
@pytest.mark.parametrize("consistent", [True, False])
def test_random_pileups(seed, consistent):
    rng = np.random.default_rng(100 + seed)
    data = random_pileup(rng, 400, refs=1 + seed % 3, consistent=consistent, gap_p=0.03 if seed % 2 else 0.0)
    for skip in (True, False):
        assert check(data, skip, False) > 0
        assert check(data, skip, True) > 0
        assert check(data, skip, True, start_from=int(rng.integers(50, 600))) > 0


# ---- DnaWalk --------------------------------------------------------------------------------------------------------
DATA = os.path.join(os.path.dirname(__file__), "golden", "biod_test_data")


def walk_dna(body):
    from biod_b200 import _capi
    L = _capi.lib()
    a = np.frombuffer(bytes(body) + b"\0" * 8, dtype=np.uint8)
    out = np.zeros(1 << 16, dtype=np.uint8)
    n = L.biodb_debug_md_dna(a.ctypes.data, len(body), out.ctypes.data, len(out))
    assert 0 <= n <= len(out)
    return out[:n].tobytes().decode("latin1")


def check_dna(data, max_records=None):
    b = orc.Bam(data).decode()
    n = b.n_records if max_records is None else min(b.n_records, max_records)
    some = 0
    for i in range(n):
        want = b.dna(i)
        assert walk_dna(b.record_bytes(i)) == want, (i, b.tags_raw(i))
        some += bool(want)
    return some


@pytest.mark.parametrize("name", ["ex1_header.bam", "illu_20_chunk.bam", "tags.bam", "bins.bam", "mg1655_chunk.bam"])
def test_dna_walk_fixtures(name):
    path = os.path.join(DATA, name)
    if not os.path.exists(path):
        pytest.skip("fixture not present")
    with open(path, "rb") as f:
        check_dna(f.read(), 1500)


def test_dna_walk_reference_vectors():
    from test_oracle_golden import pileup_vector_bam
    assert check_dna(pileup_vector_bam()) > 0


@pytest.mark.parametrize("consistent", [True, False])
def test_dna_walk_random_reads(consistent):
    rng = np.random.default_rng(7)
    assert check_dna(random_pileup(rng, 600, refs=2, consistent=consistent)) > 0


def test_dna_walk_malformed_md_strings():
    """MD values no aligner writes: the bidirectional quirks of mdOperations (last operation parsed from the back),
    zero-length matches at either end, deletions without '^', every byte value as a mismatch character, huge counts."""
    rng = np.random.default_rng(11)
    alphabet = "0123456789^ACGTNacgtn=RYKMxz*"
    mds = ["", "0", "00", "10", "5A", "A", "^", "^A", "^AC", "0A0", "0^AC0", "3^AC", "^AC3", "AC", "ACGT", "3AC", "3^", "^3",
           "3^^A2", "2A^", "1^A^C1", "99999999999", "4294967296", "0A0C0G0T0", "5^acgt5", "=5", "5=", "3a2", "A^C", "^C^", "1A1^",
           "12^^", "0^0", "10^AC^GT10", "z", "*5", "5*", "1 2", "1\x80\xff2", "\x01"]
    for _ in range(400):
        mds.append("".join(alphabet[k] for k in rng.integers(0, len(alphabet), int(rng.integers(1, 12)))))
    for v in range(1, 256):
        if v not in (0,):
            mds.append("2" + chr(v) + "2")
    recs = []
    for k, md in enumerate(mds):
        cig = ["12M", "4S8M", "3M2I4M1D5M", "5M10N5M2S", "2=1X9M", "6M6D6M"][k % 6]
        lq = sum(l for l, o in [(int(x[:-1]), x[-1]) for x in __import__("re").findall(r"\d+[A-Z=]", cig)] if o in "MIS=X")
        seq = "".join("ACGT"[x] for x in rng.integers(0, 4, lq))
        tags = tag_z("XA", "q") + b"MDZ" + md.encode("latin1") + b"\0" + b"NMC\x01"
        recs.append(bam_record(f"m{k}", seq, cig, 10 + k, tags=tags))
    assert check_dna(make_bam([("c0", 100000)], recs)) > 0


def test_dna_walk_tag_areas():
    """The tag walk in front of MD: every value type, arrays, an MD that is not a string, truncated areas, SEQ '*'."""
    import struct
    arr = b"XBB" + b"s" + struct.pack("<I", 3) + struct.pack("<3h", 1, -2, 3)
    areas = [
        b"",
        b"MD",
        b"MDZ",
        b"MDZ5",
        b"MDZ5\0",
        b"XAA!" + b"XcC\x07" + b"XsS\x01\x02" + b"XiI\x01\x02\x03\x04" + b"Xff\0\0\x80\x3f" + arr + b"XHH1AE3\0" + b"MDZ3A2\0",
        b"MDi\x05\0\0\0" + b"MDZ6\0",
        b"MDH0A\0",
        b"XB" + b"B" + b"q" + struct.pack("<I", 1) + b"\0" + b"MDZ6\0",
        b"XBBi" + struct.pack("<I", 1000) + b"MDZ6\0",
        b"XZZunterminated",
        b"X?!\0MDZ6\0",
        b"MDBc" + struct.pack("<I", 2) + b"\x01\x02" + b"MDZ6\0",
        b"XAAqMDZ2^AC4\0XBA!",
    ]
    recs = []
    for k, t in enumerate(areas):
        recs.append(bam_record(f"t{k}", "ACGTAC", "6M", 5 + k, tags=t))
    recs.append(bam_record("star", "", "6M", 40, tags=b"MDZ6\0"))                    # SEQ '*' with a CIGAR
    recs.append(bam_record("long", "ACGTACGTAC", "10M", 41, tags=b"MDZ4^ACGT6A20\0"))   # MD longer than the read
    assert check_dna(make_bam([("c0", 100000)], recs)) > 0


def test_admit_many_equals_admit_on_large_inputs():
    """The tight loop of admit_many against the read-by-read path on 200 k reads per shape: identical segment lists."""
    from biod_b200 import _capi
    L = _capi.lib()
    rng = np.random.default_rng(77)
    shapes = [dict(p=0.2, span=(150, 151), slack=(0, 1)),            # the benchmark's shape: 150M, 30x
              dict(p=0.5, span=(20, 400), slack=(-30, 30)),          # ragged spans, dna() longer / shorter than the span
              dict(p=0.02, span=(5, 60), slack=(-5, 1)),             # sparse: islands and zero-coverage gaps
              dict(p=0.9, span=(1, 3000), slack=(0, 1))]             # long N-skip-like spans over many short reads
    for sh in shapes:
        n = 200_000
        ref = np.sort(rng.integers(0, 3, n)).astype(np.int32)
        pos = np.zeros(n, dtype=np.int64)
        for r in range(3):
            m = ref == r
            pos[m] = np.cumsum(rng.geometric(sh["p"], int(m.sum())) - 1)
        span = rng.integers(*sh["span"], n)
        end = pos + span
        ln = np.maximum(0, span + rng.integers(*sh["slack"], n)).astype(np.int64)
        for skip in (1, 0):
            res = []
            for batch in (0, 10**9):
                seg = np.zeros(4 * (4 * n + 64), dtype=np.int64)
                k = L.biodb_debug_md_chain(ref.ctypes.data, pos.ctypes.data, end.ctypes.data, ln.ctypes.data, n, skip, batch,
                                           seg.ctypes.data, len(seg) // 4)
                assert 0 < k <= len(seg) // 4
                res.append(seg[:4 * k].reshape(k, 4))
            assert np.array_equal(res[0], res[1]), (sh, skip)
            # reads that are not part of the pileup (end = INT32_MIN in the batch arrays) are passed over: the same
            # segments as the chain over the remaining reads alone, with their indices
            keep = rng.random(n) > 0.1
            end_m = np.where(keep, end, -2**31).astype(np.int64)
            seg = np.zeros(4 * (4 * n + 64), dtype=np.int64)
            k = L.biodb_debug_md_chain(ref.ctypes.data, pos.ctypes.data, end_m.ctypes.data, ln.ctypes.data, n, skip, 10**9,
                                       seg.ctypes.data, len(seg) // 4)
            a = seg[:4 * k].reshape(k, 4)
            idx = np.nonzero(keep)[0]
            r2, p2, e2, l2 = (np.ascontiguousarray(x[idx]) for x in (ref, pos, end, ln))
            seg2 = np.zeros(4 * (4 * n + 64), dtype=np.int64)
            k2 = L.biodb_debug_md_chain(r2.ctypes.data, p2.ctypes.data, e2.ctypes.data, l2.ctypes.data, len(idx), skip, 0,
                                        seg2.ctypes.data, len(seg2) // 4)
            b = seg2[:4 * k2].reshape(k2, 4).copy()
            b[:, 2] = idx[b[:, 2]]
            assert np.array_equal(a, b), (sh, skip, "markers")
