"""Host logic of the compact read lists (biodb_pileup_params.compact_reads): the numpy expansion used by the Python
mirror inverts a straightforward encoder of the documented format (include/biod_b200.h, biodb_column_batch)."""
import numpy as np

from biod_b200.bam import expand_compact_columns, expand_compact_reads


def encode(cols):
    """cols: list of ascending read-index lists -> (col_off, last, mask, soff, sidx) as the device writes them."""
    col_off, last, mask, soff, sidx = [0], [], [], [0], []
    for c in cols:
        col_off.append(col_off[-1] + len(c))
        if not c:
            last.append(0)
            mask.append(0)
            soff.append(soff[-1])
            continue
        la = c[-1]
        m = 0
        ns = 0
        for r in c:
            d = la - r
            if d >= 64:
                ns += 1
                sidx.append(r)
            else:
                m |= 1 << d
        last.append(la)
        mask.append(m)
        soff.append(soff[-1] + ns)
    return (np.array(col_off, dtype=np.uint64), np.array(last, dtype=np.uint32), np.array(mask, dtype=np.uint64),
            np.array(soff, dtype=np.uint32), np.array(sidx, dtype=np.uint32))


def test_expand_inverts_encode():
    rng = np.random.default_rng(5)
    cols = []
    base = 1000
    for k in range(400):
        base += int(rng.integers(0, 9))
        window = sorted(set(int(x) for x in base + rng.integers(0, 64, int(rng.integers(0, 40)))))
        old = sorted(set(int(x) for x in rng.integers(0, base - 64, int(rng.integers(0, 4))))) if k % 3 == 0 and window else []
        cols.append(old + window)
    cols.append([])
    cols.append([7])
    cols.append([2**32 - 70, 2**32 - 1])
    col_off, last, mask, soff, sidx = encode(cols)
    flat = np.array([r for c in cols for r in c], dtype=np.uint32)
    out = expand_compact_reads(col_off, last, mask, soff, sidx, len(flat))
    assert np.array_equal(out, flat)


def test_expand_empty():
    z = np.zeros(0, dtype=np.uint32)
    assert len(expand_compact_reads(np.zeros(1, dtype=np.uint64), z, np.zeros(0, dtype=np.uint64), np.zeros(1, dtype=np.uint32), z, 0)) == 0


def test_expand_columns_from_runs_and_stragglers():
    rng = np.random.default_rng(9)
    cols, positions = [], []
    base, pos = 5000, 100
    for k in range(300):
        if k % 97 == 50:
            pos += int(rng.integers(2, 1000))     # a gap of zero coverage: a new run of positions
        else:
            pos += 1
        positions.append(pos)
        base += int(rng.integers(0, 5))
        window = sorted(set(int(x) for x in base + rng.integers(0, 64, int(rng.integers(1, 30)))))
        old = sorted(set(int(x) for x in rng.integers(0, base - 64, int(rng.integers(0, 3))))) if k % 4 == 0 else []
        cols.append(old + window)
    col_off, last, mask, soff, sidx = encode(cols)
    scol = np.repeat(np.arange(len(cols), dtype=np.uint32), np.diff(soff).astype(np.int64))
    positions = np.array(positions, dtype=np.uint64)
    starts = np.flatnonzero(np.concatenate([[True], np.diff(positions.astype(np.int64)) != 1]))
    run_pos = positions[starts]
    run_first = np.concatenate([starts, [len(cols)]]).astype(np.uint32)
    p2, off2, idx2 = expand_compact_columns(len(cols), last, mask, scol, sidx, run_pos, run_first)
    assert np.array_equal(p2, positions)
    assert np.array_equal(off2, col_off)
    assert np.array_equal(idx2, np.array([r for c in cols for r in c], dtype=np.uint32))
