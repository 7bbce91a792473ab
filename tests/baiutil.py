"""Test helper: a BAI index for a coordinate-sorted BAM, built from the oracle's record table the way the format
asks (SAM spec §5.2; the reference's own builder, bai/indexing.d, is row N4 and not part of this repository):
per reference the bins with their chunks (adjacent records of a bin share a chunk) and the 16 kbp linear index.
Test infrastructure only."""
import struct

from bamutil import reg2bin


def build_bai(b):
    """b: oracle.Bam (decoded).  Returns the bytes of a .bai file."""
    n_ref = len(b.ref_names)
    bins = [dict() for _ in range(n_ref)]
    lin = [dict() for _ in range(n_ref)]
    for i in range(b.n_records):
        r = int(b.ref_id[i])
        if r < 0 or r >= n_ref:
            continue
        pos, end = int(b.pos[i]), int(b.end_pos[i])
        if end <= pos:
            end = pos + 1
        sv, ev = int(b.start_vo[i]), int(b.end_vo[i])
        ch = bins[r].setdefault(reg2bin(pos, end), [])
        if ch and ch[-1][1] == sv:
            ch[-1][1] = ev
        else:
            ch.append([sv, ev])
        for w in range(pos >> 14, ((end - 1) >> 14) + 1):
            lin[r].setdefault(w, sv)
    out = [b"BAI\1", struct.pack("<i", n_ref)]
    for r in range(n_ref):
        out.append(struct.pack("<i", len(bins[r])))
        for bid in sorted(bins[r]):
            out.append(struct.pack("<Ii", bid, len(bins[r][bid])))
            for sv, ev in bins[r][bid]:
                out.append(struct.pack("<QQ", sv, ev))
        n_intv = (max(lin[r]) + 1) if lin[r] else 0
        out.append(struct.pack("<i", n_intv))
        last = 0
        for w in range(n_intv):
            last = lin[r].get(w, last)
            out.append(struct.pack("<Q", last))
    return b"".join(out)
