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


def build_bai_biod(b, check_bins=False):
    """BioD's own IndexBuilder (bam/bai/indexing.d:56-351) restated read by read in plain Python — the checker of the
    product's builder (csrc/bai_build.h).  b: oracle.Bam (decoded).  Bins are written in ascending id order (the
    reference's order is that of a D associative array).  Test infrastructure only."""
    n_refs = len(b.ref_names)
    out = [b"BAI\1", struct.pack("<i", n_refs)]                         # :264-267
    size = 37449 - 4680 + 1                                              # :262
    lin, lin_len = [0] * size, 0
    prev, first = None, True
    no_coord, beg_vo, end_vo, unmapped, mapped = 0, (1 << 64) - 1, 0, 0, 0
    chunks, cur_beg = {}, 0

    def to_lin(p):                                                       # :51-53
        return min(0 if p < 0 else p // 16384, size - 1)

    def update_linear():                                                 # :133-164
        nonlocal lin_len
        beg = to_lin(prev["pos"])
        end = beg if prev["unm"] else to_lin(prev["pos"] + (prev["end"] - prev["pos"]) - 1)
        for i in range(beg, end + 1):
            if lin[i] == 0:
                lin[i] = prev["sv"]
        lin_len = max(lin_len, end + 1)

    def update_chunks():                                                 # :219-246
        nonlocal cur_beg
        cs = chunks.setdefault(prev["bin"], [])
        if not cs or (cs[-1][1] >> 16) != (cur_beg >> 16):
            cs.append([cur_beg, prev["ev"]])
        else:
            cs[-1][1] = prev["ev"]
        cur_beg = prev["ev"]

    def dump_reference():                                                # :186-216, :166-184
        nonlocal lin, lin_len, chunks, cur_beg, beg_vo, end_vo, unmapped, mapped
        out.append(struct.pack("<i", len(chunks) + 1))
        for bid in sorted(chunks):
            out.append(struct.pack("<Ii", bid, len(chunks[bid])))
            for c in chunks[bid]:
                out.append(struct.pack("<QQ", *c))
        out.append(struct.pack("<IiQQQQ", 37450, 2, beg_vo, end_vo, mapped, unmapped))
        out.append(struct.pack("<i", lin_len))
        last = 0
        for v in lin[:lin_len]:
            last = v = v or last
            out.append(struct.pack("<Q", v))
        lin, lin_len, chunks = [0] * size, 0, {}
        cur_beg = prev["ev"]
        beg_vo = end_vo = cur_beg
        unmapped = mapped = 0

    for i in range(b.n_records):                                         # put(), :281-322
        r = dict(ref=int(b.ref_id[i]), pos=int(b.pos[i]), end=int(b.end_pos[i]), bin=int(b.bin[i]),
                 unm=bool(int(b.flag[i]) & 4), sv=int(b.start_vo[i]), ev=int(b.end_vo[i]))
        if not first and r["ref"] != -1 and not prev["ref"] < r["ref"]:
            assert r["ref"] == prev["ref"] and r["pos"] >= prev["pos"], "BAM file is not coordinate-sorted"
        if r["ref"] >= 0 and r["pos"] >= 0:
            if first:
                prev, first, cur_beg = r, False, r["sv"]
                out.extend([struct.pack("<ii", 0, 0)] * r["ref"])
            else:
                if check_bins:
                    assert r["bin"] == reg2bin(r["pos"], r["end"]), "Bin is set incorrectly"
                if r["ref"] > prev["ref"]:
                    update_linear()
                    update_chunks()
                    dump_reference()
                    out.extend([struct.pack("<ii", 0, 0)] * (r["ref"] - prev["ref"] - 1))
                if r["ref"] == prev["ref"]:
                    update_linear()
                    if r["bin"] != prev["bin"]:
                        update_chunks()
                prev = r
        if r["ref"] == -1:                                               # updateMetadata at scope exit, :117-131
            no_coord += 1
        else:
            if r["unm"]:
                unmapped += 1
            else:
                mapped += 1
            if beg_vo == (1 << 64) - 1:
                beg_vo = r["sv"]
            end_vo = r["ev"]
    if not first:                                                        # finish(), :325-339
        update_linear()
        update_chunks()
        dump_reference()
    out.extend([struct.pack("<ii", 0, 0)] * (n_refs - ((prev["ref"] if prev else -1) + 1)))
    out.append(struct.pack("<Q", no_coord))
    return b"".join(out)
