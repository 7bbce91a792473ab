"""Design aid (not product, not oracle): token statistics of the DEFLATE streams in a BGZF file and a simulation of the
lane-parallel speculative decode used by inflate_decode_kernel (inflate_tok.cu) (how many chain-repair rounds a 32-lane super-chunk needs
for a given sub-sequence size).  RFC 1951 decoder written from the RFC."""
import struct
import sys
from collections import Counter

import numpy as np

LEN_BASE = [3, 4, 5, 6, 7, 8, 9, 10, 11, 13, 15, 17, 19, 23, 27, 31, 35, 43, 51, 59, 67, 83, 99, 115, 131, 163, 195, 227, 258]
LEN_EB = [0, 0, 0, 0, 0, 0, 0, 0, 1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3, 4, 4, 4, 4, 5, 5, 5, 5, 0]
DIST_BASE = [1, 2, 3, 4, 5, 7, 9, 13, 17, 25, 33, 49, 65, 97, 129, 193, 257, 385, 513, 769, 1025, 1537, 2049, 3073, 4097, 6145,
             8193, 12289, 16385, 24577]
DIST_EB = [0, 0, 0, 0, 1, 1, 2, 2, 3, 3, 4, 4, 5, 5, 6, 6, 7, 7, 8, 8, 9, 9, 10, 10, 11, 11, 12, 12, 13, 13]
ORDER = [16, 17, 18, 0, 8, 7, 9, 6, 10, 5, 11, 4, 12, 3, 13, 2, 14, 1, 15]


def build(lens):
    """canonical code -> dict {(len, code_msb_first): sym}; and a 15-bit reversed LUT for speed"""
    maxl = max(lens) if len(lens) else 0
    cnt = [0] * 16
    for l in lens:
        cnt[l] += 1
    cnt[0] = 0
    nxt = [0] * 16
    c = 0
    for l in range(1, 16):
        c = (c + cnt[l - 1]) << 1
        nxt[l] = c
    lut = {}
    for s, l in enumerate(lens):
        if l:
            code = nxt[l]
            nxt[l] += 1
            rev = int(format(code, "0%db" % l)[::-1], 2)
            lut[(l, rev)] = s
    return lut, maxl


class Bits:
    def __init__(self, data):
        self.v = int.from_bytes(data, "little")
        self.n = len(data) * 8

    def get(self, pos, n):
        return (self.v >> pos) & ((1 << n) - 1)


def dec_sym(b, pos, table):
    lut, maxl = table
    for l in range(1, maxl + 1):
        s = lut.get((l, b.get(pos, l)))
        if s is not None:
            return s, l
    return None, 0


def token_at(b, pos, tl, td):
    """decode one token at bit pos: returns (kind, nbits, outlen, dist) kind: 0 lit 1 match 2 eob 3 invalid"""
    s, l = dec_sym(b, pos, tl)
    if s is None:
        return 3, 1, 0, 0
    if s < 256:
        return 0, l, 1, 0
    if s == 256:
        return 2, l, 0, 0
    if s > 285:
        return 3, l, 0, 0
    p = pos + l
    ln = LEN_BASE[s - 257] + b.get(p, LEN_EB[s - 257])
    p += LEN_EB[s - 257]
    d, l2 = dec_sym(b, p, td)
    if d is None or d > 29:
        return 3, p - pos + 1, 0, 0
    p += l2
    dist = DIST_BASE[d] + b.get(p, DIST_EB[d])
    p += DIST_EB[d]
    return 1, p - pos, ln, dist


def parse_block_header(b, pos):
    last = b.get(pos, 1)
    bt = b.get(pos + 1, 2)
    pos += 3
    if bt == 0:
        pos = (pos + 7) & ~7
        ln = b.get(pos, 16)
        return last, bt, pos + 32, ln, None
    if bt == 1:
        ll = [8] * 144 + [9] * 112 + [7] * 24 + [8] * 8
        return last, bt, pos, build(ll), build([5] * 32)
    hlit = b.get(pos, 5) + 257
    hdist = b.get(pos + 5, 5) + 1
    hclen = b.get(pos + 10, 4) + 4
    pos += 14
    cl = [0] * 19
    for i in range(hclen):
        cl[ORDER[i]] = b.get(pos, 3)
        pos += 3
    tc = build(cl)
    lens = []
    while len(lens) < hlit + hdist:
        s, l = dec_sym(b, pos, tc)
        pos += l
        if s < 16:
            lens.append(s)
        elif s == 16:
            lens += [lens[-1]] * (3 + b.get(pos, 2))
            pos += 2
        elif s == 17:
            lens += [0] * (3 + b.get(pos, 3))
            pos += 3
        else:
            lens += [0] * (11 + b.get(pos, 7))
            pos += 7
    return last, bt, pos, build(lens[:hlit]), build(lens[hlit:])


def sim_block(payload, S, stats):
    b = Bits(payload + b"\0" * 16)
    pos = 0
    opos = 0
    while True:
        h0 = pos
        last, bt, pos, tl, td = parse_block_header(b, pos)
        stats["hdr_bits"] += pos - h0
        stats["deflate_blocks"] += 1
        if bt == 0:
            pos += tl * 8
            opos += tl
            if last:
                break
            continue
        # true token chain
        eob = False
        while not eob:
            base = pos
            starts = [base + i * S for i in range(33)]
            # per-lane decode from a start: returns end, ntok, eob
            def run(t, limit):
                p = t
                n = 0
                while p < limit:
                    k, nb, ol, d = token_at(b, p, tl, td)
                    if k == 3:
                        return p, n, 3
                    p += nb
                    n += 1
                    if k == 2:
                        return p, n, 2
                return p, n, 0
            t = [starts[i] for i in range(32)]
            res = [run(t[i], starts[i + 1]) for i in range(32)]
            rounds = 1
            while True:
                need = []
                for i in range(1, 32):
                    tn = res[i - 1][0]
                    if tn != t[i]:
                        need.append((i, tn))
                if not need:
                    break
                for i, tn in need:
                    t[i] = tn
                    res[i] = run(tn, starts[i + 1])
                rounds += 1
            stats["rounds"][rounds] += 1
            stats["chunks"] += 1
            # commit up to first eob lane
            k = 32
            for i in range(32):
                if res[i][2] == 2:
                    k = i + 1
                    eob = True
                    break
                assert res[i][2] == 0, "invalid token on the true chain"
            # walk true tokens of committed lanes for stats
            p = t[0]
            endp = res[k - 1][0]
            while p < endp:
                kk, nb, ol, d = token_at(b, p, tl, td)
                if kk == 0:
                    stats["lit"] += 1
                    stats["lit_bits"] += nb
                elif kk == 1:
                    stats["match"] += 1
                    stats["match_bits"] += nb
                    stats["match_bytes"] += ol
                    stats["dist_hist"][min(d // 1024, 32)] += 1
                    stats["len_hist"][min(ol // 8, 33)] += 1
                opos += ol
                p += nb
            pos = endp
        if last:
            break
    return opos


def main():
    path = sys.argv[1]
    S = int(sys.argv[2]) if len(sys.argv) > 2 else 128
    nblk = int(sys.argv[3]) if len(sys.argv) > 3 else 4
    skip = int(sys.argv[4]) if len(sys.argv) > 4 else 2
    data = open(path, "rb").read()
    off = 0
    stats = dict(hdr_bits=0, deflate_blocks=0, rounds=Counter(), chunks=0, lit=0, lit_bits=0, match=0, match_bits=0,
                 match_bytes=0, dist_hist=Counter(), len_hist=Counter())
    done = 0
    i = 0
    tot_out = 0
    tot_in = 0
    while off < len(data) and done < nblk:
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        isize = struct.unpack_from("<I", data, off + bsize - 4)[0]
        if i >= skip and isize:
            out = sim_block(data[off + 18: off + bsize - 8], S, stats)
            assert out == isize, (out, isize)
            tot_out += isize
            tot_in += bsize - 26
            done += 1
        off += bsize
        i += 1
    print("S", S, "blocks", done, "in", tot_in, "out", tot_out)
    print("deflate blocks", stats["deflate_blocks"], "hdr bits avg", stats["hdr_bits"] / max(1, stats["deflate_blocks"]))
    print("literals", stats["lit"], "avg bits", stats["lit_bits"] / max(1, stats["lit"]))
    print("matches", stats["match"], "avg bits", stats["match_bits"] / max(1, stats["match"]), "avg len",
          stats["match_bytes"] / max(1, stats["match"]))
    print("chunks", stats["chunks"], "rounds", sorted(stats["rounds"].items()))
    print("dist hist (KiB)", sorted(stats["dist_hist"].items()))
    print("len hist (/8)", sorted(stats["len_hist"].items()))


if __name__ == "__main__":
    main()
