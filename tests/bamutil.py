"""Test helpers: a tiny pure-Python BAM/BGZF writer and aux-tag -> SAM text rendering.

The writer follows the on-disk conventions of the reference writer (bam/writer.d:203-268,
bgzf/compress.d:43-103, bgzf/constants.d:28-61) so that synthetic inputs look like files BioD
itself would have produced.  Test infrastructure only.
"""
import struct
import zlib

BGZF_EOF = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])
BGZF_BLOCK_SIZE = 0xFF00
CIGAR_OPS = "MIDNSHP=X"
SEQ_CODES = "=ACMGRSVTWYHKDBN"


def bgzf_block(payload: bytes, level=-1, raw_deflate=None) -> bytes:
    """One BGZF member (header 18 B + deflate + CRC32 + ISIZE)."""
    if raw_deflate is None:
        co = zlib.compressobj(level, zlib.DEFLATED, -15, 8)
        raw_deflate = co.compress(payload) + co.flush()
    bsize = len(raw_deflate) + 25
    assert bsize < 65536, "payload does not fit one BGZF block"
    hdr = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0]) + struct.pack("<H", bsize)
    return hdr + raw_deflate + struct.pack("<II", zlib.crc32(payload) & 0xFFFFFFFF, len(payload))


def bgzf_compress(data: bytes, level=-1, block_size=BGZF_BLOCK_SIZE, eof=True) -> bytes:
    out = []
    for i in range(0, len(data), block_size):
        out.append(bgzf_block(data[i:i + block_size], level))
    if eof:
        out.append(BGZF_EOF)
    return b"".join(out)


def reg2bin(beg, end):
    """bam/bai/bin.d:82-92."""
    if end == beg:
        end = beg + 1
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def parse_cigar(s):
    ops, n = [], ""
    for ch in s:
        if ch.isdigit():
            n += ch
        else:
            ops.append((int(n), ch))
            n = ""
    return ops


def bam_record(name, seq, cigar, pos, ref_id=0, flag=0, mapq=60, qual=None, tags=b"", next_ref=-1, next_pos=-1,
               tlen=0):
    """Serialise one alignment (with its 4-byte block_size prefix)."""
    if isinstance(cigar, str):
        cigar = parse_cigar(cigar)
    ref_span = sum(l for l, o in cigar if o in "MDN=X")
    if flag & 4:
        ref_span = 0
    nm = name.encode() + b"\0"
    cig = b"".join(struct.pack("<I", (l << 4) | CIGAR_OPS.index(o)) for l, o in cigar)
    codes = [SEQ_CODES.index(c) for c in seq]
    if len(codes) & 1:
        codes.append(0)
    packed = bytes((codes[i] << 4) | codes[i + 1] for i in range(0, len(codes), 2))
    if qual is None:
        qual = bytes([255] * len(seq))
    assert len(qual) == len(seq)
    b = reg2bin(pos, pos + ref_span) if pos >= 0 else 4680
    core = struct.pack("<iiIIiiii", ref_id, pos, (b << 16) | (mapq << 8) | len(nm), (flag << 16) | len(cigar),
                       len(seq), next_ref, next_pos, tlen)
    body = core + nm + cig + packed + bytes(qual) + tags
    return struct.pack("<i", len(body)) + body


def tag_z(key, value):
    return key.encode() + b"Z" + value.encode() + b"\0"


def bam_header(text, refs):
    h = b"BAM\1" + struct.pack("<i", len(text)) + text.encode()
    h += struct.pack("<i", len(refs))
    for name, ln in refs:
        nm = name.encode() + b"\0"
        h += struct.pack("<i", len(nm)) + nm + struct.pack("<i", ln)
    return h


def make_bam(refs, records, text=None, level=-1, straddle=False, block_size=BGZF_BLOCK_SIZE, eof=True):
    """Build a BAM file.  `records` are bytes from bam_record().

    straddle=False follows BamWriter.writeRecord (bam/writer.d:259-267): a record that would not
    fit in the current block starts a new one; the header gets its own block(s).
    straddle=True cuts the stream every `block_size` bytes wherever that falls.
    """
    if text is None:
        text = "@HD\tVN:1.6\tSO:coordinate\n" + "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in refs)
    hdr = bam_header(text, refs)
    if straddle:
        return bgzf_compress(hdr + b"".join(records), level, block_size, eof)
    out = [bgzf_compress(hdr, level, block_size, eof=False)]
    cur = bytearray()
    for r in records:
        if len(cur) + len(r) > block_size and cur:
            out.append(bgzf_block(bytes(cur), level))
            cur = bytearray()
        cur += r
        while len(cur) > block_size:            # oversized record: split across blocks
            out.append(bgzf_block(bytes(cur[:block_size]), level))
            cur = cur[block_size:]
    if cur:
        out.append(bgzf_block(bytes(cur), level))
    if eof:
        out.append(BGZF_EOF)
    return b"".join(out)


# ---------------------------------------------------------------------------
# aux tags -> SAM text (for the ex1_header.sam golden)
_SIZES = {"c": ("<b", 1), "C": ("<B", 1), "s": ("<h", 2), "S": ("<H", 2), "i": ("<i", 4), "I": ("<I", 4),
          "f": ("<f", 4)}


def tags_to_sam(raw: bytes):
    out, p = [], 0
    while p < len(raw):
        key = raw[p:p + 2].decode()
        t = chr(raw[p + 2])
        p += 3
        if t == "A":
            out.append(f"{key}:A:{chr(raw[p])}")
            p += 1
        elif t in "cCsSiI":
            fmt, sz = _SIZES[t]
            out.append(f"{key}:i:{struct.unpack_from(fmt, raw, p)[0]}")
            p += sz
        elif t == "f":
            out.append(f"{key}:f:{struct.unpack_from('<f', raw, p)[0]:g}")
            p += 4
        elif t in "ZH":
            e = raw.index(b"\0", p)
            out.append(f"{key}:{t}:{raw[p:e].decode()}")
            p = e + 1
        elif t == "B":
            st = chr(raw[p])
            n = struct.unpack_from("<i", raw, p + 1)[0]
            p += 5
            fmt, sz = _SIZES[st]
            vals = [struct.unpack_from(fmt, raw, p + k * sz)[0] for k in range(n)]
            p += n * sz
            out.append(f"{key}:B:{st}," + ",".join(f"{v:g}" if st == "f" else str(v) for v in vals))
        else:
            raise ValueError(f"bad tag type {t!r}")
    return out
