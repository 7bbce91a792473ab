"""GPU side of BGZF compression (row N4, first part; csrc/deflate.cu through biodb_bgzf_compress): the stream it writes
must read back — with zlib block by block (CRC32 and ISIZE of every footer checked), with the oracle's BGZF / BAM
reader and with this library's own inflate kernels — to exactly the bytes that went in: the reference's own criterion
(bgzf/outputstream.d:225-247, test/unittests.d:286-305)."""
import io
import struct
import zlib

import numpy as np
import pytest

from conftest import fixture_bytes
from oracle import oracle as orc

pytestmark = pytest.mark.gpu

EOF_BLOCK = bytes([31, 139, 8, 4, 0, 0, 0, 0, 0, 255, 6, 0, 66, 67, 2, 0, 27, 0, 3, 0, 0, 0, 0, 0, 0, 0, 0, 0])


def bgzf_read(stream):
    """Plain walk of a BGZF stream: returns (payload bytes, [isize per block])."""
    out, sizes, p = [], [], 0
    while p < len(stream):
        assert stream[p:p + 16] == EOF_BLOCK[:16], p                        # BLOCK_HEADER_START (constants.d:30-38)
        bsize = struct.unpack_from("<H", stream, p + 16)[0] + 1
        assert bsize <= 65536
        d = zlib.decompressobj(-15)
        data = d.decompress(stream[p + 18:p + bsize - 8]) + d.flush()
        assert d.eof and not d.unused_data
        crc, isize = struct.unpack_from("<II", stream, p + bsize - 8)
        assert isize == len(data) and crc == (zlib.crc32(data) & 0xFFFFFFFF)
        out.append(data)
        sizes.append(isize)
        p += bsize
    assert p == len(stream)
    return b"".join(out), sizes


def test_reference_round_trip_vector():
    # bgzf/outputstream.d:225-247
    from biod_b200 import bgzf_compress
    data = ("my very l" + "o" * 1000000 + "ng string").encode()
    for level in (-1, 0, 1):
        s = bgzf_compress(data, level)
        assert s.endswith(EOF_BLOCK)
        back, sizes = bgzf_read(s)
        assert back == data
        assert sizes[-1] == 0 and all(x == 0xFF00 for x in sizes[:-2]) and 0 < sizes[-2] <= 0xFF00
        assert (len(s) > len(data)) if level == 0 else (len(s) < len(data) // 20)


def test_sizes_and_content():
    from biod_b200 import BgzfOutputStream, bgzf_compress
    rng = np.random.default_rng(5)
    for n in (0, 1, 7, 0xFF00 - 1, 0xFF00, 0xFF00 + 1, 3 * 0xFF00, 300000):
        for kind in ("random", "text"):
            data = (rng.integers(0, 256, n, dtype=np.uint8).tobytes() if kind == "random"
                    else bytes(rng.choice(np.frombuffer(b"ACGT\n", dtype=np.uint8), n)))
            s = bgzf_compress(data)
            back, sizes = bgzf_read(s)
            assert back == data and len(sizes) == (n + 0xFF00 - 1) // 0xFF00 + 1
            assert bgzf_read(bgzf_compress(data, eof=False))[0] == data
    sink = io.BytesIO()
    out = BgzfOutputStream(sink, 1)
    out.write(b"abc" * 50000)
    out.flush()
    out.write(b"tail")
    out.close()
    assert bgzf_read(sink.getvalue())[0] == b"abc" * 50000 + b"tail" and sink.getvalue().endswith(EOF_BLOCK)
    with pytest.raises(ValueError):
        bgzf_compress(b"x", level=10)


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam"])
def test_bam_written_by_the_gpu_reads_back(name):
    # test/unittests.d:286-305 in spirit: what is written must read back as the same reads — through the oracle and
    # through this library's own inflate / record kernels
    from biod_b200 import BamReader, bgzf_compress
    o = orc.Bam(fixture_bytes(name)).decode()
    stream = bgzf_compress(bytes(o.udata))
    o2 = orc.Bam(stream).decode()
    assert o2.n_records == o.n_records and o2.header_text == o.header_text and o2.ref_names == o.ref_names
    assert bytes(o2.udata) == bytes(o.udata)
    raws = []
    for b in BamReader(stream).read_batches(copy=True):
        for i in range(b.n):
            p = int(b.rec_off[i]) + 4
            raws.append(b.data[p:p + int(b.block_size[i])].tobytes())
    assert raws == [o.record_bytes(i).tobytes() for i in range(o.n_records)]


def host_block(chunk, level=-1):
    """The host's statement of the encoder (deflate_enc.h: deflate_block_host) for one chunk."""
    from biod_b200 import _capi
    L = _capi.lib()
    a = np.frombuffer(bytes(chunk) + b"\0", dtype=np.uint8)
    out = np.zeros(len(chunk) + 64, dtype=np.uint8)
    n = int(L.biodb_debug_deflate_block(a.ctypes.data, len(chunk), out.ctypes.data, out.size, level))
    assert n > 0
    return out[:n].tobytes()


def test_device_bytes_equal_the_host_statement():
    # the warp's parse, code construction and bit packing against the plain loops the CPU suite checks with zlib: the
    # payload of every block, byte for byte
    from biod_b200 import bgzf_compress
    rng = np.random.default_rng(11)
    u = bytes(orc.Bam(fixture_bytes("ex1_header.bam")).decode().udata)
    text = (b"the quick brown fox jumps over the lazy dog " * 3000)
    streams = {
        "bam": u[:6 * 0xFF00 + 1234],
        "zeros": bytes(2 * 0xFF00 + 17),
        "random": rng.integers(0, 256, 0xFF00 + 999, dtype=np.uint8).tobytes(),
        "acgt": bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 2 * 0xFF00)),
        "text": text,
        "period3": b"abc" * 40000,
        "long_then_noise": b"x" * 70000 + rng.integers(0, 256, 5000, dtype=np.uint8).tobytes() + b"y" * 300,
        "far": (lambda unit: unit + unit + unit[:5000])(rng.integers(0, 256, 30000, dtype=np.uint8).tobytes()),
        "tiny": b"abcdefg",
        "short": b"hello hello hello hello",
        "hi_bytes": bytes(rng.integers(144, 256, 5000, dtype=np.uint8)),
    }
    for name, data in streams.items():
        for level in (-1, 0):
            s = bgzf_compress(data, level, eof=False)
            back, sizes = bgzf_read(s)
            assert back == data, name
            p = k = 0
            while p < len(s):
                bsize = struct.unpack_from("<H", s, p + 16)[0] + 1
                chunk = data[k * 0xFF00:(k + 1) * 0xFF00]
                assert s[p + 18:p + bsize - 8] == host_block(chunk, level), (name, level, k)
                p += bsize
                k += 1
