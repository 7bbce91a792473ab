"""The device's DEFLATE encoder (csrc/deflate_enc.h, row N4 first part) compiled for the host
(biodb_debug_deflate_block): whatever it writes must inflate back to the input with zlib — the round trip the
reference's own test asks of bgzfCompress (bgzf/outputstream.d:225-247).  No GPU involved."""
import zlib

import numpy as np
import pytest

from conftest import fixture_bytes
from oracle import oracle as orc


def enc(data, level=-1, cap=None):
    from biod_b200 import _capi
    L = _capi.lib()
    a = np.frombuffer(bytes(data) + b"\0", dtype=np.uint8)
    cap = len(data) + 64 if cap is None else cap
    out = np.zeros(max(cap, 1), dtype=np.uint8)
    n = int(L.biodb_debug_deflate_block(a.ctypes.data, len(data), out.ctypes.data, cap, level))
    assert 0 <= n <= cap
    return out[:n].tobytes()


def roundtrip(data, level=-1):
    raw = enc(data, level)
    assert raw, "the encoder must produce something"
    d = zlib.decompressobj(-15)
    back = d.decompress(raw) + d.flush()
    assert back == bytes(data)
    assert d.eof and not d.unused_data
    assert len(raw) <= len(data) + 5                 # never worse than a stored block
    return len(raw)


def test_small_and_edge_sizes():
    for n in list(range(0, 40)) + [255, 256, 257, 258, 259, 260, 1000, 0xFF00, 65535]:
        roundtrip(bytes((i * 7 + 3) & 0xFF for i in range(n)))
        roundtrip(b"a" * n)
        roundtrip(b"ab" * (n // 2))
    assert enc(b"x" * 65536) == b""                  # more than one stored block can hold: refused
    assert enc(b"x" * 100, cap=10) == b"" or True    # (a compressible input may fit; an incompressible one is refused below)
    rnd = np.random.default_rng(0).integers(0, 256, 1000, dtype=np.uint8).tobytes()
    assert enc(rnd, cap=1000) == b""                 # the stored form needs n + 5 bytes


def test_every_match_length_and_many_distances():
    rng = np.random.default_rng(1)
    for length in range(3, 259):
        seed = rng.integers(0, 256, length, dtype=np.uint8).tobytes()
        gap = rng.integers(0, 256, int(rng.integers(1, 50)), dtype=np.uint8).tobytes()
        roundtrip(seed + gap + seed + b"\x00\x01\x02")
    for dist in [1, 2, 3, 4, 5, 7, 8, 9, 16, 17, 31, 32, 33, 255, 256, 257, 1024, 4095, 4096, 4097, 8191, 16384, 24577, 32767,
                 32768, 32769, 40000]:
        seed = rng.integers(0, 256, 64, dtype=np.uint8).tobytes()
        fill = rng.integers(0, 256, max(0, dist - 64), dtype=np.uint8).tobytes()
        data = (seed + fill)[:dist] + seed
        roundtrip(data)


def test_levels_and_compression():
    text = ("my very l" + "o" * 60000 + "ng string").encode()       # the data of outputstream.d:229
    for level in (-1, 0, 1, 6, 9):
        n = roundtrip(text, level)
        assert (n == len(text) + 5) if level == 0 else (n < 1000)
    rnd = np.random.default_rng(2).integers(0, 256, 0xFF00, dtype=np.uint8).tobytes()
    assert roundtrip(rnd) == len(rnd) + 5                              # incompressible: stored
    hi = bytes(np.random.default_rng(3).integers(144, 256, 5000, dtype=np.uint8))   # 9-bit literals only
    roundtrip(hi)


@pytest.mark.parametrize("name", ["ex1_header.bam", "bins.bam", "mg1655_chunk.bam"])
def test_bam_payloads(name):
    # real BGZF payloads: the uncompressed stream of a fixture, cut the way BgzfOutputStream cuts it
    u = orc.Bam(fixture_bytes(name)).decode().udata
    total = comp = 0
    for i in range(0, min(len(u), 40 * 0xFF00), 0xFF00):
        chunk = bytes(u[i:i + 0xFF00])
        comp += roundtrip(chunk)
        total += len(chunk)
    assert comp < total
    if name != "mg1655_chunk.bam":
        assert comp < 0.6 * total                    # dynamic Huffman codes: close to what zlib makes of BAM bytes


def test_block_types_chosen():
    """BTYPE of the single block: stored for level 0 and incompressible input, dynamic for BAM bytes and text, fixed
    where the dynamic header would cost more than it saves."""
    u = orc.Bam(fixture_bytes("ex1_header.bam")).decode().udata
    btype = lambda raw: (raw[0] >> 1) & 3
    assert btype(enc(bytes(u[:0xFF00]))) == 2
    assert btype(enc(bytes(u[:0xFF00]), level=0)) == 0
    assert btype(enc(b"abcabcabcabc" * 3)) == 1
    assert btype(enc(np.random.default_rng(9).integers(0, 256, 4000, dtype=np.uint8).tobytes())) == 0
    roundtrip(bytes(u[:0xFF00]))
    # long codes: a geometric symbol distribution deep enough to need the 15-bit limit
    data = b"".join(bytes([k]) * max(1, 60000 >> k) for k in range(40))
    rng = np.random.default_rng(10)
    data = bytes(rng.permutation(np.frombuffer(data, dtype=np.uint8)))[:65000]
    assert btype(enc(data)) == 2
    roundtrip(data)


def test_seeded_fuzz():
    rng = np.random.default_rng(4)
    for _ in range(1500):
        n = int(rng.integers(0, 3000))
        alpha = int(rng.integers(1, 256))
        data = rng.integers(0, alpha, n, dtype=np.uint8).tobytes()
        if n > 20 and rng.random() < 0.5:                               # splice in repeats
            a, l = int(rng.integers(0, n - 10)), int(rng.integers(3, 400))
            data = data + data[a:a + l] * int(rng.integers(1, 4))
        roundtrip(data[:65535])
