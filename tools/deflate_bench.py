"""Throughput and ratio of the device BGZF compressor (row N4, first part) on BAM bytes: the uncompressed stream of a
synthetic BAM (configs[1] shape) through biodb_bgzf_compress, host buffer to host buffer, beside zlib at the reference's
default level on a thread pool (what BgzfOutputStream does on the CPU).  Prints one JSON line.

    python tools/deflate_bench.py [--reads 400000]
"""
import argparse
import json
import os
import sys
import time
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=400_000)
    ap.add_argument("--zlib-threads", type=int, default=15)
    a = ap.parse_args()
    import ctypes as C
    from concurrent.futures import ThreadPoolExecutor

    import numpy as np
    from oracle import oracle as orc
    from tools import bamgen
    from biod_b200 import _capi, bgzf_compress
    L = _capi.lib()
    bam = bamgen.generate(a.reads, 1, False, -1, bamgen.SEED_BASE + 2)
    u = bytes(orc.Bam(bam.tobytes()).decode().udata)
    src = np.frombuffer(u, dtype=np.uint8)
    cap = int(L.biodb_bgzf_compress_bound(src.size))
    dst = np.zeros(cap, dtype=np.uint8)                              # the caller's buffers, touched before the clock starts
    n = C.c_size_t()

    def call():
        st = L.biodb_bgzf_compress(-1, src.ctypes.data, src.size, -1, 1, dst.ctypes.data, cap, C.byref(n))
        assert st == 0, st
    call()                                                           # warm-up: context, pinned and device buffers
    times = []
    for _ in range(5):
        t = time.perf_counter()
        call()
        times.append(time.perf_counter() - t)
    stats = (C.c_uint64 * 4)()
    L.biodb_debug_deflate_stats.argtypes = [C.c_int32, C.c_void_p]
    L.biodb_debug_deflate_stats(-1, stats)
    first_slab = min(src.size, 2048 * 0xFF00)
    s = dst[:n.value].tobytes()
    # the stream must read back (zlib, block by block)
    p, back = 0, 0
    import struct
    while p < len(s):
        bsize = struct.unpack_from("<H", s, p + 16)[0] + 1
        d = zlib.decompressobj(-15)
        blk = d.decompress(s[p + 18:p + bsize - 8]) + d.flush()
        assert blk == u[back:back + len(blk)]
        back += len(blk)
        p += bsize
    assert back == len(u)
    tm = []
    for _ in range(2):
        t = time.perf_counter()
        bgzf_compress(u)
        tm.append(time.perf_counter() - t)
    # the reference's CPU path: task!bgzfCompress on a pool (bgzf/outputstream.d:136-173) = zlib level -1 per chunk
    chunks = [u[i:i + 0xFF00] for i in range(0, len(u), 0xFF00)]

    def zl(c):
        co = zlib.compressobj(-1, zlib.DEFLATED, -15)
        return len(co.compress(c) + co.flush()) + 26
    with ThreadPoolExecutor(a.zlib_threads) as ex:
        t0 = time.perf_counter()
        z = sum(ex.map(zl, chunks))
        tz = time.perf_counter() - t0
    t0 = time.perf_counter()
    z1 = sum(zl(c) for c in chunks[:200])
    tz1 = time.perf_counter() - t0
    zn1 = sum(len(c) for c in chunks[:200])
    print(json.dumps({"metric": "bgzf_compress", "uncompressed_bytes": len(u), "compressed_bytes": len(s),
                      "ratio": len(s) / len(u), "seconds_best": min(times), "gb_per_s_in": len(u) / min(times) / 1e9,
                      "seconds_all": times,
                      "encoder_kernel": {"first_slab_bytes": first_slab, "us": int(stats[0]), "grid_ctas": int(stats[1]),
                                         "gb_per_s_in": first_slab / max(1, int(stats[0])) / 1e3},
                      "python_mirror_gb_per_s_in": len(u) / min(tm) / 1e9,
                      "zlib_written_file_ratio": len(bam) / len(u),
                      "zlib_default_level": {"threads": a.zlib_threads, "ratio": z / len(u), "gb_per_s_in": len(u) / tz / 1e9,
                                             "one_thread_gb_per_s_in": zn1 / tz1 / 1e9, "cpus": os.cpu_count()},
                      "what": "wall clock of biodb_bgzf_compress through the C ABI, caller's (pageable) buffer to caller's buffer: "
                              "copy to pinned memory, H2D, one warp per BGZF block (window-parallel LZ77, dynamic Huffman codes), "
                              "CRC32, pack, D2H, copy out; slabs of 2048 blocks on two streams"}))


if __name__ == "__main__":
    main()
