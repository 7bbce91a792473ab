"""Throughput and ratio of the device BGZF compressor (row N4, first part) on BAM bytes: the uncompressed stream of a
synthetic BAM (configs[1] shape) through biodb_bgzf_compress, host buffer to host buffer.  Prints one JSON line.

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
    a = ap.parse_args()
    from oracle import oracle as orc
    from tools import bamgen
    from biod_b200 import bgzf_compress
    bam = bamgen.generate(a.reads, 1, False, -1, bamgen.SEED_BASE + 2)
    u = bytes(orc.Bam(bam.tobytes()).decode().udata)
    bgzf_compress(u[:1 << 20])                                       # warm-up: context, allocations
    times = []
    for _ in range(3):
        t = time.perf_counter()
        s = bgzf_compress(u)
        times.append(time.perf_counter() - t)
    t0 = time.perf_counter()
    z = sum(len(zlib.compress(u[i:i + 0xFF00], 1)) for i in range(0, min(len(u), 200 * 0xFF00), 0xFF00))
    tz = time.perf_counter() - t0
    zn = min(len(u), 200 * 0xFF00)
    print(json.dumps({"metric": "bgzf_compress", "uncompressed_bytes": len(u), "compressed_bytes": len(s),
                      "ratio": len(s) / len(u), "seconds_best": min(times), "gb_per_s_in": len(u) / min(times) / 1e9,
                      "zlib_written_file_ratio": len(bam) / len(u),
                      "zlib_level1_one_core": {"ratio": z / zn, "gb_per_s_in": zn / tz / 1e9},
                      "what": "wall clock of biodb_bgzf_compress through the Python mirror (H2D, one thread per BGZF block: greedy "
                              "LZ77, dynamic or fixed Huffman codes, CRC32, pack, D2H; plus the mirror's own byte copies)"}))


if __name__ == "__main__":
    main()
