"""Latency / throughput of BAI region reads (row N2) on a synthetic coordinate-sorted BAM: random regions of a given
width through reader.region_batches (index chunks -> inflate -> record scan -> device-side filter -> host).
Prints one JSON line.  The index is built by tests/baiutil.py from the oracle's record table (input preparation only;
nothing of it is inside the timed region).

    python tools/region_bench.py [--reads 400000] [--width 10000] [--queries 300]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reads", type=int, default=400_000)
    ap.add_argument("--width", type=int, default=10_000)
    ap.add_argument("--queries", type=int, default=300)
    ap.add_argument("--check", type=int, default=20, help="queries verified against the oracle (outside the timed region)")
    a = ap.parse_args()
    from baiutil import build_bai
    from oracle import oracle as orc
    from tools import bamgen
    data = bamgen.generate(a.reads, 1, False, -1, bamgen.SEED_BASE + 2)
    o = orc.Bam(data.tobytes()).decode()
    raw = build_bai(o)
    from biod_b200 import BamReader
    rd = BamReader(data, want_offsets=True, index=raw)
    span = int(o.end_pos.max())
    rng = np.random.default_rng(1)
    starts = rng.integers(0, max(1, span - a.width), a.queries)
    bai = orc.Bai(raw)
    for s in starts[:a.check]:
        got = sum(b.n for b in rd.region_batches(0, int(s), int(s) + a.width))
        assert got == len(orc.region_reads(o, bai, 0, int(s), int(s) + a.width)[0])
    lat, n_reads = [], 0
    t0 = time.perf_counter()
    for s in starts:
        t = time.perf_counter()
        for b in rd.region_batches(0, int(s), int(s) + a.width):
            n_reads += b.n
        lat.append(time.perf_counter() - t)
    wall = time.perf_counter() - t0
    lat = np.array(lat) * 1e3
    print(json.dumps({"metric": "bai_region_reads", "reads_in_file": a.reads, "region_width": a.width, "queries": a.queries,
                      "queries_per_sec": a.queries / wall, "reads_per_query": n_reads / a.queries,
                      "latency_ms": {"median": float(np.median(lat)), "p95": float(np.percentile(lat, 95)), "min": float(lat.min())},
                      "verified_against_oracle": a.check,
                      "what": "wall clock per query through the Python mirror of the C ABI: getChunks on the host, one pass per "
                              "chunk (H2D of the chunk's blocks, inflate, record scan, device-side BamReadFilter + compaction), "
                              "records to pinned host memory"}))


if __name__ == "__main__":
    main()
