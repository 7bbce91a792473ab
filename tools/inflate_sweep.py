"""BASELINE.json configs[4]: BGZF-inflate microbench.  A stream of 0xFF00-payload blocks holding config-2 record
bytes, uncompressed total 1, 2, 4 ... GiB, zlib levels 0 / 1 / 6, on ONE GPU; compressed bytes resident in HBM ->
inflated bytes in HBM, timed with CUDA events around biodb_dev_inflate (C ABI).  The stream is the block sequence of
a generated BAM file laid out again and again in DISTINCT device memory until the target size is reached, so every
launch reads and writes its own HBM bytes.  Prints one JSON line per (level, size).

  python tools/inflate_sweep.py [--max-gib 64] [--reads 4000000] [--levels 0,1,6]"""
import argparse
import ctypes as C
import json
import os
import struct
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def block_table(data):
    """(payload offset, cdata size, isize) of every non-empty BGZF block (inputstream.d:54-199 layout, BC at +12)."""
    off, out = 0, []
    n = len(data)
    while off + 28 <= n:
        bsize = struct.unpack_from("<H", data, off + 16)[0] + 1
        isize = struct.unpack_from("<I", data, off + bsize - 4)[0]
        if isize:
            out.append((off + 18, bsize - 26, isize))
        off += bsize
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--max-gib", type=int, default=64)
    ap.add_argument("--reads", type=int, default=4_000_000)
    ap.add_argument("--levels", default="0,1,6")
    ap.add_argument("--blocks-per-launch", type=int, default=12432)   # 3 waves of the decode kernel
    a = ap.parse_args()
    import torch
    from biod_b200 import _capi
    from tools import bamgen
    L = _capi.lib()
    dev = torch.device("cuda:0")
    peak = 6442.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:  # noqa: BLE001
        pass
    for level in [int(x) for x in a.levels.split(",")]:
        buf = np.empty(a.reads * 300 + (1 << 20), dtype=np.uint8) if level == 0 else None    # stored blocks do not shrink
        data = bamgen.generate(a.reads, 1, False, level, bamgen.SEED_BASE + 2, out=buf)
        tab = block_table(data.tobytes())
        tab = tab[1:]                                  # drop the header block
        u_file = sum(t[2] for t in tab)
        c_lo, c_hi = tab[0][0] - 18, tab[-1][0] + tab[-1][1] + 8
        c_file = c_hi - c_lo
        d_file = torch.from_numpy(np.ascontiguousarray(data[c_lo:c_hi])).to(dev)
        gib = 1
        while gib <= a.max_gib:
            copies = max(1, -(-(gib << 30) // u_file))
            free = torch.cuda.mem_get_info()[0]
            need = copies * (c_file + u_file) + (1 << 30)
            if need > free:
                print(json.dumps({"level": level, "gib": gib, "skipped": f"needs {need >> 30} GiB of HBM, {free >> 30} free"}))
                break
            comp = torch.empty(copies * c_file + 256, dtype=torch.uint8, device=dev)
            for k in range(copies):
                comp[k * c_file:(k + 1) * c_file] = d_file
            nb = len(tab) * copies
            pay = np.empty(nb, dtype=np.int64)
            csz = np.empty(nb, dtype=np.int32)
            isz = np.empty(nb, dtype=np.int32)
            p0 = np.array([t[0] - c_lo for t in tab], dtype=np.int64)
            for k in range(copies):
                pay[k * len(tab):(k + 1) * len(tab)] = p0 + k * c_file
                csz[k * len(tab):(k + 1) * len(tab)] = [t[1] for t in tab]
                isz[k * len(tab):(k + 1) * len(tab)] = [t[2] for t in tab]
            ooff = np.concatenate([[0], np.cumsum(isz[:-1], dtype=np.int64)])
            out = torch.empty(int(isz.sum(dtype=np.int64)) + 256, dtype=torch.uint8, device=dev)
            d_pay, d_csz = torch.from_numpy(pay).to(dev), torch.from_numpy(csz).to(dev)
            d_isz, d_ooff = torch.from_numpy(isz).to(dev), torch.from_numpy(ooff).to(dev)
            d_st = torch.zeros(nb, dtype=torch.int32, device=dev)

            def run():
                for b0 in range(0, nb, a.blocks_per_launch):
                    n = min(a.blocks_per_launch, nb - b0)
                    rc = L.biodb_dev_inflate(comp.data_ptr(), d_pay[b0:].data_ptr(), d_csz[b0:].data_ptr(), d_ooff[b0:].data_ptr(),
                                             d_isz[b0:].data_ptr(), n, out.data_ptr(), d_st[b0:].data_ptr(), None, None)
                    assert rc == 0
            run()                                      # warm-up
            torch.cuda.synchronize()
            assert int(d_st.abs().sum()) == 0
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()                                # biodb_dev_inflate(stream=NULL) runs on the legacy default stream,
            run()                                      # which torch's current (default) stream is
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            u, c = int(isz.sum(dtype=np.int64)), int(csz.sum(dtype=np.int64))
            print(json.dumps({"level": level, "gib": gib, "blocks": nb, "uncompressed_bytes": u, "compressed_bytes": c,
                              "ms": ms, "out_gbs": u / ms / 1e6, "algorithmic_gbs": (u + c) / ms / 1e6,
                              "frac_of_hbm_peak": (u + c) / ms / 1e6 / peak, "hbm_peak_gbs": peak}), flush=True)
            del comp, out, d_pay, d_csz, d_isz, d_ooff, d_st
            torch.cuda.empty_cache()
            gib *= 2


if __name__ == "__main__":
    main()
