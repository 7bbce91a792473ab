#!/usr/bin/env python
"""bench.py — BGZF inflate -> BAM record/CIGAR decode -> pileup on B200 ("Measurement").

A step is ONE full pass of the hot path over the synthetic BAM of BASELINE.json configs[1]
(100 M reads, 1 contig, 30x, 150 bp, CIGAR 150M): every BGZF block inflated, every record decoded, every
pileup column built.  `value` = pileup positions per second with the compressed file already resident in
HBM and the columns left in HBM; `e2e` = the same pass through the C ABI with the file in (pinned) host
memory and every column batch copied back to host memory inside the timed region.

  python bench.py --gpus N --steps K --warmup W            # ours
  python bench.py --impl reference ...                     # restated BioD CPU path on the host cores
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

CONFIGS = {
    # BASELINE.json configs[] index -> recipe
    2: dict(name="configs[1]: 100M-read synthetic BAM, 1 contig, 30x, 150bp, CIGAR 150M", reads=100_000_000, refs=1, mixed=0),
    3: dict(name="configs[2]: 100M-read synthetic BAM, 1 contig, 30x, mixed CIGAR (M/I/D/S/N, 5% indel)", reads=100_000_000, refs=1, mixed=1),
    # configs[3] is 1 G reads (~130 GB compressed): pass --reads to scale it to what the box's /dev/shm holds
    4: dict(name="configs[3]: 1G-read synthetic BAM, 24 contigs, 30x, mixed CIGAR", reads=1_000_000_000, refs=24, mixed=1),
}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="override the read count (the line then says so)")
    ap.add_argument("--level", type=int, default=-1, help="zlib level of the synthetic file (-1 = BioD writer default)")
    ap.add_argument("--blocks-per-batch", type=int, default=0, help="0 = library default (three full waves of the inflate kernel)")
    ap.add_argument("--cpu-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--straddle", action="store_true", help="htsjdk-style file: records cut across BGZF blocks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--md", action="store_true",
                    help="also time the pass with use_md_tag (reference bases from MD tags, row N1); 1 GPU only")
    ap.add_argument("--cache-dir", default=os.environ.get("BIODB_BENCH_CACHE", "/dev/shm"))
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def world_size():
    return int(os.environ.get("WORLD_SIZE", "1"))


def synth_file(args, cfg, n_reads, rank, barrier):
    """Generate (rank 0) or load the synthetic BAM; returns a numpy uint8 array."""
    from tools import bamgen
    os.makedirs(args.cache_dir, exist_ok=True)
    path = os.path.join(args.cache_dir, f"biod_b200_cfg{args.config}_{n_reads}_l{args.level}{'_s' if args.straddle else ''}.bam")
    t0 = time.time()
    made = False
    if rank == 0 and not os.path.exists(path):
        data = bamgen.generate(n_reads, cfg["refs"], bool(cfg["mixed"]), args.level, bamgen.SEED_BASE + args.config,
                               straddle=args.straddle)
        made = True
        tmp = path + ".tmp"
        try:
            data.tofile(tmp)
            os.replace(tmp, path)
        except OSError:
            # the cache directory cannot hold the file: one GPU works from memory, several need a shared file
            if os.path.exists(tmp):
                os.remove(tmp)
            if world_size() == 1:
                return data, None, time.time() - t0, made
            raise
        del data
    barrier()
    # N > 1: every rank maps the same file (one copy in the page cache); N = 1: a private copy that can be pinned
    data = np.fromfile(path, dtype=np.uint8) if world_size() == 1 else np.memmap(path, dtype=np.uint8, mode="c")
    return data, path, time.time() - t0, made


class ClockSampler:
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "200"], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for k, nm in enumerate(names):
                if len(r) > 3 + k and r[3 + k].lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def run_pass(L, capi, reader, shard=None, info=None, compact=False, use_md=False):
    """One full pileup pass (pileupColumns); shard=(rank, world) runs this rank's block-range shard of it.
    Returns (stats, n_records, n_cols, n_entries)."""
    p = capi.PileupParams()
    p.single_ref, p.skip_zero_coverage, p.end_at = 0, 1, 2**64 - 1
    p.compact_reads = int(compact)
    p.use_md_tag = int(use_md)
    pl = C.c_void_p()
    if shard is not None and shard[1] > 1:
        st = L.biodb_pileup_begin_shard(reader, C.byref(p), shard[0], shard[1], 8, C.byref(pl))
    else:
        st = L.biodb_pileup_begin(reader, C.byref(p), C.byref(pl))
    if st != capi.OK:
        raise RuntimeError(L.biodb_last_error(reader).contents.message.decode())
    cb = capi.ColumnBatch()
    while True:
        st = L.biodb_pileup_next(pl, C.byref(cb))
        if st == capi.EOF:
            break
        if st != capi.OK:
            raise RuntimeError(L.biodb_last_error(reader).contents.message.decode())
    s = capi.Stats()
    L.biodb_pileup_stats(pl, C.byref(s))
    nr, nc, ne = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.biodb_pileup_totals(pl, C.byref(nr), C.byref(nc), C.byref(ne))
    n_rec = nr.value
    if shard is not None and shard[1] > 1:
        si = capi.ShardInfo()
        L.biodb_pileup_shard_info(pl, C.byref(si))
        n_rec = si.n_own_records
        if info is not None:
            info.update({f: getattr(si, f) for f, _ in si._fields_})
    L.biodb_pileup_end(pl)
    return s, n_rec, nc.value, ne.value


def run_reads_pass(L, capi, reader):
    """One full BamReader.reads pass (inflate + record scan, no pileup).  Returns (stats, n_records)."""
    it = C.c_void_p()
    if L.biodb_reads_begin(reader, C.byref(it)) != capi.OK:
        raise RuntimeError(L.biodb_last_error(reader).contents.message.decode())
    rb = capi.RecordBatch()
    n = 0
    while True:
        st = L.biodb_reads_next(it, C.byref(rb))
        if st == capi.EOF:
            break
        if st != capi.OK:
            raise RuntimeError(L.biodb_last_error(reader).contents.message.decode())
        n += rb.n
    s = capi.Stats()
    L.biodb_reads_stats(it, C.byref(s))
    L.biodb_reads_end(it)
    return s, n


def cpu_reference(args, cfg, threads):
    """Restated BioD CPU path (oracle) on a bounded prefix of the same workload."""
    from oracle import oracle as orc
    from tools import bamgen
    n = args.cpu_sample_reads
    data = bamgen.generate(n, cfg["refs"], bool(cfg["mixed"]), args.level, bamgen.SEED_BASE + args.config)
    r = orc.cpu_baseline(data, threads, True)
    return n, r


def main():
    args = parse()
    rank, world, local = dist_env()
    cfg = CONFIGS[args.config]
    n_reads = args.reads or cfg["reads"]
    cores = os.cpu_count() or 1
    workload = cfg["name"] + ("" if not args.reads else f" — SCALED to {n_reads} reads by --reads") + \
        f", zlib level {args.level}" + (", records straddling BGZF blocks (htsjdk layout)" if args.straddle else "")

    # ------------------------------------------------------------------ reference arm (CPU) --------------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = max(1, cores - 1)   # default taskPool = totalCPUs-1 inflate workers (bgzf/inputstream.d:456,467)
        vals, rec = [], None
        from oracle import oracle as orc
        from tools import bamgen
        n = args.cpu_sample_reads
        data = bamgen.generate(n, cfg["refs"], bool(cfg["mixed"]), args.level, bamgen.SEED_BASE + args.config)
        for i in range(args.warmup + args.steps):
            r = orc.cpu_baseline(data, threads, True)
            t = max(r["t_inflate"], r["t_decode"] + r["t_pileup"])
            if i >= args.warmup:
                vals.append((r["n_columns"] / t, t, r))
        v = float(np.mean([x[0] for x in vals]))
        t = float(np.mean([x[1] for x in vals]))
        r = vals[-1][2]
        sample = (f"first {n} reads of the workload ({r['n_columns']} positions); inflate on {threads} threads "
                  f"{r['t_inflate']:.2f}s, record walk {r['t_decode']:.2f}s + pileup {r['t_pileup']:.2f}s on 1 thread; "
                  "value assumes the reference overlaps inflate with its consumer thread perfectly")
        line = {"impl": "reference", "metric": "pileup_positions_per_sec", "value": v, "unit": "positions/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": workload, "sample_reads": n},
                "records_per_sec": r["n_records"] / t,
                "cpu_baseline": {"value": v, "unit": "positions/s", "cores": threads, "kind": "port", "sample": sample},
                "e2e": {"value": v, "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": "restated BioD CPU path (libz, g++ -O3), not the D binary: no D toolchain in this image"}
        print(json.dumps(line))
        return

    # ------------------------------------------------------------------ our arm ---------------------------
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: biod_b200 has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from biod_b200 import _capi as capi
    L = capi.lib()
    data, path, t_gen, made = synth_file(args, cfg, n_reads, rank, barrier)

    def open_reader(resident, device_output, pin):
        o = capi.Options()
        L.biodb_default_options(C.byref(o))
        o.device, o.blocks_per_batch = local, args.blocks_per_batch
        o.resident_input, o.device_output, o.pin_input = int(resident), int(device_output), int(pin)
        h = C.c_void_p()
        st = L.biodb_open_memory(data.ctypes.data, data.size, C.byref(o), C.byref(h))
        if st != capi.OK:
            raise RuntimeError(L.biodb_open_error().contents.message.decode())
        return h

    # ---- value: compressed file resident in HBM, columns stay in HBM ------------------------------------
    shard = (rank, world)
    rd = open_reader(True, True, False)
    for _ in range(args.warmup):
        run_pass(L, capi, rd, shard)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.time()
    shard_info = {}
    steps = [run_pass(L, capi, rd, shard, shard_info) for _ in range(args.steps)]
    barrier()
    wall = time.time() - t0
    clocks = sampler.stop()
    md_steps = None
    if args.md and world == 1:
        # the same device-resident pass with PileupColumn.reference_base rebuilt from the MD tags (use_md_tag)
        run_pass(L, capi, rd, shard, use_md=True)
        md_steps = [run_pass(L, capi, rd, shard, use_md=True) for _ in range(args.steps)]
    L.biodb_close(rd)
    dev_ms = sum(s[0].total_ms for s in steps)
    tt = torch.tensor([dev_ms, wall * 1e3], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms = float(tt[0]), float(tt[1])
    s0, n_rec, n_col, n_ent = steps[-1]
    # the stitch (the only collective of the path): every rank learns every shard's column / entry / record counts
    # -> global column offsets and record bases; plus the halo exactness check of the shard boundaries
    from biod_b200.stitch import halo_sufficient, stitch_counts
    stc = stitch_counts(n_col, n_ent, n_rec, device="cuda")
    tot_col, tot_ent, tot_rec = stc["totals"]
    halo_ok = True
    if world > 1:
        keys = ["first_coffset", "halo_coffset", "lo_ref", "hi_ref", "lo_pos", "max_end_all", "max_end_outside_tail"]
        mine = torch.tensor([int(shard_info[k]) for k in keys], dtype=torch.int64, device="cuda")
        allv = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allv, mine)
        halo_ok = halo_sufficient([dict(zip(keys, (int(x) for x in v))) for v in allv])
    ms_per_step = dev_ms / args.steps
    value = tot_col / (ms_per_step * 1e-3)
    # roofline of the dominant kernel (inflate): algorithmic bytes = compressed in + uncompressed out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (bytes per BGZF block there x
    # blocks per launch here); None when no capture is committed
    traffic = None
    try:
        for k in json.load(open(os.path.join(ROOT, "profiles", "ncu_full_r1_summary.json"))):
            if k["kernel"].endswith("inflate_par_kernel"):
                def _b(v):
                    x, u = v.split()
                    return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                per_block = (_b(k["dram__bytes_read.sum"]) + _b(k["dram__bytes_write.sum"])) / float(k["launch__grid_size"].split()[0])
                traffic = per_block * s0.n_blocks / max(1, s0.inflate_launches)
    except Exception:  # noqa: BLE001
        traffic = None
    infl_ms = np.mean([s[0].inflate_ms for s in steps])
    infl_bytes = s0.compressed_bytes + s0.uncompressed_bytes
    achieved = infl_bytes / (infl_ms * 1e-3) / 1e9
    roofline = {"kernel": "inflate_par_kernel", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": infl_bytes / max(1, s0.inflate_launches),
                "launches_per_step": int(s0.inflate_launches),
                "avg_launch_ms": infl_ms / max(1, s0.inflate_launches),
                "share_of_step": infl_ms / ms_per_step,
                "stage_ms": {"inflate": float(infl_ms), "record_scan": float(np.mean([s[0].scan_ms for s in steps])),
                             "pileup": float(np.mean([s[0].pileup_ms for s in steps]))}}
    # the two other stages against the same HBM peak (algorithmic bytes of SURVEY.md §8d / DESIGN.md §4: 283 B read +
    # 32 B written per record; 239 B per record + 6 B per entry + 20 B per column)
    scan_ms = float(np.mean([s[0].scan_ms for s in steps]))
    pile_ms = float(np.mean([s[0].pileup_ms for s in steps]))
    scan_bytes = 315.0 * n_rec
    pile_bytes = 239.0 * n_rec + 6.0 * n_ent + 20.0 * n_col
    roofline["other_stages"] = {
        "record_scan": {"bound": "hbm", "achieved": scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms else None, "unit": "GB/s",
                        "frac": scan_bytes / (scan_ms * 1e-3) / 1e9 / peak if scan_ms else None,
                        "algorithmic_bytes": scan_bytes, "ms": scan_ms},
        "pileup": {"bound": "hbm", "achieved": pile_bytes / (pile_ms * 1e-3) / 1e9 if pile_ms else None, "unit": "GB/s",
                   "frac": pile_bytes / (pile_ms * 1e-3) / 1e9 / peak if pile_ms else None,
                   "algorithmic_bytes": pile_bytes, "ms": pile_ms}}
    line = {"metric": "pileup_positions_per_sec", "value": value, "unit": "positions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "wall_ms_per_step": wall_ms / args.steps,
            "higher_is_better": True, "scaling": "weak" if world == 1 else "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload, "reads_per_gpu": n_rec, "positions_per_gpu": n_col, "entries_per_gpu": n_ent,
                       "total_reads": tot_rec, "total_positions": tot_col, "halo_check_passed": halo_ok,
                       "compressed_bytes": int(data.size), "blocks_per_batch": args.blocks_per_batch or "library default: 3 full waves of the inflate kernel (8436 on a B200)",
                       "cache": "inputs larger than L2: 12 GB compressed / 28 GB inflated per pass vs 126 MB L2",
                       "parallelism": ("1 GPU, whole file" if world == 1 else
                                       f"{world} block-range shards of ONE file (one per GPU, halo of 8 blocks, no data-path "
                                       "collective), NCCL all-gather of counts + halo check for the column stitch"),
                       "generated_in_s": round(t_gen, 1), "generated_now": made},
            "records_per_sec": tot_rec / (ms_per_step * 1e-3),
            "inflate_out_gbs": s0.uncompressed_bytes / (infl_ms * 1e-3) / 1e9,
            "roofline": roofline, "gpu_launches": int(sum(s[0].kernel_launches for s in steps)), "clocks": clocks}

    if md_steps:
        md_ms = float(np.mean([s[0].total_ms for s in md_steps]))
        line["md_reference_bases"] = {
            "value": md_steps[-1][2] / (md_ms * 1e-3), "unit": "positions/s", "ms_per_step": md_ms,
            "pileup_stage_ms": float(np.mean([s[0].pileup_ms for s in md_steps])),
            "d2h_bytes_per_step": int(md_steps[-1][0].d2h_bytes), "h2d_bytes_per_step": int(md_steps[-1][0].h2d_bytes),
            "what": "same device-resident pass with use_md_tag: dna() length per read on the GPU, provider chain on the host "
                    "(12 B per read device->host), segment replay into reference_base[] on the GPU"}
    # diagnostics of the lane-parallel inflate kernel over everything run so far (0 blocks given up = no fallback)
    try:
        cnt = (C.c_uint64 * 8)()
        if L.biodb_debug_inflate_counters(cnt, 0) == 0:
            line["inflate_counters"] = {"blocks_given_up": int(cnt[0]), "super_chunks": int(cnt[1]), "decode_rounds": int(cnt[2]),
                                        "matches_from_l2": int(cnt[3]), "matches": int(cnt[4]), "deflate_blocks": int(cnt[5])}
    except Exception:  # noqa: BLE001
        pass

    # ---- e2e: file in pinned host memory, every column batch copied back inside the timed region -------
    if not args.no_e2e:
        rd = open_reader(False, False, True)   # pin the (possibly shared, mmap-ed) file buffer when the driver allows it
        input_pinned = bool(L.biodb_input_is_pinned(rd))
        run_pass(L, capi, rd, shard, compact=True)
        barrier()
        es = [run_pass(L, capi, rd, shard, compact=True) for _ in range(args.steps)]
        barrier()
        ex = run_pass(L, capi, rd, shard, compact=False)     # same pass with explicit read_idx lists, for comparison
        barrier()
        rp = None
        if world == 1:
            # BamReader.reads alone: every record (raw bytes + field tables) delivered to host memory
            run_reads_pass(L, capi, rd)
            rp = run_reads_pass(L, capi, rd)
        L.biodb_close(rd)
        e_ms = sum(s[0].total_ms for s in es)
        te = torch.tensor([e_ms], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
        e_ms = float(te[0]) / args.steps
        line["e2e"] = {"value": tot_col / (e_ms * 1e-3), "unit": "positions/s", "ms_per_step": e_ms,
                       "h2d_bytes_per_step": int(es[-1][0].h2d_bytes), "d2h_bytes_per_step": int(es[-1][0].d2h_bytes),
                       "records_per_sec": tot_rec / (e_ms * 1e-3), "input_pinned": input_pinned,
                       "pcie_d2h_gbs": es[-1][0].d2h_bytes / (e_ms * 1e-3) / 1e9,
                       "columns": "compact_reads (sequential, lossless): positions as runs; per column n_starting_here, "
                                  "last_read, 64-bit window mask (+ stragglers); per entry base + qual",
                       "with_explicit_read_idx": {"ms_per_step": float(ex[0].total_ms), "d2h_bytes_per_step": int(ex[0].d2h_bytes),
                                                  "value": (tot_col / (float(ex[0].total_ms) * 1e-3)) if world == 1 else None}}
        if rp is not None:
            line["reads_pass_e2e"] = {"records_per_sec": rp[1] / (float(rp[0].total_ms) * 1e-3), "ms_per_step": float(rp[0].total_ms),
                                      "h2d_bytes_per_step": int(rp[0].h2d_bytes), "d2h_bytes_per_step": int(rp[0].d2h_bytes),
                                      "what": "BamReader.reads through the C ABI: raw record bytes + field / CIGAR tables of every "
                                              "record copied to host memory (no pileup)"}

    # ---- CPU baseline on the host cores (rank 0, N=1 only) ----------------------------------------------
    if not args.no_cpu and rank == 0 and world == 1:
        threads = max(1, cores - 1)
        n, r = cpu_reference(args, cfg, threads)
        t = max(r["t_inflate"], r["t_decode"] + r["t_pileup"])
        line["cpu_baseline"] = {
            "value": r["n_columns"] / t, "unit": "positions/s", "cores": threads, "kind": "port",
            "records_per_sec": r["n_records"] / t,
            "sample": (f"first {n} reads of the workload ({r['n_columns']} positions): inflate {r['t_inflate']:.2f}s on "
                       f"{threads} threads, record walk {r['t_decode']:.2f}s + pileup {r['t_pileup']:.2f}s single-threaded "
                       "(as BioD does); value assumes perfect overlap of the two")}
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
