#!/usr/bin/env python
"""bench.py — BGZF inflate -> BAM record/CIGAR decode -> pileup on B200 ("Measurement").

A step is ONE full pass of the hot path over the synthetic BAM of BASELINE.json configs[1]
(100 M reads, 1 contig, 30x, 150 bp, CIGAR 150M): every BGZF block inflated, every record decoded, every
pileup column built.  `value` = pileup positions per second with the compressed file already resident in
HBM and the columns left in HBM; `e2e` = the same pass through the C ABI with the file in (pinned) host
memory and every column batch copied back to host memory inside the timed region.  With N GPUs the ONE file is cut
into N shards (total work fixed: "strong" scaling), the only collectives being the column-table stitch and the exact-halo
exchange (biod_b200/stitch.py).  The line also carries the other configs BASELINE.json names: configs[2] (mixed CIGAR)
at N = 1, configs[3] (24 contigs, mixed CIGAR, scaled) at N > 1.

  python bench.py --gpus N --steps K --warmup W            # ours
  python bench.py --impl reference ...                     # restated BioD CPU path on the host cores
"""
import argparse
import ctypes as C
import json
import os
import shutil
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
    # configs[3] is 1 G reads (~130 GB compressed): scaled to what a bench run can generate and hold (the line says so)
    4: dict(name="configs[3]: 1G-read synthetic BAM, 24 contigs, 30x, mixed CIGAR", reads=1_000_000_000, refs=24, mixed=1),
}
CFG4_READS_PER_GPU = 12_500_000      # configs[3] inside the default N > 1 line: N x this many reads


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS))
    ap.add_argument("--reads", type=int, default=0, help="override the read count (the line then says so)")
    ap.add_argument("--level", type=int, default=-1, help="zlib level of the synthetic file (-1 = BioD writer default)")
    ap.add_argument("--blocks-per-batch", type=int, default=0,
                    help="0 = library default (three full waves of the inflate kernel; fewer waves per batch at N > 1)")
    ap.add_argument("--cpu-sample-reads", type=int, default=2_000_000)
    ap.add_argument("--straddle", action="store_true", help="htsjdk-style file: records cut across BGZF blocks")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-shards", default="weighted", choices=["weighted", "equal"],
                    help="N > 1, e2e legs: shares of the file in proportion to each rank's measured device->host rate, or equal")
    ap.add_argument("--e2e-input", default="memory", choices=["memory", "file"],
                    help="e2e leg: the caller's buffer (biodb_open_memory; page-locked whole at N = 1, the shard's range at N > 1) "
                         "or the file itself (biodb_open: pread into the library's pinned slabs)")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the blocks for the other named configs (config3 / config4)")
    ap.add_argument("--md", action="store_true",
                    help="also time the pass with use_md_tag (reference bases from MD tags, row N1); 1 GPU only")
    ap.add_argument("--cache-dir", default=os.environ.get("BIODB_BENCH_CACHE", "/dev/shm"))
    return ap.parse_args()


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", str(rank)))
    return rank, world, local


def synth_file(args, config, cfg, n_reads, rank, world, barrier):
    """Generate (rank 0) or load the synthetic BAM; returns (numpy uint8 array, path, seconds, generated now)."""
    from tools import bamgen
    os.makedirs(args.cache_dir, exist_ok=True)
    path = os.path.join(args.cache_dir, f"biod_b200_cfg{config}_{n_reads}_l{args.level}{'_s' if args.straddle else ''}.bam")
    t0 = time.time()
    made = False
    if rank == 0 and not os.path.exists(path):
        data = bamgen.generate(n_reads, cfg["refs"], bool(cfg["mixed"]), args.level, bamgen.SEED_BASE + config,
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
            if world == 1:
                return data, None, time.time() - t0, made
            raise
        del data
    barrier()
    # N > 1: every rank maps the same file read-only (one copy in the page cache; only the shard's pages are touched);
    # N = 1: a private copy that can be pinned
    # N > 1: every rank maps the one file privately ("c": the mapping is writable, which cudaHostRegister asks for; nothing
    # is ever written) and page-locks only its own shard's byte range (pin_input = 2)
    data = np.fromfile(path, dtype=np.uint8) if world == 1 else np.memmap(path, dtype=np.uint8, mode="c")
    return data, path, time.time() - t0, made


class ClockSampler:
    """SM clock and throttle reasons of one GPU DURING the timed region: NVML polled every 10 ms on a thread (the timed
    calls release the GIL); nvidia-smi -lms as the fallback when NVML cannot be loaded."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    BITS = {"hw_slowdown": 0x8, "hw_thermal_slowdown": 0x40, "sw_thermal_slowdown": 0x20, "sw_power_cap": 0x4}

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None
        self.sm, self.mx, self.reasons, self.stop_flag, self.thread, self.h = [], None, set(), False, None, None
        try:
            import pynvml
            pynvml.nvmlInit()
            # LOCAL_RANK indexes CUDA_VISIBLE_DEVICES; NVML sees every GPU of the box
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.nv, self.h = pynvml, pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.mx = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:  # noqa: BLE001
            self.h = None

    def _sample(self):
        nv = self.nv
        self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
        try:
            get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            r = int(get(self.h))
            for nm, bit in self.BITS.items():
                if r & bit:
                    self.reasons.add(nm)
        except Exception:  # noqa: BLE001
            pass

    def _poll(self):
        while not self.stop_flag:
            try:
                self._sample()
            except Exception:  # noqa: BLE001
                return
            time.sleep(0.01)

    def start(self):
        if self.h is not None:
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
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
        if self.h is not None:
            self.stop_flag = True
            if self.thread:
                self.thread.join(timeout=1.0)
            return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.mx,
                    "reasons": sorted(self.reasons), "samples": len(self.sm), "source": "nvml, every 10 ms during the timed passes"}
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
                "reasons": sorted(reasons), "samples": len(sm), "source": "nvidia-smi -lms 200"}


def run_pass(L, capi, reader, shard=None, info=None, compact=False, use_md=False, halo_voffset=None, maq=False, calls=None):
    """One full pileup pass (pileupColumns); shard=(rank, world) runs this rank's shard of it, shard=(first, n, count) the
    shards [first, first + count) of n (halo guessed at 8 BGZF blocks, or starting at halo_voffset).
    Returns (stats, n_records, n_cols, n_entries)."""
    p = capi.PileupParams()
    p.single_ref, p.skip_zero_coverage, p.end_at = 0, 1, 2**64 - 1
    p.compact_reads = int(compact)
    p.use_md_tag = int(use_md)
    p.maq_mode = 1 if maq else 0
    pl = C.c_void_p()
    sharded = shard is not None and shard[1] > 1
    count = shard[2] if sharded and len(shard) > 2 else 1       # shard=(first, n_shards, count): a span of shards as one pass
    if sharded and halo_voffset is not None:
        st = L.biodb_pileup_begin_shard_span_at(reader, C.byref(p), shard[0], count, shard[1], int(halo_voffset), C.byref(pl))
    elif sharded:
        st = L.biodb_pileup_begin_shard_span(reader, C.byref(p), shard[0], count, shard[1], 8, C.byref(pl))
    else:
        st = L.biodb_pileup_begin(reader, C.byref(p), C.byref(pl))
    if st != capi.OK:
        raise RuntimeError(L.biodb_last_error(reader).contents.message.decode())
    cb = capi.ColumnBatch()
    n_calls = 0
    while True:
        st = L.biodb_pileup_next(pl, C.byref(cb))
        if st == capi.EOF:
            break
        if st != capi.OK:
            raise RuntimeError(L.biodb_last_error(reader).contents.message.decode())
        n_calls += int(cb.n_calls)
    if calls is not None:
        calls.append(n_calls)
    s = capi.Stats()
    L.biodb_pileup_stats(pl, C.byref(s))
    nr, nc, ne = C.c_uint64(), C.c_uint64(), C.c_uint64()
    L.biodb_pileup_totals(pl, C.byref(nr), C.byref(nc), C.byref(ne))
    n_rec = nr.value
    if sharded:
        si = capi.ShardInfo()
        L.biodb_pileup_shard_info(pl, C.byref(si))
        n_rec = si.n_own_records
        if info is not None:
            info.update({f: getattr(si, f) for f, _ in si._fields_})
            reach = (C.c_uint64 * shard[1])()
            L.biodb_pileup_shard_reach(pl, reach)
            info["reach"] = [int(x) for x in reach]
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


def cpu_model():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except OSError:
        pass
    return "unknown"


def cpu_reference_block(args, config, cfg, threads, runs=1):
    """Restated BioD CPU path (oracle) on a bounded prefix of the same workload: inflate on `threads` threads
    (BgzfInputStream's task pool), record walk + pileup on one (readrange.d, pileup.d) — assumed to overlap perfectly.
    Two pileup figures: `value` materialises every column's bases / qualities / read list (what the GPU path delivers);
    `lazy` only sweeps — BioD's PileupRange.popFront keeps and advances the live reads and the consumer touches nothing
    but column.coverage (bases are lazy in BioD, pileup.d:115-134).  The truth for a given consumer lies between."""
    from oracle import oracle as orc
    from tools import bamgen
    n = args.cpu_sample_reads
    data = bamgen.generate(n, cfg["refs"], bool(cfg["mixed"]), args.level, bamgen.SEED_BASE + config)
    rs = [orc.cpu_baseline(data, threads, True) for _ in range(runs)]
    r = rs[-1]
    t_full = float(np.mean([max(x["t_inflate"], x["t_decode"] + x["t_pileup"]) for x in rs]))
    t_lazy = float(np.mean([max(x["t_inflate"], x["t_decode"] + x["t_pileup_lazy"]) for x in rs]))
    t_reads = float(np.mean([max(x["t_inflate"], x["t_decode"]) for x in rs]))
    d_compilers = [c for c in ("ldc2", "dmd", "gdc") if shutil.which(c)]
    blk = {"value": r["n_columns"] / t_full, "unit": "positions/s", "cores": threads, "kind": "port",
           "cpu_model": cpu_model(), "host_cpus": os.cpu_count(),
           "records_per_sec": r["n_records"] / t_full,
           "lazy": {"value": r["n_columns"] / t_lazy, "unit": "positions/s",
                    "what": "sweep only (popFront + coverage), nothing materialised: the floor of what BioD's own consumer pays"},
           "reads_only_records_per_sec": r["n_records"] / t_reads,
           "d_compilers_on_box": d_compilers,
           "sample": (f"first {n} reads of the workload ({r['n_columns']} positions): inflate {r['t_inflate']:.2f}s on "
                      f"{threads} threads; on one thread record walk {r['t_decode']:.2f}s, pileup {r['t_pileup']:.2f}s materialising "
                      f"every column / {r['t_pileup_lazy']:.2f}s sweeping only; value assumes inflate overlaps the consumer perfectly"),
           "note": "restated BioD CPU path (libz, g++ -O3), not the D binary: no D toolchain in this image" +
                   ("" if not d_compilers else f" (found on the box: {d_compilers}; not used)")}
    return blk, t_full * 1e3, r


def link_shares(torch, dist, world, per_rank=4):
    """Shares of the end-to-end pass for ranks whose host links are not equally fast (on this pool's 8-GPU node four GPUs
    copy device->host at 11 GB/s and four at 18 GB/s when all copy at once, tools/pcie_bw.py): every rank copies 256 MiB
    device->host four times, all ranks at the same time; the file is cut into per_rank * world shards and rank r gets a
    run of them in proportion to its rate (at least one).  Returns (first shard of every rank, counts, GB/s per rank)."""
    n = 256 << 20
    h = torch.empty(n, dtype=torch.uint8).pin_memory()
    d = torch.empty(n, dtype=torch.uint8, device="cuda")
    h.copy_(d, non_blocking=True)
    torch.cuda.synchronize()
    dist.barrier()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(4):
        h.copy_(d, non_blocking=True)
    b.record()
    torch.cuda.synchronize()
    mine = torch.tensor([4 * n / (a.elapsed_time(b) * 1e-3) / 1e9], dtype=torch.float64, device="cuda")
    allv = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(allv, mine)
    gbs = [float(v[0]) for v in allv]
    n_fine = per_rank * world
    raw = [n_fine * g / sum(gbs) for g in gbs]
    counts = [max(1, int(round(x))) for x in raw]
    while sum(counts) != n_fine:                      # settle the rounding on the rank it distorts least
        over = sum(counts) > n_fine
        k = max((i for i in range(world) if not over or counts[i] > 1),
                key=lambda i: (counts[i] - raw[i]) if over else (raw[i] - counts[i]))
        counts[k] += -1 if over else 1
    firsts = [sum(counts[:k]) for k in range(world)]
    return firsts, counts, gbs


def measure(args, L, capi, torch, dist, config, n_reads, rank, world, local, barrier, steps, warmup, want_e2e, want_reads_pass,
            want_md):
    """The device-resident and the end-to-end measurement of one config.  Returns a dict (rank 0 uses it)."""
    from biod_b200.stitch import exact_halos, exact_halos_of_spans, gather_reach, stitch_counts
    cfg = CONFIGS[config]
    data, path, t_gen, made = synth_file(args, config, cfg, n_reads, rank, world, barrier)
    bpb = args.blocks_per_batch or (0 if world == 1 else (-2 if world == 2 else -1))

    def open_reader(resident, device_output, pin, from_file=False):
        o = capi.Options()
        L.biodb_default_options(C.byref(o))
        o.device, o.blocks_per_batch = local, bpb
        o.resident_input, o.device_output, o.pin_input = int(resident), int(device_output), int(pin)
        h = C.c_void_p()
        if from_file:        # BamReader(filename): the library reads the file itself (pread into its pinned slabs)
            st = L.biodb_open(path.encode(), C.byref(o), C.byref(h))
        else:
            st = L.biodb_open_memory(data.ctypes.data, data.size, C.byref(o), C.byref(h))
        if st != capi.OK:
            raise RuntimeError(L.biodb_open_error().contents.message.decode())
        return h

    shard = (rank, world)
    out = {"generated_in_s": round(t_gen, 1), "generated_now": made, "compressed_bytes": int(data.size)}

    def exact_halo_pass(rd, info, **kw):
        """The halo exchange of a sharded pass: which shards started their halo too late?  Those run again from the
        exact offset.  Returns (replacement result or None, report)."""
        if world == 1:
            return None, {"exact": True, "rerun_shards": [], "exchange_ms": 0.0}
        t0 = time.time()
        rows, used = gather_reach(info["reach"], info["halo_voffset"], device="cuda")
        need, redo = exact_halos(rows, used)
        ms = (time.time() - t0) * 1e3
        res = None
        if rank in redo:
            res = run_pass(L, capi, rd, shard, info, halo_voffset=need[rank], **kw)
        return res, {"exact": True, "rerun_shards": redo, "exchange_ms": ms}

    # ---- value: compressed file resident in HBM, columns stay in HBM ------------------------------------
    rd = open_reader(True, True, False)
    for _ in range(warmup):
        run_pass(L, capi, rd, shard)
    sampler = ClockSampler(local)
    barrier()
    sampler.start()
    t0 = time.time()
    shard_info = {}
    runs = [run_pass(L, capi, rd, shard, shard_info) for _ in range(steps)]
    barrier()
    wall = time.time() - t0
    out["clocks"] = sampler.stop()
    redo_res, halo_report = exact_halo_pass(rd, shard_info)
    rerun_ms = float(redo_res[0].total_ms) if redo_res else 0.0
    if redo_res:
        runs[-1] = redo_res
    md_runs = None
    if want_md and world == 1:
        run_pass(L, capi, rd, shard, use_md=True)
        md_runs = [run_pass(L, capi, rd, shard, use_md=True) for _ in range(steps)]
    L.biodb_close(rd)
    dev_ms = sum(s[0].total_ms for s in runs[:steps])
    tt = torch.tensor([dev_ms, wall * 1e3, rerun_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    dev_ms, wall_ms, rerun_ms = float(tt[0]), float(tt[1]), float(tt[2])
    s0, n_rec, n_col, n_ent = runs[-1]
    # the stitch: every rank learns every shard's column / entry / record counts -> global column offsets, record bases
    torch.cuda.synchronize()
    t0 = time.time()
    stc = stitch_counts(n_col, n_ent, n_rec, device="cuda")
    torch.cuda.synchronize()
    stitch_ms = (time.time() - t0) * 1e3
    tot_col, tot_ent, tot_rec = stc["totals"]
    # shards re-run for an exact halo count as part of every step (they are rare: a read must span the whole guess)
    ms_per_step = dev_ms / steps + rerun_ms
    out.update(ms_per_step=ms_per_step, wall_ms_per_step=wall_ms / steps, value=tot_col / (ms_per_step * 1e-3),
               records_per_sec=tot_rec / (ms_per_step * 1e-3), totals=(tot_col, tot_ent, tot_rec),
               per_gpu=(n_rec, n_col, n_ent), stats=s0, runs=runs[:steps],
               stitch={"stitch_counts_ms": stitch_ms, "halo_exchange_ms": halo_report["exchange_ms"],
                       "what": "all-gather of 3 counts per rank (column / entry / record bases) + all-gather of the reach rows; "
                               "outside the timed passes, stated here"},
               halo=dict(halo_report, rerun_ms=rerun_ms))
    if md_runs:
        md_ms = float(np.mean([s[0].total_ms for s in md_runs]))
        out["md"] = {"value": md_runs[-1][2] / (md_ms * 1e-3), "unit": "positions/s", "ms_per_step": md_ms,
                     "pileup_stage_ms": float(np.mean([s[0].pileup_ms for s in md_runs])),
                     "d2h_bytes_per_step": int(md_runs[-1][0].d2h_bytes), "h2d_bytes_per_step": int(md_runs[-1][0].h2d_bytes),
                     "what": "same device-resident pass with use_md_tag: dna() length per read on the GPU, provider chain on the "
                             "host (12 B per read device->host), segment replay into reference_base[] on the GPU"}

    # ---- e2e: file in pinned host memory, every column batch copied back inside the timed region -------
    if want_e2e:
        # N = 1: the private copy of the file is page-locked whole; N > 1: every rank maps the shared file and page-locks
        # only the byte range of its own shard (+ halo), on demand
        from_file = args.e2e_input == "file"
        rd = open_reader(False, False, 1 if world == 1 else 2, from_file)
        # N > 1: the shares of this leg follow the ranks' host-link speeds (link_shares); the device-resident leg above
        # keeps equal shards
        shares = None
        if world > 1 and args.e2e_shards == "weighted":
            firsts, counts, gbs = link_shares(torch, dist, world)
            shard = (firsts[rank], sum(counts), counts[rank])
            shares = {"shards": sum(counts), "per_rank": counts, "probe_d2h_gbs": [round(g, 1) for g in gbs]}

            def exact_halo_pass(rd, info, **kw):                # the same exchange over spans: a span's halo must reach
                t0 = time.time()                                # back to what reaches its FIRST shard
                rows, used = gather_reach(info["reach"], info["halo_voffset"], device="cuda")
                need, redo = exact_halos_of_spans(rows, used, firsts)
                ms = (time.time() - t0) * 1e3
                res = run_pass(L, capi, rd, shard, info, halo_voffset=need[rank], **kw) if rank in redo else None
                return res, {"exact": True, "rerun_shards": redo, "exchange_ms": ms}
        einfo = {}
        run_pass(L, capi, rd, shard, einfo, compact=True)
        input_pinned = bool(L.biodb_input_is_pinned(rd)) or from_file
        barrier()
        es = [run_pass(L, capi, rd, shard, einfo, compact=True) for _ in range(steps)]
        barrier()
        eredo, _ = exact_halo_pass(rd, einfo, compact=True)
        e_rerun = float(eredo[0].total_ms) if eredo else 0.0
        ex = run_pass(L, capi, rd, shard, compact=False)     # same pass with explicit read_idx lists, for comparison
        barrier()
        # the fused consumer (row N3): MAQ genotype likelihoods of every column computed on the device behind the pileup
        # (reference bases from the MD tags), only the SNP calls copied back — MaqSnpCaller.findSNPs end to end
        minfo, ncalls = {}, []
        shard_eq = (rank, world)      # (equal shards: this pass is bound by GPU and host work, not by the link)
        run_pass(L, capi, rd, shard_eq, minfo, use_md=True, maq=True)
        barrier()
        mq = [run_pass(L, capi, rd, shard_eq, minfo, use_md=True, maq=True, calls=ncalls) for _ in range(steps)]
        barrier()
        rp = None
        if want_reads_pass and world == 1:
            # BamReader.reads alone: every record (raw bytes + field tables) delivered to host memory
            run_reads_pass(L, capi, rd)
            rp = run_reads_pass(L, capi, rd)
        L.biodb_close(rd)
        e_ms = sum(s[0].total_ms for s in es)
        te = torch.tensor([e_ms, e_rerun], dtype=torch.float64, device="cuda")
        tb = torch.tensor([float(es[-1][0].h2d_bytes), float(es[-1][0].d2h_bytes)], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(te, op=dist.ReduceOp.MAX)
            dist.all_reduce(tb, op=dist.ReduceOp.SUM)
        e_ms = float(te[0]) / steps + float(te[1])
        h2d_all, d2h_all = int(tb[0]), int(tb[1])
        out["e2e"] = {"value": tot_col / (e_ms * 1e-3), "unit": "positions/s", "ms_per_step": e_ms,
                      "h2d_bytes_per_step": int(es[-1][0].h2d_bytes), "d2h_bytes_per_step": int(es[-1][0].d2h_bytes),
                      "h2d_bytes_all_ranks": h2d_all, "d2h_bytes_all_ranks": d2h_all,
                      "records_per_sec": tot_rec / (e_ms * 1e-3), "input_pinned": input_pinned, "input": args.e2e_input,
                      "shares": shares,
                      "pcie_d2h_gbs": d2h_all / world / (e_ms * 1e-3) / 1e9,      # per GPU, averaged over the ranks
                      "stage_ms": {"inflate": float(np.mean([s[0].inflate_ms for s in es])),
                                   "record_scan": float(np.mean([s[0].scan_ms for s in es])),
                                   "pileup": float(np.mean([s[0].pileup_ms for s in es]))},
                      "columns": "compact_reads (sequential, lossless): positions as runs; per column n_starting_here, "
                                 "last_read, 64-bit window mask (+ stragglers); per entry base + qual",
                      "with_explicit_read_idx": {"ms_per_step": float(ex[0].total_ms), "d2h_bytes_per_step": int(ex[0].d2h_bytes),
                                                 "value": (tot_col / (float(ex[0].total_ms) * 1e-3)) if world == 1 else None}}
        m_ms = sum(s[0].total_ms for s in mq)
        tm = torch.tensor([m_ms, float(ncalls[-1])], dtype=torch.float64, device="cuda")
        if world > 1:
            tmx = tm.clone()
            dist.all_reduce(tmx, op=dist.ReduceOp.MAX)
            dist.all_reduce(tm, op=dist.ReduceOp.SUM)
            m_ms, tot_calls = float(tmx[0]) / steps, int(tm[1])
        else:
            m_ms, tot_calls = m_ms / steps, int(ncalls[-1])
        out["maq_e2e"] = {"value": tot_col / (m_ms * 1e-3), "unit": "positions/s", "ms_per_step": m_ms,
                          "h2d_bytes_per_step": int(mq[-1][0].h2d_bytes), "d2h_bytes_per_step": int(mq[-1][0].d2h_bytes),
                          "snp_calls": tot_calls,
                          "what": "MaqSnpCaller.findSNPs(makePileup(reads, use_md_tag)) through the C ABI from host memory: inflate, "
                                  "decode, pileup, MD reference bases and MAQ likelihoods on the device; only the calls "
                                  "(18 B each) and the MD chain's inputs (12 B per read) cross PCIe device->host"
                                  + ("" if world == 1 else "; shards start their MD provider chain at their halo, as BioD's chunks do")}
        if rp is not None:
            out["reads_pass_e2e"] = {"records_per_sec": rp[1] / (float(rp[0].total_ms) * 1e-3), "ms_per_step": float(rp[0].total_ms),
                                     "h2d_bytes_per_step": int(rp[0].h2d_bytes), "d2h_bytes_per_step": int(rp[0].d2h_bytes),
                                     "what": "BamReader.reads through the C ABI: raw record bytes + field / CIGAR tables of "
                                             "every record copied to host memory (no pileup)"}
    del data
    return out


def main():
    args = parse()
    rank, world, local = dist_env()
    cfg = CONFIGS[args.config]
    n_reads = args.reads or cfg["reads"]
    cores = os.cpu_count() or 1

    def workload_name(c, n, full):
        return c["name"] + ("" if n == full else f" — SCALED to {n} reads") + \
            f", zlib level {args.level}" + (", records straddling BGZF blocks (htsjdk layout)" if args.straddle else "")

    workload = workload_name(cfg, n_reads, cfg["reads"])

    # ------------------------------------------------------------------ reference arm (CPU) --------------
    if args.impl == "reference":
        if rank != 0:
            return
        threads = max(1, cores - 1)   # default taskPool = totalCPUs-1 inflate workers (bgzf/inputstream.d:456,467)
        blk, t_ms, r = cpu_reference_block(args, args.config, cfg, threads, runs=args.warmup + args.steps)
        blk_runs = args.warmup + args.steps
        line = {"impl": "reference", "metric": "pileup_positions_per_sec", "value": blk["value"], "unit": "positions/s",
                "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms,
                "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
                "config": {"workload": workload, "sample_reads": args.cpu_sample_reads, "runs_averaged": blk_runs},
                "records_per_sec": blk["records_per_sec"],
                "cpu_baseline": blk,
                "e2e": {"value": blk["value"], "unit": "positions/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0,
                "note": blk["note"]}
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
    m = measure(args, L, capi, torch, dist, args.config, n_reads, rank, world, local, barrier, args.steps, args.warmup,
                not args.no_e2e, True, args.md)
    s0 = m["stats"]
    runs = m["runs"]
    n_rec, n_col, n_ent = m["per_gpu"]
    tot_col, tot_ent, tot_rec = m["totals"]
    ms_per_step = m["ms_per_step"]
    # roofline of the dominant kernel (inflate): algorithmic bytes = compressed in + uncompressed out
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:  # noqa: BLE001
        pass
    peak, peak_src = (peaks.get("hbm_gbs"), "measured (MEASURED_PEAKS.json)") if peaks.get("hbm_gbs") else (6650.0, "fallback")
    # DRAM traffic of the dominant kernel from the committed `ncu --set full` capture (bytes per BGZF block there x
    # blocks per launch here); None when no capture is committed
    traffic, traffic_src = None, None
    for fn in ("ncu_full_r2_summary.json", "ncu_full_r1_summary.json"):
        try:
            per_block = 0.0
            for k in json.load(open(os.path.join(ROOT, "profiles", fn))):
                if "inflate" in k["kernel"] and k.get("dominant_inflate"):
                    def _b(v):
                        x, u = v.split()
                        return float(x) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
                    per_block += (_b(k["dram__bytes_read.sum"]) + _b(k["dram__bytes_write.sum"])) / float(k["launch__grid_size"].split()[0])
            if per_block:
                traffic, traffic_src = per_block * s0.n_blocks / max(1, s0.inflate_launches), fn
                break
        except Exception:  # noqa: BLE001
            continue
    infl_ms = np.mean([s[0].inflate_ms for s in runs])
    infl_bytes = s0.compressed_bytes + s0.uncompressed_bytes
    achieved = infl_bytes / (infl_ms * 1e-3) / 1e9
    roofline = {"kernel": "inflate (decode + resolve kernels)", "bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_launch": infl_bytes / max(1, s0.inflate_launches),
                "launches_per_step": int(s0.inflate_launches),
                "avg_launch_ms": infl_ms / max(1, s0.inflate_launches),
                "share_of_step": infl_ms / ms_per_step,
                "stage_ms": {"inflate": float(infl_ms), "record_scan": float(np.mean([s[0].scan_ms for s in runs])),
                             "pileup": float(np.mean([s[0].pileup_ms for s in runs]))}}
    # the two other stages against the same HBM peak (algorithmic bytes of SURVEY.md §8d / DESIGN.md §4: 283 B read +
    # 32 B written per record; 239 B per record + 6 B per entry + 20 B per column)
    scan_ms = float(np.mean([s[0].scan_ms for s in runs]))
    pile_ms = float(np.mean([s[0].pileup_ms for s in runs]))
    scan_bytes = 315.0 * n_rec
    pile_bytes = 239.0 * n_rec + 6.0 * n_ent + 20.0 * n_col
    roofline["other_stages"] = {
        "record_scan": {"bound": "hbm", "achieved": scan_bytes / (scan_ms * 1e-3) / 1e9 if scan_ms else None, "unit": "GB/s",
                        "frac": scan_bytes / (scan_ms * 1e-3) / 1e9 / peak if scan_ms else None,
                        "algorithmic_bytes": scan_bytes, "ms": scan_ms,
                        "note": "algorithmic (283 B + 32 B per record); the kernel itself touches ~182 B per record (no SEQ / QUAL) "
                                "and its chain walk is fused into the inflate kernel — from compressed bytes the decode runs at "
                                "records_per_sec, ~1 % of HBM"},
        "pileup": {"bound": "hbm", "achieved": pile_bytes / (pile_ms * 1e-3) / 1e9 if pile_ms else None, "unit": "GB/s",
                   "frac": pile_bytes / (pile_ms * 1e-3) / 1e9 / peak if pile_ms else None,
                   "algorithmic_bytes": pile_bytes, "ms": pile_ms}}
    line = {"metric": "pileup_positions_per_sec", "value": m["value"], "unit": "positions/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "wall_ms_per_step": m["wall_ms_per_step"],
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "u8",
            "data": "synthetic",
            "config": {"workload": workload, "reads_per_gpu": n_rec, "positions_per_gpu": n_col, "entries_per_gpu": n_ent,
                       "total_reads": tot_rec, "total_positions": tot_col, "halo": m["halo"],
                       "halo_check_passed": bool(m["halo"].get("exact")),
                       "compressed_bytes": m["compressed_bytes"],
                       "blocks_per_batch": args.blocks_per_batch or ("library default: 3 full waves of the inflate kernel" if world == 1
                                                                     else f"{2 if world == 2 else 1} full wave(s) of the inflate kernel"),
                       "cache": "inputs larger than L2: 12 GB compressed / 28 GB inflated per pass vs 126 MB L2",
                       "parallelism": ("1 GPU, whole file" if world == 1 else
                                       f"{world} shards of ONE file (one per GPU: total work fixed, 'strong' scaling), halos read from "
                                       "the file, no data-path collective; NCCL all-gathers of counts (stitch) and reach rows (exact halos)"),
                       "generated_in_s": m["generated_in_s"], "generated_now": m["generated_now"]},
            "records_per_sec": m["records_per_sec"],
            "inflate_out_gbs": s0.uncompressed_bytes / (infl_ms * 1e-3) / 1e9,
            "stitch": m["stitch"],
            "roofline": roofline, "gpu_launches": int(sum(s[0].kernel_launches for s in runs)), "clocks": m["clocks"]}
    if "md" in m:
        line["md_reference_bases"] = m["md"]
    if "e2e" in m:
        line["e2e"] = m["e2e"]
    if "reads_pass_e2e" in m:
        line["reads_pass_e2e"] = m["reads_pass_e2e"]
    if "maq_e2e" in m:
        line["maq_e2e"] = m["maq_e2e"]
    # diagnostics of the lane-parallel inflate kernel over everything run so far (0 blocks given up = no fallback)
    given_up = None
    try:
        cnt = (C.c_uint64 * 8)()
        if L.biodb_debug_inflate_counters(cnt, 0) == 0:
            given_up = int(cnt[0])
            line["inflate_counters"] = {"blocks_given_up": int(cnt[0]), "super_chunks": int(cnt[1]), "decode_rounds": int(cnt[2]),
                                        "matches_from_l2": int(cnt[3]), "matches": int(cnt[4]), "deflate_blocks": int(cnt[5])}
    except Exception:  # noqa: BLE001
        pass
    # size-independent checks of the full-size pass (the oracle does not run at this scale; tests/test_gpu_configs.py
    # compares its first 2 M reads column by column): every generated read came out of the record scan, a 150M read
    # makes exactly 150 column entries, the shards' columns add up across ranks, and (this rank's) inflate kernels
    # needed the fallback for no block
    line["checks"] = {"records_equal_generated": bool(tot_rec == n_reads),
                      "entries_equal_150_per_read": bool(tot_ent == 150 * tot_rec) if args.config == 2 else None,
                      "no_block_given_up": (given_up == 0) if given_up is not None else None}

    # ---- the other configs BASELINE.json names, as blocks of the same line --------------------------------
    if not args.no_extra and args.config == 2 and not args.reads:
        xc = 3 if world == 1 else 4
        xn = CONFIGS[3]["reads"] if world == 1 else CFG4_READS_PER_GPU * world
        try:
            x = measure(args, L, capi, torch, dist, xc, xn, rank, world, local, barrier, max(1, min(2, args.steps)), 1,
                        not args.no_e2e, False, False)
            line["config3" if xc == 3 else "config4"] = {
                "workload": workload_name(CONFIGS[xc], xn, CONFIGS[xc]["reads"]), "value": x["value"], "unit": "positions/s",
                "ms_per_step": x["ms_per_step"], "records_per_sec": x["records_per_sec"],
                "total_reads": x["totals"][2], "total_positions": x["totals"][0], "halo": x["halo"],
                "e2e": x.get("e2e"), "maq_e2e": x.get("maq_e2e"), "stitch": x["stitch"], "generated_in_s": x["generated_in_s"],
                "stage_ms": {"inflate": float(np.mean([s[0].inflate_ms for s in x["runs"]])),
                             "record_scan": float(np.mean([s[0].scan_ms for s in x["runs"]])),
                             "pileup": float(np.mean([s[0].pileup_ms for s in x["runs"]]))}}
        except Exception as e:  # noqa: BLE001 - the headline must survive a failure of the extra block
            line["config3" if xc == 3 else "config4"] = {"error": repr(e)[:300]}

    # ---- CPU baseline on the host cores (rank 0, N=1 only) ----------------------------------------------
    if not args.no_cpu and rank == 0 and world == 1:
        blk, _, _ = cpu_reference_block(args, args.config, cfg, max(1, cores - 1))
        line["cpu_baseline"] = blk
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
