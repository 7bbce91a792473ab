"""Turns gpurun_out/launches_<round>.csv and gpurun_out/prof_<round>.ncu-rep into the summaries committed under
profiles/ (run in the build container; needs only the ncu CLI, no GPU)."""
import csv
import json
import subprocess
import sys
from collections import defaultdict

R = sys.argv[1] if len(sys.argv) > 1 else "r2"
KEYS = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "dram__bytes_read.sum",
        "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__warps_eligible.avg.per_cycle_active"]

# ---- launch list -> per-kernel totals and shares
rows = list(csv.reader(open(f"gpurun_out/launches_{R}.csv")))
start = next(i for i, r in enumerate(rows) if r and r[0] == "ID")
hdr = rows[start]
ik, iv = hdr.index("Kernel Name"), hdr.index("Metric Value")
agg = defaultdict(lambda: [0, 0.0])
for r in rows[start + 1:]:
    if len(r) > iv:
        agg[r[ik].split("(")[0].replace("biodb::", "").replace("<unnamed>::", "")][0] += 1
        agg[r[ik].split("(")[0].replace("biodb::", "").replace("<unnamed>::", "")][1] += float(r[iv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(f"profiles/launches_{R}_summary.csv", "w") as f:
    f.write("kernel,launches,total_ms,share_of_gpu_time,avg_us\n")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write(f"{k},{v[0]},{v[1] / 1e6:.3f},{v[1] / tot:.4f},{v[1] / v[0] / 1e3:.1f}\n")

# ---- full captures -> key metrics per kernel launch
raw = subprocess.run(["ncu", "-i", f"gpurun_out/prof_{R}.ncu-rep", "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr = rows[0]
units = rows[1]
idx = {h: i for i, h in enumerate(hdr)}
out = []
for r in rows[2:]:
    d = {"kernel": r[idx["Kernel Name"]].split("(")[0]}
    for k in KEYS:
        if k in idx:
            d[k] = f"{r[idx[k]]} {units[idx[k]]}".strip()
    st = {h.replace("smsp__pcsamp_warps_issue_stalled_", ""): float(r[idx[h]] or 0) for h in hdr
          if h.startswith("smsp__pcsamp_warps_issue_stalled_") and not h.endswith("_not_issued")}
    t = sum(st.values()) or 1
    d["stall_share"] = {k: round(v / t, 3) for k, v in sorted(st.items(), key=lambda x: -x[1])[:6]}
    out.append(d)
# the inflate stage is the decode + resolve pair: flag the largest launch of each (bench.py sums their DRAM bytes per
# BGZF block for roofline.traffic)
for name in ("inflate_decode_kernel", "inflate_resolve_kernel"):
    best = max((d for d in out if d["kernel"].endswith(name)), key=lambda d: float(d["launch__grid_size"].split()[0]), default=None)
    if best is not None:
        best["dominant_inflate"] = True
json.dump(out, open(f"profiles/ncu_full_{R}_summary.json", "w"), indent=1)
print(open(f"profiles/launches_{R}_summary.csv").read())
print(json.dumps(out, indent=1)[:3000])
