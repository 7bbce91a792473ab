"""Key metrics, stall shares, hottest source lines and per-function instruction shares of one kernel from an
`ncu --set full --import-source on` report:  python tools/ncu_regions.py <report.ncu-rep> <kernel regex> [N lines]"""
import csv, subprocess, sys, collections
rep, kre = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 40
raw = subprocess.run(['ncu', '-i', rep, '--page', 'raw', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units, r = rows[0], rows[1], rows[2]
want = ['gpu__time_duration.sum', 'launch__grid_size', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers', 'launch__registers_per_thread',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__warps_eligible.avg.per_cycle_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__cycles_elapsed.max', 'sm__cycles_active.avg', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_bytes.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__throughput.avg.pct_of_peak_sustained_elapsed', 'sm__throughput.avg.pct_of_peak_sustained_elapsed']
for i, h in enumerate(hdr):
    if h in want: print(h, units[i], r[i])
st = {h: float(r[i] or 0) for i, h in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h}
t = sum(st.values()) or 1
print({k.replace('smsp__pcsamp_warps_issue_stalled_', ''): round(v / t, 3) for k, v in sorted(st.items(), key=lambda kv: -kv[1])[:9]})
cs = subprocess.run(['ncu', '-i', rep, '--page', 'source', '--print-source', 'cuda,sass', '--csv', '--kernel-name', 'regex:' + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(cs.splitlines()))
cur = None; ok = False; agg = {}
for r in rows:
    if len(r) == 2 and r[0] == 'File Path': cur = r[1].split('/')[-1]; continue
    if len(r) > 5 and r[0] == 'Line No': ie = r.index('Instructions Executed'); sa = r.index('# Samples'); th = r.index('Thread Instructions Executed'); ok = True; continue
    if ok and len(r) > ie and r[0] != '':
        try: n = int(r[ie]); s = int(r[sa]); t = int(r[th])
        except ValueError: continue
        agg[(cur, int(r[0]))] = (n, s, t, r[1].strip()[:90])
tot = sum(v[0] for v in agg.values()) or 1; ts = sum(v[1] for v in agg.values()) or 1
print('total inst', tot, 'samples', ts)
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:N]:
    print(f"{k[0][:16]:16s} {k[1]:4d} inst {100*v[0]/tot:5.1f}% samp {100*v[1]/ts:5.1f}% thr {v[2]/max(1,v[0]):4.1f} | {v[3]}")
byfile = collections.defaultdict(lambda: [0, 0])
for (f, l), v in agg.items(): byfile[f][0] += v[0]; byfile[f][1] += v[1]
print({f: (round(100*a/tot, 1), round(100*b/ts, 1)) for f, (a, b) in byfile.items()})
if len(sys.argv) > 4:      # line ranges "name:a-b,name:a-b" of the main file
    main = sys.argv[5] if len(sys.argv) > 5 else None
    for spec in sys.argv[4].split(','):
        nm, ab = spec.split(':'); a, b = map(int, ab.split('-'))
        x = [0, 0]
        for (f, l), v in agg.items():
            if (main is None or f.startswith(main)) and a <= l <= b: x[0] += v[0]; x[1] += v[1]
        print(f"{nm:24s} inst {100*x[0]/tot:5.1f}% samp {100*x[1]/ts:5.1f}%")
