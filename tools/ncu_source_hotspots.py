"""Key metrics, stall shares and the hottest source lines of inflate_par_kernel from an `ncu --set full --import-source on`
report (run in the build container; needs only the ncu CLI):  python tools/ncu_source_hotspots.py gpurun_out/prof_r1.ncu-rep [N]"""
import csv,sys,subprocess
rep=sys.argv[1]
raw=subprocess.run(['ncu','-i',rep,'--page','raw','--csv','--kernel-name','regex:inflate_par'],capture_output=True,text=True).stdout
rows=list(csv.reader(raw.splitlines()))
hdr=rows[0]; units=rows[1]; r=rows[2]
want=['gpu__time_duration.sum','launch__occupancy_limit_shared_mem','launch__occupancy_limit_registers','launch__registers_per_thread','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','smsp__warps_eligible.avg.per_cycle_active','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio','sm__cycles_elapsed.max','sm__cycles_active.avg']
for i,h in enumerate(hdr):
    if h in want: print(h, units[i], r[i])
st={h:float(r[i]) for i,h in enumerate(hdr) if 'pcsamp_warps_issue_stalled' in h and 'not_issued' not in h}
t=sum(st.values())
print({k.replace('smsp__pcsamp_warps_issue_stalled_',''):round(v/t,3) for k,v in sorted(st.items(),key=lambda kv:-kv[1])[:8]})
cs=subprocess.run(['ncu','-i',rep,'--page','source','--print-source','cuda,sass','--csv','--kernel-name','regex:inflate_par'],capture_output=True,text=True).stdout
rows=list(csv.reader(cs.splitlines()))
cur=None; hdr=None; agg={}
for r in rows:
    if len(r)==2 and r[0]=='File Path': cur=r[1].split('/')[-1]; continue
    if len(r)>5 and r[0]=='Line No': ie=r.index('Instructions Executed'); sa=r.index('# Samples'); th=r.index('Thread Instructions Executed'); hdr=1; continue
    if hdr and len(r)>ie and r[0]!='':
        try: n=int(r[ie]); s=int(r[sa]); t=int(r[th])
        except: continue
        agg[(cur,int(r[0]))]=(n,s,t,r[1].strip()[:80])
tot=sum(v[0] for v in agg.values()); ts=sum(v[1] for v in agg.values())
print('total inst',tot,'samples',ts)
N=int(sys.argv[2]) if len(sys.argv)>2 else 45
for k,v in sorted(agg.items(), key=lambda kv:-kv[1][1])[:N]:
    print(f"{k[0][:14]:14s} {k[1]:4d} inst {100*v[0]/tot:5.1f}% samp {100*v[1]/ts:5.1f}% thr {v[2]/max(1,v[0]):4.1f} | {v[3]}")
