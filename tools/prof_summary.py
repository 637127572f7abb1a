"""Summarise an ncu report: per-kernel headline metrics + SASS exec-count buckets + hottest instructions."""
import csv, subprocess, sys, io
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
raw = subprocess.run(["ncu","-i",rep,"--page","raw","--csv","--kernel-name",f"regex:{pat}"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(raw))); hdr=rows[0]
want=['gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__warps_active.avg.pct_of_peak_sustained_active','launch__registers_per_thread','inst_executed','sm__inst_executed.avg.per_cycle_active','smsp__thread_inst_executed_per_inst_executed.ratio','sm__cycles_elapsed.max','sass__inst_executed_local_loads','sass__inst_executed_local_stores','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','smsp__warps_eligible.avg.per_cycle_active']
r=rows[2]
for w in want:
    if w in hdr: print(f"{w} = {r[hdr.index(w)]}")
for h in hdr:
    if 'warp_issue_stalled' in h and h.endswith('_per_warp_active.pct') :
        v=float(r[hdr.index(h)] or 0)
        if v>3: print(f"  stall {h.split('stalled_')[1].split('_per_warp')[0]} = {v:.1f}")
src = subprocess.run(["ncu","-i",rep,"--page","source","--csv","--kernel-name",f"regex:{pat}"],capture_output=True,text=True).stdout
rows=list(csv.reader(io.StringIO(src)))
k=[i for i,x in enumerate(rows) if x and x[0]=="Address"][0]
hdr=rows[k]; ia=hdr.index('Source'); ie=hdr.index('Instructions Executed'); isamp=hdr.index('# Samples')
c=Counter(); w=Counter(); data=[]; first=None
for x in rows[k+1:]:
    if x and x[0]=="Kernel Name": break   # second kernel instance
    if len(x)<=ie: continue
    n=int(x[ie] or 0); c[n]+=1; w[n]+=n; data.append((int(x[isamp] or 0),n,x[ia].strip()))
tot=sum(w.values()); print("total warp inst", tot)
for n,kk in sorted(w.items(), key=lambda t:-t[1])[:12]: print(f"  exec_count {n:8d}: {c[n]:4d} sass -> {kk:9d} ({100*kk/tot:4.1f}%)")
print("hottest by samples:")
for s_,n,srcl in sorted(data,key=lambda t:-t[0])[:int(sys.argv[3]) if len(sys.argv)>3 else 14]: print(f"  {s_:5d} {n:8d} {srcl[:100]}")
