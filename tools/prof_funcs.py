"""Aggregate an ncu report's per-line instruction counts by enclosing C++ function (needs -lineinfo)."""
import csv, subprocess, sys, io, re, os
from collections import Counter
rep, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv","--kernel-name",f"regex:{pat}"],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
funcs = {}
def load(path):
    if path in funcs: return funcs[path]
    m = []
    try: src = open(path).read().split("\n")
    except OSError: src = []
    cur = "?"
    for i, l in enumerate(src, 1):
        g = re.match(r"\s*(?:static |template.*|__device__ |__global__ |ARKS_HD |ARKS_OUTLINE |__forceinline__ |inline |constexpr )*[\w:<>\*& ]+?\b(\w+)\s*\([^;]*\)\s*(?:const)?\s*\{\s*$", l)
        if g and not l.strip().startswith(("if", "for", "while", "else", "switch", "return")): cur = g.group(1)
        m.append(cur)
    funcs[path] = m
    return m
agg = Counter(); samp = Counter(); thr = Counter(); cur_file = None; hdr = None; seen_fn = 0
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1]; continue
    if r[0] == "Function Name":
        continue
    if r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); isamp = r.index("# Samples"); it = r.index("Thread Instructions Executed"); continue
    if hdr and r[0].isdigit():
        ln = int(r[0]); m = load(cur_file)
        fn = m[ln-1] if 0 < ln <= len(m) else "?"
        try: agg[(os.path.basename(cur_file), fn)] += int(r[ie] or 0); samp[(os.path.basename(cur_file), fn)] += int(r[isamp] or 0); thr[(os.path.basename(cur_file), fn)] += int(r[it] or 0)
        except ValueError: pass
tot = sum(agg.values()); ts = sum(samp.values())
print("total", tot)
for k, v in agg.most_common(28): print(f"{v:10d} {100*v/tot:5.1f}%  samples {100*samp[k]/max(ts,1):5.1f}%  lanes {thr[k]/max(v,1):4.1f}  {k[0]}:{k[1]}")
