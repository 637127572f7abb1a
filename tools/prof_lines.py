"""Per-CUDA-source-line instruction totals from an ncu report (needs -lineinfo): python tools_lines.py rep kernel [top]"""
import csv, subprocess, sys, io
rep, pat = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 30
out = subprocess.run(["ncu","-i",rep,"--page","source","--print-source","cuda,sass","--csv","--kernel-name",f"regex:{pat}"],capture_output=True,text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
cur_file = None; lines = []
hdr = None
for r in rows:
    if not r: continue
    if r[0] == "File Path": cur_file = r[1].split('/')[-1]; continue
    if r[0] == "Line No": hdr = r; ie = r.index("Instructions Executed"); isamp = r.index("# Samples"); continue
    if r[0] == "Function Name": continue
    if hdr and r[0].isdigit():
        try: lines.append((int(r[ie] or 0), int(r[isamp] or 0), cur_file, int(r[0]), r[1].strip()[:110]))
        except ValueError: pass
tot = sum(l[0] for l in lines)
print("total (sum over lines, first kernel instance + duplicates):", tot)
for n, s_, f, ln, src in sorted(lines, key=lambda x: -x[0])[:top]:
    print(f"{n:10d} {100*n/tot:5.1f}% samp {s_:5d}  {f}:{ln}  {src}")
