"""Per source line of an ncu report (compiled with -lineinfo): instructions executed, stall samples, lanes per instruction.
usage: prof_lines.py <report.ncu-rep> [top N]"""
import csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass,cuda"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = next(r for r in rows if r and r[0] == "Line No")
iS, iN, iE, iT = hdr.index("Source"), hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
files = {}
cur = None
out = []
for r in rows:
    if r and r[0] == "File Path": cur = r[1]; continue
    if r and r[0] and r[0].isdigit() and len(r) > iT:
        num = lambda v: int(v) if v.strip().lstrip("-").isdigit() else 0
        out.append((cur, int(r[0]), r[iS].strip(), num(r[iN]), num(r[iE]), num(r[iT])))
tot_e = sum(o[4] for o in out); tot_s = sum(o[3] for o in out)
print("total warp inst", tot_e, "samples", tot_s)
for f, ln, s, ns, ne, nt in sorted(out, key=lambda o: -o[3])[:top]:
    print(f"{100*ns/max(tot_s,1):5.1f}% smp {100*ne/max(tot_e,1):5.1f}% inst  lanes {nt/max(ne,1):5.1f}  {f.split('/')[-1]}:{ln}  {s[:90]}")
