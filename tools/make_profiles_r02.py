"""Turn this round's captures (gpurun_out/) into the committed evidence under profiles/:

  profiles/launches_r02.csv       ncu --metrics gpu__time_duration.sum launch list of a short bench.py run
  profiles/ncu_r02.md             launch shares + one `ncu --set full` capture per dominant kernel (summary, stalls, hot lines)
  profiles/roofline_traffic.json  dram bytes per launch of those kernels (what bench.py's roofline.traffic reads)
  profiles/sass_r02.txt           the SASS lines that show the Blackwell-specific instructions (UBLKCP / SYNCS / LDG.E.256)

usage: python tools/make_profiles_r02.py <launches.csv> <full.ncu-rep>"""
import collections, csv, io, json, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = os.path.join(ROOT, "profiles")
launches, rep = sys.argv[1], sys.argv[2]
out = []
w = out.append
w("# ncu evidence, round 2 (final kernels of the round; bench workload: every body of a wave distinct)\n")
w("## Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 8 --warmup 3 --no-cpu-baseline --latency-requests 0`\n")
w("Raw CSV: `profiles/launches_r02.csv`. Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes.\n")
rows = [r for r in csv.reader(open(launches)) if r and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[re.sub(r"<.*", "", r[4].split("(")[0]).replace("void ", "").strip()].append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
w("| kernel | launches | mean us (under ncu) | share of kernel time |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda t: -sum(t[1])):
    w(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
open(os.path.join(P, "launches_r02.csv"), "w").write(open(launches).read())

raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rr = list(csv.reader(io.StringIO(raw)))
hdr = rr[0]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__occupancy_limit_shared_mem",
        "launch__grid_size", "launch__block_size", "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__t_sector_hit_rate.pct", "lts__t_sector_hit_rate.pct", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum"]
stalls = [k for k in hdr if "issue_stalled" in k and k.endswith("per_issue_active.ratio")]
num = lambda v: float(v.replace(",", "")) if v.strip() not in ("", "n/a") else 0.0
traffic_path = os.path.join(P, "roofline_traffic.json")
traffic = json.load(open(traffic_path)) if os.path.exists(traffic_path) else {}
w("\n## `ncu --set full --clock-control none --import-source on -k regex:fast_re -c 2 python tools/prof_varied.py` (first launch of each kernel: cold caches)\n")
for r in rr[2:]:
    name = re.sub(r"<.*", "", r[hdr.index("Kernel Name")].split("(")[0]).replace("void ", "").strip()
    w(f"### `{name}`\n")
    w("| metric | value |\n|---|---|")
    for k in want:
        if k in hdr:
            w(f"| {k} | {r[hdr.index(k)]} {rr[1][hdr.index(k)]} |")
    top = sorted(stalls, key=lambda k: -num(r[hdr.index(k)]))[:7]
    w("| stalls per issue (top) | " + ", ".join(f"{k.split('issue_stalled_')[1].split('_per_')[0]} {num(r[hdr.index(k)]):.2f}" for k in top) + " |")
    rd, wr = num(r[hdr.index("dram__bytes_read.sum")]), num(r[hdr.index("dram__bytes_write.sum")])
    unit = rr[1][hdr.index("dram__bytes_read.sum")]
    mul = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1}.get(unit, 1e6)
    traffic[name] = {"dram_bytes_per_launch": (rd + wr) * mul, "duration_us_under_ncu": num(r[hdr.index("gpu__time_duration.sum")])}
    w("")
traffic["_source_r02"] = "profiles/ncu_r02.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, first launch, bench workload wave of distinct bodies)"
json.dump(traffic, open(traffic_path, "w"), indent=1)
lines = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "prof_lines.py"), rep, "25"], capture_output=True, text=True).stdout
w("## Source lines by stall samples (both kernels together)\n\n```\n" + lines + "```\n")
sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "arks_b200", "libarksgw.so")], capture_output=True, text=True).stdout
cnt = collections.Counter()
keep = []
fn = None
for l in sass.splitlines():
    m = re.search(r"Function : (\S+)", l)
    if m:
        fn = m.group(1)
    for pat in ("UBLKCP", "SYNCS", "LDG.E.ENL2.256", "LDGSTS", "UTMALDG"):
        if pat in l:
            cnt[(fn, pat)] += 1
            if cnt[(fn, pat)] == 1:
                keep.append(f"{fn}: {l.strip()}")
with open(os.path.join(P, "sass_r02.txt"), "w") as f:
    f.write("# cuobjdump -sass arks_b200/libarksgw.so: instructions per kernel that only exist from sm_90 / sm_100 on\n"
            "# UBLKCP = cp.async.bulk (TMA 1-D bulk copy), SYNCS = mbarrier arrive/wait, LDG.E.ENL2.256 = 256-bit global load (sm_100), LDGSTS = cp.async (sm_80)\n")
    for (fn, pat), c in sorted(cnt.items()):
        f.write(f"{c:5d}  {pat:16s} {fn}\n")
    f.write("\n# first occurrence of each\n" + "\n".join(keep) + "\n")
w("## Blackwell / Hopper-only instructions in the binary: `profiles/sass_r02.txt`\n")
for (fn, pat), c in sorted(cnt.items()):
    if pat != "LDGSTS":
        w(f"* `{pat}` x{c} in `{re.sub('^_Z[0-9]+', '', fn)[:60]}`")
open(os.path.join(P, "ncu_r02.md"), "w").write("\n".join(out) + "\n")
print("wrote profiles/ncu_r02.md, launches_r02.csv, roofline_traffic.json, sass_r02.txt")
