"""Turn the ncu captures of a round (gpurun_out/) into the committed evidence under profiles/:
  profiles/launches_<round>.csv, profiles/ncu_<round>.md, profiles/roofline_traffic.json, bench JSON lines.
Usage: python tools/make_profiles_md.py r01  (expects gpurun_out/launches_<round>.csv, prof_<round>p / q.ncu-rep, prof_sse7.ncu-rep,
bench_<round>_n1.json, bench_<round>_reference.json, configs_<round>.jsonl)"""
import collections, csv, io, json, os, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
G, P = os.path.join(ROOT, "gpurun_out"), os.path.join(ROOT, "profiles")
rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
out = []
w = out.append
w(f"# ncu evidence, round {rnd[1:].lstrip('0')} (final kernels of the round; heterogeneous workload: every body of a wave distinct)\n")
w("## Launch list: `ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 160 python bench.py --steps 20 --warmup 5 --no-cpu-baseline`\n")
w(f"Raw CSV: `profiles/launches_{rnd}.csv`. Per-launch times under ncu are cold-cache and serialised: compare SHARES of the step, not absolutes.\n")
rows = [r for r in csv.reader(open(os.path.join(G, f"launches_{rnd}.csv"))) if r and r[0].isdigit()]
agg = collections.defaultdict(list)
for r in rows:
    agg[r[4].split("(")[0].strip()].append(float(r[-1]))
tot = sum(sum(v) for v in agg.values())
w("| kernel | launches | mean us (under ncu) | share of kernel time |\n|---|---|---|---|")
for k, v in sorted(agg.items(), key=lambda t: -sum(t[1])):
    w(f"| `{k}` | {len(v)} | {sum(v) / len(v) / 1e3:.1f} | {100 * sum(v) / tot:.1f}% |")
share = lambda pred: 100 * sum(sum(v) for k, v in agg.items() if pred(k)) / tot
b = json.load(open(os.path.join(G, f"bench_{rnd}_n1.json")))
km = b["kernels_ms"]
t = sum(km.values())
w("\nSame step timed live by `bench.py` (CUDA events on the library's stream, no profiler; `scan_*` include the three `len_*` "
  "ordering kernels queued in front of them): " + ", ".join(f"{k} {v * 1e3:.1f} us ({100 * v / t:.1f}%)" for k, v in km.items()) +
  f". ncu shares with the `len_*` kernels split evenly over the two scans: request {share(lambda k: 'scan_request' in k) + share(lambda k: k.startswith('len_')) / 2:.1f} %, "
  f"response {share(lambda k: 'scan_response' in k) + share(lambda k: k.startswith('len_')) / 2:.1f} %, admit + rank {share(lambda k: 'admit' in k or 'rank' in k):.1f} %.\n")
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "inst_executed", "sm__inst_executed.avg.per_cycle_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__warps_eligible.avg.per_cycle_active", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sass__inst_executed_local_loads", "sass__inst_executed_local_stores",
        "sm__cycles_elapsed.max", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__grid_size",
        "launch__block_size"] + [f"smsp__average_warps_issue_stalled_{x}_per_issue_active.ratio" for x in
                                 ("long_scoreboard", "short_scoreboard", "wait", "branch_resolving", "no_instruction", "not_selected")]
w("## `ncu --set full --clock-control none --import-source on`, one launch per kernel (B200, clocks not controlled)\n")
w("`-k regex:\"scan_request|limit_admit|len_\" -s 24 -c 8 python bench.py --steps 3 --warmup 3`, `-k regex:scan_response -s 3 -c 1` "
  "(same command) and `-k regex:scan_sse -s 6 -c 1 python bench_configs.py 3`\n")
traffic = {}
mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
for rep, pat in ((f"prof_{rnd}p", "scan_request"), (f"prof_{rnd}p", "limit_admit"), (f"prof_{rnd}q", "scan_response"),
                 (f"prof_{rnd}p", "len_hist"), (f"prof_{rnd}p", "len_scan"), (f"prof_{rnd}p", "len_scatter"), ("prof_sse7", "scan_sse")):
    raw = subprocess.run(["ncu", "-i", os.path.join(G, rep + ".ncu-rep"), "--page", "raw", "--csv", "--kernel-name", f"regex:{pat}"],
                         capture_output=True, text=True).stdout
    rr = list(csv.reader(io.StringIO(raw)))
    if len(rr) < 3:
        continue
    hdr, units, r = rr[0], rr[1], rr[2]
    name = r[hdr.index("Kernel Name")].split("(")[0]
    w(f"### {name}\n\n| metric | value | unit |\n|---|---|---|")
    for m in (want[:4] + want[6:8] + want[18:20] if pat.startswith("len_") else want):
        if m in hdr:
            i = hdr.index(m)
            w(f"| {m} | {r[i]} | {units[i]} |")
    rd, wr = (float(r[hdr.index(f"dram__bytes_{x}.sum")]) * mult[units[hdr.index(f"dram__bytes_{x}.sum")]] for x in ("read", "write"))
    traffic[pat] = {"dram_bytes_per_launch": rd + wr, "duration_us_under_ncu": float(r[hdr.index("gpu__time_duration.sum")])}
    w("")
open(os.path.join(P, f"ncu_{rnd}.md"), "w").write("\n".join(out) + "\n")
json.dump({"scan_request_kernel": traffic["scan_request"], "limit_admit_kernel": traffic["limit_admit"],
           "scan_response_kernel": traffic["scan_response"], "scan_sse_kernel": traffic["scan_sse"],
           "_source": f"profiles/ncu_{rnd}.md (ncu --set full, dram__bytes_read.sum + dram__bytes_write.sum, one launch per kernel, "
                      "config 2 wave of distinct bodies / config 3 chunk batch)"}, open(os.path.join(P, "roofline_traffic.json"), "w"), indent=1)
for f in (f"launches_{rnd}.csv", f"bench_{rnd}_n1.json", f"bench_{rnd}_reference.json", f"configs_{rnd}.jsonl"):
    shutil.copy(os.path.join(G, f), os.path.join(P, f))
print("wrote", os.path.join(P, f"ncu_{rnd}.md"))
