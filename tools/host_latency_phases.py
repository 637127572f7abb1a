"""Where a micro-batch's time goes in host/cpp's Batcher (open loop, varied bodies): per arrival rate the latency
percentiles, the mean batch size and the mean time per cycle spent submitting, waiting for the device and delivering.
ARKS_FAST_MIN / ARKS_GRAPH etc. are read by the library, so this script is the A/B harness for the latency path."""
import json, os, sys
import numpy as np
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import cpphost, traffic
from arks_b200.gateway import Gateway
w = traffic.Workload(10000, seed=1)
g = Gateway(0, 8192, 16 << 20); g.load_tables(w.tables)
L = cpphost.load(cpphost.build())
now = 1_700_000_000
depth = int(os.environ.get("DEPTH", "1"))
hb = cpphost.Batcher(L, g._h, max_batch=8192, max_bytes=16 << 20, linger_us=0, max_inflight=depth)
warm = w.request_batch(4000, now, seed=2, body_size=1024, n_templates=64, varied=True)  # module load, first launches, page faults
hb.set_fixed_clock(now); hb.open_loop_requests(warm, 200_000, producers=8)
out = {"env": {k: v for k, v in os.environ.items() if k.startswith("ARKS_")}, "depth": depth}
for rate in (100_000, 500_000, 1_250_000, 2_500_000):
    now += 86400; hb.set_fixed_clock(now)
    n = int(rate * 0.2)
    load = w.request_batch(n, now, seed=4, body_size=1024, n_templates=512, varied=True)
    hb.reset_tail(); b0 = hb.stats(); _, lat, wall = hb.open_loop_requests(load, rate, producers=8); b1 = hb.stats()
    l = np.sort(lat[n // 10:]) / 1e3
    cyc = max(b1["cycles"] - b0["cycles"], 1)
    out[str(rate)] = {"p50": round(float(l[len(l) // 2])), "p99": round(float(l[int(len(l) * .99)])), "p999": round(float(l[int(len(l) * .999)])),
                      "req_s": round(n / wall), "mean_batch": round(n / cyc, 1),
                      "us_submit": round((b1["ns_submit"] - b0["ns_submit"]) / cyc / 1e3, 1),
                      "us_device": round((b1["ns_device"] - b0["ns_device"]) / cyc / 1e3, 1),
                      "us_deliver": round((b1["ns_deliver"] - b0["ns_deliver"]) / cyc / 1e3, 1),
                      "tail": {k: (round(v / 1e3) if k.startswith("max") else v) for k, v in b1.items() if k.startswith(("max_ns", "slow_", "late_"))},
                      "from_call": {"p50": round(float(np.percentile(hb.last_call_latency[n // 10:], 50)) / 1e3), "p99": round(float(np.percentile(hb.last_call_latency[n // 10:], 99)) / 1e3),
                                    "p999": round(float(np.percentile(hb.last_call_latency[n // 10:], 99.9)) / 1e3)},
                      "harness": hb.open_loop_lateness()}
print(json.dumps(out), flush=True)
hb.close()
