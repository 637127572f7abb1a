"""What a config change costs the data path: open-loop request latency through host/cpp's Batcher at 1.25 M req/s, first
undisturbed, then with a config thread republishing the whole 10 000-tenant table every PERIOD_MS (Batcher::LoadTables =
arks_prepare_tables on the config thread + arks_commit_tables between two cycles). Prints both sets of percentiles, the
number of swaps that happened inside the timed window and the time one prepare / one commit took on the host."""
import json, os, sys, threading, time
import numpy as np
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import cpphost, traffic
from arks_b200.gateway import Gateway
PERIOD_MS = float(os.environ.get("PERIOD_MS", "20"))
RATE = int(os.environ.get("RATE", "1250000"))
w = traffic.Workload(10000, seed=1)
g = Gateway(0, 8192, 16 << 20); g.load_tables(w.tables)
L = cpphost.load(cpphost.build())
now = 1_700_000_000
hb = cpphost.Batcher(L, g._h, max_batch=8192, max_bytes=16 << 20, linger_us=0, max_inflight=1)
warm = w.request_batch(4000, now, seed=2, body_size=1024, n_templates=64, varied=True)
hb.set_fixed_clock(now); hb.open_loop_requests(warm, 200_000, producers=8)


def pct(lat, n):
    l = np.sort(lat[n // 10:]) / 1e3
    return {"p50": round(float(l[len(l) // 2])), "p99": round(float(l[int(len(l) * .99)])), "p999": round(float(l[int(len(l) * .999)])),
            "max": round(float(l[-1]))}


out = {"rate": RATE, "period_ms": PERIOD_MS, "tenants": 10000}
n = int(RATE * 1.0)
for leg in ("quiet", "reloading"):
    now += 86400; hb.set_fixed_clock(now)
    load = w.request_batch(min(n, 400_000), now, seed=4, body_size=1024, n_templates=512, varied=True)
    stop = threading.Event(); swaps = []

    ts = w.tables.c_struct()

    def config_thread():
        # the same objects every time (limits unchanged) so that decisions stay comparable; what is measured is the swap.
        # Straight into the C entry point (Batcher::LoadTables): no Python work competes with the load generator.
        import ctypes as C
        while not stop.is_set():
            t0 = time.perf_counter()
            rc = L.arks_host_load_tables(hb._h, C.byref(ts))
            assert rc == 0, rc
            swaps.append(time.perf_counter() - t0)
            stop.wait(PERIOD_MS / 1e3)

    th = threading.Thread(target=config_thread)
    if leg == "reloading":
        th.start()
    hb.reset_tail(); _, lat, wall = hb.open_loop_requests(load, RATE, producers=8)
    stop.set()
    if leg == "reloading":
        th.join()
    out[leg] = dict(pct(lat, len(lat)), req_s=round(len(lat) / wall), harness=hb.open_loop_lateness(),
                    tail={k: (round(v / 1e3) if k.startswith("max") else v) for k, v in hb.stats().items() if k.startswith(("max_ns", "slow_", "late_"))})
    if leg == "reloading":
        out[leg]["swaps"] = len(swaps)
        out[leg]["load_tables_ms_mean"] = round(1e3 * float(np.mean(swaps)), 2) if swaps else None
# the two halves timed apart, nothing else running
t0 = time.perf_counter(); p = g.prepare_tables(w.tables); t1 = time.perf_counter(); g.commit_tables(p); t2 = time.perf_counter()
out["prepare_ms"] = round(1e3 * (t1 - t0), 2); out["commit_host_us"] = round(1e6 * (t2 - t1), 1)
print(json.dumps(out), flush=True)
hb.close()
