"""Sweep of host/cpp Batcher settings on a GPU box: queue depth x linger, closed loop (64 stream threads) and open loop
(arrival rates); prints (p50 us, p99 us, achieved req/s, mean batch) per case. Used to pick BatcherOptions defaults."""
import sys, json, numpy as np
import os
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import cpphost, traffic
from arks_b200.gateway import Gateway
w = traffic.Workload(10000, seed=1)
g = Gateway(0, 8192, 16 << 20); g.load_tables(w.tables)
L = cpphost.load(cpphost.build())
now = 1_700_000_000
for depth in (1, 2, 4):
    for linger in (0, 30):
        hb = cpphost.Batcher(L, g._h, max_batch=8192, max_bytes=16 << 20, linger_us=linger, max_inflight=depth)
        res = {}
        hb.set_fixed_clock(now)
        load = w.request_batch(30000, now, seed=3, body_size=1024, n_templates=64)
        b0 = hb.stats(); _, lat, wall = hb.run_requests(load, threads=64); b1 = hb.stats()
        l = np.sort(lat[3000:]) / 1e3
        res["closed64"] = (round(float(l[len(l)//2])), round(float(l[int(len(l)*.99)])), round(30000 / wall), round(30000 / (b1["request_batches"] - b0["request_batches"]), 1))
        for rate in (250_000, 1_250_000, 3_000_000):
            now += 86400; hb.set_fixed_clock(now)
            n = int(rate * 0.1)
            load = w.request_batch(n, now, seed=4, body_size=1024, n_templates=64)
            b0 = hb.stats(); _, lat, wall = hb.open_loop_requests(load, rate, producers=8); b1 = hb.stats()
            l = np.sort(lat[n // 10:]) / 1e3
            res[f"open{rate}"] = (round(float(l[len(l)//2])), round(float(l[int(len(l)*.99)])), round(n / wall), round(n / (b1["request_batches"] - b0["request_batches"]), 1))
        now += 86400
        print("depth", depth, "linger", linger, json.dumps(res), flush=True)
        hb.close()
