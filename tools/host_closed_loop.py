"""Closed-loop latency of the C++ host: N stream threads, one blocking HandleRequestBody at a time each (single-shape and\nheterogeneous ~1 KiB requests)."""
import sys, os, numpy as np
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import cpphost, traffic
from arks_b200.gateway import Gateway
w = traffic.Workload(10000, seed=1)
g = Gateway(0, 8192, 16 << 20); g.load_tables(w.tables)
L = cpphost.load(cpphost.build())
hb = cpphost.Batcher(L, g._h, max_batch=8192, max_bytes=16 << 20)
now = 1_700_000_000
for varied in (False, True):
    for streams in (1, 8, 64):
        hb.set_fixed_clock(now)
        load = w.request_batch(3000 * streams if streams < 64 else 60000, now, seed=3, body_size=1024, n_templates=256, varied=varied)
        b0 = hb.stats(); _, lat, wall = hb.run_requests(load, threads=streams); b1 = hb.stats()
        l = np.sort(lat[load.n // 10:]) / 1e3
        print("varied" if varied else "uniform", "streams", streams, "p50", round(float(l[len(l)//2]),1), "p99", round(float(l[int(len(l)*.99)]),1), "req/s", round(load.n / wall), "mean batch", round(load.n / (b1["request_batches"] - b0["request_batches"]), 1), flush=True)
        now += 86400
hb.close()
