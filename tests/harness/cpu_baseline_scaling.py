"""Threaded-oracle scaling on the box's host cores (the bench's reference arm): req/s for several thread counts."""
import os, sys, time
_R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import numpy as np
import orklib
from arks_b200 import traffic
w = traffic.Workload(10000, seed=1)
o = orklib.Oracle(w.tables)
now = 1_700_000_000
req = w.request_batch(65536, now, seed=1, varied=True, n_templates=4096)
a = o.request_batch(req, threads=32)
resp = w.response_batch(a, now + 1, seed=2, varied=True, n_templates=4096)
for th in (1, 8, 16, 32, 64, 128):
    tr = tp = 0.0
    for k in range(4):
        now += 86400; req.now_unix = now; resp.now_unix = now + 1
        t0 = time.perf_counter(); o.request_batch(req, threads=th); t1 = time.perf_counter(); o.response_batch(resp, threads=th); t2 = time.perf_counter()
        if k: tr += t1 - t0; tp += t2 - t1
    print(f"threads {th:4d}: request batch {tr/3*1e3:7.2f} ms  response batch {tp/3*1e3:7.2f} ms  -> {65536/((tr+tp)/3)/1e6:.2f} M req/s", flush=True)
