import sys, json, numpy as np
import os
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import traffic
from arks_b200.gateway import Gateway
w = traffic.Workload(10000, seed=1)
g = Gateway(0, 65536, 80 << 20); g.load_tables(w.tables)
now = 1_700_000_000
req = w.request_batch(65536, now, seed=3, body_size=1024, n_templates=256)
a = g.handle_request_body(req)
resp = w.response_batch(a, now + 1, seed=4)
resp.flags[:] |= 2
for on in (False, True, False, True):
    g.enable_metrics(on)
    g.set_profiling(True)
    ts = []
    for it in range(12):
        now += 86400
        req.now_unix = now; resp.now_unix = now + 1
        g.handle_request_body(req); k1 = g.last_kernel_ms()
        g.handle_response_body(resp); k2 = g.last_kernel_ms()
        ts.append((k1[0], k1[1], k2[0]))
    t = np.median(np.array(ts[2:]), axis=0)
    print("metrics", on, "scan_request %.1f us  limit_admit %.1f us  scan_response %.1f us" % tuple(t * 1e3))
