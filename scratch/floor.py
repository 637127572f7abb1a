import sys, os, numpy as np
sys.path[:0] = ['.', 'tests']
import torch, orklib
from arks_b200 import traffic
from arks_b200.gateway import Gateway
w = traffic.Workload(10000, seed=0xA2C5)
o = orklib.Oracle(w.tables)
req = w.request_batch(65536, 1_700_000_000, seed=1000, n_templates=512)
a = o.request_batch(req)
resp = w.response_batch(a, 1_700_000_001, seed=2000)
n = min(resp.n, 65536)
g = Gateway(0, 65536, 80 << 20); g.load_tables(w.tables)
g.select_slot(0); g.stage_request(req); g.stage_response(resp)
g.set_profiling(True)
now = 1_700_100_000
rs, ps = [], []
for i in range(12):
    g.run_request(now); rs.append(g.last_kernel_ms()[0]); g.run_response(now + 1); ps.append(g.last_kernel_ms()[0]); now += 86400
print(os.environ.get("ARKS_LIB", "default"), "scan_request us", round(np.median(rs[2:]) * 1e3, 1), "scan_response us", round(np.median(ps[2:]) * 1e3, 1))
