"""one varied request wave + its responses through the library (for ncu captures of the diverged case)"""
import sys, numpy as np
import os
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import traffic
from arks_b200.gateway import Gateway
N = 65536
w = traffic.Workload(10000, seed=1)
g = Gateway(0, N, 120 << 20); g.load_tables(w.tables)
now = 1_700_000_000
NT = int(os.environ.get("PROF_TEMPLATES", "0"))  # 0: every body distinct (the bench workload)
req = w.request_batch(N, now, seed=5, body_size=1024, n_templates=NT, varied=True)
for k in range(4):
    req.now_unix = now + 86400 * k
    a = g.handle_request_body(req)
    if k == 0:
        resp = w.response_batch(a, now + 1, seed=6, body_size=600, varied=True, n_templates=NT)
    resp.now_unix = now + 86400 * k + 1
    g.handle_response_body(resp)
print("done", int((a.reason == 0).sum()))
