"""Kernel-only timing of the two-stage scan on the bench workload (65 536 distinct ~1 KiB requests, their ~600 B completions):
mean device time of the fast-path kernels alone and of the whole scan stages, one context per variant of the library's env
knobs (ARKS_WALK, ARKS_REGROUP, ARKS_CARVEOUT, ... read at arks_create). usage: fast_ab.py "A=1 B=2" "A=0" ...; one JSON line
per variant."""
import json, os, sys
import numpy as np
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import traffic
from arks_b200.gateway import Gateway
N = 65536
w = traffic.Workload(10000, seed=0xA2C5)
now0 = 1_700_000_000
req = w.request_batch(N, now0, seed=1000, body_size=1024, n_templates=0, varied=True)
resp = None
for variant in (sys.argv[1:] or [""]):
    for k in [k for k in os.environ if k.startswith("ARKS_")]:
        del os.environ[k]
    for kv in variant.split():
        k, v = kv.split("=", 1); os.environ[k] = v
    g = Gateway(0, N, 120 << 20); g.load_tables(w.tables)
    now = now0
    req.now_unix = now
    a = g.handle_request_body(req)
    if resp is None:
        resp = w.response_batch(a, now + 1, seed=3000, body_size=600, varied=True, n_templates=0)
    resp.now_unix = now + 1
    g.handle_response_body(resp)
    g.select_slot(0); g.stage_request(req); g.stage_response(resp)
    g.set_profiling(True)
    rq, rs, fq, fs, ad = [], [], [], [], []
    for k in range(40):
        now += 86400
        g.run_request(now); ms = g.last_kernel_ms(); rq.append(ms[0]); ad.append(ms[1]); fq.append(ms[2])
        g.run_response(now + 1); ms = g.last_kernel_ms(); rs.append(ms[0]); fs.append(ms[1])
    m = lambda v: round(float(np.mean(v[5:])) * 1e3, 1)
    print(json.dumps({"variant": variant, "us": {"fast_request": m(fq), "request_stage": m(rq), "limit_admit": m(ad),
                                                  "fast_response": m(fs), "response_stage": m(rs)}, "declined": int(g.last_declined)}), flush=True)
    del g
