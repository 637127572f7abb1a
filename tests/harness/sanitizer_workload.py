"""A small pass over every kernel (request scan on all three paths — run with ARKS_FAST_MIN=1024 ARKS_WARP_MAX=512 so that these
small batches reach the fast and the warp path —, admit, hot-group ranking, JSON / SSE / mixed response scans, generation
swap, quota sync) for compute-sanitizer: `compute-sanitizer --tool memcheck|racecheck python tests/harness/sanitizer_workload.py`."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import numpy as np
import __graft_entry__ as ge; ge.build()
import orklib
from arks_b200 import abi, traffic
from arks_b200.abi import ResponseBatch
from arks_b200.gateway import Gateway
w = traffic.Workload(n_tenants=50, seed=3, zipf_alpha=1.2)
g = Gateway(0, 8192, 16 << 20); g.load_tables(w.tables); g.enable_metrics(True)
o = orklib.Oracle(w.tables)
now = 1_700_000_000
for n in (37, 5000):  # one body per warp / length-ordered full warps (and a hot tenant: > 256 arrivals of one qos)
    req = w.request_batch(n, now, seed=n, stream_frac=0.4, noise_frac=0.1, varied=True)
    a, b = g.handle_request_body(req), o.request_batch(req)
    assert all(np.array_equal(v, b.fields()[k]) for k, v in a.fields().items())
    resp = w.response_batch(a, now + 1, seed=n + 1, noise_frac=0.1, varied=True)
    kinds = {"mixed": np.arange(resp.n), "json": np.flatnonzero(~(resp.flags & 1).astype(bool)), "sse": np.flatnonzero((resp.flags & 1).astype(bool))}
    for name, sel in kinds.items():
        if len(sel) == 0: continue
        sel = sel[:8192]
        bodies = [bytes(resp.bodies[resp.body_off[i]:resp.body_off[i] + resp.body_len[i]]) for i in sel]
        rb = ResponseBatch.from_lists(bodies, resp.qos[sel], resp.flags[sel], now + 1)
        c, d = g.handle_response_body(rb), o.response_batch(rb)
        assert all(np.array_equal(v, d.fields()[k]) for k, v in c.fields().items()), name
    now += 61
    # a generation swap between the batches: the carry kernels move every counter by key (here: the same tables again)
    g.commit_tables(g.prepare_tables(w.tables)); o.reload(w.tables)
    assert np.array_equal(g.snapshot_rate(now), o.snapshot_rate(now)) and np.array_equal(g.snapshot_quota(), o.snapshot_quota())
pres = np.zeros(w.tables.n_quotas, np.uint32); used = np.zeros((w.tables.n_quotas, 3), np.int64)
g.sync_quota_usage(pres, used)
assert np.array_equal(g.snapshot_metrics(), o.snapshot_metrics())
print("sanitizer workload ok")
