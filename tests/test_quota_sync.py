"""A14 / N2: syncQuotaUsage, the 10-second loop that mirrors quota usage into ArksQuota.status and
"restores" the store from the CR (pkg/gateway/qosconfig/arks_impl.go:217-300).
CPU: the oracle's restatement, including the reference's behaviour on an outdated store (SetUsage with Request == 0
zeroes it). GPU: arks_sync_quota_usage == oracle for random CR statuses, both modes."""
import numpy as np
import pytest

import orklib
from arks_b200 import traffic
from arks_b200.tables import Tables, simple_endpoint, simple_quota, simple_token

NOW = 1_700_000_000


def _tables():
    q = [simple_quota("team-a", "default", [("prompt", 1000), ("response", 1000), ("total", 5000)]),
         simple_quota("team-b", "default", [("total", 100)]),
         simple_quota("team-c", "default", [("total", 100), ("prompt", 50), ("total", 70)])]  # a type listed twice
    t = [simple_token("tok-a", "default", "sk-a", "m", [("rpm", 100)], quota="team-a"),
         simple_token("tok-b", "default", "sk-b", "m", [("rpm", 100)], quota="team-b"),
         simple_token("tok-c", "default", "sk-c", "m", [("rpm", 100)], quota="team-c")]
    return Tables(t, q, [simple_endpoint("m", "default", 1, [("b0", 1)])])


def test_oracle_sync_semantics():
    t = _tables()
    o = orklib.Oracle(t)
    o.set_quota_usage(0, [10, 20, 30])
    o.set_quota_usage(1, [0, 0, 7])
    o.set_quota_usage(2, [5, 0, 9])
    present = np.array([0b000, 0b100, 0b101], np.uint32)
    used = np.array([[0, 0, 0], [0, 0, 50], [5, 0, 3]], np.int64)
    act = o.sync_quota_usage(present, used)
    # team-a: no status yet -> three entries appended from the store
    assert act[0] == 1 and present[0] == 0b111 and used[0].tolist() == [10, 20, 30]
    # team-b: CR says 50, store says 7 -> "set usage to quotaService if outdated" ... with Request == 0: the store is zeroed
    assert act[1] == 2 and used[1].tolist() == [0, 0, 50] and o.snapshot_quota()[1].tolist() == [0, 0, 0]
    # team-c: total 3 < 9 -> CR raised; prompt equal -> nothing; the store is left alone
    assert act[2] == 1 and used[2].tolist() == [5, 0, 9] and o.snapshot_quota()[2].tolist() == [5, 0, 9]
    # repaired mode: the store is raised to the CR instead
    o.set_quota_usage(1, [0, 0, 7])
    present, used = np.array([0, 0b100, 0], np.uint32), np.array([[0, 0, 0], [0, 0, 50], [0, 0, 0]], np.int64)
    act = o.sync_quota_usage(present, used, restore=True)
    assert act[1] == 2 and o.snapshot_quota()[1].tolist() == [0, 0, 50]


@pytest.mark.gpu
@pytest.mark.parametrize("restore", [False, True])
def test_device_sync_matches_oracle(gwmod, restore):
    w = traffic.Workload(n_tenants=500, seed=13)
    g = gwmod.Gateway(0, 4096, 8 << 20)
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    # some traffic so that the stores are not all zero
    req = w.request_batch(4000, NOW, seed=1)
    a = g.handle_request_body(req)
    o.request_batch(req)
    resp = w.response_batch(a, NOW + 1, seed=2)
    g.handle_response_body(resp)
    o.response_batch(resp)
    assert np.array_equal(g.snapshot_quota(), o.snapshot_quota()) and g.snapshot_quota().sum() > 0
    rng = np.random.default_rng(3)
    n = w.tables.n_quotas
    cur = g.snapshot_quota()
    for rnd in range(3):
        present = rng.integers(0, 8, n).astype(np.uint32)
        used = (cur + rng.integers(-40, 40, (n, 3)) * rng.integers(0, 2, (n, 3))).astype(np.int64)
        p2, u2 = present.copy(), used.copy()
        act_g = g.sync_quota_usage(present, used, restore=restore)
        act_o = o.sync_quota_usage(p2, u2, restore=restore)
        assert np.array_equal(act_g, act_o) and np.array_equal(present, p2) and np.array_equal(used, u2)
        assert np.array_equal(g.snapshot_quota(), o.snapshot_quota())
        cur = g.snapshot_quota()
    assert (act_o & 2).any() and (act_o & 1).any()
