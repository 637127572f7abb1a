"""N4, opt-in: the prompt's BPE count is charged to the token-type rules at request time and reconciled
with the upstream's `usage` at response time (include/arks_gateway.h: arks_set_precharge). The reference counts nothing at
request time (pkg/gateway/check.go:124-126), so this is new semantics with its own oracle mode: the estimates of a micro-batch
are charged when the batch commits, checks inside the batch see the counters as of its start.

CPU: the oracle's mode on a hand-made scenario. GPU: the CUDA path (BPE kernels + limit_admit + accounting) against the oracle
fed with the counts the device produced (which tests/test_bpe.py pins against HF tokenizers)."""
import os

import numpy as np
import pytest

import orklib
from arks_b200 import abi, bpe, traffic
from arks_b200.abi import RequestBatch, ResponseBatch
from arks_b200.tables import Tables, simple_endpoint, simple_quota, simple_token

NOW = 1_700_000_000
BUILD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build")


def small_tables():
    return Tables([simple_token("u", "ns", "tk", "m", [("tpm", 100), ("rpm", 50), ("tpd", 1000)], "q")],
                  [simple_quota("q", "ns", [("total", 10**6)])], [simple_endpoint("m", "ns")])


def reqs(n, now):
    return RequestBatch.from_lists([b'{"model":"m","messages":[{"role":"user","content":"hello there"}]}'] * n, [b"tk"] * n, now)


def test_oracle_precharge_mode_charges_at_commit_and_reconciles():
    t = small_tables()
    o = orklib.Oracle(t)
    o.set_precharge(True)
    o.set_estimates([40, 40, 40])
    a = o.request_batch(reqs(3, NOW))
    assert a.reason.tolist() == [0, 0, 0]            # inside the batch the counters are the batch's start: all three pass
    assert o.snapshot_rate(NOW)[0].tolist() == [3, 0, 120, 120]  # rpm, rpd(absent), tpm, tpd: 3 x 40 charged at commit
    b = o.request_batch(reqs(1, NOW + 1))            # no estimates handed over: nothing charged, but tpm is over now
    assert b.reason.tolist() == [abi.R_RATE_LIMIT] and (int(b.cur_usage[0]), int(b.limit_max[0])) == (120, 100)
    body = b'{"model":"m","usage":{"prompt_tokens":25,"completion_tokens":5,"total_tokens":30}}'
    resp = ResponseBatch.from_lists([body], [0], [abi.RESP_END_OF_STREAM], NOW + 2)
    resp.precharged = np.array([40], np.uint32)
    c = o.response_batch(resp)
    assert c.counted.tolist() == [1]
    assert o.snapshot_rate(NOW + 2)[0].tolist() == [3, 0, 110, 110]  # += 30 - 40
    assert o.snapshot_quota()[0].tolist() == [0, 0, 30]                 # quotas bill what the upstream reports
    # off: estimates are ignored, decisions are the reference's
    o2 = orklib.Oracle(t)
    o2.set_estimates([40, 40, 40])
    o2.request_batch(reqs(3, NOW))
    assert o2.snapshot_rate(NOW)[0].tolist() == [3, 0, 0, 0]


@pytest.mark.gpu
def test_gpu_precharge_matches_the_oracle_mode(gwmod):
    tok, text = bpe.standin_tokenizer(20_000, cache_dir=BUILD)
    w = traffic.Workload(n_tenants=300, seed=19)
    g = gwmod.Gateway(0, 16384, 32 << 20)
    g.load_tables(w.tables)
    g.load_bpe(bpe.load_tokenizer(text))
    g.set_precharge(True)
    o = orklib.Oracle(w.tables)
    o.set_precharge(True)
    est_prev = None
    denied_by_tokens = 0
    for step, n in enumerate((3000, 9000, 700)):  # warp path, fast path, warp path again
        now = NOW + step
        req = w.request_batch(n, now, seed=30 + step, varied=True)
        a = g.handle_request_body(req)
        assert a.bpe_count.max() > 10
        o.set_estimates(a.bpe_count)               # the counts themselves are pinned against tokenizers in tests/test_bpe.py
        want = o.request_batch(req)
        for k, v in a.fields().items():
            assert np.array_equal(v, want.fields()[k]), (step, k)
        assert np.array_equal(g.snapshot_rate(now), o.snapshot_rate(now)), step
        rules = np.array([w.tables.rl_rule[w.tables.qos_rl_off[q] + d] if r == abi.R_RATE_LIMIT else 255
                          for q, d, r in zip(a.qos, a.detail, a.reason)])
        denied_by_tokens += int(np.isin(rules, (2, 3)).sum())
        # every admitted stream answers; the host hands the estimate back
        resp = w.response_batch(a, now, seed=60 + step, varied=True)
        ok = np.flatnonzero(a.reason == abi.R_OK)
        assert resp.n == len(ok)
        resp.precharged = np.ascontiguousarray(np.where(a.bpe_count[ok] == bpe.UNCOUNTED, 0, a.bpe_count[ok]), np.uint32)
        c = g.handle_response_body(resp)
        d = o.response_batch(resp)
        for k, v in c.fields().items():
            assert np.array_equal(v, d.fields()[k]), (step, k)
        assert np.array_equal(g.snapshot_rate(now), o.snapshot_rate(now)), step
        assert np.array_equal(g.snapshot_quota(), o.snapshot_quota())
    # and with the switch off the same traffic is decided as the reference decides it
    g.set_precharge(False)
    o.set_precharge(False)
    req = w.request_batch(5000, NOW + 10, seed=99, varied=True)
    a, want = g.handle_request_body(req), o.request_batch(req)
    for k, v in a.fields().items():
        assert np.array_equal(v, want.fields()[k]), k
