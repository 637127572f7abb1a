"""Pins the oracle against every fixture the reference holds for the gateway hot path.

The reference (Go, not buildable here) has no decision-level tests; what it does hold is used verbatim:
  * examples/quickstart/quickstart.yaml:56-110   token sk-test123456, rpm 5 / tpm 40000 / rpd 100 / tpd 1000000,
                                                 quota prompt 100000 / response 500000 / total 600000, endpoint qwen-7b
  * README.md:161-192                            request body and response with usage 25 / 20 / 45
  * config/samples/arks_v1_arkstoken.yaml        rpm 60, rpd 1000
  * docs/gateway-usage.md:62-78                  static weights 60 / 40 + discovered 5
  * pkg/gateway/ratelimiter/redis_impl_test.go:51-107   rpd 10/100; rpm 10/100 + tpm 20/200
  * pkg/gateway/quota/redis_impl_test.go:144-170        10 vs 100 -> not over, 90 vs 80 -> over
  * pkg/gateway/quota/redis_impl_test.go:234-281        10 concurrent +1 -> 10
Key formats follow ratelimiter/cache_key.go:42-71 and quota/cache_key.go:40-58.
"""
import json
import os

import numpy as np
import pytest

import orklib
from arks_b200 import abi
from arks_b200.abi import RequestBatch, ResponseBatch
from arks_b200.tables import Tables

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def load(name):
    with open(os.path.join(GOLD, name)) as f:
        return json.load(f)


@pytest.fixture()
def quickstart():
    fx = load("quickstart.json")
    return Tables(fx["tokens"], fx["quotas"], fx["endpoints"], {("default", "qwen-7b"): ["arks-application-qwen-7b"]}), fx


def test_quickstart_config1_counters(quickstart):
    """BASELINE config 1: one ArksToken, one /v1/chat/completions request -> rpm=1 rpd=1 tpm=45 tpd=45, quota 25/20/45."""
    tables, fx = quickstart
    o = orklib.Oracle(tables)
    now = 1_700_000_000
    req = RequestBatch.from_lists([fx["request_body"].encode()], [b"sk-test123456"], now, pick_rand=[3])
    r = o.request_batch(req)
    assert r.reason[0] == abi.R_OK and r.qos[0] == 0 and r.token[0] == 0 and r.flags[0] == 0 and r.pick[0] == 0
    resp = ResponseBatch.from_lists([fx["response_body"].encode()], [0], [abi.RESP_END_OF_STREAM], now + 1)
    p = o.response_batch(resp)
    assert p.reason[0] == abi.R_OK and p.counted[0] == 1
    assert p.usage[0].tolist() == [25, 20, 45]
    assert o.snapshot_rate(now + 1)[0].tolist() == [1, 1, 45, 45]
    assert o.snapshot_quota()[0].tolist() == [25, 20, 45]


def test_quickstart_rpm_5_then_429(quickstart):
    """rpm 5 (quickstart.yaml:98-99): requests 1..5 admitted, the 6th is 429 on rule index 0 with currentUsage 5."""
    tables, fx = quickstart
    o = orklib.Oracle(tables)
    body = fx["request_body"].encode()
    r = o.request_batch(RequestBatch.from_lists([body] * 7, [b"sk-test123456"] * 7, 1_700_000_000))
    assert r.reason.tolist() == [0] * 5 + [abi.R_RATE_LIMIT] * 2
    assert r.detail[5] == 0 and r.cur_usage[5] == 5 and r.limit_max[5] == 5
    # next minute window: the rpm key changes (cache_key.go:73-80), rpd keeps counting
    r2 = o.request_batch(RequestBatch.from_lists([body], [b"sk-test123456"], 1_700_000_000 + 60))
    assert r2.reason[0] == abi.R_OK
    assert o.snapshot_rate(1_700_000_060)[0].tolist() == [1, 6, 0, 0]


def test_status_code_map(quickstart):
    """(http status, x-error header) per failure — SURVEY.md §8a status map, handle_request.go:97-171."""
    tables, fx = quickstart
    o = orklib.Oracle(tables)
    ok = fx["request_body"].encode()
    cases = [
        (b"{not json", b"sk-test123456", abi.R_REQUEST_BODY, 400, "x-error-request-body-processing"),
        (b'{"messages":[]}', b"sk-test123456", abi.R_NO_MODEL, 400, "x-error-no-model-in-request"),
        (ok, b"sk-unknown", abi.R_TOKEN_NOT_FOUND, 500, "x-error-token"),
        (b'{"model":"other"}', b"sk-test123456", abi.R_MODEL_NOT_IN_TOKEN, 500, "x-error-token"),
        (b'{"model":"qwen-7b","stream":true}', b"sk-test123456", abi.R_STREAM_OPTIONS, 400,
         "x-error-no-stream-options-include-usage"),
        (b'{"model":"qwen-7b","stream":true,"stream_options":{"include_usage":true}}', b"sk-test123456", abi.R_OK, 200,
         None),
    ]
    r = o.request_batch(RequestBatch.from_lists([c[0] for c in cases], [c[1] for c in cases], 1_700_000_000))
    for i, c in enumerate(cases):
        assert r.reason[i] == c[2], (i, c)
        assert abi.REASON_HTTP[int(r.reason[i])] == (c[3], c[4])
    assert r.flags[5] == 1


def test_model_without_endpoint_is_400():
    from arks_b200.tables import simple_token
    t = Tables([simple_token("u", "ns", "tk", "ghost", [("rpm", 5)])], [], [])
    o = orklib.Oracle(t)
    r = o.request_batch(RequestBatch.from_lists([b'{"model":"ghost"}'], [b"tk"], 1_700_000_000))
    assert r.reason[0] == abi.R_NO_MODEL_BACKENDS


def test_sample_token_rpm60_rpd1000():
    """config/samples/arks_v1_{arkstoken,arksquota,arksendpoint}.yaml: rpm 60, rpd 1000, basic-quota."""
    fx = load("sample_token.json")
    t = Tables(fx["tokens"], fx["quotas"], fx["endpoints"])
    o = orklib.Oracle(t)
    body = ('{"model":"%s"}' % fx["model"]).encode()
    tok = fx["tokens"][0]["spec"]["token"].encode()
    r = o.request_batch(RequestBatch.from_lists([body] * 61, [tok] * 61, 1_700_000_000))
    assert r.reason[:60].tolist() == [0] * 60 and r.reason[60] == abi.R_RATE_LIMIT and r.detail[60] == 0


def test_limiter_vectors_from_reference_tests():
    """redis_impl_test.go:51-107: DoLimit(rpd +10 / limit 100) and DoLimit(rpm +10, tpm +20): the values the
    reference left unchecked (`// TODO: check values`) are what INCRBY must yield."""
    from arks_b200.tables import simple_endpoint, simple_token
    t = Tables([simple_token("123", "", "tk", "m1", [("rpd", 100), ("rpm", 100), ("tpm", 200)])], [],
               [simple_endpoint("m1", "")])
    o = orklib.Oracle(t)
    body = b'{"model":"m1"}'
    now = 1_700_000_000
    o.request_batch(RequestBatch.from_lists([body] * 10, [b"tk"] * 10, now))  # ten +1 increments == one +10
    resp = b'{"model":"m1","usage":{"prompt_tokens":5,"completion_tokens":15,"total_tokens":20}}'
    o.response_batch(ResponseBatch.from_lists([resp], [0], [abi.RESP_END_OF_STREAM], now))
    assert o.snapshot_rate(now)[0].tolist() == [10, 10, 20, 0]
    assert orklib.rate_key("test", "", "123", "m1", 1, now) == "test:namespace=.user=123.model=m1.rpd:1699920000"


def test_quota_truth_table_and_concurrent_incr():
    """quota/redis_impl_test.go:144-170 (10/100 -> not over; 90/80 -> over) and :234-281 (10 x +1 -> 10)."""
    from arks_b200.tables import simple_endpoint, simple_quota, simple_token
    t = Tables([simple_token("user1", "ns", "t1", "m", [], "q1"), simple_token("user2", "ns", "t2", "m", [], "q2")],
               [simple_quota("q1", "ns", [("total", 100)]), simple_quota("q2", "ns", [("total", 80)])],
               [simple_endpoint("m", "ns")])
    o = orklib.Oracle(t)
    o.incr_quota_usage(0, [0, 0, 10])
    o.incr_quota_usage(1, [0, 0, 90])
    r = o.request_batch(RequestBatch.from_lists([b'{"model":"m"}'] * 2, [b"t1", b"t2"], 1_700_000_000))
    assert r.reason.tolist() == [abi.R_OK, abi.R_QUOTA] and r.cur_usage[1] == 90 and r.limit_max[1] == 80
    for _ in range(10):
        o.incr_quota_usage(0, [1, 0, 0])
    assert o.snapshot_quota()[0].tolist() == [10, 0, 10]
    # usage == limit still admits (strict >, quota/redis_impl.go:98-103)
    o.set_quota_usage(0, [0, 0, 100])
    r = o.request_batch(RequestBatch.from_lists([b'{"model":"m"}'], [b"t1"], 1_700_000_000))
    assert r.reason[0] == abi.R_OK
    assert orklib.quota_key("arks-quota", "ns", "q1", 2) == "arks-quota:namespace=ns.quotaname=q1.type=total."


def test_window_start_matches_go_truncate():
    """time.Unix(now,0).Truncate(W).Unix() == floor(now/W)*W for W in {60, 86400} (cache_key.go:73-80)."""
    rng = np.random.default_rng(7)
    for now in [0, 59, 60, 61, 86399, 86400, 1_700_000_123, 2_000_000_000, -1, -61] + rng.integers(0, 2**33, 200).tolist():
        for rule, w in ((0, 60), (1, 86400), (2, 60), (3, 86400)):
            assert orklib.window_start(now, rule) == (now // w) * w


def test_weighted_pick_gateway_usage_doc():
    """docs/gateway-usage.md:62-78: static 60 / 40 then discovered 5; cumulative walk of r mod 105."""
    w = [60, 40, 5]
    got = [orklib.weighted_pick(w, r) for r in range(105)]
    assert got == [0] * 60 + [1] * 40 + [2] * 5
    assert orklib.weighted_pick(w, 105 + 61) == 1
    assert orklib.weighted_pick([0, 0], 9) == -1


def test_readme_bodies_parse():
    fx = load("quickstart.json")
    rc, model, stream, so, iu = orklib.parse_request_body(fx["request_body"].encode())
    assert (rc, model, stream, so, iu) == (0, b"qwen-7b", 0, 0, 0)
    rc, mlen, usage = orklib.parse_response_body(fx["response_body"].encode())
    assert (rc, mlen, usage) == (0, 7, (25, 20, 45))


def test_sse_fixture():
    fx = load("sse_stream.json")
    chunks = [c.encode() for c in fx["chunks"]]
    got = [orklib.parse_sse_chunk(c) for c in chunks]
    assert [g[0] for g in got] == [0] * len(chunks)
    assert got[-1][1] == tuple(fx["usage"])
    assert all(g[1] == (0, 0, 0) for g in got[:-1])
    # an event split across chunks is a decode error in the reference (no carry-over, handle_response.go:114-117)
    whole = b"".join(chunks)
    cut = whole.index(b'"usage"') + 3
    assert orklib.parse_sse_chunk(whole[cut:])[0] == 1


def test_threaded_baseline_equals_serial_oracle():
    """bench.py's cpu_baseline / --impl reference run the oracle tenant-sharded over threads (ork_*_batch_mt): same
    decisions, counters and metric rows as the serial oracle, whatever the thread count"""
    import numpy as np
    import orklib
    from arks_b200 import traffic
    w = traffic.Workload(n_tenants=300, seed=9)
    now = 1_700_000_000
    serial = orklib.Oracle(w.tables)
    for threads in (1, 3, 7):
        o = orklib.Oracle(w.tables)
        s = orklib.Oracle(w.tables)
        t = now
        for wave in range(3):
            req = w.request_batch(2500, t, seed=20 + wave, stream_frac=0.3, noise_frac=0.15, varied=bool(wave & 1))
            a, b = s.request_batch(req), o.request_batch(req, threads=threads)
            assert all(np.array_equal(v, b.fields()[k]) for k, v in a.fields().items()), (threads, wave)
            resp = w.response_batch(a, t + 1, seed=40 + wave, noise_frac=0.1, varied=bool(wave & 1))
            c, d = s.response_batch(resp), o.response_batch(resp, threads=threads)
            assert all(np.array_equal(v, d.fields()[k]) for k, v in c.fields().items()), (threads, wave)
            assert np.array_equal(s.snapshot_rate(t + 1), o.snapshot_rate(t + 1))
            assert np.array_equal(s.snapshot_quota(), o.snapshot_quota())
            assert np.array_equal(s.snapshot_metrics(), o.snapshot_metrics())
            t += 45
    del serial
