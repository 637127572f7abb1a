"""N3: the gateway's Prometheus series that are functions of the request stream
(pkg/gateway/metrics/metrics.go:24-98), accumulated next to the counters.
CPU: the oracle's restatement of the three call sites (check.go:145, handle_response.go:99-109, gateway.go:129) on the
quickstart fixture + the text exposition. GPU: device rows == oracle rows after mixed waves, and across a table reload."""
import json
import os

import numpy as np
import pytest

import orklib
from arks_b200 import abi, metrics, traffic
from arks_b200.abi import RequestBatch, ResponseBatch
from arks_b200.tables import Tables

FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quickstart.json")))
NOW = 1_700_000_000


def test_oracle_metrics_quickstart_and_exposition():
    t = Tables(FX["tokens"], FX["quotas"], FX["endpoints"])
    o = orklib.Oracle(t)
    body, tok = FX["request_body"].encode(), b"sk-test123456"
    r = o.request_batch(RequestBatch.from_lists([body] * 7, [tok] * 7, NOW))
    assert r.reason.tolist() == [0] * 5 + [abi.R_RATE_LIMIT] * 2          # rpm 5
    m = o.snapshot_metrics()
    assert m[0, abi.METRIC_HITS:abi.METRIC_HITS + 4].tolist() == [2, 0, 0, 0]
    rb = FX["response_body"].encode()                                      # usage 25 / 20 / 45
    sse_usage = b'data: {"id":"c","choices":[],"usage":{"prompt_tokens":3,"completion_tokens":70000,"total_tokens":70003}}\n\n'
    flags = [abi.RESP_END_OF_STREAM,                                       # recorded
             abi.RESP_STREAM,                                              # usage-bearing chunk that is not the last: never recorded
             abi.RESP_STREAM | abi.RESP_END_OF_STREAM,                     # recorded
             abi.RESP_STREAM | abi.RESP_END_OF_STREAM | abi.RESP_COMPLETED,  # the stream had completed earlier: not again
             0]                                                            # partial non-stream body: only the message count
    o.response_batch(ResponseBatch.from_lists([rb, sse_usage, sse_usage, sse_usage, rb[:10]], [0] * 5, flags, NOW + 1))
    m = o.snapshot_metrics()
    assert m[0, abi.METRIC_MESSAGES] == 5
    assert m[0, abi.METRIC_USAGE:abi.METRIC_USAGE + 2].tolist() == [25 + 3, 20 + 70000]
    hin = m[0, abi.METRIC_HIST_IN:abi.METRIC_HIST_IN + 18]
    hout = m[0, abi.METRIC_HIST_OUT:abi.METRIC_HIST_OUT + 18]
    assert hin.sum() == 2 and hin[5] == 1 and hin[2] == 1                  # 25 <= 32 (le index 5), 3 <= 4 (index 2)
    assert hout.sum() == 2 and hout[5] == 1 and hout[17] == 1              # 20 <= 32, 70000 -> +Inf
    text = metrics.exposition(t, m)
    assert 'gateway_rate_limit_hits_total{namespace="default",user="example-token",model="qwen-7b",rule_type="rpm"} 2' in text
    assert 'gateway_requests_total{namespace="default",user="example-token",model="qwen-7b",status="200"} 5' in text
    assert 'gateway_token_usage{namespace="default",user="example-token",token="qwen-7b",type="output"} 70020' in text
    assert 'gateway_token_distribution_bucket{namespace="default",user="example-token",token="qwen-7b",type="input",le="4"} 1' in text
    assert 'gateway_token_distribution_count{namespace="default",user="example-token",token="qwen-7b",type="input"} 2' in text
    # survives a reload by (namespace, user, model)
    o.reload(t)
    assert np.array_equal(o.snapshot_metrics(), m)


@pytest.mark.gpu
def test_device_metrics_match_oracle(gwmod):
    w = traffic.Workload(n_tenants=60, seed=21)
    g = gwmod.Gateway(0, 8192, 16 << 20)
    g.load_tables(w.tables)
    g.enable_metrics(True)
    o = orklib.Oracle(w.tables)
    rng = np.random.default_rng(5)
    now = NOW
    for wave in range(5):
        req = w.request_batch(6000, now, seed=300 + wave, stream_frac=0.5, noise_frac=0.1)
        a = g.handle_request_body(req)
        b = o.request_batch(req)
        assert np.array_equal(a.reason, b.reason)
        resp = w.response_batch(a, now + 1, seed=400 + wave, noise_frac=0.1)
        # every combination of the two stream-position bits the metrics read
        extra = rng.choice([0, abi.RESP_END_OF_STREAM, abi.RESP_COMPLETED, abi.RESP_END_OF_STREAM | abi.RESP_COMPLETED],
                           resp.n).astype(np.uint8)
        resp.flags[:] = np.where(resp.flags & abi.RESP_STREAM, resp.flags | extra, resp.flags)
        for lo in range(0, resp.n, 8192):
            part = ResponseBatch(resp.bodies, resp.body_off[lo:lo + 8192].copy(), resp.body_len[lo:lo + 8192].copy(),
                                 resp.qos[lo:lo + 8192].copy(), resp.flags[lo:lo + 8192].copy(), now + 1)
            c, d = g.handle_response_body(part), o.response_batch(part)
            assert np.array_equal(c.reason, d.reason) and np.array_equal(c.usage, d.usage)
        gm, om = g.snapshot_metrics(), o.snapshot_metrics()
        assert np.array_equal(gm, om), f"wave {wave}: metric rows differ at {np.argwhere(gm != om)[:5]}"
        now += 61
    assert om[:, abi.METRIC_HITS:abi.METRIC_HITS + 4].sum() > 0 and om[:, abi.METRIC_HIST_IN:abi.METRIC_HIST_IN + 18].sum() > 0
    # reload with the tenants in another order: rows follow their (namespace, user, model) key
    w2 = traffic.Workload(n_tenants=60, seed=21)
    g.load_tables(w2.tables)
    o.reload(w2.tables)
    assert np.array_equal(g.snapshot_metrics(), o.snapshot_metrics())
    # all-SSE batch (scan_sse_kernel) and metrics off: nothing moves
    g.enable_metrics(False)
    before = g.snapshot_metrics()
    sse = ResponseBatch.from_lists([b'data: {"choices":[],"usage":{"prompt_tokens":1,"completion_tokens":2,"total_tokens":3}}\n\n'] * 64,
                                   [0] * 64, [abi.RESP_STREAM | abi.RESP_END_OF_STREAM] * 64, now)
    g.handle_response_body(sse)
    assert np.array_equal(g.snapshot_metrics(), before)
    g.enable_metrics(True)
    g.handle_response_body(sse)
    after = g.snapshot_metrics()
    assert after[0, abi.METRIC_MESSAGES] - before[0, abi.METRIC_MESSAGES] == 64
    assert after[0, abi.METRIC_USAGE] - before[0, abi.METRIC_USAGE] == 64


def test_host_metrics_exposition_buckets_and_truncation():
    """the wall-clock series kept by the host: bucket bounds of metrics.go:42,52, durations truncated to whole milliseconds"""
    m = metrics.HostMetrics()
    m.record_request("ns", "u", "m", 0.0999, 200)   # 99 ms -> 0.099 s -> le 0.1
    m.record_request("ns", "u", "m", 0.1004, 200)   # 100 ms -> 0.1 s -> le 0.1 (bounds are inclusive)
    m.record_request("ns", "u", "m", 61.0, 500)     # +Inf only
    m.record_resp_processing("ns", "u", "m", 0.0019)  # 1 ms
    m.record_resp_processing("ns", "u", "m", 7.5)     # 7500 ms -> +Inf
    text = m.exposition()
    lab = 'namespace="ns",user="u",model="m"'
    assert f'gateway_requests_total{{{lab},status="500"}} 1' in text and 'status="200"' not in text
    assert f'gateway_request_duration_seconds_bucket{{{lab},le="0.1"}} 2' in text
    assert f'gateway_request_duration_seconds_bucket{{{lab},le="60"}} 2' in text
    assert f'gateway_request_duration_seconds_bucket{{{lab},le="+Inf"}} 3' in text
    assert f"gateway_request_duration_seconds_sum{{{lab}}} 61.199" in text
    assert f'gateway_response_process_duration_milliseconds_bucket{{{lab},le="1"}} 1' in text
    assert f'gateway_response_process_duration_milliseconds_bucket{{{lab},le="5000"}} 1' in text
    assert f"gateway_response_process_duration_milliseconds_sum{{{lab}}} 7501" in text
    assert f"gateway_response_process_duration_milliseconds_count{{{lab}}} 2" in text
    assert metrics.HostMetrics().exposition() == ""
