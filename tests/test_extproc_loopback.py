"""BASELINE config 1: ext_proc gRPC loopback, 1 ArksToken, 1 /v1/chat/completions request.
Pass = 4 ext_proc replies with the header sets of pkg/gateway/handle_request.go / handle_response.go and counters
rpm=1 rpd=1 tpm=45 tpd=45, quota 25/20/45 (README.md:161-192, examples/quickstart/quickstart.yaml:56-110).

CPU: the engine behind the server is the oracle (plumbing check; the product has no CPU path).
GPU: the engine is the CUDA library through the C ABI."""
import json
import os

import pytest

import orklib
from arks_b200 import abi, extproc, gateway
from arks_b200.extproc import PB
from arks_b200.tables import Tables

FX = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quickstart.json")))
NOW = 1_700_000_000


class OracleEngine:
    def __init__(self, tables):
        self.o = orklib.Oracle(tables)

    def handle_request_body(self, b):
        return self.o.request_batch(b)

    def handle_response_body(self, b):
        return self.o.response_batch(b)

    def snapshot_rate(self, now):
        return self.o.snapshot_rate(now)

    def snapshot_quota(self):
        return self.o.snapshot_quota()


def hdrs(pairs, end=False):
    m = PB["ProcessingRequest"]()
    for k, v in pairs:
        h = m.request_headers.headers.headers.add()
        h.key, h.raw_value = k, v.encode()
    m.request_headers.end_of_stream = end
    return m


def resp_hdrs(pairs):
    m = PB["ProcessingRequest"]()
    for k, v in pairs:
        h = m.response_headers.headers.headers.add()
        h.key, h.raw_value = k, v.encode()
    return m


def body(b, which, eos=True):
    m = PB["ProcessingRequest"]()
    getattr(m, which).body = b
    getattr(m, which).end_of_stream = eos
    return m


def set_headers(mut):
    return {o.header.key: (o.header.raw_value or o.header.value.encode()).decode() for o in mut.set_headers}


def run_loopback(engine, tables, make_batcher=None):
    import __graft_entry__ as ge
    ge.build()  # arks_extract_bearer lives in the C ABI (pure host function)
    srv = extproc.ExtProcServer(engine, tables, gateway.extract_bearer, clock=lambda: NOW,
                                batcher=make_batcher(lambda: NOW) if make_batcher else None)
    tick = iter(range(0, 10**9, 250))  # the server's monotonic clock: every reading 0.25 s after the one before
    srv.monotonic = lambda: next(tick) / 1000.0
    server, port = extproc.serve(srv, port=0)
    ch, stub = extproc.client_stub(port)
    try:
        msgs = [hdrs([(":method", "POST"), (":path", "/v1/chat/completions"), ("authorization", "Bearer sk-test123456"),
                      ("content-type", "application/json")]),
                body(FX["request_body"].encode(), "request_body"),
                resp_hdrs([(":status", "200"), ("content-type", "application/json")]),
                body(FX["response_body"].encode()[:100], "response_body", eos=False),
                body(FX["response_body"].encode()[100:], "response_body", eos=True)]
        replies = list(stub(iter(msgs)))
        assert [r.WhichOneof("response") for r in replies] == ["request_headers", "request_body", "response_headers",
                                                               "response_body", "response_body"]
        assert set_headers(replies[0].request_headers.response.header_mutation) == {"x-went-into-req-headers": "true"}
        assert replies[0].request_headers.response.clear_route_cache
        assert set_headers(replies[1].request_body.response.header_mutation) == {
            "model": "qwen-7b", "namespace": "default", "username": "example-token"}
        h3 = set_headers(replies[2].response_headers.response.header_mutation)
        assert h3 == {"x-went-into-resp-headers": "true", ":status": "200", "content-type": "application/json"}
        assert len(replies[4].response_body.response.header_mutation.set_headers) == 0
        assert engine.snapshot_rate(NOW)[0].tolist() == [1, 1, 45, 45]
        assert engine.snapshot_quota()[0].tolist() == [25, 20, 45]
        # wall-clock series (collector.go:35-38,47-49): RecordRequest on both response-body messages, the processing time once
        text = srv.metrics.exposition()
        lab = 'namespace="default",user="example-token",model="qwen-7b"'
        assert f"gateway_request_duration_seconds_count{{{lab}}} 2" in text
        assert f'gateway_request_duration_seconds_bucket{{{lab},le="0.5"}} 0' in text
        assert f"gateway_response_process_duration_milliseconds_count{{{lab}}} 1" in text
        assert "gateway_requests_total" not in text  # status 200 is counted on the device

        # upstream answers 503: the body is passed through, RecordRequest carries the status label (gateway.go:122-129)
        r = list(stub(iter([hdrs([("authorization", "Bearer sk-test123456")]), body(FX["request_body"].encode(), "request_body"),
                            resp_hdrs([(":status", "503")]), body(b"upstream down", "response_body")])))
        assert r[3].immediate_response.status.code == 503
        assert f'gateway_requests_total{{{lab},status="503"}} 1' in srv.metrics.exposition()

        # no bearer -> 401 x-error-token on the headers message (handle_request.go:48-56)
        r = list(stub(iter([hdrs([("content-type", "application/json")])])))
        assert r[0].WhichOneof("response") == "immediate_response" and r[0].immediate_response.status.code == 401
        assert set_headers(r[0].immediate_response.headers)["x-error-token"] == "true"
        assert json.loads(r[0].immediate_response.body)["error"]["code"] == 401

        # rpm 5: requests 3..5 pass, the 6th is 429 x-error-rate-limit with currentUsage 5 / limitMax 5
        codes = []
        for _ in range(4):
            r = list(stub(iter([hdrs([("Authorization", "Bearer sk-test123456")]),
                                body(FX["request_body"].encode(), "request_body")])))
            codes.append(r[1].WhichOneof("response"))
        assert codes == ["request_body"] * 3 + ["immediate_response"]
        assert r[1].immediate_response.status.code == 429
        assert "x-error-rate-limit" in set_headers(r[1].immediate_response.headers)
        detail = json.loads(json.loads(r[1].immediate_response.body)["error"]["message"])
        # RateLimitResponse.JSON(), ratelimiter/types.go:98-114
        assert (detail["ruleName"], detail["overLimit"], detail["currentUsage"], detail["limitMax"]) == ("rpm", True, 5, 5)
        assert detail["expiresAt"] == "2023-11-14T22:14:00Z"  # end of NOW's minute window

        # streaming request + SSE chunks
        sse = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "sse_stream.json")))
        srv.batcher.clock = lambda: NOW + 60
        sreq = b'{"model":"qwen-7b","stream":true,"stream_options":{"include_usage":true},"messages":[]}'
        msgs = [hdrs([("authorization", "Bearer sk-test123456")]), body(sreq, "request_body"), resp_hdrs([(":status", "200")])]
        msgs += [body(c.encode(), "response_body", eos=(i == len(sse["chunks"]) - 1)) for i, c in enumerate(sse["chunks"])]
        replies = list(stub(iter(msgs)))
        assert all(r.WhichOneof("response") != "immediate_response" for r in replies)
        assert engine.snapshot_rate(NOW + 60)[0].tolist() == [1, 6, 45, 90]
    finally:
        ch.close()
        server.stop(0)
        srv.batcher.close()


def test_loopback_plumbing_cpu_oracle_engine():
    t = Tables(FX["tokens"], FX["quotas"], FX["endpoints"])
    run_loopback(OracleEngine(t), t)


class ShimEngine:
    """snapshots of the oracle that sits behind the ABI shim (tests/abi_shim.c) the compiled batcher is linked against"""

    def __init__(self, cpu):
        import ctypes as C
        self.cpu, self.C = cpu, C
        cpu.shim.arks_shim_oracle.restype = C.c_void_p
        cpu.shim.arks_shim_oracle.argtypes = [C.c_void_p]
        self.o = orklib.Oracle.__new__(orklib.Oracle)
        self.o.tables, self.o.h = cpu.tables, cpu.shim.arks_shim_oracle(cpu.ctx)

    def snapshot_rate(self, now):
        return orklib.Oracle.snapshot_rate(self.o, now)

    def snapshot_quota(self):
        return orklib.Oracle.snapshot_quota(self.o)


def test_loopback_through_the_compiled_batcher_cpu():
    """the same gRPC loopback, batched by host/cpp's Batcher (linked against the oracle-backed ABI shim)"""
    from test_cpp_host import CpuEngine
    from arks_b200 import cpphost
    t = Tables(FX["tokens"], FX["quotas"], FX["endpoints"])
    cpu = CpuEngine(t, max_batch=256, max_bytes=1 << 20)
    cpu.tables = t
    eng = ShimEngine(cpu)
    try:
        run_loopback(eng, t, make_batcher=lambda clock: extproc.CompiledBatcher(cpu.b, clock))
    finally:
        eng.o.h = None  # the shim owns the oracle


@pytest.mark.gpu
def test_loopback_gpu_compiled_batcher():
    from arks_b200 import cpphost
    t = Tables(FX["tokens"], FX["quotas"], FX["endpoints"])
    g = gateway.Gateway(0, 4096, 8 << 20)
    g.load_tables(t)
    hb = cpphost.Batcher(cpphost.load(cpphost.build()), g._h, max_batch=256, max_bytes=1 << 20)
    run_loopback(g, t, make_batcher=lambda clock: extproc.CompiledBatcher(hb, clock))


@pytest.mark.gpu
def test_loopback_gpu_engine():
    t = Tables(FX["tokens"], FX["quotas"], FX["endpoints"])
    g = gateway.Gateway(0, 4096, 8 << 20)
    g.load_tables(t)
    run_loopback(g, t)


def test_rows_and_calls_that_arrive_after_close_are_answered_not_held():
    """a stream whose message reaches the Python batcher after close() gets the engine-failure row (-> 500, like an INCRBY
    error in the reference, handle_request.go:199-205) instead of waiting forever; a table swap asked for then raises"""
    tables = Tables(FX["tokens"], FX["quotas"], FX["endpoints"])
    b = extproc.Batcher(OracleEngine(tables), clock=lambda: NOW)
    assert int(b.request(FX["request_body"].encode(), b"sk-test123456")["reason"]) == abi.R_OK
    b.close()
    out = b.request(b"{}", b"t")
    assert int(out["reason"]) == 255 and "closed" in out["error"]
    assert int(b.response(b"{}", 0, abi.RESP_END_OF_STREAM)["reason"]) == 255
    with pytest.raises(RuntimeError):
        b.between_batches(lambda: 1)
