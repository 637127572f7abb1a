"""The rest of the gateway's wire surface next to ext_proc Process: the protobuf descriptors checked
against what ships in the image, grpc.health.v1 (pkg/gateway/gateway.go:261-279) and GET /v1/models
(pkg/gateway/http_handler.go:18-60) over real sockets on 127.0.0.1."""
import json
import os
import sys
import urllib.error
import urllib.request

import pytest

from arks_b200 import extproc
from arks_b200.tables import Tables, simple_endpoint, simple_quota, simple_token

sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tools"))


def test_core_descriptors_match_the_ones_grpcio_embeds():
    import verify_descriptors
    r = verify_descriptors.report()
    if r["source"] is None or not r["verified"]:
        pytest.skip("grpcio's C core carries no embedded envoy descriptors here")
    assert r["mismatch"] == {}
    assert {"HeaderValue", "HeaderMap", "HeaderValueOption", "HttpStatus"} <= set(r["verified"])
    assert r["verified"]["HeaderValue"] == {"key": 1, "value": 2, "raw_value": 3}


def tables():
    toks = [simple_token("alice", "team-a", "sk-alice", "qwen", [("rpm", 5)], quota="q"),
            {"metadata": {"name": "bob", "namespace": "team-a"},
             "spec": {"token": "sk-bob", "qos": [{"arksEndpoint": {"name": "qwen"}}, {"arksEndpoint": {"name": "模型-β"}}]}},
            {"metadata": {"name": "carol", "namespace": "team-b"}, "spec": {"token": "sk-carol", "qos": []}},
            simple_token("alice-2", "team-b", "sk-alice", "other", [])]  # same spec.token as alice: the first object wins
    return Tables(toks, [simple_quota("q", "team-a", [("total", 100)])], [simple_endpoint("qwen", "team-a", [("svc", 1)])])


def test_health_service_over_grpc():
    import grpc
    srv = extproc.ExtProcServer(engine=None, tables=tables(), extract_bearer=lambda hs: b"", batcher=object())
    server, port = extproc.serve(srv, port=0)
    try:
        ch = grpc.insecure_channel(f"127.0.0.1:{port}")
        HP = extproc.HEALTH_PB
        check = ch.unary_unary("/grpc.health.v1.Health/Check", request_serializer=HP["HealthCheckRequest"].SerializeToString,
                               response_deserializer=HP["HealthCheckResponse"].FromString)
        assert check(HP["HealthCheckRequest"](service="envoy.service.ext_proc.v3.ExternalProcessor"), timeout=10).status == extproc.SERVING
        assert check(HP["HealthCheckRequest"](), timeout=10).status == extproc.SERVING  # any service name: SERVING
        lst = ch.unary_unary("/grpc.health.v1.Health/List", request_serializer=HP["HealthListRequest"].SerializeToString,
                             response_deserializer=HP["HealthListResponse"].FromString)
        assert dict(lst(HP["HealthListRequest"](), timeout=10).statuses) == {}
        watch = ch.unary_stream("/grpc.health.v1.Health/Watch", request_serializer=HP["HealthCheckRequest"].SerializeToString,
                                response_deserializer=HP["HealthCheckResponse"].FromString)
        with pytest.raises(grpc.RpcError) as e:
            list(watch(HP["HealthCheckRequest"](), timeout=10))
        assert e.value.code() == grpc.StatusCode.UNIMPLEMENTED and e.value.details() == "watch is not implemented"
        # the status enum travels as the varint 1 in field 1
        assert HP["HealthCheckResponse"](status=1).SerializeToString() == b"\x08\x01"
        ch.close()
    finally:
        server.stop(0)


def get(port, auth=None, path="/v1/models", method="GET"):
    req = urllib.request.Request(f"http://127.0.0.1:{port}{path}", method=method, headers={} if auth is None else {"Authorization": auth})
    try:
        with urllib.request.urlopen(req, timeout=10) as r:
            return r.status, r.headers.get("Content-Type"), r.read()
    except urllib.error.HTTPError as e:
        return e.code, e.headers.get("Content-Type"), e.read()


def test_v1_models_over_http():
    t = tables()
    httpd, port = extproc.serve_http(lambda: t, port=0)
    try:
        plain = "text/plain; charset=utf-8"
        assert get(port) == (401, plain, b"Unauthorized\n")
        assert get(port, "sk-alice") == (401, plain, b"Unauthorized\n")       # no "Bearer " prefix
        assert get(port, "bearer sk-alice") == (401, plain, b"Unauthorized\n")  # the prefix is case-sensitive here
        assert get(port, "Bearer ") == (401, plain, b"Unauthorized\n")
        assert get(port, "Bearer nobody") == (500, plain, b"error in getting model list\n")
        st, ct, body = get(port, "Bearer sk-alice")
        assert (st, ct) == (200, "application/json")
        assert body == b'{"object":"list","data":[{"id":"qwen","created":0,"object":"model","owned_by":""}]}'  # alice, not alice-2
        st, ct, body = get(port, "Bearer sk-bob", method="POST")  # the mux pattern carries no method
        assert st == 200 and [m["id"] for m in json.loads(body)["data"]] == ["qwen", "模型-β"]
        assert "模型-β".encode() in body  # encoding/json leaves non-ASCII as UTF-8
        assert get(port, "Bearer sk-carol")[2] == b'{"object":"list","data":null}'  # nil slice
        assert get(port, "Bearer sk-alice", path="/v1/model")[0] == 404
        assert get(port, "Bearer sk-alice", path="/v1/models?x=1")[0] == 200
    finally:
        httpd.shutdown()


def test_metrics_listener_and_graceful_shutdown():
    """/metrics over HTTP and Server.GracefullyShutdown (gateway.go:158-173,194-259): an ext_proc stream that is open when the
    shutdown starts is allowed to finish, new ones are refused, the HTTP listeners close."""
    import threading
    import time

    import grpc
    from arks_b200 import metrics
    hm = metrics.HostMetrics()
    hm.record_request("ns", "u", "m", 0.25, 503)
    httpd_m, port_m = extproc.serve_metrics(lambda: hm.exposition(), port=0)
    t = tables()
    httpd, port_h = extproc.serve_http(lambda: t, port=0)
    st, ct, body = get(port_m, path="/metrics")
    assert st == 200 and ct.startswith("text/plain; version=0.0.4") and b'gateway_requests_total{namespace="ns",user="u",model="m",status="503"} 1' in body
    assert get(port_m, path="/other")[0] == 404
    srv = extproc.ExtProcServer(engine=None, tables=t, extract_bearer=lambda hs: b"tok", batcher=object())
    server, port = extproc.serve(srv, port=0)
    ch, stub = extproc.client_stub(port)
    gate = threading.Event()

    def slow_stream():  # a stream that sends its headers, then stays open for a while
        m = extproc.PB["ProcessingRequest"]()
        m.request_headers.end_of_stream = False
        yield m
        gate.wait(5)

    replies = []
    th = threading.Thread(target=lambda: replies.extend(stub(slow_stream())))
    th.start()
    time.sleep(0.3)
    box = {}
    stopper = threading.Thread(target=lambda: box.update(errors=extproc.gracefully_shutdown(server, httpd, httpd_m, timeout_s=5.0)))
    stopper.start()
    time.sleep(0.3)
    with pytest.raises(grpc.RpcError):  # no new streams once the shutdown has begun
        list(stub(iter([extproc.PB["ProcessingRequest"]()]), timeout=2))
    gate.set()  # ... but the open one runs to its end
    th.join(10)
    stopper.join(10)
    assert box["errors"] == [] and len(replies) == 1 and replies[0].WhichOneof("response") == "request_headers"
    with pytest.raises(Exception):
        get(port_h)
    ch.close()
