"""`python -m arks_b200` (arks_b200/__main__.py), the counterpart of cmd/gateway/main.go:171-231: provider first, then the three
listeners, graceful shutdown. CPU: everything except the engine is the product's; the engine is the stand-in of
tests/test_watch_provider.py (the library's ConfigStore + the oracle). Without a CUDA device main() itself must refuse."""
import copy
import io
import json
import os
import subprocess
import sys
import time

import pytest
import yaml

from arks_b200 import extproc
from arks_b200.__main__ import Assembly, parse_args, read_objects
from arks_b200.extproc import HEALTH_PB as HP
from arks_b200.tables import simple_token
from test_extproc_loopback import body, hdrs, resp_hdrs, set_headers
from test_watch_provider import LiveEngine, kinded
from test_wire_surface import get

HERE = os.path.dirname(os.path.abspath(__file__))
FX = json.load(open(os.path.join(HERE, "golden", "quickstart.json")))


def crds(skip=()):
    out = []
    for kind, key in (("ArksToken", "tokens"), ("ArksQuota", "quotas"), ("ArksEndpoint", "endpoints")):
        out += [dict(kinded(kind, o), apiVersion="arks.ai/v1") for o in FX[key] if o["metadata"]["name"] not in skip]
    return out


def wait_for(cond, what, timeout=10.0):
    end = time.monotonic() + timeout
    while not cond():
        assert time.monotonic() < end, what
        time.sleep(0.02)


def test_flags_keep_the_reference_names_and_defaults():
    a = parse_args([])
    assert (a.grpc_port, a.http_port, a.metrics_port) == (50052, 8080, 9110)  # cmd/gateway/main.go:97-98,115
    a = parse_args(["--server.grpc-port", "1", "--server.http-port", "2", "--metrics.port", "3"])
    assert (a.grpc_port, a.http_port, a.metrics_port) == (1, 2, 3)


def test_objects_file_forms(tmp_path):
    y = tmp_path / "o.yaml"
    y.write_text(yaml.safe_dump_all(crds() + [{"apiVersion": "v1", "kind": "ConfigMap", "metadata": {"name": "x"}}]))
    j = tmp_path / "o.json"
    j.write_text(json.dumps({"kind": "List", "items": crds()}))
    assert [o["kind"] for o in read_objects(str(y))] == ["ArksToken", "ArksQuota", "ArksEndpoint"]
    assert read_objects(str(j)) == read_objects(str(y))


def test_the_assembled_process(tmp_path):
    objects, events = tmp_path / "objects.yaml", tmp_path / "events.jsonl"
    objects.write_text(yaml.safe_dump_all(crds()))
    adam = dict(kinded("ArksToken", simple_token("adam", "default", "sk-adam", "qwen-7b", [("rpm", 10)], quota="basic-quota")),
                apiVersion="arks.ai/v1")
    events.write_text(json.dumps({"type": "ADDED", "object": adam}) + "\n\nnot json\n"
                      + json.dumps({"type": "BOOKMARK", "object": {"kind": "ArksToken", "metadata": {"resourceVersion": "7"}}}) + "\n")
    args = parse_args(["--server.grpc-port", "0", "--server.http-port", "0", "--metrics.port", "0", "--server.bind", "127.0.0.1",
                       "--provider.objects", str(objects), "--provider.events", str(events), "--batcher", "python"])
    status = io.StringIO()
    eng = LiveEngine()
    asm = Assembly(eng, args, status_out=status)
    asm.loop.sync_every_s = 0.3
    asm.start()
    try:
        assert eng.generation >= 1  # the first generation was there before the listeners
        wait_for(lambda: asm.srv.tables.token_user == ["adam", "example-token"], "the event stream was not applied")
        ch, stub = extproc.client_stub(asm.grpc_port)
        r = list(stub(iter([hdrs([("authorization", "Bearer sk-adam")]), body(FX["request_body"].encode(), "request_body"),
                            resp_hdrs([(":status", "200")]), body(FX["response_body"].encode(), "response_body")])))
        assert set_headers(r[1].request_body.response.header_mutation)["username"] == "adam"
        check = ch.unary_unary("/grpc.health.v1.Health/Check", request_serializer=HP["HealthCheckRequest"].SerializeToString,
                               response_deserializer=HP["HealthCheckResponse"].FromString)
        assert check(HP["HealthCheckRequest"](), timeout=10).status == extproc.SERVING
        st, ct, text = get(asm.http_port, auth="Bearer sk-test123456")
        assert st == 200 and json.loads(text)["data"][0]["id"] == "qwen-7b"
        st, ct, text = get(asm.metrics_port, path="/metrics")
        assert st == 200 and b'gateway_request_duration_seconds_count{namespace="default",user="adam",model="qwen-7b"} 1' in text
        assert b'gateway_requests_total{namespace="default",user="adam",model="qwen-7b",status="200"}' in text  # the engine's rows
        # the status loop: the restore pass ran before the listeners opened (it entered the quota's types into the status, all
        # zero); the ticker then writes one line per ArksQuota whose status moved
        assert not asm.loop.restore_next and '"used":0' in status.getvalue().splitlines()[0]
        wait_for(lambda: '"used":45' in status.getvalue(), "no status update was written")
        line = json.loads(status.getvalue().splitlines()[-1])
        assert (line["kind"], line["metadata"]) == ("ArksQuota", {"namespace": "default", "name": "basic-quota"})
        assert {s["type"]: s["used"] for s in line["status"]["quotaStatus"]} == {"prompt": 25, "response": 20, "total": 45}
        # the list changes on disk: example-token is gone (adam is now part of the list)
        objects.write_text(yaml.safe_dump_all(crds(skip=("example-token",)) + [adam]))
        os.utime(objects, (time.time() + 5, time.time() + 5))
        wait_for(lambda: asm.srv.tables.token_user == ["adam"], "the relist was not applied")
        st, ct, text = get(asm.http_port, auth="Bearer sk-test123456")
        assert st == 500  # GetModelsByToken fails for a token that no longer exists
        ch.close()
    finally:
        assert asm.shutdown() == []


def test_main_refuses_without_a_cuda_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    p = subprocess.run([sys.executable, "-m", "arks_b200", "--server.grpc-port", "0"], capture_output=True, text=True, timeout=300,
                       cwd=os.path.dirname(HERE))
    assert p.returncode != 0 and "no CPU fallback" in p.stderr


def test_a_restart_restores_the_quota_counters_before_the_first_request(tmp_path):
    """ArksQuota.status carries what the previous process billed; with --provider.restore-on-start (default) the counters are
    raised to it before the listeners open, so the very first request already meets the quota (the reference's first pass comes
    ten seconds later and zeroes the store instead, arks_impl.go:217-225,286-288)"""
    objs = crds()
    for o in objs:
        if o["kind"] == "ArksQuota":
            limits = {i["type"]: int(i["value"]) for i in o["spec"]["quotas"]}
            o["status"] = {"quotaStatus": [{"type": t, "used": v + 1, "lastUpdateTime": "2025-01-01T00:00:00Z"} for t, v in limits.items()]}
    objects = tmp_path / "objects.yaml"
    objects.write_text(yaml.safe_dump_all(objs))
    args = parse_args(["--server.grpc-port", "0", "--server.http-port", "0", "--metrics.port", "0", "--server.bind", "127.0.0.1",
                       "--provider.objects", str(objects), "--batcher", "python"])
    asm = Assembly(LiveEngine(), args, status_out=io.StringIO()).start()
    try:
        assert not asm.loop.restore_next
        ch, stub = extproc.client_stub(asm.grpc_port)
        r = list(stub(iter([hdrs([("authorization", "Bearer sk-test123456")]), body(FX["request_body"].encode(), "request_body")])))
        im = r[1].immediate_response
        assert im.status.code == 429 and "x-error-quota" in set_headers(im.headers)
        ch.close()
    finally:
        assert asm.shutdown() == []
