"""N2, host half: arks_b200.provider.ArksProvider -- watch events -> object upserts -> one generation
per burst, and the quota status loop (pkg/gateway/qosconfig/arks_impl.go:104-189, 217-300).

CPU: the provider drives a stand-in for the Gateway's config-plane methods made of the LIBRARY's own ConfigStore (host build
of csrc/config_store.h) and the oracle for counters and decisions; what it publishes must decide traffic exactly like tables
built from a fresh list of the surviving objects. (The same methods on the GPU: tests/test_gpu_config_plane.py.)"""
import copy
import datetime
import random

import numpy as np
import pytest

import orklib
from hostmachine import ConfigStore
from arks_b200 import abi, traffic
from arks_b200.provider import ArksProvider, object_key
from arks_b200.tables import Tables, simple_endpoint, simple_quota, simple_token

NOW = 1_700_000_000
T0 = datetime.datetime(2026, 1, 2, 3, 4, 5, tzinfo=datetime.timezone.utc)


class StoreGateway:
    """the config-plane methods of arks_b200.gateway.Gateway over ConfigStore + oracle"""

    def __init__(self):
        self.store, self.o, self.tables, self.commits = ConfigStore(), None, None, 0

    def upsert_token(self, ns, name, token, qos):
        self.store.upsert("token", {"metadata": {"namespace": ns, "name": name}, "spec": {"token": token, "qos": [
            {"arksEndpoint": {"name": m}, "quota": {"name": q},
             "rateLimits": [{"type": abi.RULE_NAMES[r], "value": v} for r, v in rls]} for m, q, rls in qos]}})

    def upsert_quota(self, ns, name, items):
        names = sorted(abi.QUOTA_TYPES, key=abi.QUOTA_TYPES.get)
        self.store.upsert("quota", {"metadata": {"namespace": ns, "name": name},
                                    "spec": {"quotas": [{"type": names[t], "value": v} for t, v in items]}})

    def upsert_endpoint(self, ns, name, weights):
        self.store.upsert("endpoint", {"metadata": {"namespace": ns, "name": name}, "weights": list(weights)})

    def delete_object(self, kind, ns, name):
        if not self.store.erase(kind, {"metadata": {"namespace": ns, "name": name}}):
            raise KeyError((kind, ns, name))

    def config_prepare(self):
        return (self.store.flatten(), None)

    def commit_tables(self, prepared):
        flat, names = prepared
        if self.o is None:
            self.o = orklib.Oracle(flat)
        else:
            self.o.reload(flat)  # counters move by (namespace, user, model) / (namespace, quota) like arks_commit_tables
        self.o.tables = names     # row counts for the snapshot helpers
        self.tables = names
        self.commits += 1

    def sync_quota_usage(self, present, used, restore=False):
        return self.o.sync_quota_usage(present, used, restore=restore)


def kinded(kind, o):
    o = copy.deepcopy(o)
    o["kind"] = kind
    return o


def events_of(w):
    tokens, quotas, endpoints = w.objects
    return ([{"type": "ADDED", "object": kinded("ArksToken", t)} for t in tokens]
            + [{"type": "ADDED", "object": kinded("ArksQuota", q)} for q in quotas]
            + [{"type": "ADDED", "object": kinded("ArksEndpoint", e)} for e in endpoints])


def fresh_tables(tokens, quotas, endpoints):
    return Tables(sorted(tokens, key=object_key), sorted(quotas, key=object_key), sorted(endpoints, key=object_key))


class FreshList:
    """the oracle loaded from tables built from a list of the surviving objects (counters carried across reloads by key)"""

    def __init__(self):
        self.o = None

    def load(self, tables):
        if self.o is None:
            self.o = orklib.Oracle(tables)
        else:
            self.o.reload(tables)
        return self


def same_decisions(g, ref, w, seed):
    req = w.request_batch(4096, NOW, seed=seed)
    a, b = g.o.request_batch(req), ref.o.request_batch(req)
    for k, v in a.fields().items():
        assert np.array_equal(v, b.fields()[k]), k
    return a


def test_a_burst_of_events_is_one_generation_that_decides_like_a_fresh_list():
    w = traffic.Workload(n_tenants=300, seed=31)
    ev = events_of(w)
    random.Random(3).shuffle(ev)
    g = StoreGateway()
    p = ArksProvider(g)
    assert not p.flush()  # nothing to publish yet
    assert all(p.apply(e) for e in ev)
    assert p.flush() and g.commits == 1 and p.published == 1 and not p.flush()
    tokens, quotas, endpoints = (list(x) for x in w.objects)
    ref = fresh_tables(tokens, quotas, endpoints)
    assert g.tables.token_user == ref.token_user and g.tables.qos_model_name == ref.qos_model_name
    assert g.tables.endpoint_backends == ref.endpoint_backends and np.array_equal(g.tables.qos_quota, ref.qos_quota)
    fl = FreshList().load(ref)
    a = same_decisions(g, fl, w, 1)
    assert (a.reason == 0).any()

    # what the informers see every 10 s: the provider's own status writes, resyncs of unchanged objects, bookmarks
    q0 = kinded("ArksQuota", quotas[0])
    q0["status"] = {"quotaStatus": [{"type": "total", "used": 5}]}
    q0["metadata"]["resourceVersion"] = "12345"
    assert not p.apply({"type": "MODIFIED", "object": q0})
    assert not p.apply({"type": "MODIFIED", "object": kinded("ArksToken", tokens[0])})
    assert not p.apply({"type": "BOOKMARK", "object": {"kind": "ArksToken", "metadata": {"resourceVersion": "9"}}})
    assert not p.apply({"type": "ADDED", "object": {"kind": "ConfigMap", "metadata": {"name": "x"}}})
    assert not p.flush() and g.commits == 1
    assert p.objects["quota"][object_key(q0)]["status"]["quotaStatus"][0]["used"] == 5  # the cache follows the status

    # a spec change, a deletion of a quota in use (its users answer 500 from now on) and of a token, in one burst
    t1 = kinded("ArksToken", tokens[1])
    for q in t1["spec"]["qos"]:
        q["rateLimits"] = [{"type": "rpm", "value": 1}]
    tokens[1] = t1
    assert p.apply({"type": "MODIFIED", "object": t1})
    dq = quotas.pop(2)
    assert p.apply({"type": "DELETED", "object": kinded("ArksQuota", dq)})
    assert not p.apply({"type": "DELETED", "object": kinded("ArksQuota", dq)})  # already gone
    dt = tokens.pop(5)
    assert p.apply({"type": "DELETED", "object": kinded("ArksToken", dt)})
    assert p.flush() and g.commits == 2
    a = same_decisions(g, fl.load(fresh_tables(tokens, quotas, endpoints)), w, 2)
    assert (a.reason == abi.R_QUOTA_CONFIG).any() and (a.reason == abi.R_TOKEN_NOT_FOUND).any()


def test_relist_deletes_what_is_no_longer_listed():
    w = traffic.Workload(n_tenants=60, seed=32)
    tokens, quotas, endpoints = (list(x) for x in w.objects)
    g = StoreGateway()
    p = ArksProvider(g)
    assert p.replace("ArksToken", [kinded("ArksToken", t) for t in tokens]) == len(tokens)
    p.replace("ArksQuota", [kinded("ArksQuota", q) for q in quotas])
    p.replace("ArksEndpoint", [kinded("ArksEndpoint", e) for e in endpoints])
    assert p.flush()
    fl = FreshList().load(fresh_tables(tokens, quotas, endpoints))
    same_decisions(g, fl, w, 1)
    # the watch expired; the new list misses ten tokens and changes one endpoint's weights
    tokens = tokens[10:]
    e0 = copy.deepcopy(endpoints[0])
    e0["spec"]["routeConfigs"] = [{"name": "a", "weight": 3}, {"name": "b", "weight": 0}]
    endpoints[0] = e0
    assert p.replace("ArksToken", [kinded("ArksToken", t) for t in tokens]) == 10
    assert p.replace("ArksEndpoint", [kinded("ArksEndpoint", e) for e in endpoints]) == 1
    assert p.replace("ArksQuota", [kinded("ArksQuota", q) for q in quotas]) == 0
    assert p.flush() and g.commits == 2
    same_decisions(g, fl.load(fresh_tables(tokens, quotas, endpoints)), w, 2)
    assert g.tables.endpoint_backends[sorted(map(object_key, endpoints)).index(object_key(e0))] == ["a", "b"]


def test_an_object_the_gateway_cannot_read_keeps_its_previous_version():
    g = StoreGateway()
    p = ArksProvider(g)
    tok = kinded("ArksToken", simple_token("alice", "default", "sk-a", "m", [("rpm", 2)], quota="q"))
    for o in (tok, kinded("ArksQuota", simple_quota("q", "default", [("total", 10)])),
              kinded("ArksEndpoint", simple_endpoint("m", "default", 1, [("b0", 1)]))):
        assert p.apply({"type": "ADDED", "object": o})
    assert p.flush()
    bad = copy.deepcopy(tok)
    bad["spec"]["qos"][0]["rateLimits"] = [{"type": "rps", "value": 1}]  # ratelimiter/types.go:46 panics on such a unit
    assert not p.apply({"type": "MODIFIED", "object": bad})
    assert p.rejected and p.rejected[0][:2] == ("token", ("default", "alice")) and not p.flush()
    assert g.tables.qos_rule_names == [["rpm"]]


def _run(g, w, seed, now):
    req = w.request_batch(2048, now, seed=seed)
    a = g.o.request_batch(req)
    g.o.response_batch(w.response_batch(a, now, seed=seed + 1))


def test_quota_status_loop():
    w = traffic.Workload(n_tenants=40, seed=33)
    g = StoreGateway()
    clock = [T0]
    p = ArksProvider(g, clock=lambda: clock[0])
    for e in events_of(w):
        p.apply(e)
    p.flush()
    assert len(p.sync_quota_status()) == g.tables.n_quotas  # no status yet: every spec type is appended (with 0 used)
    _run(g, w, 1, NOW)
    usage = g.o.snapshot_quota()
    assert usage.sum() > 0
    patches = p.sync_quota_status()
    by_key = {(u["namespace"], u["name"]): u for u in patches}
    t = g.tables
    for q in range(t.n_quotas):
        k = (t.strings[t.quota_ns_str[q]].decode(), t.strings[t.quota_name_str[q]].decode())
        o = p.objects["quota"][k]
        want = {}
        for it in o["spec"]["quotas"]:
            want.setdefault(it["type"], int(usage[q, abi.QUOTA_TYPES[it["type"]]]))
        got = o["status"]["quotaStatus"]
        assert [s["type"] for s in got] == list(want) and [s["used"] for s in got] == list(want.values())  # spec order, one per type
        if any(want.values()):
            assert by_key[k]["status"]["quotaStatus"] == got
            assert all(s["lastUpdateTime"] == "2026-01-02T03:04:05Z" for s in got if s["used"])
    # nothing moved: nothing to send
    assert p.sync_quota_status() == []
    # more usage, ten seconds later: only the entries that moved carry the new time
    clock[0] = T0 + datetime.timedelta(seconds=10)
    before = copy.deepcopy({k: o.get("status") for k, o in p.objects["quota"].items()})
    _run(g, w, 5, NOW + 10)
    patches = p.sync_quota_status()
    assert patches
    for u in patches:
        old = {s["type"]: s for s in before[(u["namespace"], u["name"])]["quotaStatus"]}
        for s in u["status"]["quotaStatus"]:
            moved = s["used"] != old[s["type"]]["used"]
            assert s["used"] >= old[s["type"]]["used"]
            assert s["lastUpdateTime"] == ("2026-01-02T03:04:15Z" if moved else old[s["type"]]["lastUpdateTime"])
    assert np.array_equal(g.o.snapshot_quota() >= usage, np.ones_like(usage, bool))


@pytest.mark.parametrize("restore", [False, True])
def test_status_ahead_of_the_counters(restore):
    """a restart: the CRs remember usage, the counters are empty. The reference's loop zeroes (SetUsage with Request == 0,
    arks_impl.go:286-288); restore=True is the start-up pass that raises the counters to the CR"""
    g = StoreGateway()
    p = ArksProvider(g, clock=lambda: T0)
    q = kinded("ArksQuota", simple_quota("q", "default", [("prompt", 100), ("total", 1000)]))
    q["status"] = {"quotaStatus": [{"type": "total", "used": 700, "lastUpdateTime": "2025-12-31T00:00:00Z"},
                                   {"type": "total", "used": 9999}]}  # a second entry of a type is never looked at
    for o in (q, kinded("ArksToken", simple_token("alice", "default", "sk-a", "m", [("rpm", 2)], quota="q")),
              kinded("ArksEndpoint", simple_endpoint("m", "default", 1, [("b0", 1)]))):
        p.apply({"type": "ADDED", "object": o})
    p.flush()
    g.o.set_quota_usage(0, [3, 0, 5])
    patches = p.sync_quota_status(restore=restore)
    # prompt has no status entry yet: appended from the counters; total stays what the CR says
    assert len(patches) == 1
    st = patches[0]["status"]["quotaStatus"]
    assert [(s["type"], s["used"]) for s in st] == [("total", 700), ("total", 9999), ("prompt", 3)]
    assert st[0]["lastUpdateTime"] == "2025-12-31T00:00:00Z" and st[2]["lastUpdateTime"] == "2026-01-02T03:04:05Z"
    assert g.o.snapshot_quota()[0].tolist() == ([3, 0, 700] if restore else [0, 0, 0])


def test_loop_coalesces_bursts_and_ticks_the_status_every_ten_seconds():
    from arks_b200.provider import ProviderLoop
    w = traffic.Workload(n_tenants=30, seed=34)
    g = StoreGateway()
    p = ArksProvider(g, clock=lambda: T0)
    sent = []
    loop = ProviderLoop(p, write_status=sent.append, debounce_s=0.05, max_delay_s=1.0, sync_every_s=10.0)
    ev = events_of(w)
    for e in ev[:50]:
        loop.offer(e)
    assert loop.step(100.0) == pytest.approx(100.05) and g.commits == 0  # waits for the burst to end
    for e in ev[50:]:
        loop.offer(e)
    assert loop.step(100.04) == pytest.approx(100.09) and g.commits == 0
    loop.step(100.095)
    assert g.commits == 1 and g.tables.n_tokens == 30  # one generation for the whole burst
    # a steady trickle (an event every 40 ms) is published after max_delay all the same
    tokens = [kinded("ArksToken", t) for t in w.objects[0]]
    t = 101.0
    for i in range(30):
        tk = copy.deepcopy(tokens[i])
        tk["spec"]["qos"][0]["rateLimits"] = [{"type": "rpm", "value": 1000 + i}]
        loop.offer({"type": "MODIFIED", "object": tk})
        loop.step(t)
        t += 0.04
    assert g.commits == 2
    loop.step(t + 0.05)
    assert g.commits == 3 and not p.dirty
    # the ticker: nothing before ten seconds, restore mode on the first pass only
    assert sent == [] and loop.restore_next
    _run(g, w, 1, NOW)
    loop.step(109.9)
    assert sent == []
    loop.step(110.0)
    assert len(sent) == 1 and len(sent[0]) == g.tables.n_quotas and not loop.restore_next
    loop.step(115.0)
    loop.step(120.0)
    assert len(sent) == 1  # nothing moved, nothing sent
    _run(g, w, 3, NOW + 120)  # the next minute: rpm has room again
    loop.step(130.0)
    assert len(sent) == 2
    # run(): the same loop on a thread
    import threading
    th = threading.Thread(target=loop.run)
    th.start()
    tk = copy.deepcopy(tokens[0])
    tk["spec"]["qos"][0]["rateLimits"] = [{"type": "rpm", "value": 123}]
    loop.offer({"type": "MODIFIED", "object": tk})
    import time
    deadline = time.monotonic() + 5
    while g.commits < 4 and time.monotonic() < deadline:
        time.sleep(0.01)
    loop.stop.set()
    th.join()
    assert g.commits == 4 and g.tables.qos_rule_names[sorted(map(object_key, tokens)).index(object_key(tk))] == ["rpm"]


@pytest.mark.gpu
def test_provider_over_the_library(gwmod):
    """the same events through the real Gateway (ConfigStore + generations + counters in HBM) and through the stand-in:
    decisions, usage and the status updates agree"""
    w = traffic.Workload(n_tenants=500, seed=35)
    gpu, cpu = gwmod.Gateway(0, 8192, 64 << 20), StoreGateway()
    pg, pc = ArksProvider(gpu, clock=lambda: T0), ArksProvider(cpu, clock=lambda: T0)
    ev = events_of(w)
    random.Random(9).shuffle(ev)
    tokens = [kinded("ArksToken", t) for t in w.objects[0]]
    quotas = [kinded("ArksQuota", q) for q in w.objects[1]]
    t1 = copy.deepcopy(tokens[3])
    t1["spec"]["qos"][0]["rateLimits"] = [{"type": "rpm", "value": 1}]
    bursts = [ev, [{"type": "MODIFIED", "object": t1}, {"type": "DELETED", "object": quotas[7]},
                   {"type": "DELETED", "object": tokens[11]}, {"type": "MODIFIED", "object": tokens[5]}]]
    for step, burst in enumerate(bursts):
        for e in burst:
            assert pg.apply(copy.deepcopy(e)) == pc.apply(copy.deepcopy(e))
        assert pg.flush() and pc.flush()
        assert gpu.tables.token_user == cpu.tables.token_user and gpu.tables.n_quotas == cpu.tables.n_quotas
        now = NOW + 61 * step
        req = w.request_batch(8192, now, seed=step + 1)
        a, b = gpu.handle_request_body(req), cpu.o.request_batch(req)
        for k, v in a.fields().items():
            assert np.array_equal(v, b.fields()[k]), (step, k)
        resp = w.response_batch(a, now + 1, seed=step + 50)
        c, d = gpu.handle_response_body(resp), cpu.o.response_batch(resp)
        for k, v in c.fields().items():
            assert np.array_equal(v, d.fields()[k]), (step, k)
        assert np.array_equal(gpu.snapshot_quota(), cpu.o.snapshot_quota())
        ug, uc = pg.sync_quota_status(restore=step == 0), pc.sync_quota_status(restore=step == 0)
        assert ug == uc and len(ug) > 0
    assert (b.reason == abi.R_QUOTA_CONFIG).any()


class LiveEngine(StoreGateway):
    """StoreGateway + the two batch calls + a generation counter, i.e. what the Python ext_proc server needs of a Gateway.
    Response rows of an older generation are re-mapped by (namespace, user, model) like the library does."""

    def __init__(self):
        super().__init__()
        self.generation, self.names = 0, {}

    def commit_tables(self, prepared):
        super().commit_tables(prepared)
        self.generation += 1
        self.names[self.generation] = prepared[1]

    def handle_request_body(self, b):
        return self.o.request_batch(b)

    def snapshot_metrics(self):
        return self.o.snapshot_metrics()

    def handle_response_body(self, b):
        cur = self.names[self.generation]
        key = lambda t, q: (t.token_namespace[int(t.qos_token[q])], t.token_user[int(t.qos_token[q])], t.qos_model_name[q])
        index = {key(cur, q): q for q in range(cur.n_qos)}
        for i in range(b.n):
            if b.gen is not None and int(b.gen[i]) != self.generation and b.qos[i] >= 0:
                b.qos[i] = index.get(key(self.names[int(b.gen[i])], int(b.qos[i])), -1)
        return self.o.response_batch(b)


def test_ext_proc_server_follows_the_generations():
    """the names in routing headers and metric labels are those of the generation the request was decided on, while the
    provider publishes new generations between batches of a running server"""
    import json
    import os
    import __graft_entry__ as ge
    from arks_b200 import extproc, gateway
    from test_extproc_loopback import body, hdrs, resp_hdrs, set_headers
    ge.build()
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quickstart.json")))
    eng = LiveEngine()
    srv = extproc.ExtProcServer(eng, None, gateway.extract_bearer, clock=lambda: NOW)
    p = ArksProvider(eng, publish=srv.publisher(eng))
    for kind, key in (("ArksToken", "tokens"), ("ArksQuota", "quotas"), ("ArksEndpoint", "endpoints")):
        p.replace(kind, [kinded(kind, o) for o in fx[key]])
    assert p.flush() and eng.generation == 1 and srv.tables.token_user == ["example-token"]
    server, port = extproc.serve(srv, port=0)
    ch, stub = extproc.client_stub(port)

    def stream(token, split=None):
        msgs = [hdrs([("authorization", "Bearer " + token)]), body(fx["request_body"].encode(), "request_body"),
                resp_hdrs([(":status", "200")]), body(fx["response_body"].encode(), "response_body")]
        if split is None:
            return list(stub(iter(msgs)))

        def gen():  # the configuration changes while this stream waits for its upstream
            yield msgs[0]
            yield msgs[1]
            split()
            yield msgs[2]
            yield msgs[3]
        return list(stub(gen()))

    try:
        r = stream("sk-test123456")
        assert set_headers(r[1].request_body.response.header_mutation)["username"] == "example-token"
        # a second user whose name sorts first: every index of the old generation now means somebody else
        adam = kinded("ArksToken", simple_token("adam", "default", "sk-adam", "qwen-7b", [("rpm", 10)], quota="basic-quota"))

        def add_adam():
            assert p.apply({"type": "ADDED", "object": adam}) and p.flush()
        r = stream("sk-test123456", split=add_adam)
        assert eng.generation == 2 and srv.tables.token_user == ["adam", "example-token"]
        assert set_headers(r[1].request_body.response.header_mutation)["username"] == "example-token"
        assert r[3].WhichOneof("response") == "response_body"
        r = stream("sk-adam")
        assert set_headers(r[1].request_body.response.header_mutation) == {"model": "qwen-7b", "namespace": "default", "username": "adam"}
        r = stream("sk-test123456")
        assert set_headers(r[1].request_body.response.header_mutation)["username"] == "example-token"
        # usage went to the right rows (three streams of example-token, one of adam; they share the quota)
        rate = eng.o.snapshot_rate(NOW)
        assert rate[:, 0].tolist() == [1, 3] and eng.o.snapshot_quota()[0].tolist() == [100, 80, 180]
        text = srv.metrics.exposition()
        assert 'gateway_request_duration_seconds_count{namespace="default",user="example-token",model="qwen-7b"} 3' in text
        assert 'gateway_request_duration_seconds_count{namespace="default",user="adam",model="qwen-7b"} 1' in text
        # the token is deleted: its next request is refused with the reference's reply
        assert p.apply({"type": "DELETED", "object": adam}) and p.flush() and eng.generation == 3
        r = stream("sk-adam")
        assert r[1].immediate_response.status.code == 500  # "error to get qos by token", handle_request.go:117-125
        assert "x-error-token" in set_headers(r[1].immediate_response.headers)
    finally:
        ch.close()
        extproc.gracefully_shutdown(server)
        srv.batcher.close()


def test_generations_published_under_load_never_mix_names():
    """client streams keep running while the provider publishes a generation every few milliseconds (a user that sorts first
    comes and goes, so every index flips each time): each stream's routing headers name the user its token belongs to"""
    import json
    import os
    import threading
    import time
    import __graft_entry__ as ge
    from arks_b200 import extproc, gateway
    from test_extproc_loopback import body, hdrs, resp_hdrs, set_headers
    ge.build()
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quickstart.json")))
    eng = LiveEngine()
    srv = extproc.ExtProcServer(eng, None, gateway.extract_bearer, clock=lambda: NOW)
    p = ArksProvider(eng, publish=srv.publisher(eng))
    tok = kinded("ArksToken", fx["tokens"][0])
    tok["spec"]["qos"][0]["rateLimits"] = []  # nothing to run into: every stream is admitted
    tok["spec"]["qos"][0]["quota"] = {"name": ""}
    users = [dict(copy.deepcopy(tok), metadata={"name": f"user-{i}", "namespace": "default"}) for i in range(4)]
    for i, u in enumerate(users):
        u["spec"]["token"] = f"sk-{i}"
    p.replace("ArksToken", users)
    p.replace("ArksEndpoint", [kinded("ArksEndpoint", o) for o in fx["endpoints"]])
    assert p.flush()
    server, port = extproc.serve(srv, port=0)
    stop, bad, done = threading.Event(), [], [0]

    def client(i):
        ch, stub = extproc.client_stub(port)
        try:
            while not stop.is_set():
                r = list(stub(iter([hdrs([("authorization", f"Bearer sk-{i}")]), body(fx["request_body"].encode(), "request_body"),
                                    resp_hdrs([(":status", "200")]), body(fx["response_body"].encode(), "response_body")])))
                h = set_headers(r[1].request_body.response.header_mutation) if r[1].WhichOneof("response") == "request_body" else r[1]
                if h != {"model": "qwen-7b", "namespace": "default", "username": f"user-{i}"} or len(r) != 4:
                    bad.append((i, h))
                done[0] += 1
        except Exception as e:  # noqa: BLE001
            bad.append((i, repr(e)))
        finally:
            ch.close()

    threads = [threading.Thread(target=client, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    first = dict(copy.deepcopy(users[0]), metadata={"name": "aaa-first", "namespace": "default"})
    first["spec"]["token"] = "sk-first"
    end, flips = time.monotonic() + 1.5, 0
    try:
        while time.monotonic() < end:
            assert p.apply({"type": "ADDED" if flips % 2 == 0 else "DELETED", "object": first}) and p.flush()
            flips += 1
            time.sleep(0.003)
    finally:
        stop.set()
        for t in threads:
            t.join()
        extproc.gracefully_shutdown(server)
        srv.batcher.close()
    assert not bad, bad[:3]
    assert flips > 20 and done[0] > 20 and eng.generation == flips + 1
    rate = eng.o.snapshot_rate(NOW)
    assert rate.shape[0] == eng.tables.n_qos


def test_loop_survives_a_refused_generation_and_retries():
    from arks_b200.provider import ProviderLoop
    import threading
    import time
    g = StoreGateway()
    fail = [2]
    real = g.commit_tables

    def flaky(prepared):
        if fail[0]:
            fail[0] -= 1
            raise RuntimeError("commit refused")
        real(prepared)
    g.commit_tables = flaky
    p = ArksProvider(g, publish=lambda names: g.commit_tables((g.config_prepare()[0], names)))
    loop = ProviderLoop(p, debounce_s=0.01)
    loop.offer({"type": "ADDED", "object": kinded("ArksToken", simple_token("alice", "default", "sk-a", "m", [("rpm", 2)]))})
    with pytest.raises(RuntimeError):
        loop.step(1.0) and loop.step(1.02)
    assert p.dirty and g.commits == 0  # nothing was lost
    th = threading.Thread(target=loop.run)
    th.start()
    end = time.monotonic() + 10
    while g.commits == 0 and time.monotonic() < end:
        time.sleep(0.05)
    loop.stop.set()
    th.join()
    assert g.commits == 1 and len(loop.errors) == 1 and g.tables.token_user == ["alice"]
