"""The config plane without a data-path stall: arks_prepare_tables / arks_commit_tables and the object-level
arks_upsert_* / arks_delete_* + arks_config_prepare, interleaved with 64 k-request waves, against the oracle reloaded at
the same points. Reference: informer events arriving while requests are in flight (qosconfig/arks_impl.go:104-189)."""
import copy
import threading

import numpy as np
import pytest

import orklib
from arks_b200 import abi, traffic
from arks_b200.abi import RequestResult
from arks_b200.tables import Tables, endpoint_backends

pytestmark = pytest.mark.gpu
NOW = 1_700_000_000
WAVE = 65536


def same(a, b, ctx=""):
    for k, v in a.fields().items():
        w = b.fields()[k]
        if not np.array_equal(v, w):
            bad = np.nonzero((v != w).reshape(len(v), -1).any(axis=1))[0]
            raise AssertionError(f"{ctx} field {k}: {len(bad)} mismatches, first at {bad[:5]}: gpu={v[bad[:5]]} oracle={w[bad[:5]]}")


def state_same(g, o, now):
    assert np.array_equal(g.snapshot_rate(now), o.snapshot_rate(now)), "rate counters differ"
    assert np.array_equal(g.snapshot_quota(), o.snapshot_quota()), "quota usage differs"


def key(o):
    return (o["metadata"].get("namespace", "default"), o["metadata"]["name"])


def sorted_tables(tokens, quotas, endpoints):
    return Tables(sorted(tokens, key=key), sorted(quotas, key=key), sorted(endpoints, key=key))


def push(g, kind, obj):
    md = obj["metadata"]
    ns, name = md.get("namespace", "default"), md["name"]
    if kind == "token":
        g.upsert_token(ns, name, obj["spec"]["token"],
                       [(q["arksEndpoint"]["name"], (q.get("quota") or {}).get("name", ""),
                         [(abi.RULES[r["type"]], int(r["value"])) for r in q.get("rateLimits") or []]) for q in obj["spec"].get("qos") or []])
    elif kind == "quota":
        g.upsert_quota(ns, name, [(abi.QUOTA_TYPES[i["type"]], int(i["value"])) for i in obj["spec"]["quotas"]])
    else:
        g.upsert_endpoint(ns, name, endpoint_backends(obj)[1])


def wave(g, o, w, now, seed, ctx):
    req = w.request_batch(WAVE, now, seed=seed)
    a = g.handle_request_body(req)
    same(a, o.request_batch(req), ctx + " requests")
    resp = w.response_batch(a, now, seed=seed + 1000)
    same(g.handle_response_body(resp), o.response_batch(resp), ctx + " responses")
    return a


def test_prepare_commit_between_waves_and_behind_a_queued_batch(gwmod):
    w = traffic.Workload(10_000, seed=21)
    g = gwmod.Gateway(0, WAVE, int(WAVE * 1200))
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    gen0 = g.generation
    wave(g, o, w, NOW, 1, "gen 0")
    # next generation: some tenants gone, limits changed (another seed reshuffles them), prepared while nothing waits
    w2 = traffic.Workload(9_000, seed=21)
    stale = g.prepare_tables(w2.tables)
    p = g.prepare_tables(w2.tables)
    assert g.generation == gen0  # nothing visible yet
    wave(g, o, w, NOW + 1, 2, "prepared, not committed")
    # the commit goes in BEHIND a batch that is already queued: that batch is decided on the old tables
    req = w.request_batch(WAVE, NOW + 2, seed=3)
    g.select_slot(0)
    g.submit_request_async(req)
    g.commit_tables(p)
    assert g.generation == gen0 + 1
    a = g.wait_request(0, RequestResult.empty(req.n))
    same(a, o.request_batch(req), "queued before the commit")
    o.reload(w2.tables)
    state_same(g, o, NOW + 2)  # every counter arrived in the new arrays, by key
    wave(g, o, w2, NOW + 3, 4, "gen 1")
    state_same(g, o, NOW + 3)
    # a handle prepared against an older generation is refused and can be dropped
    with pytest.raises(Exception):
        g.commit_tables(stale)
    g.discard_prepared(stale)
    assert g.generation == gen0 + 1
    wave(g, o, w2, NOW + 4, 5, "after the refused commit")
    state_same(g, o, NOW + 4)


def test_object_level_upserts_and_deletes_interleaved_with_waves(gwmod):
    w = traffic.Workload(4_000, seed=22)
    tokens, quotas, endpoints = (list(x) for x in w.objects)
    g = gwmod.Gateway(0, WAVE, int(WAVE * 1200))
    for t in tokens:
        push(g, "token", t)
    for q in quotas:
        push(g, "quota", q)
    for e in endpoints:
        push(g, "endpoint", e)
    ref = sorted_tables(tokens, quotas, endpoints)
    g.commit_tables(g.config_prepare())
    g.tables = ref
    o = orklib.Oracle(ref)
    wave(g, o, w, NOW, 1, "store gen 1")
    state_same(g, o, NOW)
    rng = np.random.default_rng(7)
    for step in range(3):
        # a handful of informer events: tighter limits on some tokens, a quota deleted (its users now answer 500),
        # a token deleted, one brought back, an endpoint's weights changed
        for i in rng.choice(len(tokens), 20, replace=False):
            t = copy.deepcopy(tokens[i])
            for q in t["spec"]["qos"]:
                for r in q.get("rateLimits") or []:
                    r["value"] = max(1, int(r["value"]) // 2)
            tokens[i] = t
            push(g, "token", t)
        dq = quotas.pop(int(rng.integers(len(quotas))))
        g.delete_object("quota", *key(dq))
        dt = tokens.pop(int(rng.integers(len(tokens))))
        g.delete_object("token", *key(dt))
        if step == 1:
            tokens.append(dt)
            push(g, "token", dt)
        e = copy.deepcopy(endpoints[int(rng.integers(len(endpoints)))])
        for rc in e["spec"].get("routeConfigs") or []:
            rc["weight"] = int(rng.integers(0, 5))
        endpoints = [e if key(x) == key(e) else x for x in endpoints]
        push(g, "endpoint", e)
        with pytest.raises(Exception):
            g.delete_object("token", "no-such-namespace", "nobody")
        ref = sorted_tables(tokens, quotas, endpoints)
        g.commit_tables(g.config_prepare())
        g.tables = ref
        o.reload(ref)
        state_same(g, o, NOW + step)
        a = wave(g, o, w, NOW + step, 10 + step, f"store step {step}")
        assert (a.reason == abi.R_QUOTA_CONFIG).any() or step == 0
        state_same(g, o, NOW + step)
    q0 = ref.token_namespace[0], ref.token_user[0], ref.qos_model_name[0]
    assert g.find_qos(*q0) == 0 and g.find_qos("x", "y", "z") == -1
    assert g.find_quota(*key(quotas[0])) >= 0 and g.find_quota(*key(dq)) == -1


def test_prepare_on_a_config_thread_while_batches_run(gwmod):
    """the build runs on its own thread and stream; decisions made meanwhile are the old generation's, bit for bit"""
    w = traffic.Workload(10_000, seed=23)
    w2 = traffic.Workload(10_000, seed=24)
    g = gwmod.Gateway(0, WAVE, int(WAVE * 1200))
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    box = {}

    def config_thread():
        box["p"] = [g.prepare_tables(w2.tables) for _ in range(3)]

    th = threading.Thread(target=config_thread)
    th.start()
    for k in range(6):
        wave(g, o, w, NOW + k, 30 + k, f"while preparing {k}")
    th.join()
    for p in box["p"][:-1]:
        g.discard_prepared(p)
    g.commit_tables(box["p"][-1])
    o.reload(w2.tables)
    state_same(g, o, NOW + 6)
    wave(g, o, w2, NOW + 6, 40, "after the threaded prepare")
    state_same(g, o, NOW + 6)
