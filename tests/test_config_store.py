"""Object-level config plane, host half (arks_b200/csrc/config_store.h): upserts / deletes in any order flatten to the
same tables `arks_b200.tables.Tables` builds from the surviving objects in (namespace, name) order, judged by the oracle's
decisions on the same traffic (reference: the informer cache of pkg/gateway/qosconfig/arks_impl.go:104-189)."""
import random

import numpy as np
import orklib
import hostmachine as hm
from hostmachine import ConfigStore

from arks_b200 import traffic
from arks_b200.tables import Tables, endpoint_backends


def key(o):
    return (o["metadata"].get("namespace", "default"), o["metadata"]["name"])


def as_endpoint(e):
    return {"metadata": e["metadata"], "weights": endpoint_backends(e)[1]}


def reference_tables(tokens, quotas, endpoints):
    return Tables(sorted(tokens, key=key), sorted(quotas, key=key), sorted(endpoints, key=key))


def decisions(tables_like, w, now, seed):
    o = orklib.Oracle(tables_like)
    req = w.request_batch(4096, now, seed=seed)
    a = o.request_batch(req)
    return a


def test_store_flattens_like_tables_after_shuffled_upserts_and_deletes():
    w = traffic.Workload(n_tenants=200, seed=11)
    tokens, quotas, endpoints = (list(x) for x in w.objects)
    r = random.Random(5)
    st = ConfigStore()
    events = [("token", t) for t in tokens] + [("quota", q) for q in quotas] + [("endpoint", e) for e in endpoints]
    r.shuffle(events)
    for kind, obj in events:
        st.upsert(kind, as_endpoint(obj) if kind == "endpoint" else obj)
    # re-apply some objects (an update event with the same content), delete a few, put one back
    for kind, obj in r.sample(events, 40):
        st.upsert(kind, as_endpoint(obj) if kind == "endpoint" else obj)
    gone_tok = r.sample(tokens, 7)
    gone_quota = r.sample(quotas, 3)
    for t in gone_tok:
        assert st.erase("token", t)
    for q in gone_quota:
        assert st.erase("quota", q)
    assert not st.erase("token", gone_tok[0])  # not there any more
    st.upsert("token", gone_tok[0])
    live_tok = [t for t in tokens if t not in gone_tok[1:]]
    live_quota = [q for q in quotas if q not in gone_quota]

    ref = reference_tables(live_tok, live_quota, endpoints)
    flat = st.flatten()
    ts = flat.c_struct()
    assert (ts.n_tokens, ts.n_qos, ts.n_quotas, ts.n_endpoints) == (ref.n_tokens, ref.n_qos, ref.n_quotas, ref.n_endpoints)
    assert np.array_equal(np.ctypeslib.as_array(ts.qos_quota, (ts.n_qos,)), ref.qos_quota)  # incl. ARKS_QUOTA_MISSING for the deleted quotas
    assert (ref.qos_quota == -2).any()
    assert np.array_equal(np.ctypeslib.as_array(ts.backend_weight, (ts.n_backends,)), ref.backend_weight)
    now = 1_700_000_000
    a, b = decisions(flat, w, now, 3), decisions(ref, w, now, 3)
    for k, v in a.fields().items():
        assert np.array_equal(v, b.fields()[k]), k


def test_empty_store_and_empty_objects_flatten():
    st = ConfigStore()
    ts = st.flatten().c_struct()
    assert (ts.n_tokens, ts.n_qos, ts.n_quotas, ts.n_endpoints, ts.n_str) == (0, 0, 0, 0, 0)
    st.upsert("token", {"metadata": {"name": "u", "namespace": "n"}, "spec": {"token": "sk-1", "qos": []}})
    st.upsert("quota", {"metadata": {"name": "q", "namespace": "n"}, "spec": {"quotas": []}})
    st.upsert("endpoint", {"metadata": {"name": "m", "namespace": "n"}, "weights": []})
    ts = st.flatten().c_struct()
    assert (ts.n_tokens, ts.n_qos, ts.n_quotas, ts.n_endpoints, ts.n_backends) == (1, 0, 1, 1, 0)
    assert orklib.Oracle(st.flatten()) is not None


def test_snapshot_shape_check_accepts_every_producer_and_names_what_is_wrong():
    """arks_prepare_tables / arks_load_tables check a caller-built snapshot before anything follows its indices
    (config_store.h: tables_shape_error; the reference's objects come validated from the API server). Every producer in this
    repository must pass; a snapshot with one index out of place is refused with the array named."""
    import copy
    import json
    import os
    import numpy as np
    from arks_b200 import traffic
    from arks_b200.tables import Tables
    fx = json.load(open(os.path.join(os.path.dirname(__file__), "golden", "quickstart.json")))
    good = [Tables(fx["tokens"], fx["quotas"], fx["endpoints"]), Tables([], [], []),
            traffic.Workload(n_tenants=300, seed=9).tables, traffic.Workload(50, seed=3, n_backends=16).tables]
    for t in good:
        assert hm.tables_shape_error(t.c_struct()) is None
    # the library's own flatten (object-level config plane), full and empty
    w = traffic.Workload(n_tenants=40, seed=2)
    st = ConfigStore()
    assert hm.tables_shape_error(st.flatten().c_struct()) is None
    for kind, objs in zip(("token", "quota", "endpoint"), w.objects):
        for o in objs:
            st.upsert(kind, as_endpoint(o) if kind == "endpoint" else o)
    assert hm.tables_shape_error(st.flatten().c_struct()) is None

    def broken(field, edit):
        t = copy.deepcopy(good[2])
        edit(getattr(t, field))
        return hm.tables_shape_error(t.c_struct())
    assert "tok_qos_off" in broken("tok_qos_off", lambda a: a.__setitem__(3, a[3] + 1000))
    assert "tok_qos_off" in broken("tok_qos_off", lambda a: a.__setitem__(0, 1))
    assert "qos_rl_off" in broken("qos_rl_off", lambda a: a.__setitem__(-1, a[-1] + 1))
    assert "quota_item_off" in broken("quota_item_off", lambda a: a.__setitem__(5, 0))
    assert "ep_backend_off" in broken("ep_backend_off", lambda a: a.__setitem__(-1, a[-1] - 1))
    assert "string id" in broken("tok_ns_str", lambda a: a.__setitem__(7, 1 << 30))
    assert "string id" in broken("qos_model_str", lambda a: a.__setitem__(0, 0xFFFFFFFF))
    assert "string pool" in broken("str_off", lambda a: a.__setitem__(4, a[5] + 1))
    assert "negative" in broken("backend_weight", lambda a: a.__setitem__(1, -1))
