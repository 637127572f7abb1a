"""A second, independent statement of the reference's decision logic, and the oracle held to it on random scenarios.

`oracle/ork_core.c` restates check.go / redis_impl.go / handle_*.go in C over dense counter arrays; its fixtures come from the
reference (tests/test_oracle_golden.py), but the reference's own tests pin almost nothing of the decision logic (SURVEY.md
section 8c). This file is the same logic written the way the Go code is written -- a string-keyed Redis (a dict: GET of a missing
key = 0, INCRBY creates at 0), an informer cache of objects, the handlers in the order of their Go statements -- and shares no
code or data layout with the oracle: document parsing comes from `json.loads` (tests/pymodel.py), keys are the reference's
strings. Tables with shared quotas, missing ArksQuotas, duplicated quota items, tokens without limits, several tokens per
namespace and window roll-overs are drawn at random; every decision, counter value carried in a 429 and every Redis value at
the end must agree. TEST INFRASTRUCTURE.

Reference (paths relative to the reference tree):
  HandleRequestBody          pkg/gateway/handle_request.go:83-249
  HandleResponseBody         pkg/gateway/handle_response.go:80-268
  checkRateLimit & friends   pkg/gateway/check.go:31-156
  CheckLimit / DoLimit       pkg/gateway/ratelimiter/redis_impl.go:47-168, rules rate_limiter.go:31-68, key cache_key.go:42-80
  GetUsage / IncrUsage       pkg/gateway/quota/redis_impl.go:38-107, key cache_key.go:40-58
  GetQosByToken ...          pkg/gateway/qosconfig/arks_impl.go:303-376, QosToQuotaRequests types.go:45-72
"""
import json
import random

import numpy as np
import pytest

import orklib
import pymodel
from arks_b200 import abi
from arks_b200.abi import RequestBatch, ResponseBatch
from arks_b200.tables import Tables, simple_endpoint, simple_quota

RULES = {"rpm": ("request", 60), "rpd": ("request", 86400), "tpm": ("token", 60), "tpd": ("token", 86400)}  # rate_limiter.go:31-68


class GoGateway:
    """the handlers, statement by statement; `redis` is the one store both services share in the reference"""

    def __init__(self, tokens, quotas, endpoints):
        self.tokens, self.quotas, self.endpoints = tokens, quotas, endpoints
        self.redis = {}

    # ---- qosconfig/arks_impl.go
    def get_qos_by_token(self, token, model):
        items = [t for t in self.tokens if t["spec"]["token"] == token]
        if not items:
            return None, "token not found"
        obj = items[0]
        for qos in obj["spec"].get("qos") or []:
            if qos["arksEndpoint"]["name"] == model:
                return {"user": obj["metadata"]["name"], "namespace": obj["metadata"]["namespace"], "model": model,
                        "quota": (qos.get("quota") or {}).get("name", ""),
                        "limits": [(r["type"], int(r["value"])) for r in qos.get("rateLimits") or []]}, None
        return None, "model not found"

    def get_quota_config(self, ns, name):
        for q in self.quotas:
            if q["metadata"]["namespace"] == ns and q["metadata"]["name"] == name:
                return [(i["type"], int(i["value"])) for i in q["spec"]["quotas"]]
        return None

    def get_model_list(self, ns):
        return [e["metadata"]["name"] for e in self.endpoints if e["metadata"]["namespace"] == ns]

    # ---- ratelimiter/cache_key.go:42-80, quota/cache_key.go:40-58
    @staticmethod
    def rate_key(qos, rule, now):
        w = RULES[rule][1]
        return "arks:namespace=%s.user=%s.model=%s.%s:%d" % (qos["namespace"], qos["user"], qos["model"], rule, now // w * w)

    @staticmethod
    def quota_key(ns, name, ty):
        return "arks:namespace=%s.quotaname=%s.type=%s." % (ns, name, ty)

    # ---- handle_request.go:83-249 (the parsed body comes from json.loads)
    def handle_request_body(self, f, token, now):
        """f: pymodel.request_fields(body). -> (reason, detail, cur_usage, limit_max, qos or None, stream)"""
        if f["err"]:
            return abi.R_REQUEST_BODY, 0, 0, 0, None, False
        model = f["model"].decode()
        if model == "":
            return abi.R_NO_MODEL, 0, 0, 0, None, False
        qos, err = self.get_qos_by_token(token, model)
        if err:
            return (abi.R_TOKEN_NOT_FOUND if err == "token not found" else abi.R_MODEL_NOT_IN_TOKEN), 0, 0, 0, None, False
        if model not in self.get_model_list(qos["namespace"]):
            return abi.R_NO_MODEL_BACKENDS, 0, 0, 0, qos, False
        stream = f["stream"] == 2
        if stream and not (f["so_present"] and f["include_usage"] == 2):
            return abi.R_STREAM_OPTIONS, 0, 0, 0, qos, stream
        # checkRateLimit, check.go:108-156: request-type rules ask for 1, token-type rules for 0
        for i, (rule, limit) in enumerate(qos["limits"]):
            cur = self.redis.get(self.rate_key(qos, rule, now), 0)
            if cur + (1 if RULES[rule][0] == "request" else 0) > limit:
                return abi.R_RATE_LIMIT, i, cur, limit, qos, stream
        # checkTokenQuotaLimit, check.go:75-106
        if qos["quota"] != "":
            conf = self.get_quota_config(qos["namespace"], qos["quota"])
            if conf is None:
                return abi.R_QUOTA_CONFIG, 0, 0, 0, qos, stream
            for i, (ty, limit) in enumerate(conf):
                cur = self.redis.get(self.quota_key(qos["namespace"], qos["quota"], ty), 0)
                if cur > limit:
                    return abi.R_QUOTA, i, cur, limit, qos, stream
        # doRequestRateLimit, check.go:31-44
        for rule, _ in qos["limits"]:
            if RULES[rule][0] == "request":
                k = self.rate_key(qos, rule, now)
                self.redis[k] = self.redis.get(k, 0) + 1
        return abi.R_OK, 0, 0, 0, qos, stream

    # ---- handle_response.go:134-268, complete non-streamed body with :status 200
    def handle_response_body(self, f, qos, now):
        """f: pymodel.response_fields(body). -> (reason, counted, usage)"""
        if f["err"]:
            return abi.R_RESPONSE_UNMARSHAL, 0, [0, 0, 0]
        if f["model_len"] == 0:
            return abi.R_RESPONSE_UNKNOWN, 0, [0, 0, 0]
        prompt, completion, total = f["usage"]
        if total == 0:
            return abi.R_OK, 0, [prompt, completion, total]
        for rule, _ in qos["limits"]:  # doTokenRateLimit, check.go:47-59
            if RULES[rule][0] == "token":
                k = self.rate_key(qos, rule, now)
                self.redis[k] = self.redis.get(k, 0) + total
        if qos["quota"] != "":  # doTokenQuotaLimit, check.go:62-72
            conf = self.get_quota_config(qos["namespace"], qos["quota"])
            if conf is None:
                return abi.R_QUOTA_CONFIG_RESP, 1, [prompt, completion, total]
            count = {"prompt": prompt, "response": completion, "total": total}
            for ty, _ in conf:
                k = self.quota_key(qos["namespace"], qos["quota"], ty)
                self.redis[k] = self.redis.get(k, 0) + count.get(ty, 0)
        return abi.R_OK, 1, [prompt, completion, total]


def random_objects(r: random.Random):
    """a small cluster: 3 namespaces, tokens with 0-4 limits in random order, quotas shared / missing / with duplicated items"""
    tokens, quotas, endpoints = [], [], []
    for ns in ("alpha", "beta", "gamma"):
        models = ["m%d" % i for i in range(r.randint(1, 3))]
        for m in models[: r.randint(max(1, len(models) - 1), len(models))] + (["ghost"] if r.random() < 0.3 else []):
            if m != "ghost":
                endpoints.append(simple_endpoint(m, ns, default_weight=r.randint(0, 5), routes=[("s%d" % i, r.randint(0, 9)) for i in range(r.randint(0, 3))]))
        qnames = ["q%d" % i for i in range(r.randint(0, 2))]
        for qn in qnames:
            items = [(r.choice(["prompt", "response", "total"]), r.choice([0, 30, 200, 5000])) for _ in range(r.randint(0, 4))]
            quotas.append(simple_quota(qn, ns, items))
        for u in range(r.randint(1, 4)):
            qos = []
            for m in r.sample(models + ["other"], r.randint(1, len(models) + 1)):
                # 0-4 limits in any order; one entry in four repeats a rule (every entry is checked and INCRBY'd on its own)
                rules = r.sample(list(RULES), r.randint(0, 4)) if r.random() < 0.75 else [r.choice(list(RULES)) for _ in range(r.randint(1, 4))]
                lim = [{"type": t, "value": r.choice([0, 1, 3, 8, 60, 2000])} for t in rules]
                quota = r.choice(qnames + ["", "", "missing"]) if qnames or r.random() < 0.5 else ""
                qos.append({"arksEndpoint": {"name": m}, "rateLimits": lim, "quota": {"name": quota}})
            tokens.append({"metadata": {"name": "u%d" % u, "namespace": ns}, "spec": {"token": "sk-%s-%d" % (ns, u), "qos": qos}})
    r.shuffle(tokens)
    return tokens, quotas, endpoints


def random_request(r: random.Random, tokens):
    tok = r.choice(tokens)["spec"]["token"] if r.random() < 0.93 else "sk-nobody"
    d = {"model": r.choice(["m0", "m1", "m2", "other", "ghost", ""]) if r.random() < 0.97 else None, "messages": [{"role": "user", "content": "hi"}]}
    if r.random() < 0.3:
        d["stream"] = r.choice([True, False, None])
        if r.random() < 0.7:
            d["stream_options"] = r.choice([{"include_usage": True}, {"include_usage": False}, {}, None])
    body = json.dumps(d).encode()
    if r.random() < 0.03:
        body = body[:-1] + b"]"  # the object closed by the wrong bracket: an error for any JSON parser
    return tok.encode(), body


@pytest.mark.parametrize("seed", range(40))
def test_oracle_against_the_go_shaped_model(seed):
    run_scenario(orklib.Oracle, seed)


def run_scenario(make_engine, seed):
    """make_engine(tables) -> anything with handle-or-oracle batch calls (the oracle here; the CUDA library in
    tests/test_z_gpu_reference_model.py)"""
    r = random.Random(1000 + seed)
    tokens, quotas, endpoints = random_objects(r)
    tables = Tables(tokens, quotas, endpoints)
    o, go = make_engine(tables), GoGateway(tokens, quotas, endpoints)
    request_batch = getattr(o, "request_batch", None) or o.handle_request_body
    response_batch = getattr(o, "response_batch", None) or o.handle_response_body
    key_of = {(tables.token_namespace[int(tables.qos_token[q])], tables.token_user[int(tables.qos_token[q])], tables.qos_model_name[q]): q
              for q in reversed(range(tables.n_qos))}  # first entry with the key
    now = 1_700_000_000 + r.randint(0, 86400)
    seen = [0] * 17
    for step in range(40):
        now += r.choice([0, 1, 7, 30, 61, 3600, 90000])  # same window, next minute, next day
        reqs = [random_request(r, tokens) for _ in range(r.randint(1, 60))]
        rand = [r.getrandbits(63) for _ in reqs]
        got = request_batch(RequestBatch.from_lists([b for _, b in reqs], [t for t, _ in reqs], now, pick_rand=np.array(rand, np.uint64)))
        admitted = []
        for i, (tok, body) in enumerate(reqs):
            f = {"err": 1} if body.endswith(b"]") else pymodel.request_fields(body)
            if f is None:
                pytest.fail("the generator left the subset json.loads referees")
            reason, detail, cur, lim, qos, stream = go.handle_request_body(f, tok.decode(), now)
            seen[reason] += 1
            assert (int(got.reason[i]), int(got.detail[i]), int(got.cur_usage[i]), int(got.limit_max[i])) == (reason, detail, cur, lim), (seed, step, i, body, tok)
            if reason == abi.R_OK:
                q = key_of[(qos["namespace"], qos["user"], qos["model"])]
                assert int(got.qos[i]) == q and bool(got.flags[i] & 1) == stream
                assert (tables.token_namespace[int(got.token[i])], tables.token_user[int(got.token[i])]) == (qos["namespace"], qos["user"])
                # A12: the backend Envoy's weighted-cluster walk lands on for this request's random (r mod sum of weights over
                # the cumulative weights, in backendRef order: static routeConfigs; no discovered Services in this cluster)
                ep = next(e for e in endpoints if (e["metadata"]["namespace"], e["metadata"]["name"]) == (qos["namespace"], qos["model"]))
                weights = [int(rc["weight"]) for rc in ep["spec"]["routeConfigs"]]
                want, x = -1, rand[i] % sum(weights) if sum(weights) else 0
                for k, wgt in enumerate(weights):
                    if sum(weights) and x < wgt:
                        want = k
                        break
                    x -= wgt
                assert int(got.pick[i]) == want, (seed, step, weights, rand[i])
                admitted.append((q, qos))
        # the upstream answers some of the admitted requests (complete bodies; SSE has its own pins)
        resp = []
        for q, qos in admitted:
            if r.random() < 0.8:
                p, c = r.randint(0, 40), r.randint(0, 40)
                u = {"prompt_tokens": p, "completion_tokens": c, "total_tokens": r.choice([p + c, p + c, 0])}
                d = r.choice([{"model": "x", "usage": u}, {"model": "x", "usage": u, "choices": []}, {"usage": u}, {"model": "x"}, {"model": 5}])
                resp.append((q, qos, json.dumps(d).encode()))
        if resp:
            rgot = response_batch(ResponseBatch.from_lists([b for _, _, b in resp], [q for q, _, _ in resp],
                                                             [abi.RESP_END_OF_STREAM] * len(resp), now + 1))
            for i, (q, qos, body) in enumerate(resp):
                reason, counted, usage = go.handle_response_body(pymodel.response_fields(body), qos, now + 1)
                assert (int(rgot.reason[i]), int(rgot.counted[i])) == (reason, counted), (seed, step, body)
                RESP_COVERED[reason] += 1
                if reason == abi.R_OK:
                    assert rgot.usage[i].tolist() == usage
        # every counter the oracle holds equals the Redis value under the reference's key
        rate, quota = o.snapshot_rate(now + 1), o.snapshot_quota()
        for (ns, user, model), q in key_of.items():
            for k, rule in enumerate(("rpm", "rpd", "tpm", "tpd")):
                assert int(rate[q, k]) == go.redis.get(go.rate_key({"namespace": ns, "user": user, "model": model}, rule, now + 1), 0), (seed, step, ns, user, model, rule)
        for qi, qobj in enumerate(quotas):
            ns, name = qobj["metadata"]["namespace"], qobj["metadata"]["name"]
            for k, ty in enumerate(("prompt", "response", "total")):
                assert int(quota[qi, k]) == go.redis.get(go.quota_key(ns, name, ty), 0), (seed, step, ns, name, ty)
        now += 1  # the library's clock contract: a batch never falls into an earlier window than the one before it
    for k, v in enumerate(seen):
        COVERED[k] += v
    if hasattr(o, "close"):
        o.close()


COVERED = [0] * 17
RESP_COVERED = [0] * 17


def test_the_scenarios_reached_every_branch_of_the_request_path():
    """(runs after the seeds above) every request-phase outcome the handlers can produce was compared, many times"""
    if sum(COVERED) == 0:
        pytest.skip("the scenario seeds were not run in this session")
    for reason in (abi.R_OK, abi.R_REQUEST_BODY, abi.R_NO_MODEL, abi.R_TOKEN_NOT_FOUND, abi.R_MODEL_NOT_IN_TOKEN, abi.R_NO_MODEL_BACKENDS,
                   abi.R_STREAM_OPTIONS, abi.R_RATE_LIMIT, abi.R_QUOTA, abi.R_QUOTA_CONFIG):
        assert COVERED[reason] >= 20, (reason, COVERED)
    # (R_QUOTA_CONFIG_RESP needs the ArksQuota to vanish between a request and its response: tests/test_gpu_config_plane.py)
    for reason in (abi.R_OK, abi.R_RESPONSE_UNMARSHAL, abi.R_RESPONSE_UNKNOWN):
        assert RESP_COVERED[reason] >= 20, (reason, RESP_COVERED)
