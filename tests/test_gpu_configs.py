"""BASELINE.json configs 3, 4, 5 at full size against the oracle (these parity blocks lived in a bench
script), the shared-quota fold on one GPU (two contexts on cuda:0, the all-reduce done in process), and the config
plane's generation re-map (a reload between a stream's request and its response chunks)."""
import numpy as np
import pytest

import orklib
from arks_b200 import abi, traffic
from arks_b200.abi import RequestBatch, ResponseBatch
from arks_b200.tables import Tables, simple_endpoint, simple_quota, simple_token

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


def body_of(b, i):
    return bytes(b.bodies[b.body_off[i]:b.body_off[i] + b.body_len[i]])


def test_config3_sse_streams_frame_cut_and_arbitrary_cut(gwmod):
    """65 536 streams x 4 SSE chunks. Faithful mode: chunks cut on frame boundaries (every chunk decodes on its own).
    Carry-over mode: the same streams re-cut at seeded arbitrary offsets — the reference decodes every chunk in isolation
    (handle_response.go:113-133), so a chunk that starts or ends inside a frame is a broken event: 500 x-error-streaming
    for the ones whose fragment is not JSON, and the usage frame is only counted when it arrives whole."""
    w = traffic.Workload(10_000, seed=0xA2C5)
    g = gwmod.Gateway(0, WAVE, int(WAVE * 1200))
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    req = w.request_batch(WAVE, NOW, seed=31, stream_frac=1.0, n_templates=2048, varied=True)
    a = g.handle_request_body(req)
    same(a, o.request_batch(req), "config 3 requests")
    ok = np.flatnonzero(a.reason == abi.R_OK)
    assert len(ok) > 50_000 and np.all(a.flags[ok] & 1)
    # 2 048 distinct streams of 4 chunks; every admitted request is answered by one of them
    rng = np.random.default_rng(32)
    streams = [traffic.sse_response_chunks(rng, int(rng.integers(50, 401)), int(rng.integers(1, 513)), 4096) for _ in range(2048)]
    pick = rng.integers(0, 2048, len(ok))
    now = NOW + 1
    for c in range(4):  # chunk c of every stream is one response batch (the 4 chunks of a stream arrive one after another)
        resp = ResponseBatch.from_lists([streams[p][c] for p in pick], a.qos[ok], [abi.RESP_STREAM] * len(ok), now)
        r = g.handle_response_body(resp)
        same(r, o.response_batch(resp), f"config 3 frame-cut chunk {c}")
        assert not np.any(r.reason == abi.R_STREAMING)
        assert (r.counted.sum() > 0) == (c == 3)  # only the last chunk carries the usage frame
    state_same(g, o, now)
    # carry-over mode: re-cut every stream at 3 seeded arbitrary offsets
    recut = []
    for s in streams:
        whole = b"".join(s)
        cuts = np.sort(rng.integers(1, len(whole) - 1, 3))
        recut.append([whole[:cuts[0]], whole[cuts[0]:cuts[1]], whole[cuts[1]:cuts[2]], whole[cuts[2]:]])
    n_stream_err = n_counted = 0
    for c in range(4):
        resp = ResponseBatch.from_lists([recut[p][c] for p in pick], a.qos[ok], [abi.RESP_STREAM] * len(ok), now + 1)
        r = g.handle_response_body(resp)
        same(r, o.response_batch(resp), f"config 3 arbitrary-cut chunk {c}")
        n_stream_err += int((r.reason == abi.R_STREAMING).sum())
        n_counted += int(r.counted.sum())
    assert n_stream_err > len(ok)        # split events fail their chunk, as in the reference
    assert 0 < n_counted < len(ok)       # some usage frames survive the cut whole, many do not
    state_same(g, o, now + 1)


def test_config4_zipf_quota_exhaustion_flip_indices(gwmod):
    """100 000 ArksQuotas, Zipf(1.1) popularity, all item-type mixes; limits low enough that quotas and windows run out
    mid-run. Bit-exact decisions, and explicitly: the request index at which every tenant first flips to deny."""
    w = traffic.Workload(100_000, seed=0xA2C6, zipf_alpha=1.1)
    g = gwmod.Gateway(0, WAVE, int(WAVE * 1200))
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    first_deny_gpu, first_deny_ork = {}, {}
    biggest = 0
    for k in range(3):
        now = NOW + 5 * k
        req = w.request_batch(WAVE, now, seed=41 + k, n_templates=2048, varied=True)
        a, b = g.handle_request_body(req), o.request_batch(req)
        same(a, b, f"config 4 wave {k}")
        for res, first in ((a, first_deny_gpu), (b, first_deny_ork)):
            den = np.flatnonzero((res.reason == abi.R_RATE_LIMIT) | (res.reason == abi.R_QUOTA))
            for i in den:
                first.setdefault(int(res.token[i]), (k, int(i), int(res.reason[i]), int(res.detail[i])))
        biggest = max(biggest, int(np.bincount(a.token[a.token >= 0]).max()))
        resp = w.response_batch(a, now + 1, seed=51 + k, varied=True, n_templates=2048)
        n = min(resp.n, WAVE)
        resp = ResponseBatch(resp.bodies, resp.body_off[:n], resp.body_len[:n], resp.qos[:n], resp.flags[:n], resp.now_unix)
        same(g.handle_response_body(resp), o.response_batch(resp), f"config 4 wave {k} responses")
        state_same(g, o, now + 1)
    assert first_deny_gpu == first_deny_ork and len(first_deny_gpu) > 100
    assert any(v[2] == abi.R_QUOTA for v in first_deny_gpu.values()) and any(v[2] == abi.R_RATE_LIMIT for v in first_deny_gpu.values())
    assert biggest > 4096  # the hot tenant's group goes through rank_hot_groups_kernel


def test_config5_weight_upsert_between_every_batch(gwmod):
    """1 000 ArksEndpoints x 16 backends; the weights of one endpoint are replaced between every two batches; the pick
    of every admitted request equals the cumulative walk over the weights in force for ITS batch."""
    rng = np.random.default_rng(5)
    w = traffic.Workload(1000, seed=0xA2C7, n_backends=16)
    g = gwmod.Gateway(0, WAVE, int(WAVE * 1200))
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    req = w.request_batch(WAVE, NOW, seed=61, n_templates=2048, varied=True)
    picks = []
    for rnd in range(8):
        ep, wts = int(rng.integers(1000)), rng.integers(0, 100, 16)
        if rnd == 3:
            wts[:] = 0  # an endpoint whose backends all have weight 0: no pick
        g.update_endpoint_weights(ep, wts)
        o.update_endpoint_weights(ep, wts)
        req.now_unix = NOW + 86_400 * rnd  # fresh windows: the admission pattern repeats, the picks must not
        a = g.handle_request_body(req)
        same(a, o.request_batch(req), f"config 5 round {rnd}")
        picks.append(a.pick.copy())
    assert any(not np.array_equal(picks[0], p) for p in picks[1:])


def test_shared_quota_fold_on_one_gpu_with_traffic_between_export_and_fold(gwmod):
    """Two replicas of one tenant's tables on cuda:0 (what two GPUs hold when a hot tenant is split). Fold epoch:
    export on both, sum (the all-reduce, done here with torch on the device), fold on both. A response batch that runs
    BETWEEN a replica's export and its fold must survive into the next epoch (it used to be dropped)."""
    import torch
    toks = [simple_token("u%d" % i, "hot", "tk-%d" % i, "m", [("tpm", 10**9)], "shared") for i in range(4)]
    t = Tables(toks, [simple_quota("shared", "hot", [("prompt", 10**9), ("response", 10**9), ("total", 10**9)])], [simple_endpoint("m", "hot")])
    reps = [gwmod.Gateway(0, 1024, 1 << 20, share_quota=True) for _ in range(2)]
    for g in reps:
        g.load_tables(t)
    truth = np.zeros(3, np.int64)
    rng = np.random.default_rng(3)

    def serve(g, n, now):
        u = rng.integers(1, 50, (n, 2))
        bodies = [b'{"model":"m","usage":{"prompt_tokens":%d,"completion_tokens":%d,"total_tokens":%d}}' % (p, c, p + c) for p, c in u]
        r = g.handle_response_body(ResponseBatch.from_lists(bodies, rng.integers(0, 4, n), [abi.RESP_END_OF_STREAM] * n, now))
        assert np.all(r.counted == 1)
        truth[:] += [u[:, 0].sum(), u[:, 1].sum(), u.sum()]

    dev = torch.device("cuda", 0)
    for epoch in range(3):
        serve(reps[0], 300, NOW + epoch)
        serve(reps[1], 200, NOW + epoch)
        own = [torch.zeros((1, 3), dtype=torch.int64, device=dev) for _ in reps]
        for g, buf in zip(reps, own):
            g.export_quota_delta_dev(buf.data_ptr())
        serve(reps[0], 77, NOW + epoch)  # lands after replica 0's export, before its fold
        reduced = own[0] + own[1]
        torch.cuda.synchronize()
        for g in reps:
            g.fold_quota_delta_dev(reduced.data_ptr())
        # both replicas have seen everything exported so far; replica 0 additionally its own late batch
        q0, q1 = reps[0].snapshot_quota()[0], reps[1].snapshot_quota()[0]
        late = truth - q1
        assert np.all(late > 0) and np.array_equal(q0, truth), (q0, q1, truth)
    # the late increments were not lost: one more epoch without traffic makes the replicas equal
    own = [torch.zeros((1, 3), dtype=torch.int64, device=dev) for _ in reps]
    for g, buf in zip(reps, own):
        g.export_quota_delta_dev(buf.data_ptr())
    reduced = own[0] + own[1]
    torch.cuda.synchronize()
    for g in reps:
        g.fold_quota_delta_dev(reduced.data_ptr())
    assert np.array_equal(reps[0].snapshot_quota()[0], truth) and np.array_equal(reps[1].snapshot_quota()[0], truth)
    # unfolded increments survive a config reload (they travel with their quota's key)
    serve(reps[1], 50, NOW + 10)
    reps[1].load_tables(Tables(toks[:3], [simple_quota("other", "hot", [("total", 5)])] +
                               [simple_quota("shared", "hot", [("prompt", 10**9), ("response", 10**9), ("total", 10**9)])],
                               [simple_endpoint("m", "hot")]))
    assert np.array_equal(reps[1].take_quota_delta()[1], truth - reps[0].snapshot_quota()[0])


def test_reload_between_request_and_response_bills_the_same_key(gwmod):
    """a stream keeps (generation, qos) from its request; after ArksTokens are added / removed /
    reordered the same index names another tenant. Responses carrying the old generation are re-mapped by
    (namespace, user, model); a stream whose qos entry is gone is answered ARKS_R_QOS_GONE and bills nobody."""
    mk = lambda i: simple_token("user-%d" % i, "ns-%d" % i, "tk-%d" % i, "m", [("tpm", 10**6), ("rpm", 100)], "q")
    quotas = lambda ids: [simple_quota("q", "ns-%d" % i, [("total", 10**6)]) for i in ids]
    eps = lambda ids: [simple_endpoint("m", "ns-%d" % i) for i in ids]
    ids0 = [0, 1, 2, 3, 4, 5]
    t0 = Tables([mk(i) for i in ids0], quotas(ids0), eps(ids0))
    g = gwmod.Gateway(0, 256, 1 << 20)
    g.load_tables(t0)
    gen0 = g.generation
    req = RequestBatch.from_lists([b'{"model":"m"}'] * 6, [b"tk-%d" % i for i in ids0], NOW)
    a = g.handle_request_body(req)
    assert np.all(a.reason == 0) and a.qos.tolist() == [0, 1, 2, 3, 4, 5]
    # reload: tenant 1 removed, a new tenant 9 inserted at the front, the rest reversed
    ids1 = [9, 5, 4, 3, 2, 0]
    t1 = Tables([mk(i) for i in ids1], quotas(ids1), eps(ids1))
    g.load_tables(t1)
    assert g.generation == gen0 + 1
    usage = [(10 * (i + 1), i + 1) for i in range(6)]
    bodies = [b'{"model":"m","usage":{"prompt_tokens":%d,"completion_tokens":%d,"total_tokens":%d}}' % (p, c, p + c) for p, c in usage]
    resp = ResponseBatch.from_lists(bodies, a.qos, [abi.RESP_END_OF_STREAM] * 6, NOW + 1, gen=[gen0] * 6)
    r = g.handle_response_body(resp)
    assert r.reason.tolist() == [0, abi.R_QOS_GONE, 0, 0, 0, 0] and r.counted.tolist() == [1, 0, 1, 1, 1, 1]
    quota = g.snapshot_quota()[:, 2]
    for k, i in enumerate(ids1):  # every tenant was billed ITS stream's tokens, whatever its row is called now
        want = 0 if i == 9 else sum(usage[i])
        assert quota[k] == want, (i, quota[k], want)
    # two reloads later the first generation is still resolvable; without a generation the index is taken at face value
    g.load_tables(Tables([mk(i) for i in ids0], quotas(ids0), eps(ids0)))
    r = g.handle_response_body(ResponseBatch.from_lists(bodies[:1], [0], [abi.RESP_END_OF_STREAM], NOW + 2, gen=[gen0]))
    assert r.reason.tolist() == [0] and g.snapshot_quota()[0, 2] == 2 * sum(usage[0])
    # rows that never had a qos entry / out of range / from the future: per-row verdict, the batch goes through
    r = g.handle_response_body(ResponseBatch.from_lists(bodies[:4], [-1, 77, 2, 3], [abi.RESP_END_OF_STREAM] * 4, NOW + 3,
                                                        gen=[g.generation, g.generation, g.generation + 5, g.generation]))
    assert r.reason.tolist() == [abi.R_QOS_GONE, abi.R_QOS_GONE, abi.R_QOS_GONE, 0]


def test_library_side_nccl_fold_single_rank_communicator(gwmod):
    """arks_comm_* / arks_fold_quota_allreduce on a one-rank communicator (all a one-GPU lease can hold; two ranks run in
    tests/test_multi_rank.py on >= 2 GPUs): libnccl is dlopen()ed, the shared rows go through ncclAllReduce in place, and
    with nobody else to hear from the fold must leave every quota where the oracle has it and hand every delta back."""
    w = traffic.Workload(2_000, seed=41)
    g = gwmod.Gateway(0, WAVE, int(WAVE * 1200), share_quota=True)
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    shared = np.arange(0, w.tables.n_quotas, 7, dtype=np.uint32)  # a subset, like the replicated hot tenants
    g.comm_init(0, 1, g.comm_unique_id(), shared)
    for epoch in range(3):
        req = w.request_batch(16384, NOW + epoch, seed=50 + epoch)
        a = g.handle_request_body(req)
        same(a, o.request_batch(req), f"epoch {epoch} requests")
        resp = w.response_batch(a, NOW + epoch, seed=60 + epoch)
        same(g.handle_response_body(resp), o.response_batch(resp), f"epoch {epoch} responses")
        before = g.snapshot_quota()
        g.fold_quota_allreduce(wait=(epoch != 1))  # epoch 1: stream-ordered only, the next call is behind it
        assert np.array_equal(g.snapshot_quota(), before)
        state_same(g, o, NOW + epoch)
        assert not g.take_quota_delta().any()  # every increment was exported and given up
    g.comm_set_shared(None)  # every quota, identical numbering: the other calling convention
    g.fold_quota_allreduce()
    state_same(g, o, NOW + 2)
    with pytest.raises(Exception):
        g.comm_set_shared(np.array([w.tables.n_quotas], np.uint32))  # out of range
