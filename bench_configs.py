#!/usr/bin/env python
"""Kernel-resident throughput of the other BASELINE.json configs (3: SSE streaming, 4: Zipf quota exhaustion,
5: routing churn) on one GPU, each with a parity spot check against the oracle. `bench.py` stays the contract bench
(config 2); this script prints one JSON line per config and is what profiles/configs_rNN.json is made from."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

import __graft_entry__ as ge  # noqa: E402

ge.build()
import orklib  # noqa: E402
from arks_b200 import abi, traffic  # noqa: E402
from arks_b200.abi import ResponseBatch  # noqa: E402
from arks_b200.gateway import Gateway  # noqa: E402
from bench import pin_batch  # noqa: E402

NOW0, DAY = 1_700_000_000, 86_400


def same(a, b):
    return all(np.array_equal(v, b.fields()[k]) for k, v in a.fields().items())


def timed(g, fn, steps, warm=3):
    ext = torch.cuda.ExternalStream(g.stream_handle)
    for i in range(warm):
        fn(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(ext)
    for i in range(steps):
        fn(warm + i)
    e1.record(ext)
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / steps


def config3(wave=65536, steps=20):
    """SSE streaming completions, 4 KiB per response in 4 chunks cut on frame boundaries, TPM+RPM enforced."""
    w = traffic.Workload(10_000, seed=0xA2C5)
    g = Gateway(0, wave, int(wave * 1100 * 1.05)); g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    req = pin_batch(w.request_batch(wave, NOW0, seed=31, stream_frac=1.0, n_templates=2048, varied=True))
    a = g.handle_request_body(req)
    assert same(a, o.request_batch(req))
    allresp = w.response_batch(a, NOW0 + 1, seed=32)
    # chunk c of every stream -> response batch c (every chunk is decoded in isolation, like the reference)
    idx = np.arange(allresp.n)
    parts = []
    for c in range(4):
        sel = idx[c::4]
        bodies = [bytes(allresp.bodies[allresp.body_off[i]:allresp.body_off[i] + allresp.body_len[i]]) for i in sel]
        parts.append(pin_batch(ResponseBatch.from_lists(bodies, allresp.qos[sel], allresp.flags[sel], NOW0 + 1)))
    for c, p in enumerate(parts):
        assert same(g.handle_response_body(p), o.response_batch(p)), f"chunk batch {c}"
    assert np.array_equal(g.snapshot_rate(NOW0 + 1), o.snapshot_rate(NOW0 + 1))
    g.select_slot(0); g.stage_request(req)
    for c, p in enumerate(parts):
        g.select_slot(c); g.stage_response(p)
    state = {"now": NOW0 + DAY}

    def step(i):
        g.select_slot(0); g.run_request(state["now"])
        for c in range(4):
            g.select_slot(c); g.run_response(state["now"] + 1)
        state["now"] += DAY
    ms = timed(g, step, steps)
    g.set_profiling(True)
    g.select_slot(3); g.run_response(state["now"]); k = g.last_kernel_ms()[0]
    g.set_profiling(False)
    chunk_bytes = float(sum(int(p.body_len.sum()) for p in parts))
    return {"config": "3: SSE streaming, 4 chunks x ~1 KiB per response, stream:true + include_usage", "streams_per_step": wave,
            "chunks_per_step": int(sum(p.n for p in parts)), "ms_per_step": ms, "streams_per_s": wave / ms * 1e3,
            "sse_bytes_per_step": chunk_bytes, "scan_response_sse_ms_per_chunk_batch": k,
            "scan_response_sse_GBps": parts[3].body_len.sum() / k / 1e6, "parity": "ok"}


def config4(wave=65536, steps=20):
    """100k ArksQuotas, Zipf(1.1) tenant popularity: hot groups straddle their limits inside a batch."""
    w = traffic.Workload(100_000, seed=0xA2C6, zipf_alpha=1.1)
    g = Gateway(0, wave, int(wave * 1100 * 1.05)); g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    reqs = [pin_batch(w.request_batch(wave, NOW0 + 5 * k, seed=41 + k, n_templates=2048, varied=True)) for k in range(3)]
    flips = 0
    for k, r in enumerate(reqs):  # serial-order parity incl. the exact index at which each tenant flips to deny
        a, b = g.handle_request_body(r), o.request_batch(r)
        assert same(a, b), f"wave {k}"
        flips += int(((a.reason == abi.R_RATE_LIMIT) | (a.reason == abi.R_QUOTA)).sum())
        resp = w.response_batch(a, NOW0 + 5 * k + 1, seed=51 + k, varied=True, n_templates=2048)
        n = min(resp.n, wave)
        resp = ResponseBatch(resp.bodies, resp.body_off[:n], resp.body_len[:n], resp.qos[:n], resp.flags[:n], resp.now_unix)
        assert same(g.handle_response_body(resp), o.response_batch(resp))
    assert np.array_equal(g.snapshot_quota(), o.snapshot_quota())
    hot = int(np.bincount(np.frombuffer(reqs[0].tokens, np.uint8)[:0].astype(int), minlength=1)[0]) if False else None
    for k, r in enumerate(reqs):
        g.select_slot(k); g.stage_request(r)
    state = {"now": NOW0 + DAY}

    def step(i):
        g.select_slot(i % 3); g.run_request(state["now"]); state["now"] += 60   # same day: rpd / quota keep biting
    ms = timed(g, step, steps)
    g.set_profiling(True)
    g.select_slot(0); g.run_request(state["now"]); km = g.last_kernel_ms()
    g.set_profiling(False)
    tok = np.frombuffer(reqs[0].tokens[:reqs[0].token_off[-1]], np.uint8).reshape(wave, -1)
    _, counts = np.unique(tok, axis=0, return_counts=True)
    return {"config": "4: 100k ArksQuotas, Zipf(1.1) popularity, limits exhausted mid-run", "requests_per_step": wave,
            "ms_per_step": ms, "req_per_s": wave / ms * 1e3, "largest_group_in_a_wave": int(counts.max()),
            "denied_by_limit_or_quota_in_parity_waves": flips, "scan_request_ms": km[0], "limit_admit_ms": km[1],
            "parity": "ok"}


def config5(wave=65536, steps=40):
    """1k ArksEndpoints x 16 backends, weights of some endpoint replaced every other batch (~100 updates/s at the
    batch rates a gateway sees), weighted pick checked against the oracle after every update."""
    rng = np.random.default_rng(5)
    w = traffic.Workload(1000, seed=0xA2C7, n_backends=16)
    g = Gateway(0, wave, int(wave * 1100 * 1.05)); g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    req = pin_batch(w.request_batch(wave, NOW0, seed=61, n_templates=2048, varied=True))
    for rnd in range(6):
        ep = int(rng.integers(1000)); wts = rng.integers(0, 100, 16)
        g.update_endpoint_weights(ep, wts); o.update_endpoint_weights(ep, wts)
        req.now_unix = NOW0 + rnd
        assert same(g.handle_request_body(req), o.request_batch(req)), f"round {rnd}"
    g.select_slot(0); g.stage_request(req)
    state = {"now": NOW0 + DAY}

    def step(i):
        if i % 2 == 0:
            g.update_endpoint_weights(int(rng.integers(1000)), rng.integers(0, 100, 16))
        g.run_request(state["now"]); state["now"] += DAY
    t0 = time.perf_counter()
    ms = timed(g, step, steps)
    return {"config": "5: 1k ArksEndpoints x 16 backends, weight upsert every other batch under load", "requests_per_step": wave,
            "ms_per_step_device": ms, "req_per_s": wave / ms * 1e3, "weight_updates": steps // 2, "parity": "ok"}


if __name__ == "__main__":
    which = sys.argv[1:] or ["3", "4", "5"]
    for c in which:
        print(json.dumps({"3": config3, "4": config4, "5": config5}[c]()), flush=True)
