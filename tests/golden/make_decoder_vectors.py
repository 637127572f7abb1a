#!/usr/bin/env python
"""Generates tests/golden/decoder_vectors.json: request / response / SSE documents with the fields the reference's
handlers must extract from them, derived WITHOUT the oracle: Python's json (RFC 8259 validity, escape decoding) and
openai-python's SSEDecoder (the Stainless SSE decoder, as in openai-go) + the Go struct-binding rules restated in
tests/pymodel.py. Only documents inside the subset on which those independent parsers and the Go libraries agree are
kept (pymodel returns None otherwise). tests/test_decoder_pins.py checks the oracle (CPU) and the CUDA path (-m gpu)
against these vectors.

    python tests/golden/make_decoder_vectors.py        # rewrites the JSON (deterministic: seeded generators)
"""
import base64
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]

import numpy as np  # noqa: E402

import pymodel  # noqa: E402
from arks_b200 import traffic  # noqa: E402
from jsonfuzz import Gen  # noqa: E402

N_EACH = 600


def b64(b):
    return base64.b64encode(b).decode()


def collect(make, fields, n):
    out, tried = [], 0
    while len(out) < n:
        doc = make()
        tried += 1
        f = fields(doc)
        if f is None:
            continue
        if "model" in f:
            f["model"] = b64(f["model"])
        out.append({"doc": b64(doc), **f})
    return out, tried


def main():
    rng = np.random.default_rng(0xD0C5)
    gen = Gen(0xD0C5)
    req_real = lambda: traffic.chat_request_body_varied(rng, 450, stream=bool(rng.random() < 0.3))
    resp_real = lambda: traffic.chat_response_body_varied(rng, int(rng.integers(50, 401)), int(rng.integers(1, 513)), 420)

    def sse_real():
        chunks = traffic.sse_response_chunks(rng, int(rng.integers(50, 401)), int(rng.integers(1, 513)), 1200, 2)
        c = chunks[int(rng.integers(len(chunks)))]
        k = rng.random()
        if k < 0.2:
            c = c.replace(b"\n", b"\r\n")
        elif k < 0.4:  # a cut at an arbitrary offset (BASELINE config 3's carry-over mode): the reference decodes every chunk
            c = c[int(rng.integers(0, len(c))):]  # in isolation, so a split frame is a broken event
        elif k < 0.5:
            c = c[:int(rng.integers(0, len(c)))]
        return c

    mix = lambda a, b: (lambda: a() if rng.random() < 0.7 else b())
    vec, stats = {}, {}
    for name, make, fields in (("request", mix(gen.request, req_real), pymodel.request_fields),
                               ("response", mix(gen.response, resp_real), pymodel.response_fields),
                               ("sse", mix(gen.sse_chunk, sse_real), pymodel.sse_fields)):
        vec[name], tried = collect(make, fields, N_EACH)
        stats[name] = {"kept": len(vec[name]), "generated": tried, "errors": sum(v["err"] for v in vec[name])}
    # SSE event split alone (SSEDecoder vs the oracle's ork_sse_events), including chunks whose JSON is not in the subset
    split = []
    while len(split) < N_EACH:
        c = gen.sse_chunk() if rng.random() < 0.7 else sse_real()
        ev = pymodel.sse_split(c)
        if ev is None:
            continue
        split.append({"doc": b64(c), "events": [[t, d] for t, d in ev]})
    vec["sse_split"] = split
    vec["_stats"] = stats
    vec["_how"] = "tests/golden/make_decoder_vectors.py: json.loads + openai._streaming.SSEDecoder + tests/pymodel.py"
    with open(os.path.join(HERE, "decoder_vectors.json"), "w") as f:
        json.dump(vec, f, separators=(",", ":"))
    print(stats)


if __name__ == "__main__":
    main()
