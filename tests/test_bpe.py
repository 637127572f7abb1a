"""The on-device BPE token counter against HF `tokenizers` (its oracle: the reference has no tokenizer).

CPU: the host build of arks_b200/csrc/bpe.cuh (the same scanner, pre-tokenizer and merge loop the kernels run).
GPU (-m gpu): the kernels through the C ABI, bit-exact counts on 64 k generated bodies.
The vocabulary is a seeded stand-in trained offline with `tokenizers` (arks_b200.bpe.standin_tokenizer): the real
Qwen2.5 files are not on disk and there is no network; same pre-tokenizer pattern, same algorithm."""
import json
import os
import random

import numpy as np
import pytest

import hostmachine as hm
from arks_b200 import bpe, traffic

BUILD = os.path.join(os.path.dirname(__file__), "_build")
tokenizers = pytest.importorskip("tokenizers")


@pytest.fixture(scope="module")
def standin():
    tok, text = bpe.standin_tokenizer(20_000, cache_dir=BUILD)
    tables = bpe.load_tokenizer(text)
    hm.bpe_load(tables)
    return tok, tables


def oracle_count(tok, body: bytes) -> int:
    return sum(len(tok.encode(s, add_special_tokens=False).ids) for s in bpe.content_strings(body))


def test_unicode_table_is_what_the_regex_engine_does():
    """the committed class table against the Split pre-tokenizer itself, on a sample of code points"""
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(__file__)), "tools"))
    import gen_bpe_unicode as g
    split = g.splitter()
    cls = bpe.unicode_classes()
    r = random.Random(7)
    sample = list(range(0x300)) + [r.randrange(0x110000) for _ in range(6000)] + [0x85, 0xA0, 0x1680, 0x2028, 0x3000, 0xFEFF, 0x200B]
    for cp in sample:
        if 0xD800 <= cp <= 0xDFFF:
            continue
        got = (cls[cp >> 1] >> (4 * (cp & 1))) & 3
        assert got == g.classify(cp, split), hex(cp)


def test_pretokenizer_matches_the_split_pattern(standin):
    from tokenizers import Regex, pre_tokenizers
    split = pre_tokenizers.Split(Regex(bpe.QWEN2_PATTERN), behavior="isolated", invert=False)
    r = random.Random(1)
    pool = list("abcXYZ019 \t\n\r'.,!?-_()\"\\/sStTmMdDrReEvVlL") + ["é", "你", "好", "—", "🙂", " ", "　", " ", "́", "ſ", "K", "٣", "²",
                                                                   "½", "ǅ", "ʰ", "ª", "\x0b", "\x0c", "\x85", "​", "﻿", "\x1f", "\x7f"]
    for t in range(60000):
        if t % 3 == 0:
            s = "".join(chr(r.choice([r.randint(1, 0x2FFF), r.randint(0x3000, 0xD7FF), r.randint(0xE000, 0xFFFF),
                                      r.randint(0x10000, 0x2FFFF), r.randint(32, 126)])) for _ in range(r.randint(0, 10)))
        else:
            s = "".join(r.choice(pool) for _ in range(r.randint(0, 14)))
        want = [len(s[:e].encode()) for _, (b, e) in split.pre_tokenize_str(s)]
        assert hm.bpe_pretokenize(s.encode()) == want, repr(s)


def test_counts_match_tokenizers_on_generated_bodies(standin):
    tok, _ = standin
    rng = np.random.default_rng(3)
    for _ in range(1500):
        b = traffic.chat_request_body_varied(rng, 900, stream=bool(rng.random() < 0.3))
        assert hm.bpe_count(b) == oracle_count(tok, b), b
    for _ in range(600):
        b = traffic.chat_response_body_varied(rng, 10, 20, 600)
        assert hm.bpe_count(b) == oracle_count(tok, b), b


def test_counts_match_tokenizers_on_odd_text(standin):
    tok, _ = standin
    r = random.Random(5)
    pool = list("abc XYZ 0123  \t'.,!?-()") + ["é", "你好", "—", "🙂", " the", " and", "ing", "\n", "\n\n", "  ", "'s", "'RE", "\\", "\""]
    for _ in range(3000):
        text = "".join(r.choice(pool) for _ in range(r.randint(0, 60)))
        body = json.dumps({"model": "m", "messages": [{"role": "user", "content": text}, {"content": text[::-1]}],
                           "x": {"content": 5, "y": [{"content": "nested " + text[:9]}]}},
                          ensure_ascii=bool(r.random() < 0.5)).encode()
        assert hm.bpe_count(body) == oracle_count(tok, body), body
    # escaped spelling of the key, surrogate pairs, every simple escape
    b = b'{"c\\u006fntent":"a\\ud83d\\ude42b\\n\\t\\"\\\\\\/\\b\\f\\r","content":"\\u4f60\\u597d"}'
    assert hm.bpe_count(b) == oracle_count(tok, b)
    # uncounted instead of wrong: a lone surrogate, a piece longer than the device handles, invalid UTF-8
    assert hm.bpe_count(b'{"content":"\\ud800x"}') == bpe.UNCOUNTED
    assert hm.bpe_count(b'{"content":"' + b"a" * 200 + b'"}') == bpe.UNCOUNTED
    assert hm.bpe_count(b'{"content":"\xff"}') == bpe.UNCOUNTED
    assert hm.bpe_count(b'{"model":"m"}') == 0
    # SSE chunks: the same scanner over the data lines
    rng = np.random.default_rng(4)
    for ch in traffic.sse_response_chunks(rng, 10, 20, 4096):
        want = sum(len(tok.encode(s, add_special_tokens=False).ids) for s in bpe.sse_content_strings(ch))
        assert hm.bpe_count(ch) == want


def test_nfc_tokenizers_decline_unsafe_text(standin):
    _, tables = standin
    try:
        hm.bpe_load(bpe.BpeTables(tables.byte_id, tables.left, tables.right, tables.merged, tables.cp_class, bpe.BPE_NFC))
        assert hm.bpe_count('{"content":"café"}'.encode()) == bpe.UNCOUNTED  # e + combining acute: NFC would change it
        assert hm.bpe_count('{"content":"café 你好"}'.encode()) != bpe.UNCOUNTED
    finally:
        hm.bpe_load(tables)


@pytest.mark.gpu
def test_gpu_counts_match_tokenizers_on_64k_bodies(gwmod):
    """the kernels through the C ABI: bit-exact token counts on 65 536 generated requests (escapes, UTF-8, \\u sequences) and
    on completions / SSE chunks; decisions are untouched by the side output"""
    import orklib
    from arks_b200 import abi
    from arks_b200.abi import RequestBatch, ResponseBatch
    tok, text = bpe.standin_tokenizer(20_000, cache_dir=BUILD)
    tables = bpe.load_tokenizer(text)
    w = traffic.Workload(n_tenants=500, seed=9)
    g = gwmod.Gateway(0, 65536, 96 << 20)
    g.load_tables(w.tables)
    o = orklib.Oracle(w.tables)
    req = w.request_batch(65536, 1_700_000_000, seed=77, varied=True, n_templates=4096, stream_frac=0.2)
    plain = g.handle_request_body(req)
    assert np.all(plain.bpe_count == 0)  # no vocabulary: the column is 0
    want0 = o.request_batch(req)
    g.load_bpe(tables)
    req.now_unix += 86_400
    a = g.handle_request_body(req)
    want1 = o.request_batch(req)
    for got, want in ((plain, want0), (a, want1)):  # the decisions are the reference's, with or without the side output
        for k, v in got.fields().items():
            assert np.array_equal(v, want.fields()[k]), k
    # every distinct body once through tokenizers
    want = {}
    for i in range(req.n):
        body = bytes(req.bodies[req.body_off[i]:req.body_off[i] + req.body_len[i]])
        if body not in want:
            want[body] = oracle_count(tok, body)
        assert a.bpe_count[i] == want[body], (i, body[:200], int(a.bpe_count[i]), want[body])
    assert len(want) >= 4000 and a.bpe_count.max() > 100
    # completions and SSE chunks
    ok = abi.RequestResult.empty(3000)
    ok.reason[:] = 0
    ok.qos[:] = np.arange(3000) % w.tables.n_qos
    ok.flags[:] = (np.arange(3000) % 3 == 0)
    resp = w.response_batch(ok, 1_700_086_500, seed=78, varied=True, sse_total=2048)
    c = g.handle_response_body(resp)
    for i in range(resp.n):
        body = bytes(resp.bodies[resp.body_off[i]:resp.body_off[i] + resp.body_len[i]])
        strings = bpe.sse_content_strings(body) if resp.flags[i] & abi.RESP_STREAM else bpe.content_strings(body)
        assert c.bpe_count[i] == sum(len(tok.encode(s, add_special_tokens=False).ids) for s in strings), body[:200]
    # bodies the device declines are reported as such, not miscounted
    odd = RequestBatch.from_lists([b'{"model":"m","messages":[{"content":"' + b"z" * 300 + b'"}]}', b'{"content":"\\ud800"}',
                                   b'{"content":"ok then"}'], [w.token_strings[0]] * 3, 1_700_086_600)
    r = g.handle_request_body(odd)
    assert r.bpe_count[0] == bpe.UNCOUNTED and r.bpe_count[1] == bpe.UNCOUNTED
    assert r.bpe_count[2] == len(tok.encode("ok then", add_special_tokens=False).ids)
