"""A13: the replies of a failed phase are the reference's, value for value. tests/golden/
error_replies.json is transcribed from the Go sources; both implementations are held to it: the compiled host
(host/cpp RequestErrorReply / ResponseErrorReply, through its C surface) and the Python mirror the ext_proc loopback uses."""
import ctypes as C
import json
import os

import pytest

from arks_b200 import abi, cpphost, replies
from arks_b200.tables import Tables, simple_endpoint, simple_quota, simple_token

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "error_replies.json")))
NOW = 1_700_000_000


def fixture_tables():
    return Tables([simple_token("alice", "team-a", "sk-x", "qwen-7b", [("rpm", 5), ("tpd", 900)], "q-main")],
                  [simple_quota("q-main", "team-a", [("prompt", 100), ("total", 600)])], [simple_endpoint("qwen-7b", "team-a")])


def rows():
    return [pytest.param(r, id=r["reason"] + "-" + r["ref"].split(" ")[0]) for r in GOLD["rows"]]


def model_span(body: bytes):
    k = body.find(b'"model":"')
    if k < 0:
        return 0, 0
    s = e = k + 9
    while body[e:e + 1] != b'"':  # the closing quote: an escaped one does not end the string
        e += 2 if body[e:e + 1] == b"\\" else 1
    return s, (e - s) | (0x80000000 if b"\\" in body[s:e] else 0)


@pytest.mark.parametrize("row", rows())
def test_python_mirror(row):
    t = fixture_tables()
    reason = getattr(abi, "R_" + row["reason"])
    want = (row["status"], row["header"], row["value"], row["message"])
    if row.get("phase") == "response":
        assert replies.response_error_reply(reason, t, 0, row.get("chunk", "x").encode()) == want
    else:
        body = row.get("body", '{"model":"qwen-7b"}').encode()
        off, ln = model_span(body)
        model = replies.decode_model(body[off:off + (ln & 0x7FFFFFFF)], bool(ln >> 31)) if reason != abi.R_NO_MODEL else ""
        got = replies.request_error_reply(reason, row.get("detail", 0), row.get("cur_usage", 0), row.get("limit_max", 0), NOW, t, 0,
                                          row.get("token", "sk-x").encode(), model)
        assert got == want


@pytest.fixture(scope="module")
def host():
    from test_cpp_host import CpuEngine
    t = fixture_tables()
    eng = CpuEngine(t, max_batch=16, max_bytes=1 << 16)
    eng.b.set_names(t)
    yield eng
    eng.close()


@pytest.mark.parametrize("row", rows())
def test_compiled_host(row, host):
    reason = getattr(abi, "R_" + row["reason"])
    want = (row["status"], row["header"], row["value"], row["message"])
    if row.get("phase") == "response":
        d = cpphost.ResponseDecision()
        d.reason = reason
        assert host.b.response_error_reply(d, 0, row.get("chunk", "x").encode()) == want
    else:
        body = row.get("body", '{"model":"qwen-7b"}').encode()
        d = cpphost.RequestDecision()
        d.reason, d.detail, d.cur_usage, d.limit_max, d.now_unix, d.qos = (reason, row.get("detail", 0), row.get("cur_usage", 0),
                                                                           row.get("limit_max", 0), NOW, 0)
        d.model_off, d.model_len = model_span(body) if reason != abi.R_NO_MODEL else (0, 0)
        assert host.b.request_error_reply(d, row.get("token", "sk-x").encode(), body) == want


def test_every_failure_reason_has_a_row():
    covered = {r["reason"] for r in GOLD["rows"]}
    failing = {n[2:] for n in dir(abi) if n.startswith("R_")} - {"OK", "PENDING", "QOS_GONE"}
    assert failing <= covered, failing - covered


def test_error_body_shape():
    assert replies.error_body('say "hi"\n', 429) == b'{"error":{"message":"say \\"hi\\"\\n","code":429}}'


def test_both_shapers_agree_on_random_decisions(host):
    """beyond the golden rows: for random failed decisions (every reason, rule / item index, usage, window time, odd tokens
    and model spellings) the compiled host and the Python mirror produce the same four strings"""
    import random
    r = random.Random(17)
    t = fixture_tables()
    reasons = [getattr(abi, n) for n in dir(abi) if n.startswith("R_") and n[2:] not in ("OK", "PENDING", "QOS_GONE")]
    request_side = [x for x in reasons if x not in (abi.R_STREAMING, abi.R_RESPONSE_UNMARSHAL, abi.R_RESPONSE_UNKNOWN, abi.R_QUOTA_CONFIG_RESP)]
    # (no \\n in a name: the C surface hands the four strings back newline-separated)
    models = ['qwen-7b', 'caf\\u00e9', 'a\\"b', 'x\\ry', '', 'm/with/slash', 'emoji \\ud83d\\ude00', 'tab\\there', 'lone \\ud83d!', 'rev \\ude00\\ud83d',
              'half \\ud83d\\t', 'sol\\/idus', 'caf\xe9 raw', 'esc\\t+raw \xff']
    for _ in range(3000):
        reason = r.choice(request_side)
        detail = r.randrange(2)
        cur, lim = r.randrange(0, 10**12), r.randrange(0, 10**12)
        now = NOW + r.randrange(0, 10**6)
        token = r.choice([b"sk-x", b"", b"sk-\xff\xfe", "sk-é".encode(), b"a" * 200])
        m = r.choice(models)
        body = ('{"messages":[],"model":"%s"}' % m).encode("latin-1")  # \xNN in the list above: bytes that are not UTF-8
        off, ln = model_span(body)
        d = cpphost.RequestDecision()
        d.reason, d.detail, d.cur_usage, d.limit_max, d.now_unix, d.qos = reason, detail, cur, lim, now, 0
        d.model_off, d.model_len = (off, ln) if reason != abi.R_NO_MODEL else (0, 0)
        got = host.b.request_error_reply(d, token, body)
        model = replies.decode_model(body[off:off + (ln & 0x7FFFFFFF)], bool(ln >> 31)) if reason != abi.R_NO_MODEL else ""
        want = replies.request_error_reply(reason, detail, cur, lim, now, t, 0, token, model)
        assert got == want, (reason, detail, token, m, got, want)
    for _ in range(500):
        reason = r.choice([abi.R_STREAMING, abi.R_RESPONSE_UNMARSHAL, abi.R_RESPONSE_UNKNOWN, abi.R_QUOTA_CONFIG_RESP])
        chunk = r.choice([b"x", b"", b'{"error":"boom"}', b"data: [DONE]\n\n", "café".encode(), b"\xff\x00bin"])
        d = cpphost.ResponseDecision()
        d.reason = reason
        assert host.b.response_error_reply(d, 0, chunk) == replies.response_error_reply(reason, t, 0, chunk), (reason, chunk)
