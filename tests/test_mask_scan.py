"""The fast path of the scan kernels (arks_b200/csrc/mask_scan.cuh) on CPU: its host driver is exactly what one lane of
fast_request_kernel / fast_response_kernel does (mask pass, table-driven token pass, member extraction).

Contract: the fast path is a filter. Whenever it ACCEPTS a document its fields equal what the exact engine (and the oracle)
extract; whenever it declines, nothing is claimed. It must accept the traffic it was built for (every generated chat
request / completion) or it is useless as a fast path."""
import random

import numpy as np
import pytest

import hostmachine as hm
import orklib
from arks_b200 import traffic
from jsonfuzz import Gen


def check_request(b):
    f = hm.fast_request(b)
    if f is None:
        return 0
    e = hm.engine_request(b)
    assert e is not None, ("fast path accepted a document the engine rejects", b)
    assert f == e, (b, f, e)
    rc, model, st, so, iu = orklib.parse_request_body(b)
    assert rc == 0 and (st, so, iu) == f[1:], (b, f)
    if not f[0][2]:
        assert model == b[f[0][0]:f[0][0] + f[0][1]], (b, f, model)
    return 1


def check_response(b):
    f = hm.fast_response(b)
    if f is None:
        return 0
    e = hm.engine_response(b)
    assert e is not None, ("fast path accepted a document the engine rejects", b)
    assert f == e, (b, f, e)
    rc, ml, usage = orklib.parse_response_body(b)
    assert rc == 0 and tuple(usage) == f[1] and (ml > 0) == (f[0][1] > 0), (b, f, usage)
    return 1


@pytest.mark.parametrize("seed", [81, 82])
def test_fuzzed_requests_never_disagree(seed):
    g = Gen(seed)
    took = sum(check_request(g.request()) for _ in range(40000))
    assert took > 2500  # a good part of the hostile documents is plain enough for the fast path to take part


@pytest.mark.parametrize("seed", [83, 84])
def test_fuzzed_responses_never_disagree(seed):
    g = Gen(seed)
    took = sum(check_response(g.response()) for _ in range(40000))
    assert took > 1500


def test_generated_traffic_is_accepted():
    rng = np.random.default_rng(5)
    n = 3000
    reqs = [traffic.chat_request_body_varied(rng, 1024, stream=bool(rng.random() < 0.3)) for _ in range(n)]
    assert sum(check_request(b) for b in reqs) == n
    assert sum(check_request(traffic.chat_request_body(rng, 1024, stream=bool(i & 1))) for i in range(300)) == 300
    resps = [traffic.chat_response_body_varied(rng, int(rng.integers(50, 401)), int(rng.integers(1, 513)), 600) for _ in range(n)]
    assert sum(check_response(b) for b in resps) == n
    assert sum(check_response(traffic.chat_response_body(rng, 7, 9, 600)) for _ in range(300)) == 300


def test_structure_corner_cases():
    r = random.Random(3)
    # documents that straddle the 32-byte chunks at every phase of a token
    pad = lambda n: '"p":"' + "x" * n + '",'
    for n in list(range(0, 80)) + list(range(980, 1060)) + [2000, 2010, 2030]:
        b = ('{' + pad(n) + '"model":"qwen-7b","stream":true,"stream_options":{"include_usage":true},"n":[1,2.5e-3,{"a":null}],"t":false}').encode()
        if len(b) <= 2048:
            assert check_request(b) == 1, n
        else:
            assert hm.fast_request(b) is None
        b2 = ('{' + pad(n) + '"usage":{"prompt_tokens":12,"completion_tokens":345,"total_tokens":357,"d":{"x":[1]}},"model":"m"}').encode()
        if len(b2) <= 2048:
            assert check_response(b2) == 1, n
    # backslash runs across chunk boundaries (and one that fills whole chunks)
    for n in range(20, 45):
        for k in range(1, 9):
            s = "y" * n + "\\\\" * k + ('\\"' if r.random() < 0.5 else "")
            b = ('{"a":"' + s + '","model":"m"}').encode()
            assert check_request(b) == 1, (n, k)
    for n in (0, 1, 7, 31, 32, 33):
        for k in (31, 32, 33, 64, 65, 66):
            b = ('{"a":"' + "z" * n + "\\" * k + ('' if k % 2 == 0 else 'n') + '","model":"m"}').encode()
            assert check_request(b) == 1, (n, k)
    # things the fast path must decline (the exact engine decides them)
    for b in [b'', b'null', b'[]', b'{"model":"m"} x', b'{"model":"m"}\x00', b'{"model":5}', b'{"model":"a","model":"b"}',
              b'{"mod\\u0065l":"m"}', b'{"model":"m","x":-01}', b'{"model":"m","x":"\x01"}', b'{"model":"m","x":"\\q"}',
              b'{"model":"m","x":[1,]}', b'{"model":"m",}', b'{"model":"m" "x":1}', b'{"model":"m","x":tru}',
              b'{"model":"m","x":1.}', b'{"model":"m","x":"unterminated}', b'{"model":"m"}}', b'{"model":"m"',
              b'{"stream_options":{"include_usage":1},"model":"m"}', b'{"stream":"true","model":"m"}',
              b'{"model":"m","d":' + b'[' * 40 + b']' * 40 + b'}']:
        assert hm.fast_request(b) is None, b
    # the member log only takes keys as long as a name that is read: 40 other members are no reason to decline, an escape
    # in ANY top-level key is (it may spell a name at another raw length), and so are more than 8 keys of a matching length
    many = ",".join('"param_%02d":%d' % (i, i) for i in range(40))
    assert check_request(('{' + many + ',"model":"m","stream":true}').encode()) == 1
    assert hm.fast_request(('{' + many + ',"model":"m","stream":true}').encode())[1] == 2
    assert hm.fast_request(b'{"te\\u006dperature":1,"model":"m"}') is None
    assert check_request(b'{"d":{"te\\u006dperature":1},"model":"m"}') == 1  # deeper keys are nobody's business
    five = ",".join('"k%04d":1' % i for i in range(9))
    assert hm.fast_request(('{' + five + ',"model":"m"}').encode()) is None
    assert check_request(('{' + ",".join('"k%04d":1' % i for i in range(7)) + ',"model":"m"}').encode()) == 1
    # and things it must get right
    assert hm.fast_request(b' {"MODEL" : "M\\u00e9" , "Stream":null, "STREAM_OPTIONS":null} \n')[1:] == (0, 0, 0)
    assert hm.fast_request(b'{"stream_options":{"x":{"include_usage":false},"include_usage":true},"model":"m"}')[1:] == (0, 1, 2)
    assert hm.fast_response(b'{"usage":null,"model":"m"}')[1] == (0, 0, 0)
    assert hm.fast_response(b'{"usage":{"total_tokens":"7"},"model":"m"}') is None
    assert hm.fast_response(b'{"usage":{"total_tokens":7,"prompt_tokens":null},"model":"m"}') is None
