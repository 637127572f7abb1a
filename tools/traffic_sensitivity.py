"""Which property of the request / response bodies costs the scan kernels time? Times one 64k wave per variant."""
import sys, json, numpy as np
import os
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import __graft_entry__ as ge; ge.build()
from arks_b200 import traffic, abi
from arks_b200.abi import RequestBatch, ResponseBatch
from arks_b200.gateway import Gateway
N = 65536
w = traffic.Workload(10000, seed=1)
g = Gateway(0, N, 120 << 20); g.load_tables(w.tables)
rng = np.random.default_rng(7)
toks = [w.token_strings[int(t)] for t in rng.integers(0, 10000, N)]
now = [1_700_000_000]

SORT = "sorted" in sys.argv

def time_req(name, make, nt=N):
    templ = [make() for _ in range(nt)]
    bodies = templ if nt == N else [templ[int(k)] for k in rng.integers(0, nt, N)]
    if SORT: bodies.sort(key=len)
    if "sigsort" in sys.argv: bodies.sort(key=lambda x: (x[:48], len(x)))
    b = RequestBatch.from_lists(bodies, toks, now[0])
    g.set_profiling(True)
    ts = []
    for _ in range(6):
        now[0] += 86400; b.now_unix = now[0]
        a = g.handle_request_body(b); ts.append(g.last_kernel_ms()[0])
    L = np.array([len(x) for x in templ])
    print(f"REQ  {name:34s} scan {np.median(ts[1:])*1e3:7.1f} us   mean {L.mean():6.0f} B  max {L.max():5d}  ok {(a.reason==0).mean():.2f}", flush=True)

def time_resp(name, make, nt=N, flags=abi.RESP_END_OF_STREAM):
    templ = [make() for _ in range(nt)]
    bodies = templ if nt == N else [templ[int(k)] for k in rng.integers(0, nt, N)]
    if SORT: bodies.sort(key=len)
    b = ResponseBatch.from_lists(bodies, rng.integers(0, 10000, N).astype(np.int32), [flags] * N, now[0])
    g.set_profiling(True)
    ts = []
    for _ in range(6):
        now[0] += 86400; b.now_unix = now[0]
        c = g.handle_response_body(b); ts.append(g.last_kernel_ms()[0])
    L = np.array([len(x) for x in templ])
    print(f"RESP {name:34s} scan {np.median(ts[1:])*1e3:7.1f} us   mean {L.mean():6.0f} B  max {L.max():5d}  ok {(c.reason==0).mean():.2f}", flush=True)

T = traffic
QUICK = "quick" in sys.argv
if QUICK:
    time_req("uniform 1024", lambda: T.chat_request_body(rng, 1024), nt=2048)
    time_req("varied (bench)", lambda: T.chat_request_body_varied(rng, 1024))
    time_resp("uniform 600", lambda: T.chat_response_body(rng, 100, 50, 600), nt=2048)
    time_resp("uniform shape, lengths 330-1000", lambda: T.chat_response_body(rng, 100, 50, int(rng.integers(330, 1000))))
    time_resp("varied (bench)", lambda: T.chat_response_body_varied(rng, 100, 50, 600))
    sys.exit(0)
time_req("uniform 1024", lambda: T.chat_request_body(rng, 1024))
time_req("uniform, lengths 400-1650", lambda: T.chat_request_body(rng, int(rng.integers(400, 1650))))
def esc_uniform():
    head = '{"model":"qwen-7b","messages":[{"role":"user","content":"'
    return (head + T._prose(rng, 1024 - len(head) - 6) + '"}]}').encode()
time_req("uniform shape, prose with escapes", esc_uniform)
def multi_msg():
    k = int(rng.integers(1, 7))
    msgs = ",".join('{"role":"user","content":"%s"}' % T._text(rng, 900 // k) for _ in range(k))
    return ('{"model":"qwen-7b","messages":[%s]}' % msgs).encode()
time_req("1-6 messages, plain text", multi_msg)
def params_only():
    p = ',"temperature":0.7,"max_tokens":%d,"top_p":0.9' % int(rng.integers(16, 4096)) if rng.random() < 0.5 else ""
    head = '{"model":"qwen-7b"%s,"messages":[{"role":"user","content":"' % p
    return (head + T._text(rng, 1024 - len(head) - 6) + '"}]}').encode()
time_req("optional numeric params", params_only)
time_req("varied (bench)", lambda: T.chat_request_body_varied(rng, 1024))
time_resp("uniform 600", lambda: T.chat_response_body(rng, 100, 50, 600))
time_resp("uniform shape, lengths 330-1000", lambda: T.chat_response_body(rng, 100, 50, int(rng.integers(330, 1000))))
time_resp("varied (bench)", lambda: T.chat_response_body_varied(rng, 100, 50, 600))
