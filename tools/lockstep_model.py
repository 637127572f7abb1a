"""Why heterogeneous bodies are slow on a lane-per-body scan: a CPU model of the lock step. For every body the engine's
own work profile (advances and events per 128-byte window, from the host build of json_engine.cuh); bodies are grouped 32
to a warp the way the library does it (length order) and the warp's cost per window is the MAXIMUM over its lanes.
Prints, per workload: advances one body needs, advance slots a warp spends per body-row (sum over windows of the max),
the ratio, and the same for events."""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import numpy as np
import hostmachine as hm
from arks_b200 import traffic

def model(name, bodies, kind, order):
    if order:
        bodies = sorted(bodies, key=len)
    prof = [hm.work_profile(b, kind) for b in bodies]
    per_body = np.array([p[0].sum() for p in prof]); ev_body = np.array([p[1].sum() for p in prof])
    slots, evslots, nw = 0, 0, 0
    for g in range(0, len(prof) - 31, 32):
        grp = prof[g:g + 32]
        wmax = max(len(p[0]) for p in grp)
        A = np.zeros((32, wmax)); E = np.zeros((32, wmax))
        for i, p in enumerate(grp):
            A[i, :len(p[0])] = p[0]; E[i, :len(p[1])] = p[1]
        slots += A.max(axis=0).sum(); evslots += E.max(axis=0).sum(); nw += 1
    print(f"{name:46s} advances/body {per_body.mean():6.0f}  warp slots {slots / nw:7.0f}  ratio {slots / nw / per_body.mean():4.1f}x   "
          f"events/body {ev_body.mean():5.0f}  warp event rounds >= {evslots / nw:5.0f}")

rng = np.random.default_rng(5)
N = 4096
uni = [traffic.chat_request_body(rng, 1024) for _ in range(N)]
var = [traffic.chat_request_body_varied(rng, 1024) for _ in range(N)]
model("requests, one shape 1024 B", uni, 0, True)
model("requests, client applications, arrival order", var, 0, False)
model("requests, client applications, length order", var, 0, True)
ru = [traffic.chat_response_body(rng, 100, 50, 600) for _ in range(N)]
rl = [traffic.chat_response_body(rng, 100, 50, int(rng.integers(330, 1000))) for _ in range(N)]
rv = [traffic.chat_response_body_varied(rng, 100, 50, 600) for _ in range(N)]
model("completions, one shape 600 B", ru, 1, True)
model("completions, one shape, lengths 330-1000, arrival", rl, 1, False)
model("completions, one shape, lengths 330-1000, ordered", rl, 1, True)
model("completions, three dialects, length order", rv, 1, True)

# --- which lane assignment would shrink the lock-step loss? (keys a cheap device pre-pass could compute) ---
def model_key(name, bodies, kind, key):
    bodies = sorted(bodies, key=key)
    prof = [hm.work_profile(b, kind) for b in bodies]
    per_body = np.array([p[0].sum() for p in prof])
    slots, nw = 0, 0
    for g in range(0, len(prof) - 31, 32):
        grp = prof[g:g + 32]
        wmax = max(len(p[0]) for p in grp)
        A = np.zeros((32, wmax))
        for i, p in enumerate(grp):
            A[i, :len(p[0])] = p[0]
        slots += A.max(axis=0).sum(); nw += 1
    print(f"  order by {name:52s} warp slots {slots / nw:6.0f}  ratio {slots / nw / per_body.mean():4.2f}x")

if "--orders" in sys.argv:
    def qwin(b):  # quotes per 128-byte window
        a = np.frombuffer(b, np.uint8) == 0x22
        return tuple(np.add.reduceat(a, np.arange(0, len(b), 128)).tolist())
    print("requests, client applications:")
    model_key("length", var, 0, len)
    model_key("number of quotes, then length", var, 0, lambda b: (b.count(b'"'), len(b)))
    model_key("length, then number of quotes", var, 0, lambda b: (len(b) >> 5, b.count(b'"')))
    model_key("quotes per window (lexicographic)", var, 0, qwin)
    model_key("windows, then quotes per window", var, 0, lambda b: ((len(b) + 127) // 128,) + qwin(b))
    model_key("total advances (oracle of the work, not computable)", var, 0, lambda b: int(hm.work_profile(b, 0)[0].sum()))
    model_key("per-window advances (ideal, not computable)", var, 0, lambda b: tuple(hm.work_profile(b, 0)[0].tolist()))
