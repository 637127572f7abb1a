"""A longer run of the filter contracts than the CPU suite affords: the host drivers of the lane-per-document fast path
(mask_scan.cuh) and of the warp-per-document latency path (warp_scan.cuh) — the same ARKS_HD code the kernels execute —
against the exact engine and the oracle on hostile documents of fresh seeds. Any disagreement raises.

usage: python tests/harness/long_fuzz.py [first_seed=1000] [seeds=8] [documents_per_seed=100000]"""
import os, sys
_R = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, _R); sys.path.insert(0, os.path.join(_R, "tests"))
import test_mask_scan as lane, test_warp_scan as warp
from jsonfuzz import Gen

first, seeds, n = (int(sys.argv[i]) if len(sys.argv) > i else d for i, d in ((1, 1000), (2, 8), (3, 100000)))
took = [0, 0, 0, 0]
for s in range(first, first + seeds):
    g = Gen(s)
    for _ in range(n):
        q, r = g.request(), g.response()
        took[0] += lane.check_request(q); took[1] += lane.check_response(r)
        took[2] += warp.check_request(q); took[3] += warp.check_response(r)
    print(f"seed {s}: no disagreement; accepted so far lane req/resp {took[0]}/{took[1]}, warp req/resp {took[2]}/{took[3]}", flush=True)
