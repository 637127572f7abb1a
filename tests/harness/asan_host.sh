#!/bin/sh
# Memory / undefined-behaviour check for the compiled host (host/cpp) and the oracle-backed ABI shim: the same drive as
# tsan_host.sh (12 blocking stream threads, 10 response threads, 3 open-loop producers, queue depth 1 and 2) under
# AddressSanitizer + UBSan. Prints reports, if any, then "asan run done". CPU only.
set -e
R=$(cd "$(dirname "$0")/../.." && pwd)
T=${TMPDIR:-/tmp}/arks_asan
mkdir -p "$T"
make -C "$R/oracle" -s libarks_oracle.so
gcc -O1 -g -fsanitize=address,undefined -shared -fPIC -o "$T/libarksgw_shim.so" "$R/tests/abi_shim.c" -L"$R/oracle" -larks_oracle -Wl,-rpath,"$R/oracle"
g++ -O1 -g -fsanitize=address,undefined -std=c++20 -shared -fPIC -pthread -o "$T/libarkshost.so" "$R/host/cpp/arks_host.cc" -L"$T" -larksgw_shim -Wl,-rpath,"$T"
cat > "$T/run.py" <<PY
import sys, ctypes as C, numpy as np
sys.path.insert(0, "$R"); sys.path.insert(0, "$R/tests")
from arks_b200 import cpphost, traffic, abi
shim = C.CDLL("$T/libarksgw_shim.so", mode=C.RTLD_GLOBAL)
L = cpphost.load("$T/libarkshost.so")
w = traffic.Workload(40, seed=11)
ts = w.tables.c_struct(); ctx = C.c_void_p()
shim.arks_shim_create.argtypes = [C.c_void_p, C.POINTER(C.c_void_p)]
assert shim.arks_shim_create(C.byref(ts), C.byref(ctx)) == 0
for depth in (1, 2):
    b = cpphost.Batcher(L, ctx, max_batch=256, max_bytes=1 << 20, max_inflight=depth)
    b.set_fixed_clock(1_700_000_000)
    req = w.request_batch(3000, 1_700_000_000, seed=5, stream_frac=0.3, noise_frac=0.1)
    d, lat, wall = b.run_requests(req, threads=12)
    ok = np.flatnonzero(d["reason"] == 0)
    adm = abi.RequestResult.empty(len(ok)); adm.reason[:] = 0; adm.qos[:] = d["qos"][ok]; adm.flags[:] = d["flags"][ok]
    resp = w.response_batch(adm, 1_700_000_001, seed=6)
    b.run_responses(resp, threads=10)
    b.open_loop_requests(req, 100000, producers=3)
    # a config thread republishes the tables (and their names, NameBook) while streams run and replies are shaped
    import threading
    stop = threading.Event()
    def config():
        while not stop.is_set():
            b.load_tables(w.tables)
    def shaper():
        k = 0
        while not stop.is_set():
            dd = cpphost.RequestDecision(); dd.reason, dd.qos, dd.gen = 8, 0, k % 40
            b.request_error_reply(dd, b"sk", b"{}"); k += 1
    ths = [threading.Thread(target=config), threading.Thread(target=shaper)]
    for t in ths: t.start()
    d2, _, _ = b.run_requests(req, threads=12)
    stop.set()
    for t in ths: t.join()
    assert (d2["reason"] != 255).all()
    print("depth", depth, b.stats())
    b.close()
print("asan run done")
PY
ASAN_OPTIONS=detect_leaks=0 UBSAN_OPTIONS=print_stacktrace=1 LD_PRELOAD=$(gcc -print-file-name=libasan.so) python "$T/run.py"
