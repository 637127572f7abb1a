"""ctypes wrapper around oracle/libarks_oracle.so — the CPU restatement of the reference's Go path.

TEST INFRASTRUCTURE: only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this module (the product path never does).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

from arks_b200 import abi
from arks_b200.abi import (ArksRequestBatch, ArksRequestResult, ArksResponseBatch, ArksResponseResult, ArksTables,
                           RequestBatch, RequestResult, ResponseBatch, ResponseResult)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
LIB_PATH = os.path.join(ORACLE_DIR, "libarks_oracle.so")


def build_oracle(force: bool = False) -> str:
    srcs = [os.path.join(ORACLE_DIR, f) for f in os.listdir(ORACLE_DIR) if f.endswith((".c", ".h"))]
    srcs.append(os.path.join(ROOT, "include", "arks_gateway.h"))
    stale = force or not os.path.exists(LIB_PATH) or any(
        os.path.getmtime(s) > os.path.getmtime(LIB_PATH) for s in srcs)
    if stale:
        subprocess.check_call(["make", "-C", ORACLE_DIR, "-s", "-B", "libarks_oracle.so"])
    return LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build_oracle())
        L.ork_create.restype = C.c_void_p
        L.ork_create.argtypes = [C.POINTER(ArksTables)]
        L.ork_destroy.argtypes = [C.c_void_p]
        L.ork_reload.argtypes = [C.c_void_p, C.POINTER(ArksTables)]
        L.ork_set_precharge.argtypes = [C.c_void_p, C.c_int]
        L.ork_set_precharge.restype = None
        L.ork_set_estimates.argtypes = [C.c_void_p, abi.u32p, C.c_uint32]
        L.ork_set_estimates.restype = None
        L.ork_update_endpoint_weights.argtypes = [C.c_void_p, C.c_uint32, C.c_uint32, abi.i32p]
        for f in ("ork_request_batch",):
            getattr(L, f).argtypes = [C.c_void_p, C.POINTER(ArksRequestBatch), C.POINTER(ArksRequestResult)]
        L.ork_request_batch_mt.argtypes = [C.c_void_p, C.POINTER(ArksRequestBatch), C.POINTER(ArksRequestResult),
                                           C.c_int]
        L.ork_response_batch.argtypes = [C.c_void_p, C.POINTER(ArksResponseBatch), C.POINTER(ArksResponseResult)]
        L.ork_response_batch_mt.argtypes = [C.c_void_p, C.POINTER(ArksResponseBatch), C.POINTER(ArksResponseResult),
                                            C.c_int]
        L.ork_snapshot_quota.argtypes = [C.c_void_p, abi.i64p]
        L.ork_set_quota_usage.argtypes = [C.c_void_p, C.c_uint32, abi.i64p]
        L.ork_incr_quota_usage.argtypes = [C.c_void_p, C.c_uint32, abi.i64p]
        L.ork_snapshot_rate.argtypes = [C.c_void_p, C.c_int64, abi.i64p]
        L.ork_snapshot_metrics.argtypes = [C.c_void_p, abi.i64p]
        L.ork_sync_quota_usage.argtypes = [C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), abi.i64p, C.c_int]
        L.ork_parse_request_body.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t,
                                             C.POINTER(C.c_size_t), C.POINTER(C.c_int), C.POINTER(C.c_int),
                                             C.POINTER(C.c_int)]
        L.ork_parse_response_body.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t), abi.i64p]
        L.ork_parse_sse_chunk.argtypes = [C.c_char_p, C.c_size_t, abi.i64p]
        L.ork_sse_events.argtypes = [C.c_char_p, C.c_size_t, C.c_char_p, C.c_size_t, C.POINTER(C.c_size_t)]
        L.ork_window_start.restype = C.c_int64
        L.ork_window_start.argtypes = [C.c_int64, C.c_int]
        L.ork_rate_key.restype = C.c_size_t
        L.ork_rate_key.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_int64, C.c_char_p,
                                   C.c_size_t]
        L.ork_quota_key.restype = C.c_size_t
        L.ork_quota_key.argtypes = [C.c_char_p, C.c_char_p, C.c_char_p, C.c_int, C.c_char_p, C.c_size_t]
        L.ork_weighted_pick.restype = C.c_int32
        L.ork_weighted_pick.argtypes = [abi.i32p, C.c_uint32, C.c_uint64]
        _lib = L
    return _lib


class Oracle:
    def __init__(self, tables):
        self.tables = tables
        self._ts = tables.c_struct()
        self.h = lib().ork_create(C.byref(self._ts))
        if not self.h:
            raise ValueError("oracle rejected the tables")

    def __del__(self):
        if getattr(self, "h", None):
            lib().ork_destroy(self.h)
            self.h = None

    def set_precharge(self, on: bool):
        lib().ork_set_precharge(self.h, int(bool(on)))

    def set_estimates(self, est):
        """prompt token counts of the NEXT request batch (N4; the oracle has no tokenizer of its own)"""
        self._est = np.ascontiguousarray(est, np.uint32)
        lib().ork_set_estimates(self.h, abi.ptr(self._est, abi.u32p), len(self._est))

    def reload(self, tables):
        ts = tables.c_struct()
        rc = lib().ork_reload(self.h, C.byref(ts))
        if rc:
            raise ValueError(f"ork_reload rc={rc}")
        self.tables, self._ts = tables, ts

    def update_endpoint_weights(self, ep, weights):
        w = np.ascontiguousarray(weights, np.int32)
        return lib().ork_update_endpoint_weights(self.h, ep, len(w), abi.ptr(w, abi.i32p))

    def request_batch(self, b: RequestBatch, threads: int = 0) -> RequestResult:
        r = RequestResult.empty(b.n)
        bs, rs = b.c_struct(), r.c_struct()
        if threads:
            rc = lib().ork_request_batch_mt(self.h, C.byref(bs), C.byref(rs), threads)
        else:
            rc = lib().ork_request_batch(self.h, C.byref(bs), C.byref(rs))
        if rc:
            raise RuntimeError(f"ork_request_batch rc={rc}")
        return r

    def response_batch(self, b: ResponseBatch, threads: int = 0) -> ResponseResult:
        r = ResponseResult.empty(b.n)
        bs, rs = b.c_struct(), r.c_struct()
        if threads:
            rc = lib().ork_response_batch_mt(self.h, C.byref(bs), C.byref(rs), threads)
        else:
            rc = lib().ork_response_batch(self.h, C.byref(bs), C.byref(rs))
        if rc:
            raise RuntimeError(f"ork_response_batch rc={rc}")
        return r

    def snapshot_quota(self) -> np.ndarray:
        out = np.zeros((self.tables.n_quotas, 3), np.int64)
        lib().ork_snapshot_quota(self.h, abi.ptr(out, abi.i64p))
        return out

    def sync_quota_usage(self, status_present, status_used, restore=False) -> np.ndarray:
        """syncQuotaUsage, quota by quota in list order, same in/out convention as Gateway.sync_quota_usage"""
        action = np.zeros(self.tables.n_quotas, np.uint8)
        for q in range(self.tables.n_quotas):
            p = C.c_uint32(int(status_present[q]))
            u = np.ascontiguousarray(status_used[q], np.int64)
            action[q] = lib().ork_sync_quota_usage(self.h, q, C.byref(p), abi.ptr(u, abi.i64p), 1 if restore else 0)
            status_present[q] = p.value
            status_used[q] = u
        return action

    def snapshot_metrics(self) -> np.ndarray:
        out = np.zeros((self.tables.n_qos, abi.METRIC_COLS), np.int64)
        lib().ork_snapshot_metrics(self.h, abi.ptr(out, abi.i64p))
        return out

    def snapshot_rate(self, now) -> np.ndarray:
        out = np.zeros((self.tables.n_qos, 4), np.int64)
        lib().ork_snapshot_rate(self.h, int(now), abi.ptr(out, abi.i64p))
        return out

    def set_quota_usage(self, q, usage):
        u = np.ascontiguousarray(usage, np.int64)
        return lib().ork_set_quota_usage(self.h, q, abi.ptr(u, abi.i64p))

    def incr_quota_usage(self, q, delta):
        u = np.ascontiguousarray(delta, np.int64)
        return lib().ork_incr_quota_usage(self.h, q, abi.ptr(u, abi.i64p))


def parse_request_body(body: bytes):
    """-> (err, model bytes, stream tri-state, stream_options present, include_usage tri-state)"""
    buf = C.create_string_buffer(4096)
    ml, st, so, iu = C.c_size_t(), C.c_int(), C.c_int(), C.c_int()
    rc = lib().ork_parse_request_body(body, len(body), buf, 4096, C.byref(ml), C.byref(st), C.byref(so), C.byref(iu))
    return rc, buf.raw[:min(ml.value, 4096)], st.value, so.value, iu.value


def parse_response_body(body: bytes):
    ml = C.c_size_t()
    u = np.zeros(3, np.int64)
    rc = lib().ork_parse_response_body(body, len(body), C.byref(ml), abi.ptr(u, abi.i64p))
    return rc, ml.value, tuple(int(x) for x in u)


def parse_sse_chunk(body: bytes):
    u = np.zeros(3, np.int64)
    rc = lib().ork_parse_sse_chunk(body, len(body), abi.ptr(u, abi.i64p))
    return rc, tuple(int(x) for x in u)


def sse_events(body: bytes):
    """-> (n or negative error, [(type bytes, data bytes, n_data_lines)]) as the oracle's decoder dispatches them"""
    import struct
    cap = 64 + 16 * len(body) + 2 * len(body)
    buf = C.create_string_buffer(cap)
    used = C.c_size_t()
    n = lib().ork_sse_events(body, len(body), buf, cap, C.byref(used))
    out, p, raw = [], 0, buf.raw[:used.value]
    while p < len(raw):
        tl, dl, nl = struct.unpack_from("<III", raw, p)
        out.append((raw[p + 12:p + 12 + tl], raw[p + 12 + tl:p + 12 + tl + dl], nl))
        p += 12 + tl + dl
    return n, out


def window_start(now, rule):
    return lib().ork_window_start(int(now), int(rule))


def rate_key(prefix, ns, user, model, rule, now):
    out = C.create_string_buffer(512)
    n = lib().ork_rate_key(prefix.encode(), ns.encode(), user.encode(), model.encode(), rule, int(now), out, 512)
    return out.raw[:n].decode()


def quota_key(prefix, ns, quota, ty):
    out = C.create_string_buffer(512)
    n = lib().ork_quota_key(prefix.encode(), ns.encode(), quota.encode(), ty, out, 512)
    return out.raw[:n].decode()


def weighted_pick(weights, r):
    w = np.ascontiguousarray(weights, np.int32)
    return lib().ork_weighted_pick(abi.ptr(w, abi.i32p), len(w), int(r) & (2**64 - 1))
