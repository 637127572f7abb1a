"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/arks_gateway.h declares, refuses to
run without a CUDA device (no CPU fallback), and its pure-host entry point matches the oracle."""
import ctypes as C
import os
import re

import pytest

import orklib
from arks_b200 import abi, gateway

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def built():
    import __graft_entry__ as ge
    ge.build()


def header_functions():
    src = open(os.path.join(ROOT, "include", "arks_gateway.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(arks_[a-z_0-9]+)\s*\(", src)))


def test_exports_every_declared_symbol():
    L = gateway.lib()
    names = header_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(L, n), f"libarksgw.so does not export {n}"
    assert sorted(gateway.EXPORTED) == names
    assert L.arks_abi_version() == 2


def test_no_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(gateway.ArksError) as e:
        gateway.Gateway(0, 1024, 1 << 20)
    assert e.value.code == abi.E_NO_DEVICE


HEADER_CASES = [
    ([("content-type", "application/json"), ("authorization", "Bearer sk-test123456")], b"sk-test123456"),
    ([("Authorization", "Bearer abc")], b"abc"),
    ([("AUTHORIZATION", "Bearer  two-spaces")], b" two-spaces"),          # no trimming
    ([("authorization", "bearer abc")], b""),                              # prefix is case-sensitive
    ([("authorization", "Basic x"), ("authorization", "Bearer second")], b"second"),
    ([("authorization", "Bearer "), ("authorization", "Bearer later")], b""),  # first prefixed header ends the scan
    ([("x-authorization", "Bearer abc")], b""),
    ([], b""),
]


@pytest.mark.parametrize("headers,want", HEADER_CASES)
def test_extract_bearer(headers, want):
    """HandleRequestHeaders, pkg/gateway/handle_request.go:38-46 — product (C ABI, host-only) and oracle agree."""
    assert gateway.extract_bearer(headers) == want
    n = len(headers)
    kb = [k.encode() for k, _ in headers]
    vb = [v.encode() for _, v in headers]
    KA = (C.c_char_p * max(n, 1))(*kb)
    VA = (C.c_char_p * max(n, 1))(*vb)
    KL = (C.c_size_t * max(n, 1))(*[len(k) for k in kb])
    VL = (C.c_size_t * max(n, 1))(*[len(v) for v in vb])
    tok = C.c_void_p()
    f = orklib.lib().ork_extract_bearer
    f.restype = C.c_size_t
    ln = f(KA, KL, VA, VL, C.c_size_t(n), C.byref(tok))
    assert (C.string_at(tok.value, ln) if ln else b"") == want
