"""CPU: the C-ABI library loads and exports exactly the symbols include/linnaeus_b200.h
declares (no compute calls without a GPU), and the ctypes table matches the header."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "linnaeus_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    decls = {}
    for m in re.finditer(r"\b(?:int|const char\*)\s+(lnx_\w+)\s*\(([^;]*?)\)\s*;", src, flags=re.S):
        args = m.group(2).strip()
        n = 0 if args in ("", "void") else len([a for a in args.split(",") if a.strip()])
        decls[m.group(1)] = n
    return decls


@pytest.fixture(scope="module")
def lib():
    from linnaeus_b200 import _build, _lib

    if not os.path.exists(_lib.LIB_PATH):
        _build.build(verbose=False)
    return _lib.load()


def test_header_declares_expected_surface():
    d = _declared()
    assert len(d) >= 28
    for name in ("lnx_gemm", "lnx_layernorm_fwd", "lnx_layernorm_bwd", "lnx_dwconv7_fwd", "lnx_dwconv7_wgrad", "lnx_attn_fwd", "lnx_attn_bwd",
                 "lnx_rope_qk_fwd", "lnx_rope_qk_bwd", "lnx_loss_fwd", "lnx_loss_reduce", "lnx_loss_bwd", "lnx_adamw", "lnx_sumsq"):
        assert name in d


def test_every_declared_symbol_is_exported(lib):
    raw = ctypes.CDLL(lib._name)
    for name in _declared():
        assert hasattr(raw, name), f"{name} declared in the header but not exported"


def test_ctypes_table_matches_header(lib):
    from linnaeus_b200 import _lib

    d = _declared()
    assert set(_lib.SIGNATURES) == set(d)
    for name, argtypes in _lib.SIGNATURES.items():
        assert len(argtypes) == d[name], (name, len(argtypes), d[name])


def test_error_strings_and_version(lib):
    from linnaeus_b200 import _lib

    assert lib.lnx_version() >= 100
    assert _lib.strerror(0) == "ok"
    assert "shape" in _lib.strerror(-1)
    assert "null" in _lib.strerror(-6)


def test_argument_validation_needs_no_gpu(lib):
    """Null / shape checks run before any CUDA call, so they are testable on CPU."""
    assert lib.lnx_layernorm_fwd(None, None, None, None, None, None, None, 4, 32, 1e-5, 0, None) == -6
    assert lib.lnx_gemm(1, None, 8, 0, None, 8, 0, None, 1, 4, 4, 4, None, 0, None, None, None, None, None, 0, None, 0, 0, None) == -6
    assert lib.lnx_sumsq(None, 8, None, None, None) == -6
    buf = (ctypes.c_float * 64)()
    p = ctypes.addressof(buf)
    assert lib.lnx_layernorm_fwd(p, p, p, None, p, None, None, 0, 32, 1e-5, 0, None) == -1  # rows == 0
    assert lib.lnx_dwconv7_fwd(p, p, 0, None, None, p, 1, 8, 8, 33, 0, None) == -1  # C % 32 != 0
    assert lib.lnx_layernorm_fwd(p, p, p, None, p, None, None, 1, 32, 1e-5, 7, None) == -2  # bad dtype


def test_missing_library_fails_loudly(monkeypatch):
    from linnaeus_b200 import _lib

    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/liblinnaeus_b200.so")
    with pytest.raises(_lib.LnxError):
        _lib.load()
