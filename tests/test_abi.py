"""CPU: the C-ABI library builds, loads without a GPU and exports exactly what include/istvt_b200.h declares;
compute entry points reject bad arguments without touching a device."""
import ctypes
import os
import re

import pytest

from helpers import ROOT, pkg

HEADER = os.path.join(ROOT, "include", "istvt_b200.h")


def _declared():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(istvt_[a-z0-9_]+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as ge
    ge.build()
    return pkg()._lib.lib()


def test_header_symbols_exported(lib):
    names = _declared()
    assert len(names) >= 17
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"


def test_python_signatures_cover_header():
    assert sorted(pkg()._lib.SIGNATURES.keys()) == _declared()


def test_error_strings_and_version(lib):
    assert lib.istvt_abi_version() == 1
    assert b"invalid argument" in lib.istvt_error_string(-1)
    assert lib.istvt_error_string(0) == b"ok"
    assert lib.istvt_launch_count() >= 0


def test_argument_validation_without_gpu(lib):
    """Null pointers / bad sizes are rejected before any CUDA call."""
    assert lib.istvt_layernorm_fwd(None, 1, None, None, None, 0, 4, 728, 1e-5, None) == -1
    assert lib.istvt_gemm_fwd(None, 8, None, 8, None, 8, 0, 128, 128, 64, None, None, 0, 0, None) == -1
    assert lib.istvt_attn_spatial_fwd(None, None, None, 0, 1, 362, 8, 0.125, None) == -1
    assert lib.istvt_dwconv3x3_fwd(None, None, None, 0, 1, 8, 8, 8, 0, None) == -1
    assert lib.istvt_attn_joint_fwd(None, None, 0, 1, 2167, 8, 0.125, None) == -1
    assert lib.istvt_token_build_fwd(None, 0, None, None, None, 1, 361, 728, 1, None) == -1
    assert lib.istvt_mean_rows_fwd(None, None, 1, 7, 728, None) == -1
    buf = ctypes.create_string_buffer(64)
    p = ctypes.cast(buf, ctypes.c_void_p)
    assert lib.istvt_layernorm_fwd(p, 1, p, p, p, 0, 4, 730, 1e-5, None) == -1      # dim % 4
    assert lib.istvt_gemm_fwd(p, 7, p, 8, p, 8, 0, 128, 128, 64, None, None, 0, 0, None) == -1   # lda % 8
    assert lib.istvt_attn_joint_fwd(p, p, 0, 1, 0, 8, 0.125, None) == -1                         # no tokens
    assert lib.istvt_attn_joint_fwd(p, p, 0, 1, 64, 8, -1.0, None) == -1                         # scale <= 0
    assert lib.istvt_token_build_fwd(p, 0, p, None, p, 1, 361, 730, 1, None) == -1               # dim % 4


def test_missing_library_fails_loudly(monkeypatch):
    _lib = pkg()._lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libistvt_b200.so")
    with pytest.raises(RuntimeError, match="not built"):
        _lib.lib()
