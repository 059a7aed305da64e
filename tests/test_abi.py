"""CPU tier: the C-ABI shared library loads, exports exactly what include/medt_axial.h
declares, and the ctypes mirrors agree with the header (no GPU, no compute calls)."""
import ctypes
import os
import re

import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(REPO, "include", "medt_axial.h")


def _declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mt_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    from medt_b200 import _cabi
    lib = _cabi.load_library()
    names = _declared_functions()
    assert "mt_axial_forward" in names and "mt_axial_backward" in names
    for n in names:
        assert hasattr(lib, n), f"{n} declared in medt_axial.h but not exported"
    assert sorted(_cabi.EXPORTS) == names


def test_version_and_error_string():
    from medt_b200 import _cabi
    lib = _cabi.load_library()
    assert lib.mt_version() == 1
    assert isinstance(lib.mt_last_error(), bytes)


def test_struct_layout_matches_header():
    from medt_b200 import _cabi
    src = open(HEADER).read()
    body = re.search(r"typedef struct MtAxialParams \{(.*?)\} MtAxialParams;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\*\s*([a-z_0-9]+)\s*;", body)
    assert fields == [f for f, _ in _cabi.MtAxialParams._fields_]
    body = re.search(r"typedef struct MtAxialGrads \{(.*?)\} MtAxialGrads;", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\*\s*([a-z_0-9]+)\s*;", body)
    assert fields == [f for f, _ in _cabi.MtAxialGrads._fields_]
    assert ctypes.sizeof(_cabi.MtAxialDesc) == 44
    assert ctypes.sizeof(_cabi.MtAxialParams) == 18 * 8


def test_size_queries_and_argument_errors():
    """Pure host logic of the ABI: geometry validation and buffer sizing (no launches)."""
    from medt_b200 import _cabi
    lib = _cabi.load_library()
    d = _cabi.MtAxialDesc(4, 16, 64, 64, 0, 1, 0, 1, 1, 1e-5, 0.1)
    sv = lib.mt_axial_saved_bytes(ctypes.byref(d))
    B, C, L = 4 * 64, 16, 64
    assert sv >= 4 * (B * 2 * C * L + B * 2 * C * L + B * 8 * L)
    assert lib.mt_axial_bwd_workspace_bytes(ctypes.byref(d)) >= 4 * (B * C * L + B * 2 * C * L)
    bad = _cabi.MtAxialDesc(4, 24, 64, 64, 0, 1, 0, 1, 1, 1e-5, 0.1)       # C % 16 != 0
    assert lib.mt_axial_saved_bytes(ctypes.byref(bad)) == 0
    p = _cabi.MtAxialParams()
    rc = lib.mt_axial_forward(ctypes.byref(bad), ctypes.byref(p), None, None, None, None)
    assert rc == 1 and b"MtAxialDesc" in lib.mt_last_error()
    rc = lib.mt_axial_forward(ctypes.byref(d), ctypes.byref(p), None, None, None, None)
    assert rc == 1 and b"null" in lib.mt_last_error()
    odd = _cabi.MtAxialDesc(4, 48, 8, 8, 0, 1, 2, 1, 1, 1e-5, 0.1)         # C/16 == 3: not covered
    rc = lib.mt_axial_forward(ctypes.byref(odd), ctypes.byref(p), None, None, None, None)
    assert rc == 2


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from medt_b200 import _cabi
    monkeypatch.setattr(_cabi, "_LIB", None)
    monkeypatch.setattr(_cabi, "_HERE", str(tmp_path))
    with pytest.raises(RuntimeError, match="no CPU"):
        _cabi.load_library()
