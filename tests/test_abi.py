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
    assert lib.mt_version() == _cabi.ABI_VERSION == 2
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


def test_misaligned_buffers_are_rejected_before_any_launch():
    """Activation buffers are read with 128-bit loads / TMA bulk copies: the ABI refuses pointers that are not
    16-byte aligned (argument check only -- nothing is dereferenced or launched)."""
    from medt_b200 import _cabi
    lib = _cabi.load_library()
    d = _cabi.MtAxialDesc(1, 16, 8, 8, 0, 1, 2, 0, 1, 1e-5, 0.1)           # nopos, eval
    p = _cabi.MtAxialParams()
    fake = ctypes.c_void_p(0x1000)
    for name, _ in _cabi.MtAxialParams._fields_:
        setattr(p, name, fake)
    ok, off = ctypes.c_void_p(0x2000), ctypes.c_void_p(0x2004)
    rc = lib.mt_axial_forward(ctypes.byref(d), ctypes.byref(p), off, ok, ok, None)
    assert rc == 1 and b"16-byte aligned" in lib.mt_last_error()
    rc = lib.mt_axial_forward(ctypes.byref(d), ctypes.byref(p), ok, ok, off, None)
    assert rc == 1 and b"16-byte aligned" in lib.mt_last_error()
    p.w_qkv = ctypes.c_void_p(0x1008)
    rc = lib.mt_axial_forward(ctypes.byref(d), ctypes.byref(p), ok, ok, ok, None)
    assert rc == 1 and b"w_qkv" in lib.mt_last_error()
    rc = lib.mt_up2x_relu_add_forward(1, 1, 4, 4, ok, None, off, ok, ok, None)
    assert rc == 1 and b"aligned" in lib.mt_last_error()


def test_dense_returns_contiguous_aligned_tensors():
    import torch
    from medt_b200 import _cabi
    base = torch.arange(4 * 9, dtype=torch.float32)
    view = base[1:33].view(2, 16)                                          # contiguous, storage offset 4 bytes
    assert view.is_contiguous() and view.data_ptr() % 16 != 0
    t = _cabi.dense(view)
    assert t.data_ptr() % 16 == 0 and t.is_contiguous() and torch.equal(t, view)
    assert _cabi.dense(base) is base and _cabi.dense(None) is None
    assert _cabi.dense(base.view(6, 6).t()).is_contiguous()


def test_ctypes_bindings_match_the_header_prototypes():
    """Every prototype of include/medt_axial.h against the ctypes argtypes of the host binding: same arity, pointers
    bound as pointers, int as c_int, float as c_float -- an ABI change that forgets one side fails here, not on a GPU."""
    import re
    from medt_b200 import _cabi
    lib = _cabi.load_library()
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    h = re.sub(r"/\*.*?\*/", "", open(os.path.join(here, "include", "medt_axial.h")).read(), flags=re.S)
    protos = re.findall(r"\b(?:int|size_t|const char\s*\*)\s*(mt_\w+)\s*\(([^)]*)\)\s*;", h)
    assert sorted(n for n, _ in protos) == sorted(_cabi.EXPORTS)
    for name, args in protos:
        params = [] if args.strip() in ("", "void") else [a.strip() for a in args.split(",")]
        at = getattr(lib, name).argtypes
        if not params:
            assert not at, name
            continue
        assert at is not None and len(at) == len(params), (name, len(params), at)
        for prm, ct in zip(params, at):
            if "*" in prm:
                assert ct is ctypes.c_void_p or hasattr(ct, "contents") or issubclass(ct, ctypes._Pointer), (name, prm, ct)
            elif prm.startswith("float"):
                assert ct is ctypes.c_float, (name, prm, ct)
            elif prm.startswith("int"):
                assert ct is ctypes.c_int, (name, prm, ct)
            else:
                raise AssertionError(f"unhandled parameter type in {name}: {prm}")


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    from medt_b200 import _cabi
    monkeypatch.setattr(_cabi, "_LIB", None)
    monkeypatch.setattr(_cabi, "_HERE", str(tmp_path))
    with pytest.raises(RuntimeError, match="no CPU"):
        _cabi.load_library()
