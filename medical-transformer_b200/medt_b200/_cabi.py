"""ctypes binding of include/medt_axial.h (the reference-side stub INTEGRATION.md shows).

Struct layouts mirror the header field for field; test_abi.py checks that every symbol the
header declares is exported and that sizeof() agrees."""
from __future__ import annotations

import ctypes as C

import torch
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
_LOCK = threading.Lock()

MT_VARIANT = {"gated": 0, "pos": 1, "nopos": 2}


class MtAxialDesc(C.Structure):
    _fields_ = [("N", C.c_int32), ("C", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("axis", C.c_int32), ("stride", C.c_int32), ("variant", C.c_int32),
                ("training", C.c_int32), ("ngroups", C.c_int32),
                ("eps", C.c_float), ("momentum", C.c_float)]


_PARAM_FIELDS = ["w_qkv", "relative", "f_qr", "f_kr", "f_sv", "f_sve",
                 "bn_qkv_w", "bn_qkv_b", "bn_qkv_rm", "bn_qkv_rv",
                 "bn_sim_w", "bn_sim_b", "bn_sim_rm", "bn_sim_rv",
                 "bn_out_w", "bn_out_b", "bn_out_rm", "bn_out_rv"]
_GRAD_FIELDS = ["dx", "dw_qkv", "drelative", "dgates", "dbn_qkv_w", "dbn_qkv_b",
                "dbn_sim_w", "dbn_sim_b", "dbn_out_w", "dbn_out_b"]


class MtAxialParams(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _PARAM_FIELDS]


class MtAxialGrads(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in _GRAD_FIELDS]


ABI_VERSION = 2        # == MT_ABI_VERSION of include/medt_axial.h

EXPORTS = ["mt_version", "mt_last_error", "mt_last_launch_count", "mt_axial_saved_bytes",
           "mt_axial_bwd_workspace_bytes", "mt_axial_forward", "mt_axial_backward",
           "mt_axial_forward_stages", "mt_axial_backward_stages",
           "mt_up2x_relu_add_forward", "mt_up2x_relu_add_backward", "mt_bn_act_forward", "mt_bn_act_backward",
           "mt_conv1x1_forward", "mt_conv1x1_backward", "mt_conv2d_forward", "mt_conv2d_backward", "mt_set_conv_tc", "mt_head_ce_forward", "mt_head_ce_backward"]


def lib_path() -> str:
    # MEDT_LIB_PATH: developer override used for same-box A/B runs of two builds
    return os.environ.get("MEDT_LIB_PATH") or os.path.join(_HERE, "libmedt_b200.so")


def load_library():
    """Load libmedt_b200.so (built in-tree by __graft_entry__.build() / csrc/Makefile).
    Fails loudly: the product has no fallback path."""
    global _LIB
    with _LOCK:
        if _LIB is not None:
            return _LIB
        path = lib_path()
        if not os.path.exists(path):
            raise RuntimeError(
                f"medt_b200: {path} is missing -- build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                "(or `make -C medical-transformer_b200/csrc`). There is no CPU / eager fallback.")
        lib = C.CDLL(path)
        lib.mt_version.restype = C.c_int
        if int(lib.mt_version()) != ABI_VERSION:
            raise RuntimeError(f"medt_b200: {path} has ABI version {int(lib.mt_version())}, this host code binds "
                               f"version {ABI_VERSION} (include/medt_axial.h) -- rebuild the library")
        lib.mt_last_error.restype = C.c_char_p
        lib.mt_last_launch_count.restype = C.c_int
        lib.mt_axial_saved_bytes.restype = C.c_size_t
        lib.mt_axial_saved_bytes.argtypes = [C.POINTER(MtAxialDesc)]
        lib.mt_axial_bwd_workspace_bytes.restype = C.c_size_t
        lib.mt_axial_bwd_workspace_bytes.argtypes = [C.POINTER(MtAxialDesc)]
        lib.mt_axial_forward.restype = C.c_int
        lib.mt_axial_forward.argtypes = [C.POINTER(MtAxialDesc), C.POINTER(MtAxialParams), C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_void_p]
        lib.mt_axial_backward.restype = C.c_int
        lib.mt_axial_backward.argtypes = [C.POINTER(MtAxialDesc), C.POINTER(MtAxialParams), C.c_void_p, C.c_void_p,
                                          C.c_void_p, C.c_void_p, C.POINTER(MtAxialGrads), C.c_void_p]
        lib.mt_axial_forward_stages.restype = C.c_int
        lib.mt_axial_forward_stages.argtypes = lib.mt_axial_forward.argtypes + [C.c_int]
        lib.mt_axial_backward_stages.restype = C.c_int
        lib.mt_axial_backward_stages.argtypes = lib.mt_axial_backward.argtypes + [C.c_int]
        VP, I, F_ = C.c_void_p, C.c_int, C.c_float
        lib.mt_up2x_relu_add_forward.restype = C.c_int
        lib.mt_up2x_relu_add_forward.argtypes = [I, I, I, I, VP, VP, VP, VP, VP, VP]
        lib.mt_up2x_relu_add_backward.restype = C.c_int
        lib.mt_up2x_relu_add_backward.argtypes = [I, I, I, I, VP, VP, VP, VP, VP]
        lib.mt_bn_act_forward.restype = C.c_int
        lib.mt_bn_act_forward.argtypes = [I, I, I, I, I, I, F_, F_, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP]
        lib.mt_bn_act_backward.restype = C.c_int
        lib.mt_bn_act_backward.argtypes = [I, I, I, I, I, I, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP, VP]
        lib.mt_conv1x1_forward.restype = C.c_int
        lib.mt_conv1x1_forward.argtypes = [I, I, I, I, I, I, I, VP, VP, VP, VP]
        lib.mt_conv1x1_backward.restype = C.c_int
        lib.mt_conv1x1_backward.argtypes = [I, I, I, I, I, I, I, VP, VP, VP, VP, VP, VP]
        lib.mt_conv2d_forward.restype = C.c_int
        lib.mt_conv2d_forward.argtypes = [I, I, I, I, I, I, I, I, I, VP, VP, VP, VP]
        lib.mt_conv2d_backward.restype = C.c_int
        lib.mt_conv2d_backward.argtypes = [I, I, I, I, I, I, I, I, I, VP, VP, VP, VP, VP, VP]
        lib.mt_set_conv_tc.restype = C.c_int
        lib.mt_set_conv_tc.argtypes = [I]
        lib.mt_head_ce_forward.restype = C.c_int
        lib.mt_head_ce_forward.argtypes = [I, I, I, VP, VP, VP, VP, VP, VP, VP]
        lib.mt_head_ce_backward.restype = C.c_int
        lib.mt_head_ce_backward.argtypes = [I, I, I, VP, VP, VP, VP, VP, VP, VP, VP, VP]
        _LIB = lib
        report = os.environ.get("MEDT_LAUNCH_REPORT")
        if report:          # evidence for callers that cannot be instrumented (an UNMODIFIED train.py run through run_reference_script.py)
            import atexit
            import json

            def _write_report(path=report, lib_file=path):
                try:
                    with open(path, "w") as f:
                        json.dump({"library": lib_file, "abi": ABI_VERSION, "kernels_launched": LAUNCHES["kernels"],
                                   "calls": LAUNCHES.get("calls", 0)}, f)
                except OSError:
                    pass
            atexit.register(_write_report)
        return lib


def abi_version() -> int:
    return int(load_library().mt_version())


def check(rc: int, what: str):
    if rc != 0:
        msg = load_library().mt_last_error().decode(errors="replace")
        kind = {1: ValueError, 2: NotImplementedError}.get(rc, RuntimeError)
        raise kind(f"{what} failed (status {rc}): {msg}")


# launch accounting for bench.py's gpu_launches
LAUNCHES = {"kernels": 0}


def count_launches():
    LAUNCHES["kernels"] += int(load_library().mt_last_launch_count())
    LAUNCHES["calls"] = LAUNCHES.get("calls", 0) + 1


def dense(t):
    """Contiguous AND 16-byte aligned (the kernels use 128-bit loads and TMA bulk copies on activation buffers;
    a contiguous view at an odd storage offset is copied)."""
    if t is None:
        return None
    t = t.contiguous()
    if t.data_ptr() & 15:
        t = t.clone(memory_format=torch.contiguous_format)
    return t
