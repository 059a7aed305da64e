"""TEST INFRASTRUCTURE -- builds the CPU-emulated copy of the product library (tests/emu/prep.py + the shim) and calls its
C ABI with numpy buffers.  Used only by tests/test_emu_kernels.py; the product never loads this library."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
import prep  # noqa: E402

PARAM_FIELDS = ["w_qkv", "relative", "f_qr", "f_kr", "f_sv", "f_sve", "bn_qkv_w", "bn_qkv_b", "bn_qkv_rm", "bn_qkv_rv",
                "bn_sim_w", "bn_sim_b", "bn_sim_rm", "bn_sim_rv", "bn_out_w", "bn_out_b", "bn_out_rm", "bn_out_rv"]
GRAD_FIELDS = ["dx", "dw_qkv", "drelative", "dgates", "dbn_qkv_w", "dbn_qkv_b", "dbn_sim_w", "dbn_sim_b", "dbn_out_w", "dbn_out_b"]
VARIANT = {"gated": 0, "pos": 1, "nopos": 2}


class Desc(C.Structure):
    _fields_ = [(k, C.c_int32) for k in ("N", "C", "H", "W", "axis", "stride", "variant", "training", "ngroups")] + \
               [("eps", C.c_float), ("momentum", C.c_float)]


class Params(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in PARAM_FIELDS]


class Grads(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in GRAD_FIELDS]


def build(out_dir=None, cxx="g++"):
    if os.environ.get("EMU_LIB"):                            # a prebuilt variant (e.g. -fsanitize=address)
        return os.environ["EMU_LIB"]
    """Rewrite + compile (only what changed) -> path of libmedt_emu.so."""
    out_dir = out_dir or os.path.join(HERE, "_build")
    srcs = prep.main(out_dir)
    inc = os.path.join(HERE, "include")
    flags = ["-std=c++20", "-O1", "-g", "-fPIC", "-Wno-unknown-pragmas", "-I", inc, "-I", out_dir]
    headers = [s for s in srcs if s.endswith(".cuh")] + [os.path.join(inc, "cuda_runtime.h"), os.path.join(ROOT, "include", "medt_axial.h")]
    hm = max(os.path.getmtime(h) for h in headers)
    objs, jobs = [], []
    for src in [s for s in srcs if s.endswith(".cpp")] + [os.path.join(HERE, "emu.cpp")]:
        obj = os.path.join(out_dir, os.path.basename(src)[:-4] + ".o")
        objs.append(obj)
        if not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hm):
            jobs.append((src, subprocess.Popen([cxx] + flags + ["-c", src, "-o", obj], stderr=subprocess.PIPE, text=True)))
    for src, j in jobs:
        err = j.communicate()[1]
        if j.returncode:
            raise RuntimeError(f"emu build failed for {src}:\n{err[-3000:]}")
    lib = os.path.join(out_dir, "libmedt_emu.so")
    if jobs or not os.path.exists(lib):
        subprocess.run([cxx, "-shared", "-o", lib] + objs, check=True)
    return lib


def load(path):
    lib = C.CDLL(path)
    lib.mt_last_error.restype = C.c_char_p
    lib.mt_axial_saved_bytes.restype = C.c_size_t
    lib.mt_axial_bwd_workspace_bytes.restype = C.c_size_t
    return lib


def _buf(a):
    return a.ctypes.data_as(C.c_void_p)


def _aligned(n_bytes):
    if os.environ.get("EMU_TIGHT"):                          # exact-size allocations: for runs under AddressSanitizer
        raw = np.zeros(n_bytes, dtype=np.uint8)
        assert raw.ctypes.data % 16 == 0
        return raw
    raw = np.zeros(n_bytes + 64, dtype=np.uint8)
    off = (-raw.ctypes.data) % 64
    return raw[off:off + n_bytes]


def f32(a):
    out = _aligned(4 * int(np.prod(np.shape(a))) if np.ndim(a) else 4).view(np.float32)
    out[...] = np.asarray(a, dtype=np.float32).reshape(-1)
    return out.reshape(np.shape(a)) if np.ndim(a) else out


def run_axial(lib, x, p, axis, stride, variant, training, ngroups=1, dout=None, eps=1e-5, momentum=0.1):
    """One forward (+ backward when dout is given) through the emulated C ABI.  `p` as oracle/axial_np.py takes it.
    Returns out, grads (dict or None), running (dict of the updated running statistics)."""
    n, c, h, w = x.shape
    L = h if axis == 0 else w
    gp = c // 8
    d = Desc(n, c, h, w, axis, stride, VARIANT[variant], int(training), ngroups, eps, momentum)
    keep = {}
    P = Params()
    keep["w_qkv"] = f32(np.asarray(p["w_qkv"]).reshape(2 * c, c))
    if variant != "nopos":
        keep["relative"] = f32(p["relative"])
    if variant == "gated":
        for k, v in zip(("f_qr", "f_kr", "f_sv", "f_sve"), p["gates"]):
            keep[k] = f32(np.float32(v))
    for k in PARAM_FIELDS[6:]:
        keep[k] = f32(p[k])
    for k in PARAM_FIELDS:
        setattr(P, k, keep[k].ctypes.data if k in keep else None)
    xs = f32(x)
    out = f32(np.zeros((n, c, h // stride, w // stride)))
    saved = _aligned(int(lib.mt_axial_saved_bytes(C.byref(d))))
    saved[:] = 0xA5                                         # stale garbage: the library must initialise what it reads
    rc = lib.mt_axial_forward(C.byref(d), C.byref(P), _buf(xs), _buf(out), _buf(saved), None)
    if rc:
        raise RuntimeError(f"mt_axial_forward rc={rc}: {lib.mt_last_error().decode()}")
    running = {k: np.array(keep[k]) for k in PARAM_FIELDS if k.endswith(("_rm", "_rv"))}
    grads = None
    if dout is not None:
        ws = _aligned(int(lib.mt_axial_bwd_workspace_bytes(C.byref(d))))
        ws[:] = 0x5A
        nsim, nout = (8, c) if variant == "nopos" else (24, 2 * c)
        g = {"dx": f32(np.zeros_like(x)), "dw_qkv": f32(np.zeros((2 * c, c))), "dgates": f32(np.zeros(4)),
             "dbn_qkv_w": f32(np.zeros(2 * c)), "dbn_qkv_b": f32(np.zeros(2 * c)), "dbn_sim_w": f32(np.zeros(nsim)),
             "dbn_sim_b": f32(np.zeros(nsim)), "dbn_out_w": f32(np.zeros(nout)), "dbn_out_b": f32(np.zeros(nout))}
        if variant != "nopos":
            g["drelative"] = f32(np.zeros((2 * gp, 2 * L - 1)))
        for v in g.values():
            v[...] = np.float32(777.0)                      # "each buffer fully overwritten"
        G = Grads()
        for k in GRAD_FIELDS:
            setattr(G, k, g[k].ctypes.data if k in g else None)
        ds = f32(dout)
        rc = lib.mt_axial_backward(C.byref(d), C.byref(P), _buf(xs), _buf(ds), _buf(saved), _buf(ws), C.byref(G), None)
        if rc:
            raise RuntimeError(f"mt_axial_backward rc={rc}: {lib.mt_last_error().decode()}")
        grads = {k: np.array(v) for k, v in g.items()}
    return np.array(out), grads, running
