"""Per-op timing of mt_axial_forward / mt_axial_backward (whole C-ABI calls, CUDA events on the launching stream) for the
axial-op geometries of the BASELINE configurations.  One JSON line per run; MT_DEBUG_FLAGS is read once per process, so an
A/B of two code paths is two processes on the same box:

    MT_DEBUG_FLAGS=32 python tools/time_ops.py --tag chain     # round-1 launch chain
    python tools/time_ops.py --tag coop                         # persistent cooperative kernels

--px 128 (default): gated 128 px N=4 ops; --px 512: the L = 256 / 128 / 64 ops of the 512-px configurations."""
import argparse
import ctypes
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch  # noqa: E402

import bench  # noqa: E402

# (name, variant, C, N, H, W, axis, stride, ngroups)
OPS_128 = [
    ("l1.0.h C16 L64 B256", "gated", 16, 4, 64, 64, 0, 1, 1),
    ("l1.0.w C16 L64 B256", "gated", 16, 4, 64, 64, 1, 1, 1),
    ("l2.0.h C32 L64 B256", "gated", 32, 4, 64, 64, 0, 1, 1),
    ("l2.0.w C32 L64 B256 s2", "gated", 32, 4, 64, 64, 1, 2, 1),
    ("l2.1.h C32 L32 B128", "gated", 32, 4, 32, 32, 0, 1, 1),
    ("l3.0.w C64 L32 B128 s2", "gated", 64, 4, 32, 32, 1, 2, 1),
    ("l3.1.h C64 L16 B64", "gated", 64, 4, 16, 16, 0, 1, 1),
    ("l3.1.w C64 L16 B64", "gated", 64, 4, 16, 16, 1, 1, 1),
    ("l4.0.w C128 L16 B64 s2", "gated", 128, 4, 16, 16, 1, 2, 1),
    ("medt.local C16 L16 B1024 g16", "nopos", 16, 64, 16, 16, 0, 1, 16),
    ("medt.local C128 L2 B128 g16", "nopos", 128, 64, 2, 2, 1, 2, 16),
]
OPS_512 = [
    ("512 l1.0.h C16 L256 B1024", "gated", 16, 4, 256, 256, 0, 1, 1),
    ("512 l2.0.w C32 L256 B1024 s2", "gated", 32, 4, 256, 256, 1, 2, 1),
    ("512 l3.0.h C64 L128 B512", "gated", 64, 4, 128, 128, 0, 1, 1),
    ("512 l4.0.w C128 L64 B256 s2", "gated", 128, 4, 64, 64, 1, 2, 1),
]


def time_op(lib_, dev, flush, variant, C, N, H, W, axis, stride, ngroups, reps):
    from medt_b200 import _cabi
    from medt_b200.axial_op import _desc, _ptr
    L = H if axis == 0 else W
    gp = C // 8
    g = torch.Generator(device="cpu").manual_seed(1)
    rn = lambda *s: torch.randn(*s, generator=g).to(dev)                       # noqa: E731
    nsim, nout = (8, C) if variant == "nopos" else (24, 2 * C)
    x, dout = rn(N, C, H, W), rn(N, C, H // stride, W // stride)
    w = rn(2 * C, C) * (1.0 / C) ** 0.5
    rel = rn(2 * gp, 2 * L - 1) * (1.0 / gp) ** 0.5 if variant != "nopos" else None
    gates = [torch.tensor(v, device=dev) for v in (0.1, 0.1, 1.0, 0.1)] if variant == "gated" else [None] * 4
    bn = []
    for n in (2 * C, nsim, nout):
        bn += [torch.rand(n, device=dev) + 0.5, rn(n) * 0.1, torch.zeros(n, device=dev), torch.ones(n, device=dev)]
    desc = _desc(x, axis, stride, variant, True, ngroups, 1e-5, 0.1)
    params = _cabi.MtAxialParams(_ptr(w), _ptr(rel), *[_ptr(t) for t in gates], *[_ptr(t) for t in bn])
    saved = torch.empty(lib_.mt_axial_saved_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
    ws = torch.empty(lib_.mt_axial_bwd_workspace_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
    out = torch.empty_like(dout)
    gt = [torch.empty_like(x), torch.empty_like(w), None if rel is None else torch.empty_like(rel), torch.empty(4, device=dev)] + \
         [torch.empty(n, device=dev) for n in (2 * C, 2 * C, nsim, nsim, nout, nout)]
    grads = _cabi.MtAxialGrads(*[_ptr(t) for t in gt])
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def fwd():
        _cabi.check(lib_.mt_axial_forward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(out), _ptr(saved), st), "fwd")

    def bwd():
        _cabi.check(lib_.mt_axial_backward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(dout), _ptr(saved), _ptr(ws),
                                           ctypes.byref(grads), st), "bwd")
    res = {}
    fwd()
    res["fwd_launches"] = int(lib_.mt_last_launch_count())
    bwd()
    res["bwd_launches"] = int(lib_.mt_last_launch_count())
    torch.cuda.synchronize(dev)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        res[name + "_us_cold"] = round(bench.timed_steps(fn, reps, flush, dev) / reps * 1e3, 2)      # includes host launch cost
        res[name + "_us_warm"] = round(bench.timed_steps(fn, reps, lambda: None, dev) / reps * 1e3, 2)
        # GPU-side time: 10 calls captured in one CUDA graph (no host launch latency in the bracket), replayed
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            stg = ctypes.c_void_p(side.cuda_stream)
            graph = torch.cuda.CUDAGraph()
            old = st.value
            st.value = stg.value
            fn()
            torch.cuda.synchronize(dev)
            with torch.cuda.graph(graph, stream=side):
                for _ in range(10):
                    fn()
            st.value = old
        torch.cuda.synchronize(dev)
        for _ in range(2):
            graph.replay()
        res[name + "_us_graph"] = round(bench.timed_steps(graph.replay, reps, lambda: None, dev) / reps / 10 * 1e3, 2)
        if os.environ.get("MT_DEBUG_FLAGS") and int(os.environ["MT_DEBUG_FLAGS"]) & 64:
            import numpy as np
            torch.cuda.synchronize(dev)
            fn()
            torch.cuda.synchronize(dev)
            buf = (ctypes.c_ulonglong * (16 * 512))()
            (lib_.mt_debug_read_phases if name == "fwd" else lib_.mt_debug_read_phases_bwd)(buf, 16 * 512)
            a = np.frombuffer(buf, dtype=np.uint64).reshape(16, 512).astype(np.int64)
            t0 = a[0][a[0] > 0].max() - 200000           # stamps of this launch: within 200 us of the newest start stamp
            sel = a[0] >= t0
            n = int(sel.sum())
            t0 = a[0][sel].min()
            pre = "" if name == "fwd" else "bwd_"
            res[pre + "phases_us_max_over_ctas"] = [round(float((a[k][sel].max() - t0) / 1e3), 2) for k in range(10)]
            res[pre + "phases_us_min_over_ctas"] = [round(float((a[k][sel].min() - t0) / 1e3), 2) for k in range(10)]
            res[pre + "phase_ctas"] = n
    B = N * (W if axis == 0 else H)
    res["alg_mb_fwd"] = round(2 * B * C * L * 4 / 1e6, 3)
    res["exp_G_per_s_fwd_warm"] = round(8 * B * L * L / (res["fwd_us_warm"] * 1e-6) / 1e9, 1)
    res["checksum"] = float(out.double().sum().item())
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--px", type=int, default=128)
    ap.add_argument("--reps", type=int, default=20)
    ap.add_argument("--tag", default="")
    args = ap.parse_args()
    from medt_b200 import _cabi
    dev = torch.device("cuda", 0)
    lib_ = _cabi.load_library()
    flush = bench.L2Flusher(dev)
    rows = {}
    for name, *geo in (OPS_512 if args.px == 512 else OPS_128):
        try:
            rows[name] = time_op(lib_, dev, flush, *geo, reps=args.reps)
        except Exception as e:  # noqa: BLE001
            rows[name] = {"error": repr(e)[:300]}
    tot = {k: round(sum(r.get(k, 0.0) for r in rows.values()), 1) for k in ("fwd_us_warm", "bwd_us_warm", "fwd_us_cold", "bwd_us_cold")}
    print(json.dumps({"tag": args.tag, "flags": os.environ.get("MT_DEBUG_FLAGS"), "px": args.px, "ops": rows, "sum": tot}))


if __name__ == "__main__":
    main()
