"""TEST INFRASTRUCTURE -- axial-op cases run through the CPU-emulated library and compared with oracle/axial_np.py.
`python tests/emu/cases.py` prints one JSON line {case: max_error} (used by the flag tests, which need a fresh process
because MT_DEBUG_FLAGS is read once)."""
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
for p in (ROOT, HERE):
    if p not in sys.path:
        sys.path.insert(0, p)
import driver  # noqa: E402
from oracle import axial_np  # noqa: E402

# variant, C, N, H, W, axis, stride, training, ngroups
CASES = [
    ("gated", 16, 2, 16, 8, 0, 1, True, 1),      # d=1, RT=4 tiled kernels, padded pitch
    ("gated", 16, 2, 64, 4, 0, 1, True, 1),      # L=64: TMA-staged layout (LP == L), baseline layer1.0 geometry
    ("gated", 32, 2, 4, 64, 1, 2, True, 1),      # d=2, width axis, stride 2 (layer2.0)
    ("pos", 16, 4, 16, 16, 1, 1, True, 2),       # two statistic groups
    ("nopos", 16, 2, 16, 16, 0, 2, True, 1),     # MedT local branch op
    ("nopos", 32, 4, 8, 8, 1, 1, True, 4),       # grouped, no replicas
    ("gated", 64, 2, 8, 8, 0, 1, True, 1),       # d=4
    ("gated", 128, 1, 8, 4, 1, 1, True, 1),      # d=8 -> one-row-per-thread kernels
    ("gated", 16, 2, 6, 10, 0, 1, True, 1),      # ragged length -> one-row-per-thread kernels
    ("pos", 32, 2, 12, 6, 1, 2, True, 1),        # ragged, stride 2
    ("gated", 16, 2, 16, 8, 0, 1, False, 1),     # eval: running statistics
    ("nopos", 16, 2, 8, 8, 1, 1, False, 1),
]
FLAG_CASES = [0, 1, 2, 3, 6, 10]                  # subset for the MT_DEBUG_FLAGS variants (tiled / grouped / d=4 / eval)


def make_case(variant, C, N, H, W, axis, seed=0):
    rs = np.random.RandomState(seed + 7 * C + H)
    L, gp = (H if axis == 0 else W), C // 8
    nsim, nout = (8, C) if variant == "nopos" else (24, 2 * C)
    p = {"w_qkv": rs.standard_normal((2 * C, C)) * 0.3}
    for short, n in (("bn_qkv", 2 * C), ("bn_sim", nsim), ("bn_out", nout)):
        p[short + "_w"] = rs.rand(n) + 0.5
        p[short + "_b"] = rs.standard_normal(n) * 0.2
        p[short + "_rm"] = rs.standard_normal(n) * 0.1
        p[short + "_rv"] = rs.rand(n) + 0.5
    if variant != "nopos":
        p["relative"] = rs.standard_normal((2 * gp, 2 * L - 1)) * 0.5
    p = {k: np.asarray(v, dtype=np.float32).astype(np.float64) for k, v in p.items()}      # fp32-representable
    if variant == "gated":
        p["gates"] = tuple(float(np.float32(g)) for g in (0.3, -0.2, 0.9, 0.15))
    x = rs.standard_normal((N, C, H, W)).astype(np.float32).astype(np.float64)
    return rs, p, x


def run_case(lib, variant, C, N, H, W, axis, stride, training, ngroups):
    """max over all outputs of the relative L2 error vs the fp64 oracle (d bn_similarity.bias, identically zero in exact
    arithmetic, is measured against the scale of d bn_similarity.weight)."""
    rs, p, x = make_case(variant, C, N, H, W, axis)
    o_ref, cache, run_ref = axial_np.axial_forward(x, p, axis, stride, variant, training=training, ngroups=ngroups)
    dout = rs.standard_normal(o_ref.shape).astype(np.float32).astype(np.float64)
    g_ref = axial_np.axial_backward(dout, cache)
    out, g, run = driver.run_axial(lib, x, p, axis, stride, variant, training, ngroups, dout)
    rel = lambda a, b: float(np.linalg.norm(np.ravel(a) - np.ravel(b)) / max(np.linalg.norm(np.ravel(b)), 1e-30))  # noqa: E731
    errs = {"out": rel(out, o_ref)}
    for k, v in g.items():
        if k == "dbn_sim_b":
            errs[k] = float(np.abs(v).max() / max(np.abs(g_ref["dbn_sim_w"]).max(), 1e-30))
        elif k == "dgates" and variant != "gated":
            continue                                      # not written for the gate-less variants
        else:
            errs[k] = rel(v, g_ref[k])
    if training:
        for k, v in run_ref.items():
            errs[k] = rel(run[k], v)
    return errs


if __name__ == "__main__":
    lib = driver.load(driver.build())
    which = [int(a) for a in sys.argv[1:]] or list(range(len(CASES)))
    res = {}
    for i in which:
        e = run_case(lib, *CASES[i])
        res[str(i)] = max(e.values())
    print(json.dumps(res))
