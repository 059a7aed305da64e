"""Shared test helpers: load op fixtures and rebuild the parameter dict the oracle and the
C-ABI wrapper take (values come from tests/golden/fill.py, exactly as make_golden.py
filled the reference module)."""
from __future__ import annotations

import glob
import os

import numpy as np

import fill

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARIANT_OF = {"AxialAttention_dynamic": "gated", "AxialAttention": "pos", "AxialAttention_wopos": "nopos"}


def op_fixture_names():
    return sorted(os.path.basename(p)[3:-4] for p in glob.glob(os.path.join(GOLDEN, "op_*.npz")))


def load_op(name):
    z = np.load(os.path.join(GOLDEN, f"op_{name}.npz"))
    C, N, H, W, width, stride, ngroups = [int(v) for v in z["meta"]]
    variant = VARIANT_OF[str(z["cls"])]
    L = W if width else H
    gp = C // 8
    pre = f"op.{name}."
    t = lambda k, shape: fill.tensor_for(pre + k, shape)  # noqa: E731
    nsim = 8 if variant == "nopos" else 24
    nout = C if variant == "nopos" else 2 * C
    p = {"w_qkv": t("qkv_transform.weight", (2 * C, C, 1)).reshape(2 * C, C)}
    for bn, n in (("bn_qkv", 2 * C), ("bn_similarity", nsim), ("bn_output", nout)):
        short = {"bn_qkv": "bn_qkv", "bn_similarity": "bn_sim", "bn_output": "bn_out"}[bn]
        p[short + "_w"] = t(bn + ".weight", (n,))
        p[short + "_b"] = t(bn + ".bias", (n,))
        p[short + "_rm"] = t(bn + ".running_mean", (n,))
        p[short + "_rv"] = t(bn + ".running_var", (n,))
    if variant != "nopos":
        p["relative"] = t("relative", (2 * gp, 2 * L - 1))
    if variant == "gated":
        p["gates"] = tuple(float(t(g, ())) for g in ("f_qr", "f_kr", "f_sv", "f_sve"))
    cfg = dict(C=C, N=N, H=H, W=W, axis=1 if width else 0, stride=stride, ngroups=ngroups, variant=variant, L=L)
    return z, p, cfg


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b)
    return float(np.linalg.norm(a - b) / den) if den > 0 else float(np.abs(a).max())


def load_model_golden(tag):
    return np.load(os.path.join(GOLDEN, f"model_{tag}.npz"))


def build_filled_model(factory, S, dtype=None, device="cpu"):
    """Drop-in model with the deterministic fill of tests/golden/fill.py (== what make_golden.py
    loaded into the reference)."""
    import torch
    import lib
    f = lib.models.axialunet if factory == "axialunet" else getattr(lib.models.axialnet, factory)
    m = f(img_size=S, imgchan=3)
    vals = fill.fill_state_dict(m.state_dict())
    sd = m.state_dict()
    with torch.no_grad():
        for k, v in vals.items():
            sd[k].copy_(torch.as_tensor(v).to(sd[k].dtype))
    for p in m.parameters():
        p.requires_grad_(True)
    if dtype is not None:
        m = m.to(dtype)          # fill values are fp32-representable, so this cast is exact
    return m.to(device)
