"""CPU tier: the product's CUDA kernels, compiled for the host by a CUDA-on-CPU shim (tests/emu: source rewrite of the
launches / dynamic shared memory / PTX helpers, ucontext fibers for the threads of a CTA, scheduler barriers for
__syncthreads and the warp collectives) and called through the SAME C ABI with numpy buffers, against the oracle.

What this pins without a GPU: every index computation, staging loop, barrier placement, shuffle reduction, last-CTA
finaliser, the saved / workspace layout and the ABI plumbing of all kernels (tiled and one-row-per-thread layouts, all
three variants, grouped statistics, eval mode) plus the experimental MT_DEBUG_FLAGS variants.  What it cannot pin:
performance, PTX-level behaviour (TMA, ex2.approx), races that only real parallel execution shows -- the `-m gpu` tier
stays the parity gate proper."""
import ctypes as C
import json
import os
import shutil
import subprocess
import sys

import numpy as np
import pytest
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "emu"))

pytestmark = pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
TOL = 2e-5


@pytest.fixture(scope="session")
def emu_lib():
    import driver
    return driver.load(driver.build())


def _cases():
    import cases
    return list(range(len(cases.CASES)))


@pytest.mark.parametrize("idx", range(12))
def test_axial_kernels_match_oracle_under_emulation(emu_lib, idx):
    import cases
    errs = cases.run_case(emu_lib, *cases.CASES[idx])
    assert max(errs.values()) < TOL, (cases.CASES[idx], errs)


@pytest.mark.parametrize("flags", [2, 8, 16, 24])
def test_debug_flag_variants_match_oracle_under_emulation(emu_lib, flags):
    """2 = force the one-row-per-thread kernels; 8 / 16 / 24 = the experimental k_bwd2 variants (tiled wrapped-diagonal
    d-relative pass, split row / column CTAs, both).  MT_DEBUG_FLAGS is read once per process -> subprocess."""
    import cases
    env = dict(os.environ, MT_DEBUG_FLAGS=str(flags))
    r = subprocess.run([sys.executable, os.path.join(HERE, "emu", "cases.py")] + [str(i) for i in cases.FLAG_CASES],
                       env=env, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    res = json.loads(r.stdout.strip().splitlines()[-1])
    assert len(res) == len(cases.FLAG_CASES) and max(res.values()) < TOL, res


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def _f32(t):
    import driver
    return driver.f32(t.detach().numpy() if torch.is_tensor(t) else t)


@pytest.mark.parametrize("n,c,h,w,groups,relu,with_res,training", [(2, 8, 8, 8, 1, True, True, True), (6, 5, 3, 7, 3, True, False, True),
                                                                   (4, 16, 4, 4, 4, False, True, True), (2, 8, 8, 8, 1, True, False, False)])
def test_bn_act_kernels_under_emulation(emu_lib, n, c, h, w, groups, relu, with_res, training):
    torch.manual_seed(0)
    x = torch.randn(n, c, h, w, requires_grad=True)
    res = torch.randn(n, c, h, w, requires_grad=True) if with_res else None
    gamma, beta = (torch.rand(c) + 0.5).requires_grad_(True), torch.randn(c).requires_grad_(True)
    rm, rv = torch.randn(c) * 0.1, torch.rand(c) + 0.5
    rm_ref, rv_ref = rm.clone(), rv.clone()
    ys = []
    for g in range(groups):                                   # the reference: one BN call per statistic group
        sl = slice(g * n // groups, (g + 1) * n // groups)
        ys.append(F.batch_norm(x[sl], rm_ref, rv_ref, gamma, beta, training, 0.1, 1e-5))
    y_ref = torch.cat(ys)
    if with_res:
        y_ref = y_ref + res
    if relu:
        y_ref = F.relu(y_ref)
    dy = torch.randn_like(y_ref)
    leaves = [x, gamma, beta] + ([res] if with_res else [])
    g_ref = torch.autograd.grad(y_ref, leaves, dy)
    xs, rs_, ga, be, rm_a, rv_a = _f32(x), (_f32(res) if with_res else None), _f32(gamma), _f32(beta), _f32(rm), _f32(rv)
    y, sm, si = _f32(torch.zeros_like(x)), _f32(torch.zeros(2 * groups, c)), _f32(torch.zeros(groups, c))
    rc = emu_lib.mt_bn_act_forward(n, c, h * w, groups, int(training), int(relu), C.c_float(1e-5), C.c_float(0.1), _ptr(xs),
                                   _ptr(rs_) if with_res else None, _ptr(ga), _ptr(be), _ptr(rm_a), _ptr(rv_a), _ptr(y), _ptr(sm), _ptr(si), None)
    assert rc == 0
    assert np.allclose(y, y_ref.detach().numpy(), atol=1e-5, rtol=1e-5)
    if training:
        assert np.allclose(rm_a, rm_ref.numpy(), atol=1e-6) and np.allclose(rv_a, rv_ref.numpy(), atol=1e-6, rtol=1e-5)
    dx, dres, dga, dbe = _f32(torch.zeros_like(x)), _f32(torch.zeros_like(x)), _f32(torch.zeros(c)), _f32(torch.zeros(c))
    rc = emu_lib.mt_bn_act_backward(n, c, h * w, groups, int(training), int(relu), _ptr(xs), _ptr(y), _ptr(_f32(dy)), _ptr(ga), _ptr(sm), _ptr(si),
                                    _ptr(dx), _ptr(dres) if with_res else None, _ptr(dga), _ptr(dbe), None)
    assert rc == 0
    got = [dx, dga, dbe] + ([dres] if with_res else [])
    for a, b in zip(got, g_ref):
        assert np.allclose(a, b.numpy(), atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("shape,with_skip,with_bias", [((2, 5, 4, 4), True, True), ((3, 7, 1, 3), False, True), ((2, 16, 8, 16), True, False),
                                                       ((1, 6, 5, 9), False, False)])
def test_up2x_kernels_under_emulation(emu_lib, shape, with_skip, with_bias):
    torch.manual_seed(1)
    n, c, h, w = shape
    x = torch.randn(shape, requires_grad=True)
    skip = torch.randn(n, c, 2 * h, 2 * w) if with_skip else None
    bias = torch.randn(c, requires_grad=True) if with_bias else None
    ref = F.relu(F.interpolate(x + bias[None, :, None, None] if with_bias else x, scale_factor=(2, 2), mode="bilinear"))
    if with_skip:
        ref = ref + skip
    d = torch.randn_like(ref)
    g_ref = torch.autograd.grad(ref, [x] + ([bias] if with_bias else []), d)
    out = _f32(torch.zeros_like(ref))
    mask = np.zeros(ref.numel() + 64, dtype=np.uint8)
    rc = emu_lib.mt_up2x_relu_add_forward(n * c, c, h, w, _ptr(_f32(x)), _ptr(_f32(bias)) if with_bias else None,
                                          _ptr(_f32(skip)) if with_skip else None, _ptr(out), _ptr(mask), None)
    assert rc == 0 and np.allclose(out, ref.detach().numpy(), atol=1e-6, rtol=1e-6)
    din, dbias = _f32(torch.zeros_like(x)), _f32(torch.full((c,), 9.0))
    rc = emu_lib.mt_up2x_relu_add_backward(n * c, c, h, w, _ptr(_f32(d)), _ptr(mask), _ptr(din), _ptr(dbias) if with_bias else None, None)
    assert rc == 0 and np.allclose(din, g_ref[0].numpy(), atol=1e-5, rtol=1e-5)
    if with_bias:
        assert np.allclose(dbias, g_ref[1].numpy(), atol=2e-5, rtol=1e-5)


@pytest.mark.parametrize("n,ci,co,h,w,stride,relu_in", [(2, 8, 16, 8, 8, 1, False), (2, 32, 64, 8, 8, 2, False), (3, 5, 7, 6, 10, 1, True),
                                                        (2, 16, 32, 5, 7, 2, True), (1, 70, 66, 4, 4, 1, True)])
def test_conv1x1_kernels_under_emulation(emu_lib, n, ci, co, h, w, stride, relu_in):
    torch.manual_seed(3)
    x = torch.randn(n, ci, h, w, requires_grad=True)
    wt = (torch.randn(co, ci, 1, 1) * 0.2).requires_grad_(True)
    ref = F.conv2d(F.relu(x) if relu_in else x, wt, stride=stride)
    d = torch.randn_like(ref)
    g_ref = torch.autograd.grad(ref, (x, wt), d)
    xs, ws, y = _f32(x), _f32(wt.reshape(co, ci)), _f32(torch.zeros_like(ref))
    assert emu_lib.mt_conv1x1_forward(n, ci, co, h, w, stride, int(relu_in), _ptr(xs), _ptr(ws), _ptr(y), None) == 0
    assert np.allclose(y, ref.detach().numpy(), atol=1e-5, rtol=1e-5)
    dx, dw = _f32(torch.full_like(x, 5.0)), _f32(torch.full((co, ci), 5.0))
    assert emu_lib.mt_conv1x1_backward(n, ci, co, h, w, stride, int(relu_in), _ptr(xs), _ptr(ws), _ptr(_f32(d)), _ptr(dx), _ptr(dw), None) == 0
    assert np.allclose(dx, g_ref[0].numpy(), atol=2e-5, rtol=1e-4)
    assert np.allclose(dw, g_ref[1].reshape(co, ci).numpy(), atol=2e-5, rtol=1e-4)


@pytest.mark.parametrize("n,ci,co,h,w,k,stride,pad", [(2, 8, 128, 8, 8, 3, 1, 1),      # stem conv2 geometry (64 x 64 tiles)
                                                       (2, 32, 8, 8, 8, 3, 1, 1),       # conv3: 8 output channels (TY = 2 tile)
                                                       (3, 16, 16, 5, 7, 3, 1, 1),      # ragged sizes, TY = 4
                                                       (2, 24, 40, 8, 8, 3, 2, 1),      # decoder1: stride 2 (dx stays with cuDNN)
                                                       (1, 3, 8, 9, 9, 7, 2, 3)])       # 7x7 stride 2 stem conv1 (forward + dW)
def test_conv2d_gemm_kernels_under_emulation(emu_lib, n, ci, co, h, w, k, stride, pad):
    torch.manual_seed(4)
    x = torch.randn(n, ci, h, w, requires_grad=True)
    wt = (torch.randn(co, ci, k, k) * 0.2).requires_grad_(True)
    ref = F.conv2d(x, wt, stride=stride, padding=pad)
    d = torch.randn_like(ref)
    g_ref = torch.autograd.grad(ref, (x, wt), d)
    xs, ws, y = _f32(x), _f32(wt), _f32(torch.full_like(ref, 7.0))
    assert emu_lib.mt_conv2d_forward(n, ci, co, h, w, k, k, stride, pad, _ptr(xs), _ptr(ws), _ptr(y), None) == 0
    assert np.allclose(y, ref.detach().numpy(), atol=2e-5, rtol=1e-4)
    dx, dw = _f32(torch.full_like(x, 5.0)), _f32(torch.full_like(wt, 5.0))
    assert emu_lib.mt_conv2d_backward(n, ci, co, h, w, k, k, stride, pad, _ptr(xs), _ptr(ws), _ptr(_f32(d)),
                                      _ptr(dx) if stride == 1 else None, _ptr(dw), None) == 0
    if stride == 1:
        assert np.allclose(dx, g_ref[0].numpy(), atol=5e-5, rtol=1e-4)
    assert np.allclose(dw, g_ref[1].numpy(), atol=5e-5, rtol=1e-4)
    if stride != 1:            # dX of a strided convolution is not implemented: status code, no fault
        assert emu_lib.mt_conv2d_backward(n, ci, co, h, w, k, k, stride, pad, _ptr(xs), _ptr(ws), _ptr(_f32(d)), _ptr(dx), None, None) == 2


@pytest.mark.parametrize("n,c,h,w", [(2, 16, 8, 8), (3, 5, 7, 9)])
def test_head_ce_kernels_under_emulation(emu_lib, n, c, h, w):
    """adjust(relu(x)) + cross-entropy (csrc/head_ce.cu) against torch: logits, loss, dx, dW, db"""
    torch.manual_seed(6)
    x = torch.randn(n, c, h, w, requires_grad=True)
    wt = (torch.randn(2, c, 1, 1) * 0.5).requires_grad_(True)
    b = torch.randn(2, requires_grad=True)
    y = torch.randint(0, 2, (n, h, w))
    logits = F.conv2d(F.relu(x), wt, b)
    loss = F.cross_entropy(logits, y)
    g = torch.autograd.grad(loss * 1.7, (x, wt, b))
    xs, ws, bs = _f32(x), _f32(wt.reshape(2, c)), _f32(b)
    ys = np.ascontiguousarray(y.numpy().astype(np.int64))
    lg, ls = _f32(torch.zeros(n, 2, h, w)), _f32(torch.full((1,), 9.0))
    assert emu_lib.mt_head_ce_forward(n, c, h * w, _ptr(xs), _ptr(ws), _ptr(bs), _ptr(ys), _ptr(lg), _ptr(ls), None) == 0
    assert np.allclose(lg, logits.detach().numpy(), atol=1e-5) and abs(float(ls[0]) - loss.item()) < 1e-5
    dx, dw, db = _f32(torch.full_like(x, 3.0)), _f32(torch.full((2, c), 3.0)), _f32(torch.full((2,), 3.0))
    dl = _f32(torch.full((1,), 1.7))
    assert emu_lib.mt_head_ce_backward(n, c, h * w, _ptr(xs), _ptr(ws), _ptr(lg), _ptr(ys), _ptr(dl), _ptr(dx), _ptr(dw), _ptr(db), None) == 0
    assert np.allclose(dx, g[0].numpy(), atol=1e-6, rtol=1e-4)
    assert np.allclose(dw, g[1].reshape(2, c).numpy(), atol=1e-6, rtol=1e-4)
    assert np.allclose(db, g[2].numpy(), atol=1e-6, rtol=1e-4)


def _golden_names():
    sys.path.insert(0, os.path.join(HERE, "golden"))
    sys.path.insert(0, HERE)
    import helpers
    return helpers.op_fixture_names()


@pytest.mark.parametrize("mode", ["train", "eval"])
@pytest.mark.parametrize("name", _golden_names())
def test_emulated_kernels_match_the_reference_goldens(emu_lib, name, mode):
    """The same fixtures the `-m gpu` tier checks (produced by executing the REFERENCE in fp64, tests/golden/make_golden.py):
    outputs, every gradient and the updated running statistics, here through the emulated kernels."""
    import driver
    import helpers
    z, p, cfg = helpers.load_op(name)
    training = mode == "train"
    out, g, run = driver.run_axial(emu_lib, z["x"], p, cfg["axis"], cfg["stride"], cfg["variant"], training, cfg["ngroups"], z["dout"])
    rel = helpers.rel_err
    errs = {"out": rel(out, z[f"{mode}_out"]), "dx": rel(g["dx"], z[f"{mode}_dx"])}
    gk = lambda k: z[f"{mode}_grad.{k}"]  # noqa: E731
    errs["dw_qkv"] = rel(g["dw_qkv"].reshape(-1), gk("qkv_transform.weight").reshape(-1))
    for short, bn in (("bn_qkv", "bn_qkv"), ("bn_out", "bn_output")):
        errs["d" + short + "_w"] = rel(g["d" + short + "_w"], gk(bn + ".weight"))
        errs["d" + short + "_b"] = rel(g["d" + short + "_b"], gk(bn + ".bias"))
    errs["dbn_sim_w"] = rel(g["dbn_sim_w"], gk("bn_similarity.weight"))
    scale = max(float(np.abs(gk("bn_similarity.weight")).max()), 1e-30)
    errs["dbn_sim_b"] = float(np.abs(g["dbn_sim_b"] - gk("bn_similarity.bias")).max()) / scale      # identically zero in exact arithmetic
    if cfg["variant"] != "nopos":
        errs["drelative"] = rel(g["drelative"], gk("relative"))
    if cfg["variant"] == "gated":
        want = np.array([float(gk(k)) for k in ("f_qr", "f_kr", "f_sv", "f_sve")])
        errs["dgates"] = rel(g["dgates"], want)
    if training:
        for short, bn in (("bn_qkv", "bn_qkv"), ("bn_sim", "bn_similarity"), ("bn_out", "bn_output")):
            errs[short + "_rm"] = rel(run[short + "_rm"], z[f"train_buf.{bn}.running_mean"])
            errs[short + "_rv"] = rel(run[short + "_rv"], z[f"train_buf.{bn}.running_var"])
    assert max(errs.values()) < 5e-5, errs
