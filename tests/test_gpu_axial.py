"""GPU tier: the CUDA path (through the C ABI) against (a) the golden vectors produced by
the reference itself and (b) the pinned oracle at BASELINE sizes.  fp32 kernels vs fp64
truth: tolerances are stated per check."""
import numpy as np
import pytest
import torch

from helpers import load_op, op_fixture_names, rel_err

pytestmark = pytest.mark.gpu

FWD_TOL = 2e-5       # forward output, relative L2 (fp32 arithmetic + ex2.approx)
GRAD_TOL = 5e-4      # gradients, relative L2 (fp32 atomics, recomputed softmax)
STAT_TOL = 1e-5      # running statistics


def build_module(cfg, p, name):
    import lib
    ax = lib.models.axialnet
    cls = {"gated": ax.AxialAttention_dynamic, "pos": ax.AxialAttention, "nopos": ax.AxialAttention_wopos}[cfg["variant"]]
    mod = cls(cfg["C"], cfg["C"], groups=8, kernel_size=cfg["L"], stride=cfg["stride"], width=cfg["axis"] == 1)
    t = lambda a: torch.as_tensor(np.asarray(a), dtype=torch.float32)  # noqa: E731
    with torch.no_grad():
        mod.qkv_transform.weight.copy_(t(p["w_qkv"]).unsqueeze(-1))
        for bn, short in ((mod.bn_qkv, "bn_qkv"), (mod.bn_similarity, "bn_sim"), (mod.bn_output, "bn_out")):
            bn.weight.copy_(t(p[short + "_w"])); bn.bias.copy_(t(p[short + "_b"]))
            bn.running_mean.copy_(t(p[short + "_rm"])); bn.running_var.copy_(t(p[short + "_rv"]))
            bn.num_batches_tracked.fill_(3)
        if cfg["variant"] != "nopos":
            mod.relative.copy_(t(p["relative"]))
        if cfg["variant"] == "gated":
            for g, v in zip(("f_qr", "f_kr", "f_sv", "f_sve"), p["gates"]):
                getattr(mod, g).fill_(v)
    for q in mod.parameters():
        q.requires_grad_(True)
    mod._stat_groups = cfg["ngroups"]
    return mod.cuda()


def run_case(name, mode):
    z, p, cfg = load_op(name)
    mod = build_module(cfg, p, name)
    mod.train(mode == "train")
    x = torch.tensor(z["x"], dtype=torch.float32, device="cuda", requires_grad=True)
    out = mod(x)
    dout = torch.tensor(z["dout"], dtype=torch.float32, device="cuda")
    out.backward(dout)
    torch.cuda.synchronize()
    errs = {"out": (rel_err(out.detach().cpu().numpy(), z[f"{mode}_out"]), FWD_TOL),
            "dx": (rel_err(x.grad.cpu().numpy(), z[f"{mode}_dx"]), GRAD_TOL)}
    for k, q in mod.named_parameters():
        ref = z[f"{mode}_grad.{k}"]
        got = q.grad.cpu().numpy()
        if k.endswith("bn_similarity.bias"):
            errs["grad." + k] = (float(np.abs(got - ref).max()), 1e-6)      # identically zero
        elif ref.ndim == 0:
            errs["grad." + k] = (abs(float(got) - float(ref)) / max(1e-3, abs(float(ref))), 5 * GRAD_TOL)
        else:
            errs["grad." + k] = (rel_err(got, ref), GRAD_TOL)
    if mode == "train":
        for k, b in mod.named_buffers():
            if "running" in k:
                errs["buf." + k] = (rel_err(b.cpu().numpy(), z[f"train_buf.{k}"]), STAT_TOL)
            elif "num_batches" in k:
                errs["buf." + k] = (abs(int(b) - int(z[f"train_buf.{k}"])), 0.5)
    return errs


@pytest.mark.parametrize("name", op_fixture_names())
@pytest.mark.parametrize("mode", ["train", "eval"])
def test_op_matches_reference_golden(name, mode):
    errs = run_case(name, mode)
    bad = {k: v for k, v in errs.items() if not (v[0] <= v[1])}
    assert not bad, f"{name}/{mode}: " + ", ".join(f"{k}={v[0]:.3e}>{v[1]:.0e}" for k, v in bad.items())


@pytest.mark.parametrize("C,S,stride,axis,variant", [
    (16, 64, 1, 0, "gated"),      # gated layer1.0 hight_block at 128 px, N=4  (BASELINE config 1/2)
    (32, 64, 2, 1, "gated"),      # layer2.0 width_block: stride-2 epilogue
    (64, 32, 1, 0, "pos"),
    (128, 16, 2, 1, "gated"),     # layer4.0
    (16, 16, 1, 1, "nopos"),      # MedT local branch, first block
])
def test_op_matches_oracle_at_baseline_sizes(C, S, stride, axis, variant):
    """Full BASELINE sizes (N=4): CUDA vs the pinned numpy oracle run in fp64 on the same inputs."""
    from oracle import axial_np
    import fill
    rs = np.random.RandomState(C * 1000 + S)
    N, L, gp = 4, S, C // 8
    name = f"big.{variant}.{C}.{S}"
    nsim, nout = (8, C) if variant == "nopos" else (24, 2 * C)
    p = {"w_qkv": fill.tensor_for(name + ".qkv_transform.weight", (2 * C, C))}
    for bn, short, n in (("bn_qkv", "bn_qkv", 2 * C), ("bn_similarity", "bn_sim", nsim), ("bn_output", "bn_out", nout)):
        p[short + "_w"] = fill.tensor_for(f"{name}.{bn}.weight", (n,)); p[short + "_b"] = fill.tensor_for(f"{name}.{bn}.bias", (n,))
        p[short + "_rm"] = fill.tensor_for(f"{name}.{bn}.running_mean", (n,)); p[short + "_rv"] = fill.tensor_for(f"{name}.{bn}.running_var", (n,))
    if variant != "nopos":
        p["relative"] = fill.tensor_for(name + ".relative", (2 * gp, 2 * L - 1))
    if variant == "gated":
        p["gates"] = (0.1, 0.1, 1.0, 0.1)          # the reference's initial gate values (:124-127)
    x = rs.standard_normal((N, C, S, S))
    cfg = dict(C=C, N=N, H=S, W=S, axis=axis, stride=stride, ngroups=1, variant=variant, L=L)
    out_ref, cache, newrun = axial_np.axial_forward(x, p, axis, stride, variant, training=True)
    dout = rs.standard_normal(out_ref.shape)
    gref = axial_np.axial_backward(dout, cache)
    mod = build_module(cfg, p, name).train()
    xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
    out = mod(xt)
    out.backward(torch.tensor(dout, dtype=torch.float32, device="cuda"))
    torch.cuda.synchronize()
    assert rel_err(out.detach().cpu().numpy(), out_ref) < 1e-4
    assert rel_err(xt.grad.cpu().numpy(), gref["dx"]) < 2e-3
    assert rel_err(mod.qkv_transform.weight.grad.cpu().numpy()[:, :, 0], gref["dw_qkv"]) < 2e-3
    if variant != "nopos":
        assert rel_err(mod.relative.grad.cpu().numpy(), gref["drelative"]) < 2e-3
    assert rel_err(mod.bn_similarity.weight.grad.cpu().numpy(), gref["dbn_sim_w"]) < 2e-3
    assert rel_err(mod.bn_output.running_var.cpu().numpy(), newrun["bn_out_rv"]) < 1e-4
    assert rel_err(mod.bn_similarity.running_var.cpu().numpy(), newrun["bn_sim_rv"]) < 1e-4


def test_linearity_of_backward_in_grad_output():
    """Size-independent property: backward is linear in dout (checked at layer1 size)."""
    import lib
    torch.manual_seed(0)
    mod = lib.models.axialnet.AxialAttention_dynamic(16, 16, groups=8, kernel_size=64).cuda().train()
    x = torch.randn(4, 16, 64, 64, device="cuda", requires_grad=True)
    d1, d2 = torch.randn(4, 16, 64, 64, device="cuda"), torch.randn(4, 16, 64, 64, device="cuda")
    g = []
    for d in (d1, d2, 2.0 * d1 - 3.0 * d2):
        y = mod(x)
        gx, gw = torch.autograd.grad(y, (x, mod.relative), d)
        g.append((gx, gw))
    for k in range(2):
        comb = 2.0 * g[0][k] - 3.0 * g[1][k]
        assert (comb - g[2][k]).norm() / g[2][k].norm() < 1e-3


def test_softmax_shift_invariance_of_bn_similarity_bias():
    """bn_similarity.bias cannot change the output."""
    import lib
    torch.manual_seed(0)
    mod = lib.models.axialnet.AxialAttention(32, 32, groups=8, kernel_size=32).cuda().eval()
    x = torch.randn(2, 32, 32, 32, device="cuda")
    with torch.no_grad():
        y0 = mod(x)
        mod.bn_similarity.bias.add_(3.0)
        y1 = mod(x)
    assert torch.equal(y0, y1)


def test_wrong_length_raises_like_reference():
    import lib
    mod = lib.models.axialnet.AxialAttention_dynamic(16, 16, groups=8, kernel_size=16).cuda()
    with pytest.raises(RuntimeError):
        mod(torch.randn(1, 16, 8, 8, device="cuda"))


@pytest.mark.parametrize("variant,C,H,W,axis,stride", [("gated", 16, 6, 10, 0, 1), ("pos", 32, 12, 6, 1, 2),
                                                      ("nopos", 16, 10, 14, 1, 1), ("gated", 16, 56, 4, 0, 1)])
def test_ragged_lengths_match_oracle(variant, C, H, W, axis, stride):
    """Lengths that are not powers of two / multiples of 4 (the reference's default kernel_size is 56) take the
    generic one-row-per-thread kernels and the scalar staging paths."""
    from oracle import axial_np
    import fill
    rs = np.random.RandomState(H * 100 + W)
    N, L, gp = 2, (H if axis == 0 else W), C // 8
    name = f"ragged.{variant}.{C}.{H}.{W}"
    nsim, nout = (8, C) if variant == "nopos" else (24, 2 * C)
    p = {"w_qkv": fill.tensor_for(name + ".qkv_transform.weight", (2 * C, C))}
    for bn, short, n in (("bn_qkv", "bn_qkv", 2 * C), ("bn_similarity", "bn_sim", nsim), ("bn_output", "bn_out", nout)):
        for leaf, suf in (("weight", "_w"), ("bias", "_b"), ("running_mean", "_rm"), ("running_var", "_rv")):
            p[short + suf] = fill.tensor_for(f"{name}.{bn}.{leaf}", (n,))
    if variant != "nopos":
        p["relative"] = fill.tensor_for(name + ".relative", (2 * gp, 2 * L - 1))
    if variant == "gated":
        p["gates"] = (0.3, -0.2, 0.9, 0.15)
    x = rs.standard_normal((N, C, H, W))
    cfg = dict(C=C, N=N, H=H, W=W, axis=axis, stride=stride, ngroups=1, variant=variant, L=L)
    for training in (True, False):
        out_ref, cache, _ = axial_np.axial_forward(x, p, axis, stride, variant, training=training)
        dout = rs.standard_normal(out_ref.shape)
        gref = axial_np.axial_backward(dout, cache)
        mod = build_module(cfg, p, name).train(training)
        xt = torch.tensor(x, dtype=torch.float32, device="cuda", requires_grad=True)
        out = mod(xt)
        out.backward(torch.tensor(dout, dtype=torch.float32, device="cuda"))
        assert rel_err(out.detach().cpu().numpy(), out_ref) < 1e-4
        assert rel_err(xt.grad.cpu().numpy(), gref["dx"]) < 2e-3
        assert rel_err(mod.qkv_transform.weight.grad.cpu().numpy()[:, :, 0], gref["dw_qkv"]) < 2e-3
        if variant != "nopos":
            assert rel_err(mod.relative.grad.cpu().numpy(), gref["drelative"]) < 2e-3
