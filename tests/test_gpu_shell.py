"""GPU tier: the fused shell kernels against plain PyTorch fp32 ops (the reference's own
F.interpolate / BatchNorm2d / relu / add sequence, axialnet.py:285-300, :493-501, :661-700)."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(2, 5, 4, 4), (4, 16, 64, 64), (3, 7, 1, 3), (2, 32, 8, 16), (3, 6, 5, 9)])
@pytest.mark.parametrize("with_skip", [True, False])
@pytest.mark.parametrize("with_bias", [True, False])
def test_up2x_relu_add_matches_torch(shape, with_skip, with_bias):
    from medt_b200.shell_ops import up2x_relu_add
    torch.manual_seed(1)
    n, c, h, w = shape
    x = torch.randn(shape, device="cuda", requires_grad=True)
    skip = torch.randn(n, c, 2 * h, 2 * w, device="cuda", requires_grad=True) if with_skip else None
    bias = torch.randn(c, device="cuda", requires_grad=True) if with_bias else None
    ref = F.relu(F.interpolate(x + bias[None, :, None, None] if with_bias else x, scale_factor=(2, 2), mode="bilinear"))
    if with_skip:
        ref = ref + skip
    got = up2x_relu_add(x, skip, bias)
    assert torch.allclose(got, ref, atol=1e-6, rtol=1e-6)
    d = torch.randn_like(ref)
    leaves = [x] + ([skip] if with_skip else []) + ([bias] if with_bias else [])
    gref = torch.autograd.grad(ref, leaves, d)
    ggot = torch.autograd.grad(got, leaves, d)
    for a, b in zip(ggot, gref):
        assert a.shape == b.shape and torch.allclose(a, b, atol=2e-5, rtol=1e-5)


def test_decoder_step_with_fused_bias_matches_conv_with_bias():
    """`_up_relu` runs the decoder conv WITHOUT its bias and lets the fused kernel add it / reduce its gradient."""
    import lib
    torch.manual_seed(2)
    conv = torch.nn.Conv2d(8, 6, kernel_size=3, stride=1, padding=1).cuda()
    x = torch.randn(2, 8, 16, 16, device="cuda", requires_grad=True)
    skip = torch.randn(2, 6, 32, 32, device="cuda")
    ref = F.relu(F.interpolate(conv(x), scale_factor=(2, 2), mode="bilinear")) + skip
    got = lib.models.axialnet._up_relu(conv, x, skip)
    assert torch.allclose(got, ref, atol=1e-5, rtol=1e-5)
    d = torch.randn_like(ref)
    gr = torch.autograd.grad(ref, (x, conv.weight, conv.bias), d)
    gg = torch.autograd.grad(got, (x, conv.weight, conv.bias), d)
    for u, v in zip(gg, gr):
        assert (u - v).norm() <= 1e-4 * max(v.norm().item(), 1e-3)


@pytest.mark.parametrize("n,c,h,w,groups", [(4, 8, 64, 64, 1), (4, 128, 64, 64, 1), (32, 16, 4, 4, 16), (6, 5, 3, 7, 3)])
@pytest.mark.parametrize("relu,with_res,training", [(True, False, True), (True, True, True), (False, False, True), (True, True, False)])
def test_bn_act_matches_torch(n, c, h, w, groups, relu, with_res, training):
    from medt_b200.shell_ops import bn_act
    torch.manual_seed(2)
    x = (torch.randn(n, c, h, w, device="cuda") * 1.7 + 0.6).requires_grad_(True)
    res = torch.randn(n, c, h, w, device="cuda", requires_grad=True) if with_res else None
    a, b = torch.nn.BatchNorm2d(c).cuda(), torch.nn.BatchNorm2d(c).cuda()
    with torch.no_grad():
        for bn in (a, b):
            bn.weight.copy_(torch.linspace(0.5, 1.5, c)); bn.bias.copy_(torch.linspace(-0.3, 0.3, c))
            bn.running_mean.copy_(torch.linspace(-0.2, 0.4, c)); bn.running_var.copy_(torch.linspace(0.6, 1.4, c))
    a.train(training); b.train(training)
    g = groups if training else 1
    ref = torch.cat([a(xb) for xb in x.chunk(g)], 0)       # the reference's way: one BN call per patch
    if with_res:
        ref = ref + res
    if relu:
        ref = F.relu(ref)
    got = bn_act(x, b, groups, relu=relu, residual=res)
    assert torch.allclose(got, ref, atol=2e-5, rtol=2e-5)
    assert torch.allclose(a.running_mean, b.running_mean, atol=1e-6) and torch.allclose(a.running_var, b.running_var, atol=1e-5)
    assert int(a.num_batches_tracked) == int(b.num_batches_tracked)
    d = torch.randn_like(ref)
    ins_ref = [x, a.weight, a.bias] + ([res] if with_res else [])
    ins_got = [x, b.weight, b.bias] + ([res] if with_res else [])
    for u, v in zip(torch.autograd.grad(got, ins_got, d), torch.autograd.grad(ref, ins_ref, d)):
        assert (u - v).norm() <= 2e-4 * max(v.norm().item(), 1e-3), (u - v).norm() / v.norm()


@pytest.mark.parametrize("n,ci,co,h,w,stride", [(4, 8, 16, 64, 64, 1), (4, 32, 64, 64, 64, 2), (2, 128, 256, 16, 16, 2),
                                                 (64, 64, 128, 4, 4, 1), (3, 5, 7, 6, 10, 1), (2, 16, 32, 5, 7, 2)])
@pytest.mark.parametrize("relu_in", [False, True])
def test_conv1x1_matches_torch(n, ci, co, h, w, stride, relu_in):
    from medt_b200.shell_ops import conv1x1
    torch.manual_seed(3)
    conv = torch.nn.Conv2d(ci, co, kernel_size=1, stride=stride, bias=False).cuda()
    x = torch.randn(n, ci, h, w, device="cuda", requires_grad=True)
    tf32 = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    ref = conv(F.relu(x) if relu_in else x)
    got = conv1x1(x, conv, relu_in=relu_in)
    assert got.shape == ref.shape and torch.allclose(got, ref, atol=1e-5, rtol=1e-5)
    d = torch.randn_like(ref)
    gr = torch.autograd.grad(ref, (x, conv.weight), d)
    gg = torch.autograd.grad(got, (x, conv.weight), d)
    torch.backends.cudnn.allow_tf32 = tf32
    for u, v in zip(gg, gr):
        assert (u - v).norm() <= 1e-4 * max(v.norm().item(), 1e-3)


def test_shell_ops_refuse_cpu():
    from medt_b200.shell_ops import up2x_relu_add
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        up2x_relu_add(torch.randn(1, 1, 2, 2))


def test_misaligned_contiguous_views_give_the_same_result():
    """A contiguous view at an odd storage offset (4-byte aligned only) goes through `_cabi.dense()`; the raw ABI
    refuses it (tests/test_abi.py) instead of faulting in a 128-bit load."""
    import lib
    from medt_b200.shell_ops import bn_act, up2x_relu_add
    torch.manual_seed(5)
    n, c, h, w = 2, 16, 8, 8
    base = torch.randn(n * c * h * w + 1, device="cuda")
    xv = base[1:].view(n, c, h, w)
    assert xv.is_contiguous() and xv.data_ptr() % 16 != 0
    xa = xv.clone()
    mod = lib.models.axialnet.AxialAttention_dynamic(c, c, groups=8, kernel_size=h).cuda()
    bn = torch.nn.BatchNorm2d(c).cuda()
    for fn in (lambda t: mod(t), lambda t: bn_act(t, bn, 1, relu=True), lambda t: up2x_relu_add(t, None)):
        a, b = fn(xa), fn(xv)
        # same kernels, same data: only the order of the fp64 statistic atomics differs between two launches
        assert (a - b).abs().max().item() <= 1e-5 * a.abs().max().item()


@pytest.mark.parametrize("n,ci,co,h,k,stride,pad,bias", [(4, 8, 128, 64, 3, 1, 1, False),     # stem conv2
                                                         (4, 128, 8, 64, 3, 1, 1, False),     # stem conv3 (8 output channels)
                                                         (4, 256, 256, 8, 3, 2, 1, True),     # decoder1 (stride 2: dX via cuDNN)
                                                         (4, 256, 128, 8, 3, 1, 1, True),     # decoder2
                                                         (4, 128, 64, 16, 3, 1, 1, True),     # decoder3
                                                         (4, 64, 32, 32, 3, 1, 1, True),      # decoder4
                                                         (4, 32, 16, 64, 3, 1, 1, True),      # decoder5
                                                         (64, 8, 128, 16, 3, 1, 1, False),    # MedT local stem (64 patches)
                                                         (2, 16, 16, 128, 3, 1, 1, True)])    # LoGo decoderf
@pytest.mark.parametrize("tc", [1, 0])
def test_conv2d_gemm_matches_torch(n, ci, co, h, k, stride, pad, bias, tc):
    """medt_b200.shell_ops.conv2d against torch's fp32 convolution (TF32 off): forward, dX, dW (and the bias path), own forward +
    cuDNN backward and own forward + own backward -- on the tcgen05 3xTF32 tensor-core kernels (csrc/conv_tc.cu, tc = 1, the
    default routing of mt_conv2d_*) and on the FMA-pipe split-K kernels (csrc/conv_gemm.cu, tc = 0).  The same fp32-grade tolerance
    holds for both: the 3xTF32 split keeps the products at <= 2^-22 relative."""
    from medt_b200 import _cabi
    from medt_b200.shell_ops import Conv2dGemm
    lib = _cabi.load_library()
    lib.mt_set_conv_tc(tc)
    try:
        _conv2d_case(Conv2dGemm, n, ci, co, h, k, stride, pad, bias)
    finally:
        lib.mt_set_conv_tc(1)


def _conv2d_case(Conv2dGemm, n, ci, co, h, k, stride, pad, bias):
    torch.manual_seed(9)
    conv = torch.nn.Conv2d(ci, co, k, stride=stride, padding=pad, bias=bias).cuda()
    x = torch.randn(n, ci, h, h, device="cuda", requires_grad=True)
    ref = conv(x)
    d = torch.randn_like(ref)
    gr = torch.autograd.grad(ref, [x, conv.weight] + ([conv.bias] if bias else []), d)
    rel = lambda a, b: ((a - b).norm() / b.norm()).item()  # noqa: E731
    for own_bwd in (False, True):
        got = Conv2dGemm.apply(x, conv.weight, stride, pad, own_bwd)
        if bias:
            got = got + conv.bias.view(1, -1, 1, 1)
        gg = torch.autograd.grad(got, [x, conv.weight] + ([conv.bias] if bias else []), d)
        assert rel(got, ref) < 1e-5            # K = Ci*9 up to 2304 fp32 terms in a different order than cuDNN
        for a, b in zip(gg, gr):
            assert rel(a, b) < 2e-5            # dW sums up to 262 k positions in split slices (fp32 atomics)


@pytest.mark.parametrize("n,ci,co,h,k,stride,pad", [(4, 3, 8, 128, 7, 2, 3),       # the 7x7 stem geometry (generic tcgen05 kernel)
                                                    (1, 5, 20, 12, 3, 1, 1),        # Ci % 8 != 0, ragged sizes: generic kernel, FMA dW
                                                    (2, 7, 130, 20, 5, 1, 2),       # two output-channel tiles, 5x5
                                                    (3, 24, 40, 10, 3, 1, 1),       # slab kernel, tiles across image boundaries
                                                    (2, 8, 16, 130, 3, 1, 1),       # slab kernel, one stage (single smem buffer)
                                                    (1, 16, 128, 256, 3, 1, 1)])    # slab kernel, W = 256: 64-channel tiles (smem)
def test_conv2d_tensor_core_kernels_vs_fp64(n, ci, co, h, k, stride, pad):
    """csrc/conv_tc.cu (tcgen05 kind::tf32, 3xTF32 split, round-robin TMEM accumulators) against torch's fp64 convolution: forward,
    dX (stride 1) and dW stay at fp32-grade accuracy (< 2e-6; measured 1e-7 .. 5e-7, profiles/r3_conv_tc_accuracy_*.json) on
    shapes that exercise the generic gathered kernel, the slab kernel's image-boundary tiles and its single-buffer mode."""
    from medt_b200 import _cabi
    from medt_b200.shell_ops import Conv2dGemm
    _cabi.load_library().mt_set_conv_tc(1)
    torch.manual_seed(11)
    w = (torch.randn(co, ci, k, k, device="cuda") / (ci * k * k) ** 0.5).requires_grad_(True)
    x = torch.randn(n, ci, h, h, device="cuda", requires_grad=True)
    ref = F.conv2d(x.double(), w.double(), stride=stride, padding=pad)
    d = torch.randn_like(ref)
    gr = torch.autograd.grad(ref, [x, w], d)
    got = Conv2dGemm.apply(x, w, stride, pad, 1)
    gg = torch.autograd.grad(got, [x, w], d.float())
    rel = lambda a, b: ((a.double() - b).norm() / b.norm()).item()  # noqa: E731
    assert rel(got, ref) < 2e-6
    assert rel(gg[0], gr[0]) < 2e-6 and rel(gg[1], gr[1]) < 2e-6


@pytest.mark.parametrize("n,c,hw", [(4, 16, 128), (2, 16, 256), (3, 8, 33)])
def test_head_ce_matches_torch(n, c, hw):
    """medt_b200.shell_ops.head_ce == F.cross_entropy(adjust(F.relu(x)), y): loss, logits and every gradient"""
    from medt_b200.shell_ops import head_ce
    torch.manual_seed(2)
    adjust = torch.nn.Conv2d(c, 2, 1).cuda()
    x = torch.randn(n, c, hw, hw, device="cuda", requires_grad=True)
    y = torch.randint(0, 2, (n, hw, hw), device="cuda")
    logits = adjust(F.relu(x))
    loss = F.cross_entropy(logits, y)
    gr = torch.autograd.grad(loss, [x, adjust.weight, adjust.bias])
    loss2, logits2 = head_ce(x, adjust, y)
    gg = torch.autograd.grad(loss2, [x, adjust.weight, adjust.bias])
    assert abs(loss2.item() - loss.item()) < 2e-6 * max(1.0, abs(loss.item()))
    assert torch.allclose(logits2, logits, atol=1e-5, rtol=1e-5)
    for a, b in zip(gg, gr):
        assert ((a - b).norm() / b.norm()).item() < 2e-5
