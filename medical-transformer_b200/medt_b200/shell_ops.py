"""autograd wrappers of the fused shell kernels (mt_up2x_relu_add_*, mt_bn_act_*).

`up2x_relu_add(x, skip)`  == F.relu(F.interpolate(x, scale_factor=(2,2), mode='bilinear')) + skip
`bn_act(x, bn, groups, relu, residual)` == [relu](bn(x) [+ residual]) with per-group batch
statistics (groups > 1 only inside the batched LoGo local branch)."""
from __future__ import annotations

import ctypes as C

import os

import torch
import torch.nn.functional as F

from . import _cabi


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _stream(dev):
    return C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)


def _need_cuda(t, what):
    if not t.is_cuda:
        raise RuntimeError(f"medt_b200.{what}: input is on {t.device}; the B200 path has no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"medt_b200.{what}: float32 expected, got {t.dtype}")


class Up2xReluAdd(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, skip, bias):
        _need_cuda(x, "up2x_relu_add")
        lib = _cabi.load_library()
        x = _cabi.dense(x)
        n, c, h, w = x.shape
        skip_c = None if skip is None else _cabi.dense(skip)
        bias_c = None if bias is None else _cabi.dense(bias)
        out = torch.empty((n, c, 2 * h, 2 * w), dtype=torch.float32, device=x.device)
        need_bwd = x.requires_grad or (skip is not None and skip.requires_grad) or (bias is not None and bias.requires_grad)
        mask = torch.empty((n, c, 2 * h, 2 * w), dtype=torch.uint8, device=x.device) if need_bwd else None
        with torch.cuda.device(x.device):
            rc = lib.mt_up2x_relu_add_forward(n * c, c, h, w, _p(x), _p(bias_c), _p(skip_c), _p(out), _p(mask),
                                              _stream(x.device))
        _cabi.check(rc, "mt_up2x_relu_add_forward")
        _cabi.LAUNCHES["kernels"] += 1
        ctx.save_for_backward(mask)
        ctx.shape = (n, c, h, w)
        ctx.has_skip = skip is not None
        ctx.has_bias = bias is not None
        return out

    @staticmethod
    def backward(ctx, dout):
        (mask,) = ctx.saved_tensors
        n, c, h, w = ctx.shape
        lib = _cabi.load_library()
        dout = _cabi.dense(dout)
        din = dbias = None
        want_bias = ctx.has_bias and ctx.needs_input_grad[2]
        if ctx.needs_input_grad[0] or want_bias:
            din = torch.empty((n, c, h, w), dtype=torch.float32, device=dout.device)
            dbias = torch.empty((c,), dtype=torch.float32, device=dout.device) if want_bias else None
            with torch.cuda.device(dout.device):
                rc = lib.mt_up2x_relu_add_backward(n * c, c, h, w, _p(dout), _p(mask), _p(din), _p(dbias),
                                                   _stream(dout.device))
            _cabi.check(rc, "mt_up2x_relu_add_backward")
            _cabi.LAUNCHES["kernels"] += 1
        return (din if ctx.needs_input_grad[0] else None,
                dout if (ctx.has_skip and ctx.needs_input_grad[1]) else None, dbias)


def up2x_relu_add(x, skip=None, bias=None):
    """relu(bilinear_x2(x + bias[None, :, None, None])) [+ skip]; `bias` is the producing conv's bias when the conv
    itself was run without it (one elementwise pass and one reduction fewer per decoder step)."""
    return Up2xReluAdd.apply(x, skip, bias)


class BnAct(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, residual, weight, bias, rm, rv, cfg):
        groups, training, relu, eps, momentum = cfg
        _need_cuda(x, "bn_act")
        lib = _cabi.load_library()
        x = _cabi.dense(x)
        res = None if residual is None else _cabi.dense(residual)
        n, c, h, w = x.shape
        y = torch.empty_like(x)
        save_mean = torch.empty((2 * groups, c), dtype=torch.float32, device=x.device)
        save_invstd = torch.empty((groups, c), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mt_bn_act_forward(n, c, h * w, groups, int(training), int(relu), eps, momentum, _p(x), _p(res),
                                       _p(weight), _p(bias), _p(rm), _p(rv), _p(y), _p(save_mean), _p(save_invstd),
                                       _stream(x.device))
        _cabi.check(rc, "mt_bn_act_forward")
        _cabi.LAUNCHES["kernels"] += 2 if (training and groups > 1) else 1
        ctx.cfg = cfg
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, y, weight, save_mean, save_invstd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, y, weight, save_mean, save_invstd = ctx.saved_tensors
        groups, training, relu, eps, momentum = ctx.cfg
        lib = _cabi.load_library()
        dy = _cabi.dense(dy)
        n, c, h, w = x.shape
        dx = torch.empty_like(x)
        dres = torch.empty_like(x) if (ctx.has_res and ctx.needs_input_grad[1]) else None
        dg = torch.empty(c, dtype=torch.float32, device=x.device)
        db = torch.empty(c, dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mt_bn_act_backward(n, c, h * w, groups, int(training), int(relu), _p(x), _p(y), _p(dy), _p(weight),
                                        _p(save_mean), _p(save_invstd), _p(dx), _p(dres), _p(dg), _p(db), _stream(x.device))
        _cabi.check(rc, "mt_bn_act_backward")
        _cabi.LAUNCHES["kernels"] += 1
        need = ctx.needs_input_grad
        return (dx if need[0] else None, dres, dg if need[2] else None, db if need[3] else None, None, None, None)


def check_bn(bn):
    """The kernels implement the reference's BatchNorm configuration only (affine, running statistics, exponential
    momentum); anything else raises instead of being silently approximated."""
    if bn.momentum is None:
        raise NotImplementedError("medt_b200: BatchNorm(momentum=None) (cumulative average) is not implemented by the kernels")
    if not bn.affine or not bn.track_running_stats or bn.running_mean is None:
        raise NotImplementedError("medt_b200: the kernels need affine BatchNorm with track_running_stats=True")


def bn_act(x, bn, groups: int = 1, relu: bool = True, residual=None):
    """[relu](bn(x) [+ residual]); `bn` is the module's own nn.BatchNorm2d (parameters, running
    statistics and num_batches_tracked are updated exactly as `groups` successive calls would)."""
    check_bn(bn)
    training = bn.training
    g = groups if training else 1
    cfg = (int(g), bool(training), bool(relu), float(bn.eps), float(bn.momentum))
    y = BnAct.apply(x, residual, bn.weight, bn.bias, bn.running_mean, bn.running_var, cfg)
    if training:
        if getattr(bn, "_nbt_deferred", False):
            bn._stat_groups_bn = g
        else:
            with torch.no_grad():
                bn.num_batches_tracked += g
    return y


class Conv1x1(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, stride, relu_in):
        _need_cuda(x, "conv1x1")
        lib = _cabi.load_library()
        x = _cabi.dense(x)
        n, ci, h, w = x.shape
        co = weight.shape[0]
        ho, wo = (h - 1) // stride + 1, (w - 1) // stride + 1
        y = torch.empty((n, co, ho, wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mt_conv1x1_forward(n, ci, co, h, w, stride, int(relu_in), _p(x), _p(weight), _p(y),
                                        _stream(x.device))
        _cabi.check(rc, "mt_conv1x1_forward")
        _cabi.LAUNCHES["kernels"] += 1
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.relu_in = stride, int(relu_in)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        lib = _cabi.load_library()
        dy = _cabi.dense(dy)
        n, ci, h, w = x.shape
        co = weight.shape[0]
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(weight) if ctx.needs_input_grad[1] else None
        with torch.cuda.device(x.device):
            rc = lib.mt_conv1x1_backward(n, ci, co, h, w, ctx.stride, ctx.relu_in, _p(x), _p(weight), _p(dy), _p(dx),
                                         _p(dw), _stream(x.device))
        _cabi.check(rc, "mt_conv1x1_backward")
        _cabi.LAUNCHES["kernels"] += (1 if dx is not None else 0) + (1 if dw is not None else 0)
        return dx, dw, None, None


def conv1x1(x, conv, relu_in=False):
    """`conv` is the module's own bias-free nn.Conv2d(kernel_size=1) (weight (Co,Ci,1,1) used in place);
    relu_in=True computes conv(relu(x)) with the ReLU folded into the operand loads."""
    if conv.bias is not None or conv.kernel_size != (1, 1):
        return conv(F.relu(x) if relu_in else x)
    return Conv1x1.apply(x, conv.weight, int(conv.stride[0]), bool(relu_in))


_WGRAD_STREAMS = {}


class Conv2dGemm(torch.autograd.Function):
    """KxK convolution through mt_conv2d_* (tcgen05 3xTF32 tensor-core kernels of csrc/conv_tc.cu where the shape is eligible, the
    FMA-pipe split-K kernels of csrc/conv_gemm.cu otherwise).  `own_bwd`: 0 / False = the backward is cuDNN's
    (torch.ops.aten.convolution_backward), 1 / True = dX (stride 1 only) and dW on the own kernels, 2 = own dX, cuDNN dW."""

    @staticmethod
    def forward(ctx, x, weight, stride, pad, own_bwd):
        _need_cuda(x, "conv2d")
        lib = _cabi.load_library()
        x = _cabi.dense(x)
        weight = _cabi.dense(weight)
        n, ci, h, w = x.shape
        co, _, kh, kw = weight.shape
        ho, wo = (h + 2 * pad - kh) // stride + 1, (w + 2 * pad - kw) // stride + 1
        y = torch.empty((n, co, ho, wo), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mt_conv2d_forward(n, ci, co, h, w, kh, kw, stride, pad, _p(x), _p(weight), _p(y), _stream(x.device))
        _cabi.check(rc, "mt_conv2d_forward")
        _cabi.LAUNCHES["kernels"] += 1
        ctx.save_for_backward(x, weight)
        ctx.stride, ctx.pad, ctx.own_bwd = stride, pad, int(own_bwd)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight = ctx.saved_tensors
        dy = _cabi.dense(dy)
        n, ci, h, w = x.shape
        co, _, kh, kw = weight.shape
        need_dx, need_dw = ctx.needs_input_grad[0], ctx.needs_input_grad[1]

        def cudnn(mask_dx, mask_dw):
            return torch.ops.aten.convolution_backward(dy, x, weight, None, [ctx.stride] * 2, [ctx.pad] * 2, [1, 1], False, [0, 0], 1,
                                                       [bool(mask_dx), bool(mask_dw), False])
        if not ctx.own_bwd:
            dx, dw, _ = cudnn(need_dx, need_dw)
            return dx, dw, None, None, None
        lib = _cabi.load_library()
        own_dx = need_dx and ctx.stride == 1
        own_dw = need_dw and ctx.own_bwd == 1
        dx = torch.empty_like(x) if own_dx else None
        dw = torch.empty_like(weight) if own_dw else None
        side = None
        if own_dx and need_dw and not own_dw and os.environ.get("MEDT_WGRAD_STREAM"):
            # experiment (off): cuDNN's dW on a forked stream next to the own dX.  Measured on B200: 3.28 ms per gated step against
            # 3.13 ms with the dW behind the dX on the same stream (the fork / join edges cost more than the overlap returns)
            cur = torch.cuda.current_stream(x.device)
            key = (x.device.index, cur.cuda_stream)
            side = _WGRAD_STREAMS.get(key)
            if side is None:
                side = _WGRAD_STREAMS[key] = torch.cuda.Stream(device=x.device)
            side.wait_stream(cur)
            with torch.cuda.stream(side):
                dw = cudnn(False, True)[1]
        if own_dx or own_dw:
            with torch.cuda.device(x.device):
                rc = lib.mt_conv2d_backward(n, ci, co, h, w, kh, kw, ctx.stride, ctx.pad, _p(x), _p(weight), _p(dy), _p(dx),
                                            _p(dw if own_dw else None), _stream(x.device))
            _cabi.check(rc, "mt_conv2d_backward")
            _cabi.LAUNCHES["kernels"] += (1 if own_dx else 0) + (1 if own_dw else 0)
        if side is not None:
            torch.cuda.current_stream(x.device).wait_stream(side)
            dw.record_stream(torch.cuda.current_stream(x.device))
        rest_dx, rest_dw = need_dx and not own_dx, need_dw and not own_dw and side is None
        if rest_dx or rest_dw:
            gx, gw, _ = cudnn(rest_dx, rest_dw)
            dx = gx if rest_dx else dx
            dw = gw if rest_dw else dw
        return dx, dw, None, None, None


def conv2d(x, conv, with_bias=True):
    """`conv` is the module's own nn.Conv2d (weight used in place).  Routing (MEDT_OWN_CONV; measured on B200 with
    tools/time_convs.py, profiles/r2_notes.md):
      "tc" (default)  the 3x3 / stride-1 / pad-1 convolutions with Ci % 8 == 0 (every stem / decoder convolution but the 7x7 stem
                      and decoder1) run forward and dX on the tcgen05 3xTF32 slab kernel (csrc/conv_tc.cu); dW by MEDT_OWN_WGRAD
                      (default cuDNN); decoder1 (stride 2: cuDNN's fp32 forward collapses to 16 CTAs) runs its forward on the own
                      kernels and its backward on cuDNN;
      "strided"       only decoder1's forward on the own kernels (the round-2a policy);
      "all"           every eligible convolution forward + dX + dW on the own kernels;   "none"  cuDNN everywhere."""
    k = conv.kernel_size
    mode = os.environ.get("MEDT_OWN_CONV", "tc")
    ok = (k[0] == k[1] and conv.stride[0] == conv.stride[1] and isinstance(conv.padding, tuple) and conv.padding[0] == conv.padding[1]
          and conv.dilation == (1, 1) and conv.groups == 1 and conv.padding_mode == "zeros" and x.is_cuda and x.dtype == torch.float32
          and mode != "none")
    strided3 = ok and conv.stride[0] > 1 and k[0] == 3
    slab = (ok and k[0] == 3 and conv.stride[0] == 1 and conv.padding[0] == 1 and conv.in_channels % 8 == 0 and x.shape[-1] <= 256)
    if mode == "all":
        bwd = 1
    elif mode == "tc" and slab:
        bwd = 1 if os.environ.get("MEDT_OWN_WGRAD") else 2
    elif mode in ("tc", "strided") and strided3:
        bwd = 0
    else:
        ok = ok and mode == "all"
        bwd = 0
    if not ok:
        return F.conv2d(x, conv.weight, conv.bias if with_bias else None, conv.stride, conv.padding, conv.dilation, conv.groups)
    y = Conv2dGemm.apply(x, conv.weight, int(conv.stride[0]), int(conv.padding[0]), bwd)
    if with_bias and conv.bias is not None:
        y = y + conv.bias.view(1, -1, 1, 1)
    return y


class HeadCE(torch.autograd.Function):
    """loss = F.cross_entropy(adjust(relu(x)), y) for the two-class `adjust` 1x1 conv -- two kernels instead of ~10 ATen launches.
    Returns (loss, logits); logits carry no gradient (they are a by-product for logging / thresholding)."""

    @staticmethod
    def forward(ctx, x, weight, bias, y):
        _need_cuda(x, "head_ce")
        lib = _cabi.load_library()
        x = _cabi.dense(x)
        n, c, h, w = x.shape
        if weight.shape[0] != 2 or y.dtype != torch.int64:
            raise ValueError("medt_b200.head_ce: two output classes and int64 labels expected")
        y = y.contiguous()
        w2 = weight.reshape(2, c).contiguous()
        logits = torch.empty((n, 2, h, w), dtype=torch.float32, device=x.device)
        loss = torch.empty((), dtype=torch.float32, device=x.device)
        with torch.cuda.device(x.device):
            rc = lib.mt_head_ce_forward(n, c, h * w, _p(x), _p(w2), _p(bias), _p(y), _p(logits), _p(loss), _stream(x.device))
        _cabi.check(rc, "mt_head_ce_forward")
        _cabi.LAUNCHES["kernels"] += 1
        ctx.save_for_backward(x, w2, logits, y)
        ctx.wshape = tuple(weight.shape)
        ctx.mark_non_differentiable(logits)
        return loss, logits

    @staticmethod
    def backward(ctx, dloss, _dlogits):
        x, w2, logits, y = ctx.saved_tensors
        lib = _cabi.load_library()
        n, c, h, w = x.shape
        dx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        dw = torch.empty_like(w2)
        db = torch.empty(2, dtype=torch.float32, device=x.device)
        dl = dloss.reshape(1).contiguous().float()
        with torch.cuda.device(x.device):
            rc = lib.mt_head_ce_backward(n, c, h * w, _p(x), _p(w2), _p(logits), _p(y), _p(dl), _p(dx), _p(dw), _p(db), _stream(x.device))
        _cabi.check(rc, "mt_head_ce_backward")
        _cabi.LAUNCHES["kernels"] += 1
        return dx, dw.reshape(ctx.wshape), db, None


def head_ce(x, adjust, y):
    """(loss, logits) of `F.cross_entropy(adjust(F.relu(x)), y)`; `adjust` is the model's own nn.Conv2d(C, 2, 1) with bias."""
    return HeadCE.apply(x, adjust.weight, adjust.bias, y)
