"""torch.autograd wrapper of mt_axial_forward / mt_axial_backward.

Mirrors the reference operator interface: one call == one AxialAttention*.forward
(lib/models/axialnet.py:52-92, :142-189, :222-253) on an NCHW fp32 CUDA tensor, with the
module's own Parameters/buffers (so state_dict, Adam and the epoch-10 requires_grad flip of
train.py:169-171 all work on the very same tensors)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _cabi


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _desc(x, axis, stride, variant, training, ngroups, eps, momentum):
    n, c, h, w = x.shape
    return _cabi.MtAxialDesc(n, c, h, w, axis, stride, _cabi.MT_VARIANT[variant], int(training), ngroups,
                             eps, momentum)


def _check_tensor(t, name):
    if t is None:
        return
    if not t.is_cuda:
        raise RuntimeError(f"medt_b200.axial_attention: `{name}` is on {t.device}; the B200 path has no CPU fallback")
    if t.dtype != torch.float32:
        raise RuntimeError(f"medt_b200.axial_attention: `{name}` must be float32, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"medt_b200.axial_attention: `{name}` must be contiguous")


class AxialAttnFunction(torch.autograd.Function):
    """inputs: x, w_qkv, relative, f_qr, f_kr, f_sv, f_sve, bn_qkv_w, bn_qkv_b, bn_sim_w, bn_sim_b,
    bn_out_w, bn_out_b, then non-differentiable: running stats (6 tensors) and a cfg tuple."""

    @staticmethod
    def forward(ctx, x, w_qkv, relative, f_qr, f_kr, f_sv, f_sve, q_w, q_b, s_w, s_b, o_w, o_b,
                q_rm, q_rv, s_rm, s_rv, o_rm, o_rv, cfg):
        axis, stride, variant, training, ngroups, eps, momentum = cfg
        lib = _cabi.load_library()
        x = _cabi.dense(x)
        for name, t in (("x", x), ("qkv_transform.weight", w_qkv), ("relative", relative), ("f_qr", f_qr),
                        ("bn_qkv.weight", q_w), ("bn_similarity.weight", s_w), ("bn_output.weight", o_w),
                        ("bn_qkv.running_mean", q_rm), ("bn_similarity.running_var", s_rv)):
            _check_tensor(t, name)
        n, c, h, w = x.shape
        desc = _desc(x, axis, stride, variant, training, ngroups, eps, momentum)
        params = _cabi.MtAxialParams(_ptr(w_qkv), _ptr(relative), _ptr(f_qr), _ptr(f_kr), _ptr(f_sv), _ptr(f_sve),
                                     _ptr(q_w), _ptr(q_b), _ptr(q_rm), _ptr(q_rv),
                                     _ptr(s_w), _ptr(s_b), _ptr(s_rm), _ptr(s_rv),
                                     _ptr(o_w), _ptr(o_b), _ptr(o_rm), _ptr(o_rv))
        nbytes = lib.mt_axial_saved_bytes(C.byref(desc))
        if nbytes == 0:
            _cabi.check(1, "mt_axial_saved_bytes (bad geometry)")
        saved = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        out = torch.empty((n, c, h // stride, w // stride), dtype=torch.float32, device=x.device)
        stream = torch.cuda.current_stream(x.device).cuda_stream
        with torch.cuda.device(x.device):
            rc = lib.mt_axial_forward(C.byref(desc), C.byref(params), _ptr(x), _ptr(out), _ptr(saved),
                                      C.c_void_p(stream))
        _cabi.check(rc, "mt_axial_forward")
        _cabi.count_launches()
        ctx.cfg = cfg
        ctx.save_for_backward(x, w_qkv, relative, f_qr, f_kr, f_sv, f_sve, q_w, q_b, s_w, s_b, o_w, o_b,
                              q_rm, q_rv, s_rm, s_rv, o_rm, o_rv, saved)
        return out

    @staticmethod
    def backward(ctx, dout):
        (x, w_qkv, relative, f_qr, f_kr, f_sv, f_sve, q_w, q_b, s_w, s_b, o_w, o_b,
         q_rm, q_rv, s_rm, s_rv, o_rm, o_rv, saved) = ctx.saved_tensors
        axis, stride, variant, training, ngroups, eps, momentum = ctx.cfg
        lib = _cabi.load_library()
        dout = _cabi.dense(dout)
        _check_tensor(dout, "grad_output")
        desc = _desc(x, axis, stride, variant, training, ngroups, eps, momentum)
        params = _cabi.MtAxialParams(_ptr(w_qkv), _ptr(relative), _ptr(f_qr), _ptr(f_kr), _ptr(f_sv), _ptr(f_sve),
                                     _ptr(q_w), _ptr(q_b), _ptr(q_rm), _ptr(q_rv),
                                     _ptr(s_w), _ptr(s_b), _ptr(s_rm), _ptr(s_rv),
                                     _ptr(o_w), _ptr(o_b), _ptr(o_rm), _ptr(o_rv))
        dev = x.device
        new = lambda ref: None if ref is None else torch.empty_like(ref)  # noqa: E731
        dx = torch.empty_like(x)
        dw = torch.empty_like(w_qkv)
        drel = new(relative)
        dgates = torch.empty(4, dtype=torch.float32, device=dev)
        dq_w, dq_b, ds_w, ds_b, do_w, do_b = (torch.empty_like(t) for t in (q_w, q_b, s_w, s_b, o_w, o_b))
        grads = _cabi.MtAxialGrads(_ptr(dx), _ptr(dw), _ptr(drel), _ptr(dgates), _ptr(dq_w), _ptr(dq_b),
                                   _ptr(ds_w), _ptr(ds_b), _ptr(do_w), _ptr(do_b))
        ws = torch.empty(lib.mt_axial_bwd_workspace_bytes(C.byref(desc)), dtype=torch.uint8, device=dev)
        stream = torch.cuda.current_stream(dev).cuda_stream
        with torch.cuda.device(dev):
            rc = lib.mt_axial_backward(C.byref(desc), C.byref(params), _ptr(x), _ptr(dout), _ptr(saved), _ptr(ws),
                                       C.byref(grads), C.c_void_p(stream))
        _cabi.check(rc, "mt_axial_backward")
        _cabi.count_launches()
        need = ctx.needs_input_grad
        g = lambda i, t: t if (t is not None and need[i]) else None  # noqa: E731
        gate = lambda i, k: dgates[k].reshape(()) if (need[i] and variant == "gated") else None  # noqa: E731
        return (g(0, dx), g(1, dw), g(2, drel), gate(3, 0), gate(4, 1), gate(5, 2), gate(6, 3),
                g(7, dq_w), g(8, dq_b), g(9, ds_w), g(10, ds_b), g(11, do_w), g(12, do_b),
                None, None, None, None, None, None, None)


def axial_attention(x, mod, ngroups: int = 1):
    """Run one AxialAttention* module `mod` (from lib.models.axialnet) on x through the CUDA path."""
    variant = mod._variant
    L = x.shape[3] if mod.width else x.shape[2]
    if variant != "nopos" and L != mod.kernel_size:
        # same failure class as the reference's einsum size mismatch (axialnet.py:155-158)
        raise RuntimeError(f"axial length {L} does not match kernel_size {mod.kernel_size} of the position table")
    gated = variant == "gated"
    bq, bs, bo = mod.bn_qkv, mod.bn_similarity, mod.bn_output
    from .shell_ops import check_bn
    for bn in (bq, bs, bo):
        check_bn(bn)
    training = mod.training
    cfg = (1 if mod.width else 0, int(mod.stride), variant, bool(training), int(ngroups), float(bq.eps), float(bq.momentum))
    out = AxialAttnFunction.apply(
        x, mod.qkv_transform.weight, getattr(mod, "relative", None),
        mod.f_qr if gated else None, mod.f_kr if gated else None, mod.f_sv if gated else None,
        mod.f_sve if gated else None,
        bq.weight, bq.bias, bs.weight, bs.bias, bo.weight, bo.bias,
        bq.running_mean, bq.running_var, bs.running_mean, bs.running_var, bo.running_mean, bo.running_var, cfg)
    if training:
        for bn in (bq, bs, bo):
            if getattr(bn, "_nbt_deferred", False):      # TrainStep adds all counters in one launch
                bn._stat_groups_bn = ngroups
            else:
                with torch.no_grad():
                    bn.num_batches_tracked += ngroups
    return out
