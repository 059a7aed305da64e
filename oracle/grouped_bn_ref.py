"""CPU ORACLE (test infrastructure, NOT product code) -- plain-PyTorch BatchNorm2d with per-group batch
statistics: the semantics the CUDA kernel mt_bn_act_forward must reproduce for the batched LoGo local branch.
Only tests/ import it (tests/test_dropin.py pins it against 16 sequential nn.BatchNorm2d calls).


The reference runs its 16 patches through the SAME BatchNorm modules one after another
(axialnet.py:661-700), so each call normalises with that patch's statistics and applies one
running-stat EMA step.  Batching the patches on-device must keep both effects: `groups`
consecutive image blocks get separate statistics, the running stats receive the `groups`
sequential EMA updates in patch order, and num_batches_tracked advances by `groups`."""
from __future__ import annotations

import torch
import torch.nn.functional as F


def batch_norm_grouped(x, bn, groups: int = 1):
    if groups == 1 or not bn.training:
        if bn.training and not getattr(bn, "_nbt_deferred", False):
            bn.num_batches_tracked += 1
        return F.batch_norm(x, bn.running_mean, bn.running_var, bn.weight, bn.bias, bn.training,
                            bn.momentum if bn.momentum is not None else 0.1, bn.eps)
    pn, c, h, w = x.shape
    n = pn // groups
    xg = x.view(groups, n, c, h, w)
    var, mean = torch.var_mean(xg, dim=(1, 3, 4), unbiased=False, keepdim=True)
    y = (xg - mean) * torch.rsqrt(var + bn.eps)
    y = y * bn.weight.view(1, 1, c, 1, 1) + bn.bias.view(1, 1, c, 1, 1)
    with torch.no_grad():
        m = bn.momentum if bn.momentum is not None else 0.1
        cnt = n * h * w
        k = torch.arange(groups - 1, -1, -1, device=x.device, dtype=x.dtype)
        wts = (m * (1.0 - m) ** k).view(groups, 1)                      # weight of patch p in the final EMA
        keep = (1.0 - m) ** groups
        bn.running_mean.mul_(keep).add_((wts * mean.view(groups, c)).sum(0))
        bn.running_var.mul_(keep).add_((wts * var.view(groups, c) * (cnt / max(cnt - 1, 1))).sum(0))
        if getattr(bn, "_nbt_deferred", False):
            bn._stat_groups_bn = groups
        else:
            bn.num_batches_tracked += groups
    return y.view(pn, c, h, w)
