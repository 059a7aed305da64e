"""Bring-up of csrc/conv_tc.cu on a B200: every conv shape of the network through mt_conv2d_* with the tcgen05 path in mode
`sys.argv[1]` (1 = product descriptors, 2 = LBO/SBO swapped, 0 = FMA kernels) against fp64 torch; prints one JSON line."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402

import medt_b200  # noqa: E402,F401
from medt_b200 import _cabi  # noqa: E402
from medt_b200.shell_ops import Conv2dGemm  # noqa: E402

mode = int(sys.argv[1]) if len(sys.argv) > 1 else 1
lib = _cabi.load_library()
lib.mt_set_conv_tc(mode)
SHAPES = [(4, 8, 128, 64, 3, 1, 1), (4, 128, 8, 64, 3, 1, 1), (4, 256, 256, 8, 3, 2, 1), (4, 256, 128, 8, 3, 1, 1),
          (4, 128, 64, 16, 3, 1, 1), (4, 64, 32, 32, 3, 1, 1), (4, 32, 16, 64, 3, 1, 1), (64, 8, 128, 16, 3, 1, 1),
          (2, 16, 16, 128, 3, 1, 1), (4, 3, 8, 128, 7, 2, 3), (1, 5, 20, 12, 3, 1, 1), (2, 7, 130, 20, 5, 1, 2)]
res = {}
for (n, ci, co, h, k, s, pad) in SHAPES:
    torch.manual_seed(9)
    w = torch.randn(co, ci, k, k, device="cuda") / (ci * k * k) ** 0.5
    w.requires_grad_(True)
    x = torch.randn(n, ci, h, h, device="cuda", requires_grad=True)
    ref = F.conv2d(x.double(), w.double(), stride=s, padding=pad)
    d = torch.randn_like(ref)
    gr = torch.autograd.grad(ref, [x, w], d)
    got = Conv2dGemm.apply(x, w, s, pad, True)
    gg = torch.autograd.grad(got, [x, w], d.float())
    torch.cuda.synchronize()
    rel = lambda a, b: float(((a.double() - b).norm() / b.norm()).item())  # noqa: E731
    res[f"{n}x{ci}->{co}@{h} k{k}s{s}"] = {"fwd": rel(got, ref), "dx": rel(gg[0], gr[0]), "dw": rel(gg[1], gr[1])}
print(json.dumps({"mode": mode, "res": res}))
