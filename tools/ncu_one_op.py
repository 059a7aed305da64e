"""ONE forward + ONE backward of one axial op through the default (cooperative) path, for
    ncu --set full --clock-control none --import-source on -k regex:k_axial -c 2 -o <out> python tools/ncu_one_op.py C N H W axis stride
(variant gated).  Reuses bench.op_profile's buffer set-up with reps = 1."""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch  # noqa: E402

import bench  # noqa: E402

C, N, H, W, axis, stride = [int(v) for v in sys.argv[1:7]]
dev = torch.device("cuda", 0)
r = bench.op_profile(dev, lambda: None, "gated", C, N, H, W, axis, stride, reps=1)
torch.cuda.synchronize()
print({k: r[k] for k in ("B", "C", "L", "fwd_us_warm", "bwd_us_warm")})
