"""One call of mt_conv2d_forward / backward (dX only) / backward (dW only) per conv shape, for an ncu launch list
(ncu --metrics gpu__time_duration.sum --clock-control none): which kernel runs each and how long it takes."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch  # noqa: E402

import medt_b200  # noqa: E402,F401
from medt_b200 import _cabi  # noqa: E402

lib = _cabi.load_library()
lib.mt_set_conv_tc(int(os.environ.get("TC", "1")))
SHAPES = [(4, 8, 128, 64), (4, 128, 8, 64), (4, 256, 128, 8), (4, 128, 64, 16), (4, 64, 32, 32), (4, 32, 16, 64),
          (4, 8, 128, 256), (4, 128, 8, 256), (4, 32, 16, 256), (64, 128, 8, 16), (64, 128, 64, 4)]
p = lambda t: t.data_ptr()  # noqa: E731
st = torch.cuda.current_stream().cuda_stream
for rep in range(2):
    for (n, ci, co, h) in SHAPES:
        x = torch.randn(n, ci, h, h, device="cuda")
        w = torch.randn(co, ci, 3, 3, device="cuda")
        y = torch.empty(n, co, h, h, device="cuda")
        dy = torch.randn(n, co, h, h, device="cuda")
        dx = torch.empty_like(x)
        dw = torch.empty_like(w)
        torch.cuda.synchronize()
        assert lib.mt_conv2d_forward(n, ci, co, h, h, 3, 3, 1, 1, p(x), p(w), p(y), st) == 0
        torch.cuda.synchronize()
        assert lib.mt_conv2d_backward(n, ci, co, h, h, 3, 3, 1, 1, p(x), p(w), p(dy), p(dx), None, st) == 0
        torch.cuda.synchronize()
        assert lib.mt_conv2d_backward(n, ci, co, h, h, 3, 3, 1, 1, p(x), p(w), p(dy), None, p(dw), st) == 0
        torch.cuda.synchronize()
print("done")
