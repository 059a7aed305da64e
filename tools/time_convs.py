"""cuDNN fp32 vs the own implicit-GEMM kernels on the conv shapes of the 128-px / 512-px steps: "tc" = tcgen05 3xTF32 tensor-core
kernels (csrc/conv_tc.cu), "own" = FMA-pipe split-K kernels (csrc/conv_gemm.cu); forward and backward (dX + dW) timed with CUDA
events (L2-warm, 20 reps).  One JSON line."""
import json
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch  # noqa: E402

import medt_b200  # noqa: E402,F401  (TF32 off)
from medt_b200 import _cabi  # noqa: E402
from medt_b200.shell_ops import Conv2dGemm  # noqa: E402

LIB = _cabi.load_library()


def conv2d(x, conv):
    return Conv2dGemm.apply(x, conv.weight, int(conv.stride[0]), int(conv.padding[0]), True)

torch.backends.cudnn.benchmark = False     # as in the product (TrainStep captures the step; no autotuning under capture)
SHAPES = [("conv2 8->128 @64", 4, 8, 128, 64, 1), ("conv3 128->8 @64", 4, 128, 8, 64, 1), ("dec1 256->256 @8 s2", 4, 256, 256, 8, 2),
          ("dec2 256->128 @8", 4, 256, 128, 8, 1), ("dec3 128->64 @16", 4, 128, 64, 16, 1), ("dec4 64->32 @32", 4, 64, 32, 32, 1),
          ("dec5 32->16 @64", 4, 32, 16, 64, 1), ("512: conv2 @256", 4, 8, 128, 256, 1), ("512: conv3 @256", 4, 128, 8, 256, 1),
          ("512: dec5 32->16 @256", 4, 32, 16, 256, 1), ("512: dec4 64->32 @128", 4, 64, 32, 128, 1),
          ("medt local conv2 8->128 @16 x64", 64, 8, 128, 16, 1), ("medt local conv3 128->8 @16 x64", 64, 128, 8, 16, 1),
          ("medt local dec3 128->64 @4 x64", 64, 128, 64, 4, 1), ("medt local dec5 32->16 @16 x64", 64, 32, 16, 16, 1)]


def timeit(fn, reps=20):
    """GPU time per call: `reps` calls captured in ONE CUDA graph and replayed (a Python-issued loop of these 5 - 30 us launches
    is CPU-bound and measures the host)."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(reps):
            fn()
    g.replay()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    g.replay()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / reps * 1e3


res = {}
for name, n, ci, co, h, s in SHAPES:
    conv = torch.nn.Conv2d(ci, co, 3, stride=s, padding=1, bias=False).cuda()
    x = torch.randn(n, ci, h, h).cuda().requires_grad_(True)
    r = {}
    for tag, f, tc in (("cudnn", lambda: conv(x), 1), ("own", lambda: conv2d(x, conv), 0), ("tc", lambda: conv2d(x, conv), 1)):
        LIB.mt_set_conv_tc(tc)
        y = f()
        d = torch.randn(y.shape).to(y.device)            # CPU generator: the CUDA generator is bound to live graphs after a capture
        try:
            r[tag + "_fwd_us"] = round(timeit(f), 1)
            r[tag + "_bwd_us"] = round(timeit(lambda: torch.autograd.grad(f(), [x, conv.weight], d)) - r[tag + "_fwd_us"], 1)
            r[tag + "_dx_us"] = round(timeit(lambda: torch.autograd.grad(f(), [x], d)) - r[tag + "_fwd_us"], 1)
        except Exception as e:  # noqa: BLE001
            r[tag + "_error"] = str(e)[:200]
            torch.cuda.synchronize()
    flop = 2.0 * n * (h // s) ** 2 * co * ci * 9
    for tag in ("own", "tc"):
        if tag + "_fwd_us" in r:
            r[tag + "_fwd_tflops"] = round(flop / r[tag + "_fwd_us"] / 1e6, 2)
    res[name] = r
print(json.dumps(res))
