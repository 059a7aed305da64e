"""Launch each of the 8 kernels of one axial op (gated layer1.0 hight_block shape by default)
once between cudaProfilerStart/Stop, for an `ncu --set full` capture of the hand-written
kernels.  usage: profile_stages.py [C L axis stride]"""
import ctypes
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch  # noqa: E402

import lib  # noqa: E402
from medt_b200 import _cabi  # noqa: E402
from medt_b200.axial_op import _desc, _ptr  # noqa: E402

C_ = int(sys.argv[1]) if len(sys.argv) > 1 else 16
L_ = int(sys.argv[2]) if len(sys.argv) > 2 else 64
axis = int(sys.argv[3]) if len(sys.argv) > 3 else 0
stride = int(sys.argv[4]) if len(sys.argv) > 4 else 1
dev = torch.device("cuda", 0)
torch.manual_seed(0)
mod = lib.models.axialnet.AxialAttention_dynamic(C_, C_, groups=8, kernel_size=L_, stride=stride, width=axis == 1).to(dev).train()
lib_ = _cabi.load_library()
x = torch.randn(4, C_, L_, L_, device=dev)
dout = torch.randn(4, C_, L_ // stride, L_ // stride, device=dev)
desc = _desc(x, axis, stride, "gated", True, 1, 1e-5, 0.1)
bq, bs, bo = mod.bn_qkv, mod.bn_similarity, mod.bn_output
params = _cabi.MtAxialParams(_ptr(mod.qkv_transform.weight), _ptr(mod.relative), _ptr(mod.f_qr), _ptr(mod.f_kr),
                             _ptr(mod.f_sv), _ptr(mod.f_sve), _ptr(bq.weight), _ptr(bq.bias), _ptr(bq.running_mean),
                             _ptr(bq.running_var), _ptr(bs.weight), _ptr(bs.bias), _ptr(bs.running_mean),
                             _ptr(bs.running_var), _ptr(bo.weight), _ptr(bo.bias), _ptr(bo.running_mean), _ptr(bo.running_var))
saved = torch.empty(lib_.mt_axial_saved_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
ws = torch.empty(lib_.mt_axial_bwd_workspace_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
out = torch.empty_like(dout)
gt = [torch.empty_like(t) for t in (x, mod.qkv_transform.weight, mod.relative)] + [torch.empty(4, device=dev)] + \
     [torch.empty_like(t) for t in (bq.weight, bq.bias, bs.weight, bs.bias, bo.weight, bo.bias)]
grads = _cabi.MtAxialGrads(*[_ptr(t) for t in gt])
st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
for _ in range(2):
    _cabi.check(lib_.mt_axial_forward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(out), _ptr(saved), st), "fwd")
    _cabi.check(lib_.mt_axial_backward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(dout), _ptr(saved), _ptr(ws),
                                       ctypes.byref(grads), st), "bwd")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
_cabi.check(lib_.mt_axial_forward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(out), _ptr(saved), st), "fwd")
_cabi.check(lib_.mt_axial_backward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(dout), _ptr(saved), _ptr(ws),
                                   ctypes.byref(grads), st), "bwd")
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled axial op C", C_, "L", L_, "axis", axis, "stride", stride)
