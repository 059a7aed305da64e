"""medt_b200 -- host side of the B200-native gated axial-attention path.

Python/PyTorch here is plumbing only (device memory, streams, autograd graph, NCCL); the
arithmetic of the hot path runs in libmedt_b200.so (hand-written sm_100a CUDA behind the
C ABI of include/medt_axial.h).  There is NO CPU or eager fallback: calling the op without
the library, or on a non-CUDA tensor, raises."""
import os as _os

import torch as _torch

# Parity first: with cuDNN's default TF32 convolutions the train-mode logits drift 8-13 % from
# the reference (measured, DESIGN.md "precision"); in fp32 they sit at the reference's own
# fp32-vs-fp64 noise floor (5e-4).  MEDT_ALLOW_TF32=1 opts back in.
if _os.environ.get("MEDT_ALLOW_TF32", "0") != "1":
    _torch.backends.cudnn.allow_tf32 = False
    _torch.backends.cuda.matmul.allow_tf32 = False

from ._cabi import lib_path, load_library, abi_version  # noqa: F401
from .axial_op import axial_attention, AxialAttnFunction  # noqa: F401

__all__ = ["lib_path", "load_library", "abi_version", "axial_attention", "AxialAttnFunction"]
