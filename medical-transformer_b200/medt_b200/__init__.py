"""medt_b200 -- host side of the B200-native gated axial-attention path.

Python/PyTorch here is plumbing only (device memory, streams, autograd graph, NCCL); the
arithmetic of the hot path runs in libmedt_b200.so (hand-written sm_100a CUDA behind the
C ABI of include/medt_axial.h).  There is NO CPU or eager fallback: calling the op without
the library, or on a non-CUDA tensor, raises."""
from ._cabi import lib_path, load_library, abi_version  # noqa: F401
from .axial_op import axial_attention, AxialAttnFunction  # noqa: F401

__all__ = ["lib_path", "load_library", "abi_version", "axial_attention", "AxialAttnFunction"]
