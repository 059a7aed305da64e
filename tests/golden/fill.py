"""Deterministic, RNG-version-independent tensor fill shared by the golden generator
(make_golden.py, which runs the real reference) and the tests (which must rebuild the
same weights/inputs on a box where /root/reference does not exist).

Values come from numpy's legacy RandomState (bit-stable across numpy versions), seeded
by crc32 of the tensor's state_dict key, and are shaped so that every code path is
exercised: non-trivial BN affine + running stats, unequal gates (one negative)."""
from __future__ import annotations

import zlib

import numpy as np


def _rs(name: str, salt: int = 0) -> np.random.RandomState:
    return np.random.RandomState((zlib.crc32(name.encode()) + 7919 * salt) % (2 ** 31))


def tensor_for(name: str, shape, salt: int = 0) -> np.ndarray:
    """Value for the state_dict entry `name`: float64 array whose entries are exactly
    representable in float32, so the fp64 reference run, the fp64 oracle and the fp32 CUDA
    path all see bit-identical parameters."""
    v = _tensor_for(name, shape, salt)
    return v if v.dtype == np.int64 else v.astype(np.float32).astype(np.float64)


def _tensor_for(name: str, shape, salt: int = 0) -> np.ndarray:
    rs = _rs(name, salt)
    shape = tuple(shape)
    leaf = name.split(".")[-1]
    parent = name.split(".")[-2] if "." in name else ""
    if leaf == "num_batches_tracked":
        return np.array(3, dtype=np.int64)
    if leaf == "running_mean":
        return 0.2 * rs.standard_normal(shape)
    if leaf == "running_var":
        return 0.6 + 0.8 * rs.random_sample(shape)
    if leaf == "weight" and (parent.startswith("bn") or parent in ("1",)):   # BN gamma ("downsample.1")
        return 1.0 + 0.25 * rs.standard_normal(shape)
    if leaf == "bias" and (parent.startswith("bn") or parent in ("1",)):
        return 0.1 * rs.standard_normal(shape)
    if leaf in ("f_qr", "f_kr", "f_sv", "f_sve"):
        return np.array({"f_qr": 0.35, "f_kr": -0.2, "f_sv": 0.9, "f_sve": 0.15}[leaf])
    if leaf == "relative":
        return rs.standard_normal(shape) * np.sqrt(1.0 / (shape[0] / 2))
    if leaf == "weight":                                                      # conv weights
        fan_in = int(np.prod(shape[1:])) if len(shape) > 1 else 1
        return rs.standard_normal(shape) * np.sqrt(1.0 / fan_in)
    if leaf == "bias":
        return 0.05 * rs.standard_normal(shape)
    raise KeyError(f"no fill rule for {name}")


def image_batch(n, c, s, salt=0) -> np.ndarray:
    """Synthetic images in [0,1) quantised to k/255 like to_tensor(uint8) (utils.py:92)."""
    rs = _rs("image_batch", salt)
    return (np.floor(rs.random_sample((n, c, s, s)) * 256.0) / 255.0).astype(np.float32).astype(np.float64)


def mask_batch(n, s, salt=0) -> np.ndarray:
    rs = _rs("mask_batch", salt)
    return (rs.random_sample((n, s, s)) > 0.5).astype(np.int64)


def fill_state_dict(sd, salt: int = 0):
    """Return {key: np.ndarray} for every entry of a (reference-schema) state_dict except the
    derived int64 `flatten_index` buffers, which are left as constructed."""
    out = {}
    for k, v in sd.items():
        if k.endswith("flatten_index"):
            continue
        out[k] = tensor_for(k, tuple(v.shape), salt)
    return out
