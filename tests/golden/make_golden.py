"""Generate the golden fixtures by EXECUTING THE REFERENCE ITSELF (Python, CPU).

Run here (the container that has /root/reference), never on the GPU box:
    python tests/golden/make_golden.py
It imports /root/reference/lib/models/axialnet.py unmodified (with empty stubs for the
import-only matplotlib / skimage modules), fills every parameter/buffer with the
deterministic values of fill.py, runs forward/backward, and writes
    tests/golden/op_*.npz      one axial-attention op (fp64), train + eval
    tests/golden/model_*.npz   whole-network logits / loss / grads
The fixtures are committed; the reference sources are not copied anywhere.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
REPO = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(REPO, "medical-transformer_b200", "stubs"))
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

import fill  # noqa: E402
from lib.models import axialnet as ref  # noqa: E402  (the real reference)

torch.set_num_threads(8)

OP_CASES = [
    # name,            cls,                      C,  N, H, W, width, stride
    ("gated_c16_h",   "AxialAttention_dynamic", 16, 2, 8, 8, False, 1),
    ("gated_c32_w_s2", "AxialAttention_dynamic", 32, 2, 8, 8, True, 2),
    ("gated_c64_h",   "AxialAttention_dynamic", 64, 1, 4, 4, False, 1),
    ("gated_c128_w",  "AxialAttention_dynamic", 128, 1, 4, 4, True, 1),
    ("gated_c16_l16_w", "AxialAttention_dynamic", 16, 2, 16, 16, True, 1),
    ("pos_c16_h",     "AxialAttention",         16, 2, 8, 8, False, 1),
    ("pos_c32_w_s2",  "AxialAttention",         32, 1, 8, 8, True, 2),
    ("nopos_c16_w_s2", "AxialAttention_wopos",  16, 2, 8, 8, True, 2),
    ("nopos_c32_h",   "AxialAttention_wopos",   32, 2, 4, 8, False, 1),
    ("nopos_c128_w",  "AxialAttention_wopos",   128, 2, 4, 4, True, 1),
]


def _fill_module(mod, prefix, dtype, salt=0):
    sd = mod.state_dict()
    vals = fill.fill_state_dict({prefix + k: v for k, v in sd.items()}, salt)
    with torch.no_grad():
        for k, v in sd.items():
            if k.endswith("flatten_index"):
                continue
            t = torch.as_tensor(vals[prefix + k])
            v.copy_(t.to(v.dtype) if v.dtype == torch.int64 else t.to(dtype))
    for p in mod.parameters():
        p.requires_grad_(True)


def _rand(name, shape):
    return np.random.RandomState((sum(map(ord, name)) * 2654435761) % (2 ** 31)).standard_normal(shape).astype(np.float32).astype(np.float64)


def make_op(name, cls, C, N, H, W, width, stride, ngroups=1):
    L = W if width else H
    out = {}
    for mode in ("train", "eval"):
        mod = getattr(ref, cls)(C, C, groups=8, kernel_size=L, stride=stride, width=width).double()
        _fill_module(mod, f"op.{name}.", torch.float64)
        mod.train(mode == "train")
        x = torch.tensor(_rand(name + "x", (N, C, H, W)), dtype=torch.float64, requires_grad=True)
        if ngroups == 1:
            y = mod(x)
        else:   # LoGo-style: one call per image block -> per-block batch statistics, sequential EMA
            y = torch.cat([mod(xb) for xb in x.chunk(ngroups, dim=0)], dim=0)
        dout = torch.tensor(_rand(name + "dout", tuple(y.shape)), dtype=torch.float64)
        (y * dout).sum().backward()
        out[f"{mode}_out"] = y.detach().numpy()
        out[f"{mode}_dx"] = x.grad.numpy()
        for k, p in mod.named_parameters():
            out[f"{mode}_grad.{k}"] = (p.grad if p.grad is not None else torch.zeros_like(p)).numpy()
        for k, b in mod.named_buffers():
            if "running" in k or "num_batches" in k:
                out[f"{mode}_buf.{k}"] = b.numpy().copy()
        if mode == "train":
            out["x"] = x.detach().numpy()
            out["dout"] = dout.numpy()
    out["meta"] = np.array([C, N, H, W, int(width), stride, ngroups], dtype=np.int64)
    out["cls"] = np.array(cls)
    np.savez_compressed(os.path.join(HERE, f"op_{name}.npz"), **out)
    print("wrote op", name)


SELECT_GRADS = {
    "gated": ["conv1.weight", "layer1.0.hight_block.relative", "layer1.0.hight_block.qkv_transform.weight",
              "layer1.0.hight_block.f_qr", "layer1.0.hight_block.f_kr", "layer1.0.hight_block.f_sv",
              "layer1.0.hight_block.f_sve", "layer2.0.width_block.bn_similarity.weight",
              "layer2.0.width_block.bn_output.weight", "layer4.0.width_block.relative",
              "layer3.1.hight_block.bn_qkv.weight", "adjust.bias"],
    "axialunet": ["conv1.weight", "layer1.0.hight_block.relative", "layer2.0.width_block.bn_similarity.weight",
                  "adjust.bias"],
    "MedT": ["conv1.weight", "conv1_p.weight", "layer1.0.hight_block.relative", "layer1.0.hight_block.f_sve",
             "layer1_p.0.hight_block.qkv_transform.weight", "layer2_p.0.width_block.bn_similarity.weight",
             "layer4_p.0.width_block.bn_output.weight", "decoderf.weight", "adjust.bias"],
    "logo": ["conv1.weight", "conv1_p.weight", "layer1_p.0.hight_block.relative",
             "layer3_p.1.width_block.relative", "decoderf.weight", "adjust.bias"],
}
SELECT_BUFS = {
    "gated": ["bn1.running_mean", "layer1.0.hight_block.bn_similarity.running_var",
              "layer1.0.hight_block.bn_output.running_mean", "layer2.0.width_block.bn_qkv.running_var",
              "layer1.0.bn1.num_batches_tracked"],
    "axialunet": ["bn1.running_mean", "layer1.0.hight_block.bn_similarity.running_var"],
    "MedT": ["bn1_p.running_mean", "bn1_p.num_batches_tracked", "layer1_p.0.hight_block.bn_similarity.running_var",
             "layer1_p.0.hight_block.bn_qkv.running_mean", "layer4_p.0.width_block.bn_output.running_var",
             "layer1_p.0.bn1.running_var", "layer1.0.hight_block.bn_similarity.running_var"],
    "logo": ["bn1_p.running_mean", "layer1_p.0.hight_block.bn_similarity.running_var"],
}


def make_model(tag, factory, S, N, modes=("train", "eval"), with_grads=True):
    out = {"meta": np.array([S, N], dtype=np.int64), "factory": np.array(factory)}
    xnp = fill.image_batch(N, 3, S)
    ynp = fill.mask_batch(N, S)
    for mode in modes:
        logits = {}
        for dtype in (torch.float64, torch.float32):
            torch.manual_seed(0)
            m = getattr(ref, factory)(img_size=S, imgchan=3).to(dtype)
            _fill_module(m, "", dtype)
            m.train(mode == "train")
            x = torch.tensor(xnp, dtype=dtype)
            y = torch.tensor(ynp)
            if with_grads:
                lg = m(x)
                loss = torch.nn.functional.cross_entropy(lg, y)     # metrics.py:17-20
                loss.backward()
            else:
                with torch.no_grad():
                    lg = m(x)
                loss = torch.nn.functional.cross_entropy(lg, y)
            logits[dtype] = lg.detach()
            if with_grads:
                gdict = {k: (torch.zeros_like(p) if p.grad is None else p.grad).double() for k, p in m.named_parameters()}
                if dtype == torch.float64:
                    g64 = gdict
                else:   # the reference's OWN fp32-vs-fp64 gradient gap = the noise floor for the fp32 CUDA path
                    tot = float(np.sqrt(sum(float(g.norm()) ** 2 for g in g64.values())))
                    out[f"{mode}_ref_gradnorm_fp32_vs_fp64_worst"] = np.array(max(
                        abs(float(gdict[k].norm()) - float(g64[k].norm())) / max(float(g64[k].norm()), 1e-3 * tot)
                        for k in g64))
                    out[f"{mode}_ref_grad_fp32_vs_fp64_worst"] = np.array(max(
                        float((gdict[k] - g64[k]).norm()) / max(float(g64[k].norm()), 1e-3 * tot) for k in g64))
            if dtype == torch.float64:
                out[f"{mode}_logits"] = lg.detach().numpy().astype(np.float32)
                out[f"{mode}_loss"] = np.array(loss.item())
                if with_grads:
                    names, norms = [], []
                    for k, p in m.named_parameters():
                        names.append(k)
                        norms.append(0.0 if p.grad is None else p.grad.norm().item())
                    out[f"{mode}_gradnorm_names"] = np.array(names)
                    out[f"{mode}_gradnorms"] = np.array(norms)
                    pd = dict(m.named_parameters())
                    for k in SELECT_GRADS[factory]:
                        g = pd[k].grad
                        out[f"{mode}_grad.{k}"] = (torch.zeros_like(pd[k]) if g is None else g).numpy()
                if mode == "train":
                    bd = dict(m.named_buffers())
                    for k in SELECT_BUFS[factory]:
                        out[f"train_buf.{k}"] = bd[k].numpy().copy()
        a, b = logits[torch.float64], logits[torch.float32].double()
        out[f"{mode}_ref_fp32_vs_fp64_rel"] = np.array(((a - b).norm() / a.norm()).item())
        out[f"{mode}_ref_fp32_vs_fp64_maxabs"] = np.array((a - b).abs().max().item())
        print(tag, mode, "loss", out[f"{mode}_loss"], "ref fp32-vs-fp64 rel", out[f"{mode}_ref_fp32_vs_fp64_rel"],
              "grad gap", out.get(f"{mode}_ref_grad_fp32_vs_fp64_worst"), out.get(f"{mode}_ref_gradnorm_fp32_vs_fp64_worst"))
    np.savez_compressed(os.path.join(HERE, f"model_{tag}.npz"), **out)
    print("wrote model", tag)


if __name__ == "__main__":
    only = sys.argv[1:]
    for case in OP_CASES:
        if not only or "ops" in only:
            make_op(*case)
    if not only or "ops" in only:
        # LoGo-style grouped statistics: 4 image blocks, one BN call each
        make_op("nopos_c16_h_groups4", "AxialAttention_wopos", 16, 8, 8, 8, False, 1, ngroups=4)
        make_op("pos_c16_w_s2_groups2", "AxialAttention", 16, 4, 8, 8, True, 2, ngroups=2)
    if not only or "models" in only:
        make_model("gated_128", "gated", 128, 4)
        make_model("MedT_128", "MedT", 128, 4)
        make_model("axialunet_128", "axialunet", 128, 2)
        make_model("logo_128", "logo", 128, 2)
        make_model("MedT_256", "MedT", 256, 1, modes=("eval",), with_grads=False)
        make_model("MedT_256_train", "MedT", 256, 2, modes=("train",))      # BASELINE config 3 in train mode (round 2)


def make_schema():
    """state_dict schema + init checksums of the reference factories (drop-in contract)."""
    import json
    out = {}
    for factory, S in (("gated", 128), ("axialunet", 128), ("MedT", 128), ("logo", 128), ("gated", 64), ("MedT", 256)):
        torch.manual_seed(1234)
        m = getattr(ref, factory)(img_size=S, imgchan=3)
        sd = m.state_dict()
        out[f"{factory}_{S}"] = {
            "keys": [[k, list(v.shape), str(v.dtype)] for k, v in sd.items()],
            "init_sum": {k: float(v.double().sum()) for k, v in sd.items() if v.dtype.is_floating_point},
            "requires_grad_false": [k for k, p in m.named_parameters() if not p.requires_grad],
            "trainable": int(sum(p.numel() for p in m.parameters() if p.requires_grad)),
        }
    with open(os.path.join(HERE, "schema.json"), "w") as f:
        json.dump(out, f)
    print("wrote schema")


if __name__ == "__main__" and (not sys.argv[1:] or "schema" in sys.argv[1:]):
    make_schema()
