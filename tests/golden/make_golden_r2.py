"""Round-2 fixtures, again produced by EXECUTING THE REFERENCE ITSELF here (never on the GPU box):

    python tests/golden/make_golden_r2.py [flips] [adam] [logo512]

  parity_r2.npz   for every model fixture of make_golden.py: the reference's OWN fp32-vs-fp64 mask flip rates over ALL
                  pixels (threshold mask `logit_1 >= 0.5`, train.py:143-145, and argmax) -- the bound the CUDA path's
                  all-pixel mismatch is held to ("masks are not asserted bit-exact")
  adam2_gated_64.npz / adam2_MedT_128.npz
                  two optimizer steps of the reference's training loop (train.py:111-112,156-161: Adam lr 1e-3, wd 1e-5,
                  frozen gates) in fp64: losses, every parameter's update norm, a few updated tensors; plus the
                  reference's fp32-vs-fp64 gap of the same quantities as the noise floor
  logo512_p128.npz  the LoGo patch generalisation at 512 px / patch 128: the reference cannot run this
                  (it hard-codes 32-px patches and crashes), so this one comes from the oracle's patch loop with the constant
                  substituted (oracle/torch_port.py, itself pinned at 128 px) in fp64: sub-sampled logits, loss, gradient
                  norms, selected gradients.
The reference sources are not copied anywhere."""
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
from make_golden import _fill_module  # noqa: E402

torch.set_num_threads(8)
GATES = ("f_qr", "f_kr", "f_sv", "f_sve")


def flips():
    out = {}
    for tag, factory, S, N, modes in (("gated_128", "gated", 128, 4, ("train", "eval")), ("MedT_128", "MedT", 128, 4, ("train", "eval")),
                                      ("axialunet_128", "axialunet", 128, 2, ("train", "eval")), ("logo_128", "logo", 128, 2, ("train", "eval")),
                                      ("MedT_256", "MedT", 256, 1, ("eval",))):
        xnp = fill.image_batch(N, 3, S)
        for mode in modes:
            lg = {}
            for dtype in (torch.float64, torch.float32):
                torch.manual_seed(0)
                m = getattr(ref, factory)(img_size=S, imgchan=3).to(dtype)
                _fill_module(m, "", dtype)
                m.train(mode == "train")
                with torch.no_grad():
                    lg[dtype] = m(torch.tensor(xnp, dtype=dtype)).double().numpy()
            a, b = lg[torch.float64], lg[torch.float32]
            out[f"{tag}.{mode}.thr_flip_frac"] = np.array(float(((a[:, 1] >= 0.5) != (b[:, 1] >= 0.5)).mean()))
            out[f"{tag}.{mode}.argmax_flip_frac"] = np.array(float((a.argmax(1) != b.argmax(1)).mean()))
            out[f"{tag}.{mode}.npix"] = np.array(a[:, 1].size)
            print(tag, mode, "thr flips", out[f"{tag}.{mode}.thr_flip_frac"], "argmax flips", out[f"{tag}.{mode}.argmax_flip_frac"])
    np.savez_compressed(os.path.join(HERE, "parity_r2.npz"), **out)


SELECT = {"gated": ["layer1.0.hight_block.qkv_transform.weight", "layer1.0.hight_block.relative", "layer2.0.width_block.bn_similarity.weight",
                    "layer3.1.hight_block.bn_qkv.bias", "conv1.weight", "adjust.bias"],
          "MedT": ["layer1.0.hight_block.relative", "layer1_p.0.hight_block.qkv_transform.weight", "layer2_p.0.width_block.bn_similarity.weight",
                   "conv1_p.weight", "decoderf.bias", "adjust.bias"]}


def adam2(factory, S, N, steps=2):
    xnp, ynp = fill.image_batch(N, 3, S), fill.mask_batch(N, S)
    res = {}
    for dtype in (torch.float64, torch.float32):
        torch.manual_seed(0)
        m = getattr(ref, factory)(img_size=S, imgchan=3).to(dtype)
        _fill_module(m, "", dtype)
        for n, p in m.named_parameters():                 # the reference's construction-time state: frozen gates
            if n.split(".")[-1] in GATES:
                p.requires_grad_(False)
        m.train()
        w0 = {k: p.detach().clone() for k, p in m.named_parameters()}
        opt = torch.optim.Adam(list(m.parameters()), lr=1e-3, weight_decay=1e-5)      # train.py:111-112
        x, y = torch.tensor(xnp, dtype=dtype), torch.tensor(ynp)
        losses = []
        for _ in range(steps):                                                         # train.py:156-161
            out = m(x)
            loss = torch.nn.functional.cross_entropy(out, y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            losses.append(loss.item())
        with torch.no_grad():
            losses.append(torch.nn.functional.cross_entropy(m(x), y).item())          # loss after the last update (train mode)
        res[dtype] = dict(losses=np.array(losses), delta={k: (p.detach() - w0[k]).double() for k, p in m.named_parameters()},
                          w={k: p.detach().double().numpy() for k, p in m.named_parameters()},
                          bufs={k: b.detach().double().numpy() for k, b in m.named_buffers() if "running" in k})
    r64, r32 = res[torch.float64], res[torch.float32]
    names = list(r64["delta"].keys())
    out = {"meta": np.array([S, N, steps]), "losses": r64["losses"], "losses_fp32": r32["losses"],
           "names": np.array(names), "delta_norms": np.array([float(r64["delta"][k].norm()) for k in names]),
           "ref_delta_fp32_vs_fp64_worst": np.array(max(float((r32["delta"][k] - r64["delta"][k]).norm()) /
                                                        max(float(r64["delta"][k].norm()), 1e-12) for k in names
                                                        if float(r64["delta"][k].norm()) > 0))}
    gaps = [float((r32["delta"][k] - r64["delta"][k]).norm()) / float(r64["delta"][k].norm()) for k in names
            if float(r64["delta"][k].norm()) > 0]
    out["ref_delta_fp32_vs_fp64_median"] = np.array(float(np.median(gaps)))
    cos = lambda a, b: float((a * b).sum() / (a.norm() * b.norm()))  # noqa: E731
    out["ref_select_cos_min"] = np.array(min(cos(r32["delta"][k], r64["delta"][k]) for k in SELECT[factory]))
    big = [k for k in names if r64["delta"][k].numel() >= 256]
    a32, a64 = (torch.cat([r[k].reshape(-1) for k in big])[::8] for r in (r32["delta"], r64["delta"]))      # every 8th element
    out["ref_total_cos"] = np.array(cos(a32, a64))
    out["delta_all_sub8_over_lr"] = (a64 / 1e-3).numpy().astype(np.float16)
    for k in SELECT[factory]:
        out["w." + k] = r64["w"][k]
        out["delta." + k] = r64["delta"][k].numpy()
    for k in ("bn1.running_mean", "layer1.0.hight_block.bn_similarity.running_var", "layer1.0.hight_block.bn_output.running_var"):
        out["buf." + k] = r64["bufs"][k]
    np.savez_compressed(os.path.join(HERE, f"adam2_{factory}_{S}.npz"), **out)
    print("adam2", factory, "losses", r64["losses"], "fp32", r32["losses"], "delta gap", out["ref_delta_fp32_vs_fp64_worst"])


def logo512():
    sys.path.insert(0, REPO)
    sys.path.insert(0, os.path.join(REPO, "medical-transformer_b200"))
    from oracle import torch_port
    # the drop-in's module tree (identical parameter names / shapes to the reference's, schema-pinned) carries patch_size
    for k in [k for k in sys.modules if k == "lib" or k.startswith("lib.")]:
        del sys.modules[k]
    sys.path.remove("/root/reference")
    import lib as dropin
    S, P, N = 512, 128, 1
    torch.manual_seed(0)
    m = dropin.models.axialnet.logo(img_size=S, imgchan=3, patch_size=P).double()
    vals = fill.fill_state_dict(m.state_dict())
    with torch.no_grad():
        for k, v in m.state_dict().items():
            if not k.endswith("flatten_index"):
                v.copy_(torch.as_tensor(vals[k]).to(v.dtype))
    m.train()
    x = torch.tensor(fill.image_batch(N, 3, S), dtype=torch.float64)
    y = torch.tensor(fill.mask_batch(N, S))
    lg = torch_port.forward(m, x, patch=P)
    loss = torch.nn.functional.cross_entropy(lg, y)
    loss.backward()
    names = [k for k, _ in m.named_parameters()]
    norms = [0.0 if p.grad is None else float(p.grad.norm()) for _, p in m.named_parameters()]
    out = {"meta": np.array([S, P, N]), "logits_sub4": lg.detach().numpy()[:, :, ::4, ::4].astype(np.float32), "loss": np.array(loss.item()),
           "names": np.array(names), "gradnorms": np.array(norms)}
    pd = dict(m.named_parameters())
    for k in ("conv1.weight", "conv1_p.weight", "layer1_p.0.hight_block.relative", "layer1.0.hight_block.relative", "decoderf.weight", "adjust.bias"):
        out["grad." + k] = pd[k].grad.numpy()
    np.savez_compressed(os.path.join(HERE, "logo512_p128.npz"), **out)
    print("logo512 loss", loss.item())


if __name__ == "__main__":
    what = sys.argv[1:] or ["flips", "adam", "logo512"]
    if "flips" in what:
        flips()
    if "adam" in what:
        adam2("gated", 64, 4)
        adam2("MedT", 128, 2)
    if "logo512" in what:
        logo512()
