#!/usr/bin/env python
"""Test helper: executes the UNMODIFIED reference (baseline/_ref, copied there by baseline/make_ref.py) in its OWN process,
because its package is called `lib` like the drop-in's.  Never imported by the product.

    ref_runner.py eval_logits <modelname> <imgsize> <ckpt.pth> <x.npy> <out.npy> [device]   eval-mode logits of the reference
    ref_runner.py load_strict <modelname> <imgsize> <ckpt.pth>                             strict load into the reference
    ref_runner.py write_ckpt  <modelname> <imgsize> <seed> <out.pth>                       a reference-written state_dict
"""
import importlib.util
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(REPO, "baseline", "_ref")


def ref_lib():
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != os.path.join(REPO, "medical-transformer_b200")]
    sys.path.insert(0, REF)
    for mod in ("matplotlib", "skimage"):
        if importlib.util.find_spec(mod) is None:
            sys.path.append(os.path.join(REPO, "medical-transformer_b200", "stubs"))
            break
    import lib
    assert os.path.abspath(lib.__file__).startswith(REF), lib.__file__
    return lib


def build(lib, modelname, imgsize):
    f = {"axialunet": lib.models.axialunet, "MedT": lib.models.axialnet.MedT, "gatedaxialunet": lib.models.axialnet.gated,
         "logo": lib.models.axialnet.logo}[modelname]
    return f(img_size=imgsize, imgchan=3)


def main():
    import numpy as np
    import torch
    cmd = sys.argv[1]
    lib = ref_lib()
    if cmd == "eval_logits":
        modelname, imgsize, ckpt, xin, out = sys.argv[2:7]
        dev = sys.argv[7] if len(sys.argv) > 7 else "cpu"
        torch.backends.cudnn.allow_tf32 = False
        torch.backends.cuda.matmul.allow_tf32 = False
        m = build(lib, modelname, int(imgsize))
        sd = torch.load(ckpt, map_location="cpu")
        if all(k.startswith("module.") for k in sd):
            sd = {k[7:]: v for k, v in sd.items()}
        m.load_state_dict(sd, strict=True)
        m.to(dev).eval()
        with torch.no_grad():
            lg = m(torch.from_numpy(np.load(xin)).to(dev))
        np.save(out, lg.cpu().numpy())
    elif cmd == "load_strict":
        modelname, imgsize, ckpt = sys.argv[2:5]
        m = build(lib, modelname, int(imgsize))
        r = m.load_state_dict(torch.load(ckpt, map_location="cpu"), strict=True)
        print("load_strict ok", r)
    elif cmd == "write_ckpt":
        modelname, imgsize, seed, out = sys.argv[2:6]
        torch.manual_seed(int(seed))
        m = build(lib, modelname, int(imgsize))
        with torch.no_grad():                       # move the BN buffers / gates off their init values so the eval path is exercised
            for k, v in m.state_dict().items():
                if k.endswith("running_mean"):
                    v.normal_(0, 0.1)
                elif k.endswith("running_var"):
                    v.uniform_(0.5, 1.5)
                elif k.split(".")[-1] in ("f_qr", "f_kr", "f_sv", "f_sve"):
                    v.uniform_(0.05, 1.0)
        torch.save(m.state_dict(), out)
    else:
        raise SystemExit(__doc__)


if __name__ == "__main__":
    main()
