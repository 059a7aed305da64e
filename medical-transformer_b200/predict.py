#!/usr/bin/env python
"""Inference entry point with the command line of the reference's test.py (test.py:28-57) on the B200 path.

    python medical-transformer_b200/predict.py --loaddirec out/final_model.pth --val_dataset data/val --direc results \
        --batch_size 1 --modelname gatedaxialunet --imgsize 128 --gray no

The reference's own test.py cannot run as published (`aug = args.aug` at test.py:62 reads an option its parser never
defines -> AttributeError before the model is built), so unlike train.py it cannot be executed unchanged through
run_reference_script.py; this script does what its body does (test.py:86-146): build the model by name, load the checkpoint
(with or without DataParallel's `module.` prefix), eval mode, for every image of `<val_dataset>/img` threshold the raw
logits at 0.5 and write channel 1 as a 0/255 PNG named like the input into `<direc>/`.  Images are read the way the
reference's dataset does (utils.py:144-160: cv2.imread BGR, ToTensor scaling); masks in `labelcol/` are not needed.
"""
import argparse
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
if HERE not in sys.path:
    sys.path.insert(0, HERE)


def parse(argv=None):
    ap = argparse.ArgumentParser(description="MedT inference (B200 path)")
    ap.add_argument('-j', '--workers', default=16, type=int)
    ap.add_argument('--epochs', default=100, type=int)
    ap.add_argument('-b', '--batch_size', default=1, type=int)
    ap.add_argument('--learning_rate', default=1e-3, type=float)
    ap.add_argument('--train_dataset', type=str)
    ap.add_argument('--val_dataset', type=str, required=True)
    ap.add_argument('--save_freq', type=int, default=5)
    ap.add_argument('--modelname', default='MedT', type=str)
    ap.add_argument('--cuda', default="on", type=str)
    ap.add_argument('--direc', default='./results', type=str)
    ap.add_argument('--crop', type=int, default=None)
    ap.add_argument('--device', default='cuda', type=str)
    ap.add_argument('--loaddirec', default='load', type=str)
    ap.add_argument('--imgsize', type=int, default=None)
    ap.add_argument('--gray', default='no', type=str)
    ap.add_argument('--graph', action='store_true', help="replay the eval forward + threshold as one CUDA graph")
    return ap.parse_args(argv)


def build(modelname, imgsize, imgchan):
    import lib
    if modelname == "axialunet":
        return lib.models.axialunet(img_size=imgsize, imgchan=imgchan)
    if modelname == "MedT":
        return lib.models.axialnet.MedT(img_size=imgsize, imgchan=imgchan)
    if modelname == "gatedaxialunet":
        return lib.models.axialnet.gated(img_size=imgsize, imgchan=imgchan)
    if modelname == "logo":
        return lib.models.axialnet.logo(img_size=imgsize, imgchan=imgchan)
    raise SystemExit(f"unknown --modelname {modelname!r} (axialunet | MedT | gatedaxialunet | logo)")


def main(argv=None):
    import cv2
    import numpy as np
    import torch
    from medt_b200.infer import Predictor, load_checkpoint
    args = parse(argv)
    gray = args.gray == "yes"
    model = build(args.modelname, args.imgsize, 1 if gray else 3).to("cuda")
    load_checkpoint(model, args.loaddirec)
    src = os.path.join(args.val_dataset, 'img')
    names = sorted(os.listdir(src))
    bs = max(1, args.batch_size)
    pred = None
    written = 0
    for i in range(0, len(names), bs):
        chunk = names[i:i + bs]
        imgs = []
        for n in chunk:
            im = cv2.imread(os.path.join(src, n), 0 if gray else 1)
            if im is None:
                raise SystemExit(f"cannot read {os.path.join(src, n)}")
            if im.ndim == 2:
                im = im[:, :, None]
            imgs.append(torch.from_numpy(np.ascontiguousarray(im.transpose(2, 0, 1))).float().div(255.0))
        x = torch.stack(imgs)
        if pred is None or tuple(pred.x.shape) != tuple(x.shape):
            pred = Predictor(model, tuple(x.shape), use_graph=args.graph)
        masks = pred.predict(x)
        pred.save_png(masks, chunk, args.direc)
        written += len(chunk)
    print(f"wrote {written} masks to {args.direc}")
    return written


if __name__ == "__main__":
    main()
