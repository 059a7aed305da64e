"""The reference's training loop (train.py:132-217) on the B200 path: same epoch structure, same epoch-10 gate un-freeze,
same checkpoint files -- with the per-step body replaced by one CUDA-graph replay (`TrainStep`) fed by `PrefetchLoader`.

    fit(model, dataloader, epochs=400, lr=1e-3, direc="out/", modelname="gatedaxialunet", save_freq=10, valloader=valloader)

What is kept from train.py: Adam(lr, weight_decay=1e-5) over all parameters (:111-112), `loss = LogNLLLoss` == cross-entropy
(:156, metrics.py:17-20), the printed `epoch [e/E], loss:x` line (:163-165), `requires_grad = True` for every parameter after
epoch 10 (:169-171; the Adam state of the other parameters survives, `TrainStep.recapture()`), every `save_freq` epochs a
validation pass that writes channel 1 of the thresholded TRAIN-mode output of each validation image as a PNG (:172-213; the
reference never calls model.eval() in train.py) and `torch.save(model.state_dict())` to `<direc>/<epoch>/<modelname>.pth`
and `<direc>final_model.pth` (:215-217, string concatenation as in the reference).
What is dropped: the per-step `.cpu().numpy()` thresholding of the training output whose result is never used (:142-153)."""
from __future__ import annotations

import os

import torch

from .infer import logits_to_mask
from .loader import PrefetchLoader
from .train_step import TrainStep


def fit(model, dataloader, epochs, lr=1e-3, weight_decay=1e-5, direc="./medt", modelname="MedT", save_freq=10, valloader=None,
        batch_shape=None, device="cuda", log=print, unfreeze_epoch=10, prefetch_depth=3):
    device = torch.device(device)
    model.to(device)
    if batch_shape is None:
        x0, y0, *_ = next(iter(dataloader))
        batch_shape = tuple(x0.shape)
    ts = TrainStep(model, batch_shape, lr=lr, weight_decay=weight_decay, device=device)
    feed = PrefetchLoader(dataloader, batch_shape, depth=prefetch_depth, device=device)
    history = []
    for epoch in range(epochs):
        slots = []
        total, nb = 0.0, 0
        for x, y, _rest in feed:
            slots.append(ts.step(x, y))
            if len(slots) == ts.loss_pin.numel():           # drain the asynchronous loss ring before it wraps
                torch.cuda.current_stream(device).synchronize()
                total += float(ts.loss_pin[slots].sum()); nb += len(slots); slots = []
        torch.cuda.current_stream(device).synchronize()
        if slots:
            total += float(ts.loss_pin[slots].sum()); nb += len(slots)
        mean = total / max(nb, 1)
        history.append(mean)
        log('epoch [{}/{}], loss:{:.4f}'.format(epoch, epochs, mean))
        if epoch == unfreeze_epoch:
            for p in model.parameters():
                p.requires_grad = True
            ts.recapture()
        if (epoch % save_freq) == 0:
            fulldir = direc + "/{}/".format(epoch)
            os.makedirs(fulldir, exist_ok=True)
            if valloader is not None:
                import cv2
                bns = [m for m in model.modules() if getattr(m, "_nbt_deferred", False)]
                for b in bns:                               # outside the captured step the counters advance per call again
                    b._nbt_deferred = False
                with torch.no_grad():                       # train mode, batch 1: exactly what train.py:176-213 runs
                    for batch_idx, (xv, _yv, *rest) in enumerate(valloader):
                        name = rest[0][0] if (rest and isinstance(rest[0][0], str)) else '%s.png' % str(batch_idx + 1).zfill(3)
                        mask = logits_to_mask(model(xv.to(device)))
                        cv2.imwrite(fulldir + name, mask[0].cpu().numpy())
                for b in bns:
                    b._nbt_deferred = True
            torch.save(model.state_dict(), fulldir + modelname + ".pth")
            torch.save(model.state_dict(), direc + "final_model.pth")
    return history
