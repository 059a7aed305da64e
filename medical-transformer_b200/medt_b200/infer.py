"""Inference side of the path: what the reference's test.py:105-146 and the validation block of
train.py:176-213 do per batch -- eval-mode forward, threshold, write channel 1 as a 0/255 PNG -- without the two host
round trips per image (`.cpu().numpy()` of the full logits, thresholding in numpy).

  * `logits_to_mask`   the reference's thresholding, on whatever device the logits live on;
  * `Predictor`        pinned-host -> device copy, eval forward through the sm_100a kernels, threshold on the GPU, one
                       N*H*W-byte device -> pinned-host copy of the mask (optionally the whole thing as one CUDA graph);
  * `load_checkpoint`  accepts checkpoints written with or without nn.DataParallel's `module.` prefix
                       (train.py:104-108 wraps the model when more than one GPU is visible, test.py:100-106 loads either).
"""
from __future__ import annotations

import os

import torch


def logits_to_mask(logits: torch.Tensor) -> torch.Tensor:
    """(N, 2, H, W) float logits -> (N, H, W) uint8 in {0, 255}: the reference thresholds the RAW logits at 0.5 and
    writes channel 1 (`tmp[tmp>=0.5] = 1 ... yHaT[yHaT==1] = 255; cv2.imwrite(..., yHaT[0,1,:,:])`, test.py:124-146)."""
    return (logits[:, 1] >= 0.5).to(torch.uint8) * 255


def load_checkpoint(model: torch.nn.Module, path_or_state, strict: bool = True):
    """load_state_dict that tolerates the `module.` prefix in either direction."""
    state = torch.load(path_or_state, map_location="cpu") if isinstance(path_or_state, (str, os.PathLike)) else path_or_state
    want_prefix = isinstance(model, torch.nn.DataParallel)
    has_prefix = len(state) > 0 and all(k.startswith("module.") for k in state)
    if has_prefix and not want_prefix:
        state = {k[len("module."):]: v for k, v in state.items()}
    elif want_prefix and not has_prefix:
        state = {"module." + k: v for k, v in state.items()}
    return model.load_state_dict(state, strict=strict)


class Predictor:
    def __init__(self, model: torch.nn.Module, batch_shape, device=None, use_graph: bool = False):
        """`use_graph=True` captures copy-in-place forward + threshold once and replays it (the 128-px forward is ~130
        launches of a few microseconds); eager is the default."""
        self.model = model.eval()
        self.device = torch.device(device if device is not None else next(model.parameters()).device)
        if self.device.type != "cuda":
            raise RuntimeError("medt_b200.Predictor needs a CUDA device; there is no CPU path")
        n, c, h, w = batch_shape
        self.x = torch.zeros(batch_shape, device=self.device)
        self.mask = torch.zeros((n, h, w), dtype=torch.uint8, device=self.device)
        self.x_pin = torch.zeros(batch_shape).pin_memory()
        self.mask_pin = torch.zeros((n, h, w), dtype=torch.uint8).pin_memory()
        self.graph = None
        if use_graph:
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream(self.device))
            with torch.cuda.stream(side):                 # warm-up outside capture (allocator, cuDNN algorithm choice)
                for _ in range(3):
                    self._body()
            torch.cuda.current_stream(self.device).wait_stream(side)
            torch.cuda.synchronize(self.device)
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._body()

    @torch.no_grad()
    def _body(self):
        self.mask.copy_(logits_to_mask(self.model(self.x)))

    @torch.no_grad()
    def predict(self, x_host: torch.Tensor) -> torch.Tensor:
        """x_host: (N, C, H, W) float CPU tensor (pinned or not).  Returns the pinned (N, H, W) uint8 mask buffer,
        valid after this call returns (one stream synchronisation at the end, none in between)."""
        if x_host.is_pinned():
            self.x.copy_(x_host, non_blocking=True)
        else:
            self.x_pin.copy_(x_host)
            self.x.copy_(self.x_pin, non_blocking=True)
        if self.graph is not None:
            self.graph.replay()
        else:
            self._body()
        self.mask_pin.copy_(self.mask, non_blocking=True)
        torch.cuda.current_stream(self.device).synchronize()
        return self.mask_pin

    def save_png(self, masks: torch.Tensor, filenames, direc: str):
        """cv2.imwrite of each mask, as the reference does for element 0 of every batch (test.py:140-146)."""
        import cv2
        os.makedirs(direc, exist_ok=True)
        arr = masks.cpu().numpy()
        for i, name in enumerate(filenames):
            cv2.imwrite(os.path.join(direc, name), arr[i])
