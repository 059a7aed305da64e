import torch.nn as nn


class qkv_transform(nn.Conv1d):
    """1x1 Conv1d holding the fused q/k/v projection weight `(2C, C, 1)`; the class name is
    part of the checkpoint contract (reference lib/models/utils.py:4-5).  Its weight is
    consumed directly by the CUDA kernel; the module's own forward is never used."""
