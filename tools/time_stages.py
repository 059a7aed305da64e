"""Prints live per-kernel durations (CUDA events, cold + warm L2) of one axial op's 8 kernels."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch
import bench
dev = torch.device("cuda", 0)
m = bench.build_model("gated").to(dev).train()
fl = bench.L2Flusher(dev)
res, shp = bench.stage_profile(m, dev, fl)
print(os.environ.get("MT_DEBUG_FLAGS"), shp, {k: round(v, 1) for k, v in res.items()})
