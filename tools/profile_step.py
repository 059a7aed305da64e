"""One eager training step of a headline model between cudaProfilerStart/Stop, for
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv ...
(the launch list committed under profiles/).  Eager == the same kernels the CUDA graph replays."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch  # noqa: E402

import bench  # noqa: E402
from medt_b200.train_step import TrainStep  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "gated"
dev = torch.device("cuda", 0)
model = bench.build_model(name).to(dev).train()
ts = TrainStep(model, (4, 3, 128, 128), device=dev, use_graph=False)
x, y = bench.synthetic_batch()
ts.x.copy_(x.to(dev)); ts.y.copy_(y.to(dev))
for _ in range(3):
    ts.step_resident()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
ts.step_resident()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one", name, "step; loss", float(ts.loss))
