import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch, lib
dev = torch.device("cuda", 0)
torch.manual_seed(3000)
m = lib.models.axialnet.gated(img_size=512, imgchan=3).to(dev).train()
x = torch.rand(4, 3, 512, 512, device=dev); y = torch.randint(0, 2, (4, 512, 512), device=dev)
def step():
    for p in m.parameters(): p.grad = None
    torch.nn.functional.cross_entropy(m(x), y).backward()
for _ in range(2): step()
torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStart(); step(); torch.cuda.synchronize(); torch.cuda.cudart().cudaProfilerStop()
