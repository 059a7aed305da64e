"""BASELINE config 5 flavour: logo at 512 px with the generalised patch size (L = 256 strips), and
gated at 512 px; prints fwd+bwd step times (eager, CUDA events) and peak memory."""
import os, sys
REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [REPO, os.path.join(REPO, "medical-transformer_b200")]
import torch
import lib
dev = torch.device("cuda", 0)
for name, kw, n in (("logo", dict(img_size=512, patch_size=128), 1), ("gated", dict(img_size=512), 4), ("MedT", dict(img_size=256), 4)):
    torch.manual_seed(3000)
    f = getattr(lib.models.axialnet, name)
    m = f(imgchan=3, **kw).to(dev).train()
    opt = torch.optim.Adam(m.parameters(), lr=1e-3, weight_decay=1e-5, fused=True)
    s = kw["img_size"]
    x = torch.rand(n, 3, s, s, device=dev); y = torch.randint(0, 2, (n, s, s), device=dev)
    def step():
        opt.zero_grad(set_to_none=True)
        loss = torch.nn.functional.cross_entropy(m(x), y)
        loss.backward(); opt.step()
        return loss
    for _ in range(3): step()
    torch.cuda.synchronize(); torch.cuda.reset_peak_memory_stats()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): loss = step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(f"{name} {kw} N={n}: {ms:.2f} ms/step eager -> {n / ms * 1e3:.1f} img/s, loss {loss.item():.4f}, peak mem {torch.cuda.max_memory_allocated() / 2**20:.0f} MiB", flush=True)
