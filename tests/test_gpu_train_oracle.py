"""GPU tier, SURVEY 8(f)-3: the captured training step (forward + cross-entropy + backward + fused flat Adam in one CUDA graph)
against the REFERENCE's own optimizer steps (tests/golden/adam2_*.npz: two steps of train.py:111-112,156-161 executed with the
reference in fp64 by make_golden_r2.py, plus the reference's own fp32-vs-fp64 gap of every quantity as the noise floor).

Adam's first updates are ~lr*sign(g): parameters whose gradient is near zero flip sign under any rounding, so per-element
agreement is bounded by the reference's own fp32 behaviour, which the fixture records (cosine of the update vectors, median
per-tensor gap, loss trajectory)."""
import os

import numpy as np
import pytest
import torch

import fill
from helpers import GOLDEN, build_filled_model

pytestmark = pytest.mark.gpu
GATES = ("f_qr", "f_kr", "f_sv", "f_sve")


def _run(factory, S, N, steps, use_graph, flat):
    from medt_b200.train_step import TrainStep
    m = build_filled_model(factory, S, device="cuda")
    for n, p in m.named_parameters():                     # the reference's construction-time state (axialnet.py:124-127)
        if n.split(".")[-1] in GATES:
            p.requires_grad_(False)
    w0 = {k: p.detach().clone() for k, p in m.named_parameters()}
    ts = TrainStep(m, (N, 3, S, S), lr=1e-3, weight_decay=1e-5, use_graph=use_graph, flat_adam=flat)
    x = torch.tensor(fill.image_batch(N, 3, S), dtype=torch.float32)
    y = torch.tensor(fill.mask_batch(N, S))
    losses = []
    for _ in range(steps):
        slot = ts.step(x, y)
        torch.cuda.synchronize()
        losses.append(float(ts.loss_pin[slot]))
    with torch.no_grad():
        m.train()
        losses.append(float(torch.nn.functional.cross_entropy(m(x.cuda()), y.cuda())))
    delta = {k: (p.detach() - w0[k]).double().cpu() for k, p in m.named_parameters()}
    return m, losses, delta


@pytest.mark.parametrize("factory,S,N", [("gated", 64, 4), ("MedT", 128, 2)])
@pytest.mark.parametrize("use_graph,flat", [(True, True), (False, False)])
def test_two_optimizer_steps_track_the_reference(factory, S, N, use_graph, flat):
    z = np.load(os.path.join(GOLDEN, f"adam2_{factory}_{S}.npz"))
    assert [int(v) for v in z["meta"]] == [S, N, 2]
    m, losses, delta = _run(factory, S, N, 2, use_graph, flat)
    ref, ref32 = z["losses"], z["losses_fp32"]
    # step-0 loss is pure forward parity; later ones are held to 3x the reference's own fp32-vs-fp64 drift (+1e-3)
    assert abs(losses[0] - ref[0]) < 1e-3, (losses, ref)
    for i in (1, 2):
        assert abs(losses[i] - ref[i]) < 3 * abs(ref32[i] - ref[i]) + 1e-3 * max(1.0, abs(ref[i])), (i, losses, ref, ref32)
    # every parameter the reference updated is updated here, by the same amount (update norm), and nothing else moves
    names = z["names"].tolist()
    want = dict(zip(names, z["delta_norms"].tolist()))
    rel = []
    for k in names:
        got = float(delta[k].norm())
        if want[k] == 0.0:
            assert got == 0.0, k                          # frozen gates, never-used adjust_p / wopos conv1: untouched
        else:
            rel.append(abs(got - want[k]) / want[k])
    assert np.median(rel) < 0.05 and max(rel) < 1.0, (np.median(rel), max(rel))
    # direction of the update: cosine against the fp64 reference, bounded by what the reference's own fp32 run achieves
    cos = lambda a, b: float((a * b).sum() / (np.linalg.norm(a) * np.linalg.norm(b)))  # noqa: E731
    worst = min(cos(delta[key[6:]].numpy(), z[key]) for key in z.files if key.startswith("delta."))
    floor = float(z["ref_select_cos_min"])
    assert worst > min(0.9, floor - 3 * (1 - floor)), (worst, floor)
    big = [k for k in names if delta[k].numel() >= 256]          # whole-model update direction (all tensors of >= 256 elements)
    refd = np.load(os.path.join(GOLDEN, f"adam2_{factory}_{S}.npz"))
    tot_floor = float(refd["ref_total_cos"])
    tot = cos(torch.cat([delta[k].reshape(-1) for k in big])[::8].numpy(), refd["delta_all_sub8_over_lr"].astype(np.float64))
    assert tot > tot_floor - 3 * (1 - tot_floor), (tot, tot_floor)
    # running statistics after two steps
    bufs = dict(m.named_buffers())
    for key in z.files:
        if key.startswith("buf."):
            b = bufs[key[4:]].double().cpu().numpy()
            # the second step's statistics see weights that already moved by ~lr in rounding-dependent directions
            assert np.linalg.norm(b - z[key]) <= 1e-2 * max(np.linalg.norm(z[key]), 1e-3), key


def test_prefetch_loader_feeds_train_step_without_corruption():
    """PrefetchLoader -> TrainStep.step on pinned ring slots: every step must see exactly its own batch (the per-step loss of a
    frozen-weight model identifies the batch), with the host running ahead of the GPU and the ring wrapping several times."""
    from medt_b200.loader import PrefetchLoader
    from medt_b200.train_step import TrainStep
    S, N, steps = 64, 2, 24
    rs = np.random.RandomState(0)
    batches = [(torch.tensor(rs.rand(N, 3, S, S), dtype=torch.float32), torch.tensor(rs.randint(0, 2, (N, S, S)))) for _ in range(steps)]
    batches.append((torch.zeros(1, 3, S, S), torch.zeros(1, S, S, dtype=torch.long)))       # ragged tail: dropped
    m = build_filled_model("gated", S, device="cuda")
    ts = TrainStep(m, (N, 3, S, S), lr=0.0, weight_decay=0.0)       # lr 0: weights frozen, BN batch statistics -> loss is a pure function of the batch
    want = []
    for x, y in batches[:steps]:
        slot = ts.step(x.pin_memory(), y.pin_memory())
        torch.cuda.synchronize()
        want.append(float(ts.loss_pin[slot]))
    feed = PrefetchLoader(batches, (N, 3, S, S), depth=3)
    slots = [ts.step(x, y) for x, y, _ in feed]                     # no synchronisation inside the loop
    torch.cuda.synchronize()
    got = [float(ts.loss_pin[s]) for s in slots]
    assert feed.batches_dropped == 1 and len(got) == steps
    assert np.allclose(got, want, rtol=2e-4, atol=1e-5), (got, want)
    assert len(set(np.round(want, 5))) > steps // 2                 # the batches are distinguishable


def test_fit_loop_matches_train_py_structure(tmp_path):
    """medt_b200.loop.fit: epoch losses, epoch-10-style un-freeze (here after epoch 0), checkpoint files named as train.py:215-217."""
    from medt_b200.loop import fit
    S, N = 64, 2
    rs = np.random.RandomState(1)
    data = [(torch.tensor(rs.rand(N, 3, S, S), dtype=torch.float32), torch.tensor(rs.randint(0, 2, (N, S, S))), [f"{i:03d}.png"] * N) for i in range(4)]
    val = [(d[0][:1], d[1][:1], [d[2][0]]) for d in data[:2]]
    m = build_filled_model("gated", S, device="cuda")
    for n, p in m.named_parameters():
        if n.split(".")[-1] in GATES:
            p.requires_grad_(False)
    g0 = m.layer1[0].hight_block.f_qr.detach().clone()
    lines = []
    direc = str(tmp_path / "out")
    hist = fit(m, data, epochs=2, lr=1e-3, direc=direc, modelname="gatedaxialunet", save_freq=1, valloader=val, batch_shape=(N, 3, S, S),
               log=lines.append, unfreeze_epoch=0)
    assert len(hist) == 2 and all(np.isfinite(hist)) and lines[0].startswith("epoch [0/2], loss:")
    assert os.path.isfile(direc + "final_model.pth") and os.path.isfile(os.path.join(direc, "1", "gatedaxialunet.pth"))
    assert sorted(os.listdir(os.path.join(direc, "0"))) == ["000.png", "001.png", "gatedaxialunet.pth"]
    sd = torch.load(direc + "final_model.pth", map_location="cpu")
    assert list(sd.keys()) == list(m.state_dict().keys())
    assert not torch.equal(m.layer1[0].hight_block.f_qr.detach().cpu(), g0.cpu())      # gates train after the un-freeze
