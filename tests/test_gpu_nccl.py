"""GPU tier, two B200s, one process per GPU over NCCL: medt_b200.TrainStep with the flat gradient bucket all-reduced
inside the captured step -- the default single all-reduce after backward and the opt-in EarlyBucket overlap -- keeps the replicas
bit-identical while the ranks see different data shards, and BatchNorm statistics stay per rank (the reference's nn.DataParallel
semantics, train.py:104-108).  Skipped on a one-GPU box; bench.py asserts the same lock step at N = 2, 4, 8."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, overlap, q):
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (os.path.join(here, "tests", "golden"), here, os.path.join(here, "medical-transformer_b200")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    if overlap:
        os.environ["MEDT_OVERLAP_REDUCE"] = "1"
    else:
        os.environ.pop("MEDT_OVERLAP_REDUCE", None)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import lib
        from medt_b200.train_step import TrainStep
        torch.manual_seed(7)                                           # identical replicas
        m = lib.models.axialnet.gated(img_size=64, imgchan=3).to(torch.device("cuda", rank)).train()
        ts = TrainStep(m, (4, 3, 64, 64), device=torch.device("cuda", rank), use_graph=True)
        assert ts.world == world and (ts._early is not None) == bool(overlap)
        g = torch.Generator().manual_seed(100 + rank)                  # different data shard per rank
        losses = []
        for _ in range(3):
            x = torch.rand(4, 3, 64, 64, generator=g).pin_memory()
            y = torch.randint(0, 2, (4, 64, 64), generator=g).pin_memory()
            slot = ts.step(x, y)
            torch.cuda.synchronize()
            losses.append(float(ts.loss_pin[slot]))
        assert all(l == l and l < 10 for l in losses), losses
        if overlap:
            assert ts._early.fired                                     # the first bucket was reduced from inside backward
        flat = torch.cat([p.detach().reshape(-1) for p in m.parameters()])
        both = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])                           # replicas in lock step, bit for bit
        rm = m.bn1.running_mean.clone()
        rms = [torch.zeros_like(rm) for _ in range(world)]
        dist.all_gather(rms, rm)
        assert not torch.equal(rms[0], rms[1])                         # BN statistics stay per rank (no SyncBN in the reference)
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-800:]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
@pytest.mark.parametrize("overlap", [0, 1])
def test_trainstep_two_ranks_nccl_lock_step(overlap):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, overlap, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=600) for _ in procs]
    for p in procs:
        p.join(timeout=120)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res
