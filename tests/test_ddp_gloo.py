"""CPU tier, world_size 2 over gloo: the host logic of the data-parallel step --
ONE flat gradient bucket, mean all-reduce, gradients handed back as views, parameters without
gradients left out -- followed by an optimizer step that must keep both ranks bit-identical."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from medt_b200.train_step import FlatBucket
        torch.manual_seed(0)                                   # identical replicas
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.BatchNorm2d(4), torch.nn.ReLU(),
                                  torch.nn.Conv2d(4, 2, 1))
        frozen = torch.nn.Parameter(torch.tensor(0.1), requires_grad=False)   # like the gates before epoch 10
        params = [p for p in list(net.parameters()) + [frozen] if p.requires_grad]
        opt = torch.optim.Adam(params, lr=1e-2, weight_decay=1e-5)
        bucket = FlatBucket(params)
        assert bucket.world == world and bucket.nbytes() == 4 * sum(p.numel() for p in params)
        torch.manual_seed(100 + rank)                          # different data shard per rank
        x, y = torch.rand(4, 3, 8, 8), torch.randint(0, 2, (4, 8, 8))
        for p in params:
            p.grad = None
        torch.nn.functional.cross_entropy(net(x), y).backward()
        local = [p.grad.clone() for p in params]
        bucket.reduce_()
        gathered = [torch.zeros_like(torch.cat([g.reshape(-1) for g in local])) for _ in range(world)]
        dist.all_gather(gathered, torch.cat([g.reshape(-1) for g in local]))
        mean = sum(gathered) / world
        assert torch.allclose(bucket.flat, mean, atol=1e-7)
        assert all(p.grad.data_ptr() == v.data_ptr() for p, v in zip(params, bucket.views))   # views, no copies
        opt.step()
        flat_p = torch.cat([p.detach().reshape(-1) for p in params])
        both = [torch.zeros_like(flat_p) for _ in range(world)]
        dist.all_gather(both, flat_p)
        assert torch.equal(both[0], both[1])                   # replicas stay in lock-step
        # BN statistics stay per rank (DataParallel semantics of the reference, no SyncBN)
        rm = net[1].running_mean.clone()
        rms = [torch.zeros_like(rm) for _ in range(world)]
        dist.all_gather(rms, rm)
        assert not torch.equal(rms[0], rms[1])
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def test_flat_bucket_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def _worker_flat(rank, world, port, q):
    """The flat-parameter variant of the step (train_step.FlatGroup): gather -> mean all-reduce of the flat gradient ->
    Adam on the flat tensor; must equal a per-parameter Adam fed the rank-averaged gradients."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import copy
        from medt_b200.train_step import FlatGroup, _all_reduce_mean_
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 2, 1))
        ref = copy.deepcopy(net)
        group = FlatGroup(list(net.parameters()))
        opt = torch.optim.Adam([group.flat], lr=1e-2, weight_decay=1e-5)
        opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-5)
        for it in range(3):
            torch.manual_seed(100 + 10 * it + rank)            # different data shard per rank
            x, y = torch.rand(4, 3, 8, 8), torch.randint(0, 2, (4, 8, 8))
            for p in net.parameters():
                p.grad = None
            torch.nn.functional.cross_entropy(net(x), y).backward()
            group.gather_grads_()
            _all_reduce_mean_(group.flat.grad, world, None)
            group.publish_grads_()
            opt.step()
            # reference: per-parameter Adam on gradients averaged over the ranks' shards
            opt_ref.zero_grad(set_to_none=True)
            torch.nn.functional.cross_entropy(ref(x), y).backward()
            for p in ref.parameters():
                dist.all_reduce(p.grad)
                p.grad.div_(world)
            opt_ref.step()
            for a, b in zip(net.parameters(), ref.parameters()):
                assert torch.allclose(a.detach(), b.detach(), atol=1e-7, rtol=1e-6), it
        flat = group.flat.detach().clone()
        both = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])                   # replicas stay in lock-step
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        q.put((rank, repr(e)))
    finally:
        dist.destroy_process_group()


def _worker_early(rank, world, port, q):
    """train_step.EarlyBucket: the first bucket (gradient-arrival prefix) is reduced from inside backward by the arrival hooks, the
    remainder after it; the result must equal a per-parameter Adam fed the rank-averaged gradients, and the hook really fires."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import copy
        from medt_b200.train_step import EarlyBucket, FlatGroup
        torch.manual_seed(0)
        net = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 3, padding=1), torch.nn.ReLU(), torch.nn.Conv2d(4, 8, 3, padding=1),
                                  torch.nn.ReLU(), torch.nn.Conv2d(8, 2, 1))
        ref = copy.deepcopy(net)
        # arrival order = reverse of the forward order (the head's gradients come first)
        arrival = [p for m in reversed(list(net)) for p in m.parameters()]
        group = FlatGroup(arrival)
        eb = EarlyBucket(group, world, None, frac=0.5)
        assert 0 < eb.k < len(group.params) and eb.n_early == sum(p.numel() for p in group.params[:eb.k])
        opt = torch.optim.Adam([group.flat], lr=1e-2, weight_decay=1e-5)
        opt_ref = torch.optim.Adam(ref.parameters(), lr=1e-2, weight_decay=1e-5)
        for it in range(3):
            torch.manual_seed(100 + 10 * it + rank)            # different data shard per rank
            x, y = torch.rand(4, 3, 8, 8), torch.randint(0, 2, (4, 8, 8))
            for p in net.parameters():
                p.grad = None
            eb.arm()
            torch.nn.functional.cross_entropy(net(x), y).backward()
            assert eb.fired                                    # the prefix was reduced from inside backward
            eb.finish()
            group.publish_grads_()
            opt.step()
            opt_ref.zero_grad(set_to_none=True)
            torch.nn.functional.cross_entropy(ref(x), y).backward()
            for p in ref.parameters():
                dist.all_reduce(p.grad)
                p.grad.div_(world)
            opt_ref.step()
            for a, b in zip(net.parameters(), ref.parameters()):
                assert torch.allclose(a.detach(), b.detach(), atol=1e-7, rtol=1e-6), it
        flat = group.flat.detach().clone()
        both = [torch.zeros_like(flat) for _ in range(world)]
        dist.all_gather(both, flat)
        assert torch.equal(both[0], both[1])                   # replicas stay in lock-step
        q.put((rank, "ok"))
    except Exception as e:  # noqa: BLE001
        import traceback
        q.put((rank, repr(e) + traceback.format_exc()[-600:]))
    finally:
        dist.destroy_process_group()


def test_early_bucket_overlapped_allreduce_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_early, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_flat_group_allreduce_adam_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_flat, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, "ok"), (1, "ok")], res


def test_reference_script_launcher_paths(tmp_path, monkeypatch):
    """run_reference_script.py: the drop-in `lib` must win over the reference's, while the script's
    bare `import utils / metrics` still resolve in the reference directory."""
    import importlib.util
    import sys
    here = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("rrs", os.path.join(here, "medical-transformer_b200", "run_reference_script.py"))
    rrs = importlib.util.module_from_spec(spec); spec.loader.exec_module(rrs)
    ref = tmp_path / "Medical-Transformer"; (ref / "lib").mkdir(parents=True)
    (ref / "lib" / "__init__.py").write_text("WHO = 'reference'\n"); (ref / "metrics.py").write_text("WHO = 'reference'\n")
    (ref / "train.py").write_text("")
    monkeypatch.setattr(sys, "path", [str(ref)] + list(sys.path))
    pkg, ref_dir = rrs.setup_paths(str(ref / "train.py"))
    assert sys.path[0] == pkg and sys.path.index(ref_dir) > 0
    for m in ("lib", "metrics"):
        sys.modules.pop(m, None)
    import lib, metrics  # noqa: E401
    assert not hasattr(lib, "WHO") and lib.__file__.startswith(pkg) and metrics.WHO == "reference"
    sys.modules.pop("metrics", None)
