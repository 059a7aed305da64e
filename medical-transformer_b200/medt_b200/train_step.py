"""The user-facing training step of the B200 path.

`TrainStep(model)` is what replaces the body of the reference's inner loop
(train.py:134-162: H2D copy, forward, cross-entropy, backward, Adam) for one process per GPU:

  * the whole step -- forward, loss, backward, gradient all-reduce, Adam -- is captured ONCE in
    a CUDA graph and replayed (the 128-px step is launch-bound: ~500 kernels of a few us);
  * data parallelism is one process per GPU with a SINGLE flat fp32 gradient bucket
    all-reduced (AVG) by NCCL over NVLink/NVSwitch inside the graph -- the B200 equivalent of
    the reference's nn.DataParallel(device_ids=[0,1]) (train.py:104-108).  BatchNorm statistics
    stay per rank (DataParallel semantics: no SyncBN in the reference);
  * parameters that receive no gradient (frozen gates before train.py:169-171 flips them;
    MedT's never-used adjust_p / AxialBlock_wopos.conv1) are left out of the bucket and of the
    optimizer step exactly as torch.optim.Adam skips `grad is None` parameters; call
    `recapture()` after changing requires_grad.

  * the active parameters are re-homed as views of ONE flat fp32 buffer (16-byte alignment kept: sizes that are
    not multiples of 4 go last), so the optimizer is torch's fused Adam on a single tensor -- 1 launch instead of 6
    multi-tensor chunks (92 us of a 4.4 ms step) with bit-identical arithmetic -- and the gradient gather is one
    `torch.cat` into the matching flat gradient, which is also the NCCL bucket.  Parameters that become active later
    (the gates at epoch 10) get a second flat tensor with its own Adam step count, as torch.optim.Adam gives a
    parameter whose first gradient arrives late;

`step(x_host, y_host)` is the end-to-end call: pinned-host -> device copies of the batch,
graph replay, device -> pinned-host copy of the loss.  `step_resident()` replays on the
batch already in HBM.
"""
from __future__ import annotations

import os

import torch
import torch.distributed as dist
import torch.nn.functional as F


class FlatBucket:
    """ONE flat gradient bucket for data-parallel training: gather the per-parameter gradients
    with a single `torch.cat`, all-reduce (mean) once, hand the averaged views back.  Pure
    torch / torch.distributed, device agnostic (NCCL on GPUs; the gloo CPU test uses it too)."""

    def __init__(self, params, process_group=None):
        self.params = list(params)
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        total = sum(p.numel() for p in self.params)
        dev = self.params[0].device if self.params else "cpu"
        self.flat = torch.zeros(total, device=dev)
        self.views, off = [], 0
        for p in self.params:
            self.views.append(self.flat[off:off + p.numel()].view_as(p))
            off += p.numel()

    def nbytes(self):
        return self.flat.numel() * self.flat.element_size()

    def reduce_(self):
        """p.grad (fresh autograd tensors) -> bucket -> mean over ranks -> p.grad = bucket views."""
        torch.cat([p.grad.reshape(-1) for p in self.params], out=self.flat)
        if self.world > 1:
            if dist.get_backend(self.pg) == "nccl":
                dist.all_reduce(self.flat, op=dist.ReduceOp.AVG, group=self.pg)
            else:                                   # gloo has no AVG
                dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=self.pg)
                self.flat.div_(self.world)
        for p, v in zip(self.params, self.views):
            p.grad = v


class FlatGroup:
    """A set of parameters re-homed as views of one flat buffer (`flat`, an nn.Parameter the optimizer sees) plus the
    matching flat gradient.  Device agnostic (exercised on CPU by tests/test_dropin.py)."""

    def __init__(self, params):
        # sizes that keep 16-byte alignment first, the odd ones (gates, 2-element biases) at the end
        self.params = sorted(params, key=lambda p: p.numel() % 4 != 0)
        total = sum(p.numel() for p in self.params)
        flat = torch.zeros(total, dtype=self.params[0].dtype, device=self.params[0].device)
        self.views, self.grad_views, off = [], [], 0
        grad = torch.zeros_like(flat)
        with torch.no_grad():
            for p in self.params:
                n = p.numel()
                v = flat[off:off + n].view_as(p)
                v.copy_(p.data)
                p.data = v
                self.views.append(v)
                self.grad_views.append(grad[off:off + n].view_as(p))
                off += n
        self.flat = torch.nn.Parameter(flat)
        self.flat.grad = grad

    def gather_grads_(self):
        """fresh per-parameter autograd gradients -> the flat gradient.  One multi-tensor copy (torch._foreach_copy_: a couple of
        chunked launches) instead of `torch.cat(..., out=)`, whose batched-copy kernel needs 50 us for these ~200 small tensors."""
        if hasattr(torch, "_foreach_copy_"):
            torch._foreach_copy_(self.grad_views, [p.grad for p in self.params])
        else:
            torch.cat([p.grad.reshape(-1) for p in self.params], out=self.flat.grad)

    def publish_grads_(self):
        for p, v in zip(self.params, self.grad_views):
            p.grad = v


def _all_reduce_mean_(t, world, pg):
    if world > 1:
        if dist.get_backend(pg) == "nccl":
            dist.all_reduce(t, op=dist.ReduceOp.AVG, group=pg)
        else:                                       # gloo has no AVG
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=pg)
            t.div_(world)


class EarlyBucket:
    """Two-bucket mean all-reduce of one FlatGroup's gradient with the FIRST bucket overlapped with the rest of backward.

    The group's parameters are laid out in gradient-arrival order (TrainStep probes it): the first `k` of them -- the prefix that
    holds >= `frac` of the elements: adjust, the decoders and the deep layers, whose gradients are complete when backward has
    not yet reached layer3 -- are gathered into the flat gradient and all-reduced on a forked stream the moment the last of them
    received its gradient (post-accumulate-grad hooks count arrivals); the small remainder (layer1 / layer2 / stem) is gathered
    and reduced after backward, then the forked stream is joined.  Inside a CUDA-graph capture the fork / join become graph
    edges, so the NCCL kernel of the first bucket runs next to the attention backward kernels.  Device agnostic: without CUDA
    (the gloo CPU test) both reductions run in line at the same points of the step."""

    def __init__(self, group, world, pg, frac=0.85):
        self.g, self.world, self.pg = group, world, pg
        total = sum(p.numel() for p in group.params)
        acc, k = 0, 0
        while k < len(group.params) and acc < frac * total and group.params[k].numel() % 4 == 0:
            acc += group.params[k].numel()
            k += 1
        self.k, self.n_early = k, acc
        dev = group.flat.device
        self.stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        self.armed, self.count, self.fired = False, 0, False
        self.handles = [p.register_post_accumulate_grad_hook(self._hook) for p in group.params[:k]]

    def remove(self):
        for h in self.handles:
            h.remove()
        self.handles = []

    def arm(self):
        self.armed, self.count, self.fired = True, 0, False

    def _hook(self, _p):
        if not self.armed:
            return
        self.count += 1
        if self.count == self.k:
            self._fire()

    def _copy(self, lo, hi):
        g = self.g
        if hi > lo:
            if hasattr(torch, "_foreach_copy_"):
                torch._foreach_copy_(g.grad_views[lo:hi], [p.grad for p in g.params[lo:hi]])
            else:
                for v, p in zip(g.grad_views[lo:hi], g.params[lo:hi]):
                    v.copy_(p.grad)

    def _fire(self):
        self._copy(0, self.k)
        early = self.g.flat.grad[:self.n_early]
        if self.stream is not None:
            cur = torch.cuda.current_stream(early.device)
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                _all_reduce_mean_(early, self.world, self.pg)
        else:
            _all_reduce_mean_(early, self.world, self.pg)
        self.fired = True

    def finish(self):
        """after backward: the second bucket, then join the first"""
        self.armed = False
        g = self.g
        if not self.fired:                     # a prefix parameter received no gradient this step: one plain reduction
            g.gather_grads_()
            _all_reduce_mean_(g.flat.grad, self.world, self.pg)
            return
        self._copy(self.k, len(g.params))
        if self.n_early < g.flat.grad.numel():
            _all_reduce_mean_(g.flat.grad[self.n_early:], self.world, self.pg)
        if self.stream is not None:
            torch.cuda.current_stream(g.flat.device).wait_stream(self.stream)


class TrainStep:
    def __init__(self, model, batch_shape, lr=1e-3, weight_decay=1e-5, device=None, use_graph=True,
                 process_group=None, loss_slots=64, flat_adam=None):
        self.model = model
        self.device = torch.device(device if device is not None else next(model.parameters()).device)
        if self.device.type != "cuda":
            raise RuntimeError("medt_b200.TrainStep needs a CUDA device; there is no CPU path")
        self.lr, self.weight_decay, self.use_graph = lr, weight_decay, use_graph
        self.pg = process_group
        self.world = dist.get_world_size(process_group) if (dist.is_available() and dist.is_initialized()) else 1
        n, c, h, w = batch_shape
        self.x = torch.zeros(batch_shape, device=self.device)
        self.y = torch.zeros((n, h, w), dtype=torch.long, device=self.device)
        self.x_pin = torch.zeros(batch_shape).pin_memory()
        self.y_pin = torch.zeros((n, h, w), dtype=torch.long).pin_memory()
        self.loss = torch.zeros((), device=self.device)
        self.loss_pin = torch.zeros(loss_slots).pin_memory()
        self._slot = 0
        self._stage_ev = None
        self.graph = None
        self.opt = None
        self.steps_done = 0
        # default: flat mode (one fused-Adam launch on one tensor; the flat gradient is also the NCCL bucket), on one GPU and
        # under NCCL alike.  flat_adam=False / MEDT_NO_FLAT_ADAM=1 keeps the per-parameter optimizer + FlatBucket path.
        if flat_adam is None:
            flat_adam = not os.environ.get("MEDT_NO_FLAT_ADAM")
        self.flat_adam = bool(flat_adam)
        self.groups = []
        self._early = None
        self._arrival = {}
        self._build()

    # ------------------------------------------------------------------ setup
    def _probe_active_params(self):
        """Parameters that actually receive a gradient (one eager forward/backward)."""
        for p in self.model.parameters():
            p.grad = None
        state = {k: v.clone() for k, v in self.model.state_dict().items()}       # keep BN buffers untouched
        order, hooks = [], []
        for p in self.model.parameters():
            if p.requires_grad:
                hooks.append(p.register_post_accumulate_grad_hook(lambda q, _o=order: _o.append(id(q))))
        loss = self._loss()
        loss.backward()
        for h in hooks:
            h.remove()
        self._arrival = {pid: i for i, pid in enumerate(order)}                  # gradient arrival order (EarlyBucket layout)
        active = [p for p in self.model.parameters() if p.grad is not None]
        self.model.load_state_dict(state)
        for p in self.model.parameters():
            p.grad = None
        return active

    def _adam(self, params):
        return torch.optim.Adam(params, lr=self.lr, weight_decay=self.weight_decay, capturable=True, fused=True)

    def _sync_groups(self):
        """Flat mode: (re)build the flat groups for the current active set.  Existing groups (and their Adam state)
        are kept; newly active parameters form a new group.  A parameter that STOPPED receiving gradients (the
        reference never does that) forces a rebuild from scratch, which resets the optimizer state."""
        grouped = {id(p) for g in self.groups for p in g.params}
        act = {id(p) for p in self.active}
        if not grouped <= act:
            self.groups, self.opt, grouped = [], None, set()
            if self._early is not None:
                self._early.remove()
                self._early = None
        fresh = [p for p in self.active if id(p) not in grouped]
        fresh.sort(key=lambda p: self._arrival.get(id(p), 1 << 30))              # early gradients first (stable inside FlatGroup)
        if fresh:
            g = FlatGroup(fresh)
            self.groups.append(g)
            if self.opt is None:
                self.opt = self._adam([g.flat])
            else:
                self.opt.param_groups[0]["params"].append(g.flat)
        for pg in self.opt.param_groups:
            pg["lr"], pg["weight_decay"] = self.lr, self.weight_decay
        # opt-in (MEDT_OVERLAP_REDUCE=1): measured on 2 x B200, the overlapped variant is 14 us SLOWER per step than the single
        # all-reduce after backward (3.324 vs 3.309 ms): the attention kernels are cooperative launches sized to the whole GPU, so
        # an NCCL kernel holding SMs on the forked stream delays the next cooperative launch instead of running beside it
        if self.world > 1 and self._early is None and self.groups and os.environ.get("MEDT_OVERLAP_REDUCE"):
            self._early = EarlyBucket(self.groups[0], self.world, self.pg)

    def _build(self):
        self.model.train()
        # every BatchNorm counter becomes a 0-dim view of ONE int64 tensor: one add per step
        bns = [m for m in self.model.modules() if isinstance(m, torch.nn.modules.batchnorm._BatchNorm)]
        self._nbt = torch.stack([b.num_batches_tracked.to(self.device) for b in bns]) if bns else None
        for i, b in enumerate(bns):
            b.num_batches_tracked = self._nbt[i]
            b._nbt_deferred = True
        self.active = self._probe_active_params()          # its forward records each BN's statistic-group count
        self._nbt_inc = torch.tensor([int(getattr(b, "_stat_groups_bn", 1)) for b in bns], device=self.device)
        self.graph = None
        if self.flat_adam:
            self._build_flat()
            return
        self.bucket = FlatBucket(self.active, self.pg) if self.world > 1 else None
        # Adam state that already exists (a recapture after the epoch-10 un-freeze, train.py:169-171, or an lr change) is
        # kept PER PARAMETER: the reference's optimizer holds every parameter from the start (train.py:111), so moments and
        # step counts of the old parameters continue and only the newly active ones start at step 1
        snap = {}
        if self.opt is not None:
            for q, st in self.opt.state.items():
                snap[id(q)] = {k: v.clone() for k, v in st.items() if torch.is_tensor(v)}
        self.opt = self._adam(self.active)
        state = {k: v.clone() for k, v in self.model.state_dict().items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):               # warm-up (allocator, cuDNN autotune, lazy init); creates the state tensors
            for _ in range(3 if self.use_graph else 1):
                self._step_body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.model.load_state_dict(state)           # undo the warm-up's parameter / statistic updates
        with torch.no_grad():
            for q in self.active:
                for k, v in self.opt.state[q].items():
                    if torch.is_tensor(v):
                        old = snap.get(id(q), {}).get(k)
                        v.zero_() if old is None else v.copy_(old)
        torch.cuda.synchronize(self.device)
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.graph):
                self._step_body()
            torch.cuda.synchronize(self.device)     # capture does not execute: state is still the restored one

    def _build_flat(self):
        self.bucket = None
        self._sync_groups()
        if not self.use_graph:
            return
        # Adam state that already exists (a recapture) survives the warm-up through this snapshot; state the warm-up
        # creates for a new group is zeroed IN PLACE afterwards (the tensors must exist before capture, or their
        # creation -- a fill -- would be captured and replayed every step)
        snap = {id(g.flat): {k: v.clone() for k, v in self.opt.state[g.flat].items() if torch.is_tensor(v)}
                for g in self.groups if self.opt.state.get(g.flat)}
        state = {k: v.clone() for k, v in self.model.state_dict().items()}
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(torch.cuda.current_stream(self.device))
        with torch.cuda.stream(side):               # warm-up (allocator, cuDNN autotune, lazy init)
            for _ in range(3):
                self._step_body()
        torch.cuda.current_stream(self.device).wait_stream(side)
        torch.cuda.synchronize(self.device)
        self.model.load_state_dict(state)           # in place: the flat views stay
        with torch.no_grad():
            for g in self.groups:
                for k, v in self.opt.state[g.flat].items():
                    if torch.is_tensor(v):
                        old = snap.get(id(g.flat), {}).get(k)
                        v.zero_() if old is None else v.copy_(old)
        torch.cuda.synchronize(self.device)
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self._step_body()
        torch.cuda.synchronize(self.device)         # capture does not execute: state is still the restored one

    def recapture(self):
        """Call after flipping requires_grad (train.py:169-171) or changing lr."""
        self._build()

    def _loss(self):
        """forward + loss.  The tail `adjust(relu(x))` + cross-entropy runs as ONE fused kernel pair (csrc/head_ce.cu) when the
        model exposes `forward_features` and a two-class 1x1 `adjust` with bias (every reference factory does); MEDT_NO_FUSED_HEAD=1
        or any other model keeps `F.cross_entropy(model(x), y)`."""
        m = self.model
        adj = getattr(m, "adjust", None)
        if (hasattr(m, "forward_features") and isinstance(adj, torch.nn.Conv2d) and adj.kernel_size == (1, 1) and adj.out_channels == 2
                and adj.bias is not None and adj.in_channels <= 64 and not os.environ.get("MEDT_NO_FUSED_HEAD")):
            from .shell_ops import head_ce
            loss, _logits = head_ce(m.forward_features(self.x), adj, self.y)
            return loss
        return F.cross_entropy(m(self.x), self.y)              # LogNLLLoss == cross_entropy (metrics.py:17-20)

    def _step_body(self):
        for p in self.active:
            p.grad = None
        loss = self._loss()
        early = self._early if self.flat_adam else None
        if early is not None:
            early.arm()
        loss.backward()
        if self.flat_adam:
            for g in self.groups:
                if early is not None and g is early.g:
                    early.finish()                  # first bucket was reduced on the forked stream during backward
                else:
                    g.gather_grads_()
                    _all_reduce_mean_(g.flat.grad, self.world, self.pg)
                g.publish_grads_()
        elif self.bucket is not None:
            self.bucket.reduce_()
        self.opt.step()
        if self._nbt is not None:
            self._nbt += self._nbt_inc
        self.loss.copy_(loss.detach())

    # ------------------------------------------------------------------ stepping
    def step_resident(self):
        if self.graph is not None:
            self.graph.replay()
        else:
            self._step_body()
        self.steps_done += 1
        return self.loss

    def step(self, x_host, y_host):
        """End-to-end step from host tensors.  Returns the slot of `loss_pin` that will hold this
        step's loss once the stream has drained (no host sync here unless the inputs are pageable: then the staging
        buffer of the previous step must have been consumed before it is overwritten).
        The captured step has a fixed batch shape: a ragged last batch raises (use DataLoader(drop_last=True))."""
        if tuple(x_host.shape) != tuple(self.x.shape) or tuple(y_host.shape) != tuple(self.y.shape):
            raise ValueError(f"TrainStep was captured for x {tuple(self.x.shape)} / y {tuple(self.y.shape)}, got "
                             f"{tuple(x_host.shape)} / {tuple(y_host.shape)} (copy_ would silently broadcast)")
        staged = False
        if not (x_host.is_pinned() and y_host.is_pinned()):
            if self._stage_ev is not None:
                self._stage_ev.synchronize()        # the previous step's DMA out of x_pin / y_pin has finished
        if x_host.is_pinned():
            self.x.copy_(x_host, non_blocking=True)
        else:
            self.x_pin.copy_(x_host)
            self.x.copy_(self.x_pin, non_blocking=True)
            staged = True
        if y_host.is_pinned():
            self.y.copy_(y_host, non_blocking=True)
        else:
            self.y_pin.copy_(y_host)
            self.y.copy_(self.y_pin, non_blocking=True)
            staged = True
        if staged:
            if self._stage_ev is None:
                self._stage_ev = torch.cuda.Event()
            self._stage_ev.record(torch.cuda.current_stream(self.device))
        self.step_resident()
        slot = self._slot
        self.loss_pin[slot:slot + 1].copy_(self.loss.reshape(1), non_blocking=True)
        self._slot = (slot + 1) % self.loss_pin.numel()
        return slot

    def h2d_bytes_per_step(self):
        return self.x.numel() * self.x.element_size() + self.y.numel() * self.y.element_size()

    def d2h_bytes_per_step(self):
        return self.loss.element_size()
