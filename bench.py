#!/usr/bin/env python
"""bench.py -- training images/sec, 128x128, bs = 4 per GPU (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--model gated|MedT] [--impl b200|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A "step" is one full training step of the reference's loop (train.py:134-162) on one synthetic
batch: forward, cross-entropy, backward, [gradient all-reduce], Adam.  Synthetic data of the
reference's shape: X = rand(4,3,128,128) fp32 in [0,1), y = randint(0,2,(4,128,128)) int64,
generator seed 3000, model built under torch.manual_seed(3000) with the reference's init.

  value   img/s, whole job (N GPUs), batch resident in HBM, CUDA-graph-replayed step; per-step
          CUDA events on the launching stream, L2 flushed (256 MiB write) between timed steps,
          max over ranks.
  e2e     same metric through the public API medt_b200.TrainStep.step(x_host, y_host): pinned
          host -> device copy of the batch and device -> host copy of the loss inside the timed
          region.
  roofline / cpu_baseline / clocks / gpu_launches: see DESIGN.md "measurement".

--impl reference times the CPU restatement of the reference (oracle/torch_port.py, same ATen
op sequence as the reference) on this box's host cores, rank 0 only.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import threading
import time

REPO = os.path.dirname(os.path.abspath(__file__))
for _p in (REPO, os.path.join(REPO, "medical-transformer_b200"), os.path.join(REPO, "tests", "golden")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

IMG, BS = 128, 4
METRIC = "training images/sec (128x128, bs=4/GPU)"


def peaks():
    try:
        return json.load(open(os.path.join(REPO, "MEASURED_PEAKS.json"))), "measured"
    except Exception:  # noqa: BLE001
        return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


def synthetic_batch(n=BS, s=IMG):
    g = torch.Generator().manual_seed(3000)
    x = torch.rand(n, 3, s, s, generator=g)
    y = torch.randint(0, 2, (n, s, s), generator=g)
    return x, y


def build_model(name, s=IMG):
    import lib
    torch.manual_seed(3000)
    f = {"gated": lib.models.axialnet.gated, "MedT": lib.models.axialnet.MedT, "logo": lib.models.axialnet.logo,
         "axialunet": lib.models.axialunet}[name]
    return f(img_size=s, imgchan=3)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(self.index), "-lms", "100"], stdout=subprocess.PIPE, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:  # noqa: BLE001
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc is not None:
            time.sleep(0.15)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:  # noqa: BLE001
                self.proc.kill()

    def summary(self):
        sm = [float(r[1]) for r in self.rows if len(r) >= 9 and r[1].replace(".", "").isdigit()]
        mx = [float(r[2]) for r in self.rows if len(r) >= 9 and r[2].replace(".", "").isdigit()]
        reasons = set()
        for r in self.rows:
            if len(r) >= 9:
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------- reference arm (CPU)
REF_DIR = os.path.join(REPO, "baseline", "_ref")


def ref_available():
    return os.path.isfile(os.path.join(REF_DIR, "lib", "models", "axialnet.py"))


def _import_reference():
    """The UNMODIFIED reference from baseline/_ref (copied there by baseline/make_ref.py; git-ignored, travels with the
    gpurun snapshot).  Only the reference arm calls this, in a process that never imports the drop-in `lib`."""
    assert "lib" not in sys.modules, "the reference arm must not share a process with the drop-in package"
    drop = os.path.join(REPO, "medical-transformer_b200")
    sys.path[:] = [p_ for p_ in sys.path if os.path.abspath(p_ or ".") != drop]
    sys.path.insert(0, REF_DIR)
    import importlib.util
    for mod in ("matplotlib", "skimage"):                     # import-only in the reference (axialnet.py:8, utils.py:5)
        if importlib.util.find_spec(mod) is None:
            sys.path.append(os.path.join(drop, "stubs"))
            break
    import lib as ref_lib
    import metrics as ref_metrics
    return ref_lib, ref_metrics


def cpu_reference_steps(model_name, steps, warmup, size=IMG):
    """The reference's own CPU path: its modules (baseline/_ref/lib), its loss (metrics.LogNLLLoss) and the body of its
    training loop (train.py:134-161: forward, the unused numpy thresholding, loss, zero_grad, backward, Adam step).
    Falls back to the pinned restatement (oracle/torch_port.py) only when baseline/_ref is absent."""
    kind = "reference" if ref_available() else "port"
    if model_name == "logo" and size != 128:
        kind = "port"      # the reference hard-codes 32-px patches: `logo` raises at any other size; the port
                           # restates its loop with patch = size // 4
    if kind == "reference":
        ref_lib, ref_metrics = _import_reference()
        torch.manual_seed(3000)
        f = {"gated": ref_lib.models.axialnet.gated, "MedT": ref_lib.models.axialnet.MedT, "logo": ref_lib.models.axialnet.logo,
             "axialunet": ref_lib.models.axialunet}[model_name]
        model = f(img_size=size, imgchan=3).train()
        criterion = ref_metrics.LogNLLLoss()

        def step(opt, x, y):
            output = model(x)
            tmp2 = y.detach().cpu().numpy()
            tmp = output.detach().cpu().numpy()
            tmp[tmp >= 0.5] = 1
            tmp[tmp < 0.5] = 0
            tmp2[tmp2 > 0] = 1
            tmp2[tmp2 <= 0] = 0
            tmp2 = tmp2.astype(int)
            tmp = tmp.astype(int)
            loss = criterion(output, y)
            opt.zero_grad()
            loss.backward()
            opt.step()
            return loss.item()
    else:
        from oracle import torch_port
        patch = size // 4 if (model_name == "logo" and size != 128) else 32
        if patch != 32:
            import lib
            torch.manual_seed(3000)
            model = lib.models.axialnet.logo(img_size=size, imgchan=3, patch_size=patch).train()
        else:
            model = build_model(model_name, size).train()

        def step(opt, x, y):
            return torch_port.train_step(model, opt, x, y, patch)
    opt = torch.optim.Adam(list(model.parameters()), lr=1e-3, weight_decay=1e-5)       # train.py:111-112
    x, y = synthetic_batch(BS, size)
    # use the host thread count that is FASTEST for the reference's op mix: its ~1.2k tiny ATen ops
    # per step get slower, not faster, when oversubscribed (128 threads: > 60 s/step on the GPU host)
    cores = os.cpu_count() or 1
    best, best_t = None, float("inf")
    for nt in [c for c in (8, 16, 32, 64, 128, 256) if c <= cores] or [cores]:
        torch.set_num_threads(nt)
        if best is None:
            step(opt, x, y)            # first-touch warm-up
        t0 = time.perf_counter()
        step(opt, x, y)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = nt, dt
        elif dt > 1.5 * best_t:
            break
    torch.set_num_threads(best)
    cpu_reference_steps.threads = best
    cpu_reference_steps.kind = kind
    for _ in range(max(warmup - 1, 0)):
        step(opt, x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        step(opt, x, y)
    dt = time.perf_counter() - t0
    return BS * steps / dt, dt / steps * 1e3


def cpu_model_string():
    try:
        for line in open("/proc/cpuinfo"):
            if line.startswith("model name"):
                return line.split(":", 1)[1].strip()
    except Exception:  # noqa: BLE001
        pass
    return "unknown"


WORKLOADS = {"gated128": ("gated", 128), "medt128": ("MedT", 128), "medt256": ("MedT", 256), "gated512": ("gated", 512),
             "logo512": ("logo", 512)}


def workload_string(model, size, world=1):
    name = "gatedaxialunet" if model == "gated" else model
    return f"{name} img_size={size} batch_size=4/GPU train step (fwd+CE+bwd+{'NCCL all-reduce+' if world > 1 else ''}Adam)"


def run_reference(args, rank):
    if rank != 0:
        return
    steps = args.steps
    if args.size >= 256:
        steps = min(steps, 2)          # bounded sample: one 512-px step of the reference takes minutes on the host
    ips, ms = cpu_reference_steps(args.model, steps, min(args.warmup, 1), args.size)
    cores = os.cpu_count() or 1
    kind = cpu_reference_steps.kind
    line = {"impl": "reference", "metric": METRIC if args.size == 128 else f"training images/sec ({args.size}x{args.size}, bs=4/GPU)",
            "value": ips, "unit": "img/s", "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.model, args.size), "global_batch": BS, "parallelism": "cpu",
                       "impl_note": ("the UNMODIFIED reference from baseline/_ref (lib.models.axialnet + metrics.LogNLLLoss + the body of "
                                     "train.py:134-161) on the host cores" if kind == "reference" else
                                     "baseline/_ref absent: the reference's ATen op sequence restated in oracle/torch_port.py")},
            "cpu_baseline": {"value": ips, "unit": "img/s", "cores": cpu_reference_steps.threads, "host_cores": cores,
                             "kind": kind, "sample": f"{steps} full train steps of the bs=4 batch",
                             "cpu": cpu_model_string()},
            "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline_subprocess(model, size, steps=3, warmup=1, timeout=900):
    """The reference arm in its own process (it imports the reference's `lib`, which cannot share a process with the drop-in
    `lib`); returns its cpu_baseline object."""
    r = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--model", model, "--size", str(size),
                        "--steps", str(steps), "--warmup", str(warmup)], capture_output=True, text=True, timeout=timeout,
                       env={k: v for k, v in os.environ.items() if k not in ("RANK", "WORLD_SIZE", "LOCAL_RANK")})
    line = json.loads(r.stdout.strip().splitlines()[-1])
    cb = line["cpu_baseline"]
    cb["ms_per_step"] = line["ms_per_step"]
    return cb


# ----------------------------------------------------------------------------- product arm
class L2Flusher:
    def __init__(self, dev, mib=256):
        self.buf = torch.empty(mib * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

    def __call__(self):
        self.buf.fill_(1.0)


def timed_steps(fn, steps, flush, dev):
    """Per-step CUDA events on the current stream; L2 flush between steps is outside the brackets."""
    evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for a, b in evs:
        flush()
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize(dev)
    return sum(a.elapsed_time(b) for a, b in evs)          # ms


def op_profile(dev, flush, variant, C, N, H, W, axis, stride, reps=20):
    """Live durations of ONE axial op on the default path: the persistent cooperative forward kernel (k_axial_fwd) and backward
    kernel (k_axial_bwd; the weight-gradient GEMM a full call forks onto the side stream is left out by stage mask 0x3f), each
    launched through the C ABI profiling hook with CUDA events on the launching stream -- L2 flushed between launches (cold) and
    back to back (warm).  Inputs are synthetic; the durations are those of the real launches (valid results)."""
    from medt_b200 import _cabi
    from medt_b200.axial_op import _desc, _ptr
    lib_ = _cabi.load_library()
    L, gp = (H if axis == 0 else W), C // 8
    g = torch.Generator(device="cpu").manual_seed(1)
    rn = lambda *s: torch.randn(*s, generator=g).to(dev)                       # noqa: E731
    nsim, nout = (8, C) if variant == "nopos" else (24, 2 * C)
    x, dout = rn(N, C, H, W), rn(N, C, H // stride, W // stride)
    w = rn(2 * C, C) * (1.0 / C) ** 0.5
    rel = rn(2 * gp, 2 * L - 1) * (1.0 / gp) ** 0.5 if variant != "nopos" else None
    gates = [torch.tensor(v, device=dev) for v in (0.1, 0.1, 1.0, 0.1)] if variant == "gated" else [None] * 4
    bn = []
    for n in (2 * C, nsim, nout):
        bn += [torch.rand(n, device=dev) + 0.5, rn(n) * 0.1, torch.zeros(n, device=dev), torch.ones(n, device=dev)]
    desc = _desc(x, axis, stride, variant, True, 1, 1e-5, 0.1)
    params = _cabi.MtAxialParams(_ptr(w), _ptr(rel), *[_ptr(t) for t in gates], *[_ptr(t) for t in bn])
    saved = torch.empty(lib_.mt_axial_saved_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
    ws = torch.empty(lib_.mt_axial_bwd_workspace_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
    out = torch.empty_like(dout)
    gt = [torch.empty_like(x), torch.empty_like(w), None if rel is None else torch.empty_like(rel), torch.empty(4, device=dev)] + \
         [torch.empty(n, device=dev) for n in (2 * C, 2 * C, nsim, nsim, nout, nout)]
    grads = _cabi.MtAxialGrads(*[_ptr(t) for t in gt])
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)

    def fwd():
        _cabi.check(lib_.mt_axial_forward_stages(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(out), _ptr(saved), st, 0x3f), "fwd")

    def bwd():
        _cabi.check(lib_.mt_axial_backward_stages(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(dout), _ptr(saved), _ptr(ws),
                                                  ctypes.byref(grads), st, 0x3f), "bwd")
    res = {}
    fwd()
    res["fwd_launches"] = int(lib_.mt_last_launch_count())
    bwd()
    res["bwd_launches"] = int(lib_.mt_last_launch_count())
    torch.cuda.synchronize(dev)
    for name, fn in (("fwd", fwd), ("bwd", bwd)):
        for _ in range(3):
            fn()
        res[name + "_us_cold"] = timed_steps(fn, reps, flush, dev) / reps * 1e3
        res[name + "_us_warm"] = timed_steps(fn, reps, lambda: None, dev) / reps * 1e3
    B = N * (W if axis == 0 else H)
    res.update(B=B, C=C, L=L)
    return res


# measured ceilings of the non-tensor pipes on B200 (148 SMs x 4 SMSPs): fp32 FMA 2 x 128 lanes x 148 x 1.965 GHz, MUFU.EX2 16 / clk / SM
FMA_TFLOPS = 148 * 128 * 2 * 1.965e9 / 1e12
EXP_PER_S = 148 * 16 * 1.965e9


def op_roofline(r, pk, direction):
    """SURVEY 8(d) per-op figures for ONE axial op: bytes = 2 B C L 4 (x in, out out at the fp32 NCHW boundary; the backward reads
    dout and x and writes dx: 3 B C L 4), flops = 4 B L C^2 + 7 B L^2 C (backward 2.5x), exps = 8 B L^2 (the fused backward
    recomputes P: same count per sweep)."""
    B, C, L = r["B"], r["C"], r["L"]
    t = r[direction + "_us_cold"] * 1e-6
    nbytes = 4 * B * C * L * (2 if direction == "fwd" else 3)
    flops = (4 * B * L * C * C + 7 * B * L * L * C) * (1.0 if direction == "fwd" else 2.5)
    exps = 8 * B * L * L
    ach = nbytes / t / 1e9
    return {"alg_bytes_per_launch": nbytes, "us_per_launch_cold_l2": r[direction + "_us_cold"], "us_per_launch_warm_l2": r[direction + "_us_warm"],
            "achieved_GBps": ach, "hbm_frac": ach / pk["hbm_gbs"], "fp32_fma_frac": flops / t / 1e12 / FMA_TFLOPS,
            "exp_frac": exps / t / EXP_PER_S, "tensor_frac_of_measured_bf16": flops / t / 1e12 / pk["bf16_tflops"]}


def param_checksum(model):
    """one fp64 number per rank: sum over all parameters of sum(p) + sum(|p|) (lock-step check of the replicas)"""
    with torch.no_grad():
        return torch.stack([p.double().sum() + p.double().abs().sum() for p in model.parameters()]).sum()


def run_product(args, rank, world, local_rank):
    import torch.distributed as dist
    from medt_b200 import _cabi
    from medt_b200.train_step import TrainStep
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True          # fixed shapes: let cuDNN pick its fastest fp32 algorithms
    pk, pk_kind = peaks()
    flush = L2Flusher(dev)
    SZ = args.size

    def one_model(name, size=None):
        size = size or SZ
        x_host, y_host = synthetic_batch(BS, size)
        x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()
        log("building", name, size)
        kw = {"patch_size": size // 4} if (name == "logo" and size != 128) else {}
        import lib
        torch.manual_seed(3000)
        f = {"gated": lib.models.axialnet.gated, "MedT": lib.models.axialnet.MedT, "logo": lib.models.axialnet.logo}[name]
        model = f(img_size=size, imgchan=3, **kw).to(dev).train()
        n0 = _cabi.LAUNCHES["kernels"]
        ts = TrainStep(model, (BS, 3, size, size), lr=1e-3, weight_decay=1e-5, device=dev, use_graph=True)
        ts.x.copy_(x_host.to(dev)); ts.y.copy_(y_host.to(dev))
        n1 = _cabi.LAUNCHES["kernels"]
        log("captured", name)
        # TrainStep._build runs: probe(1) + warm-up(3) + state step(1) + capture(1) = 6 step bodies
        per_step_launches = (n1 - n0) // 6
        for _ in range(args.warmup):
            ts.step_resident()
        torch.cuda.synchronize(dev)
        if world > 1:
            dist.barrier()
        with ClockSampler(local_rank) as cs:
            ms = timed_steps(ts.step_resident, args.steps, flush, dev)
            torch.cuda.synchronize(dev)
            # back-to-back (no flush), one bracket -- reported in config for context
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(args.steps):
                ts.step_resident()
            e1.record()
            torch.cuda.synchronize(dev)
            ms_b2b = e0.elapsed_time(e1)
            # end to end through the public API: pinned host batch in, loss out, every step
            for _ in range(2):
                ts.step(x_pin, y_pin)
            torch.cuda.synchronize(dev)
            ms_e2e = timed_steps(lambda: ts.step(x_pin, y_pin), args.steps, flush, dev)
        if world > 1:
            dist.barrier()
            t = torch.tensor([ms, ms_b2b, ms_e2e], device=dev, dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms, ms_b2b, ms_e2e = t.tolist()
        loss = float(ts.loss.item())
        # replicas in lock step: every rank must hold the same parameters after the same number of steps
        lock = None
        if world > 1:
            cks = [torch.zeros((), dtype=torch.float64, device=dev) for _ in range(world)]
            dist.all_gather(cks, param_checksum(model))
            vals = [float(c) for c in cks]
            lock = {"checksums_equal": max(vals) - min(vals) <= 1e-9 * max(1.0, abs(vals[0])), "spread": max(vals) - min(vals)}
        log("timed", name, size, ms / args.steps, "ms/step")
        return dict(model=model, ts=ts, ms=ms, ms_b2b=ms_b2b, ms_e2e=ms_e2e, launches=per_step_launches,
                    clocks=cs.summary(), loss=loss, lock=lock, size=size)

    main = one_model(args.model)
    other_name = "MedT" if args.model == "gated" else "gated"
    other = one_model(other_name) if (not args.single and SZ == 128) else None
    # BASELINE configs 3 and 5 (MedT 256 px; gated / logo 512 px), mirrored under `also` -- shorter timed regions
    extra = {}
    if args.configs and SZ == 128:
        keep_steps = args.steps
        for cfg in args.configs.split(","):
            mname, msize = WORKLOADS[cfg]
            try:
                args.steps = max(3, keep_steps // 4) if msize >= 512 else max(5, keep_steps // 2)
                r = one_model(mname, msize)
                extra[cfg] = {"value": world * BS * args.steps / (r["ms"] / 1e3), "unit": "img/s", "ms_per_step": r["ms"] / args.steps,
                              "e2e": world * BS * args.steps / (r["ms_e2e"] / 1e3), "gpu_launches_per_step": r["launches"],
                              "steps": args.steps, "workload": workload_string(mname, msize, world),
                              "patch_size": msize // 4 if mname == "logo" else None}
                del r
                torch.cuda.empty_cache()
            except Exception as e:  # noqa: BLE001
                extra[cfg] = {"error": repr(e)[:300]}
        args.steps = keep_steps

    if rank != 0:
        return
    K = args.steps
    ms_step = main["ms"] / K
    value = world * BS * K / (main["ms"] / 1e3)
    line = {"metric": METRIC if SZ == 128 else f"training images/sec ({SZ}x{SZ}, bs=4/GPU)", "value": value, "unit": "img/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": workload_string(args.model, SZ, world),
                       "global_batch": BS * world, "parallelism": f"dp{world}", "l2": "flushed (256 MiB write) between timed steps",
                       "ms_per_step_back_to_back_no_flush": main["ms_b2b"] / K, "final_loss": main["loss"],
                       "cuda_graph": True, "cudnn_tf32": bool(torch.backends.cudnn.allow_tf32)},
            "e2e": {"value": world * BS * K / (main["ms_e2e"] / 1e3), "unit": "img/s",
                    "h2d_bytes_per_step": main["ts"].h2d_bytes_per_step(), "d2h_bytes_per_step": main["ts"].d2h_bytes_per_step()},
            "gpu_launches": main["launches"] * K, "gpu_launches_per_step": main["launches"],
            "clocks": main["clocks"]}
    if main["lock"] is not None:
        line["config"]["replicas_lock_step"] = main["lock"]
    if other is not None:
        line["also"] = {other_name: {"value": world * BS * K / (other["ms"] / 1e3), "unit": "img/s",
                                     "ms_per_step": other["ms"] / K,
                                     "e2e": world * BS * K / (other["ms_e2e"] / 1e3),
                                     "gpu_launches_per_step": other["launches"]}}
    if extra:
        line.setdefault("also", {}).update(extra)
    # ---- roofline of the dominant hand-written kernel (heaviest op of the step), measured live
    log("op profile")
    try:
        # heaviest op of the gated 128-px step: layer2.0 hight_block (C32 L64 B256); its backward kernel is the step's top kernel
        r = op_profile(dev, flush, "gated", 32, BS, 64, 64, 0, 1)
        rb, rf = op_roofline(r, pk, "bwd"), op_roofline(r, pk, "fwd")
        tr = None
        tpath = os.path.join(REPO, "profiles", "top_kernel_traffic.json")
        if os.path.exists(tpath):
            tr = json.load(open(tpath)).get("k_axial_bwd C32 L64 B256")
        fr = {"hbm": rb["hbm_frac"], "fp32_fma": rb["fp32_fma_frac"], "exp": rb["exp_frac"]}
        line["roofline"] = {"bound": "hbm", "kernel": "k_axial_bwd<2,true,2> (persistent cooperative backward of one axial op)",
                            "op": f"gated layer2.0 hight_block C{r['C']} L{r['L']} B{r['B']}",
                            "achieved": rb["achieved_GBps"], "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": rb["hbm_frac"],
                            "peak_source": f"MEASURED_PEAKS.json ({pk_kind})", "traffic": tr,
                            "alg_bytes_per_launch": rb["alg_bytes_per_launch"], "us_per_launch_cold_l2": rb["us_per_launch_cold_l2"],
                            "us_per_launch_warm_l2": rb["us_per_launch_warm_l2"],
                            "fractions_of_each_roof": fr, "binding": max(fr, key=fr.get) + " is the nearest roof, but none is reached at 128 px (SURVEY 8(d)): the kernel is "
                                       "instruction-fetch / latency-bound (ncu: stall_no_inst + barrier = 67 % of the warp samples on the small "
                                       "layers, DESIGN.md 5d); `bound` names the memory roof only because the contract allows hbm | tensor",
                            "forward_kernel": dict(kernel="k_axial_fwd<2,true,2>", **rf)}
        if args.configs and SZ == 128 and "512" in args.configs:
            # BASELINE config 5 geometry (L = 256): the exp roof binds there
            r5 = op_profile(dev, flush, "gated", 16, BS, 256, 256, 0, 1, reps=5)
            line["roofline"]["L256_op"] = {"op": f"512-px layer1.0 hight_block C16 L256 B{r5['B']}",
                                           "fwd": op_roofline(r5, pk, "fwd"), "bwd": op_roofline(r5, pk, "bwd")}
        # the tensor-core kernel of the step (csrc/conv_tc.cu, tcgen05 kind::tf32, 3xTF32 operand split), measured live
        line["roofline"]["tensor_core_kernel"] = conv_tc_profile(dev, flush, pk)
    except Exception as e:  # noqa: BLE001
        line["roofline"] = dict(line.get("roofline") or {}, error=repr(e))
    # ---- CPU baseline: the reference's path (oracle port) on this box's host cores, bounded sample
    log("cpu baseline, cores", os.cpu_count())
    if world == 1 and not args.no_cpu_baseline:
        try:
            cb = cpu_baseline_subprocess(args.model, SZ, 3 if SZ == 128 else 1, 1)
            cb["sample"] = f"{3 if SZ == 128 else 1} full train steps (after 1 warm-up) of the same bs=4 {SZ}x{SZ} batch, own process"
            line["cpu_baseline"] = cb
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"] = {"error": repr(e)[:300]}
    # ---- still the baseline leg: the same checker (oracle/torch_port.py, the reference's op sequence) run by stock PyTorch
    # eager fp32 on THIS GPU -- SURVEY 8(d) "reference eager on the same B200".  Context next to the CPU number, measured
    # last and fully guarded so that it can never disturb the numbers above; never the reference arm.
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"]["same_gpu_eager"] = reference_eager_gpu(args.model, dev)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"]["same_gpu_eager"] = {"error": repr(e)[:200]}
    print(json.dumps(line), flush=True)


def conv_tc_profile(dev, flush, pk, reps=10):
    """k_conv_tc3 through the C ABI (mt_conv2d_forward) on the stem's conv2 (8 -> 128 channels) at the 128-px step's 64 x 64 maps
    and at the 512-px step's 256 x 256 maps: CUDA events on the launching stream, L2 flushed before every launch.  A 3xTF32 product
    costs three kind::tf32 MMAs, so the tensor-pipe roof for fp32-grade flops is (dense TF32 peak = measured bf16 peak / 2) / 3."""
    from medt_b200 import _cabi
    lib = _cabi.load_library()
    st = torch.cuda.current_stream(dev).cuda_stream
    out = {"kernel": "k_conv_tc3<128> (3x3 conv forward: zero-padded slab, 27 tcgen05.mma per 8-channel stage, TMEM accumulators)",
           "fp32_grade_tensor_roof_TFLOPs": pk["bf16_tflops"] / 2 / 3,
           "ncu": "profiles/r3_conv_tc3_ncu_full.md (sm__pipe_tensor_cycles_active 12 - 14 % at BN = 128)"}
    for tag, h in (("conv2_64x64", 64), ("conv2_256x256", 256)):
        n, ci, co = BS, 8, 128
        x = torch.randn(n, ci, h, h, device=dev)
        w = torch.randn(co, ci, 3, 3, device=dev)
        y = torch.empty(n, co, h, h, device=dev)
        call = lambda: lib.mt_conv2d_forward(n, ci, co, h, h, 3, 3, 1, 1, x.data_ptr(), w.data_ptr(), y.data_ptr(), st)  # noqa: E731
        assert call() == 0
        torch.cuda.synchronize(dev)
        us = []
        for _ in range(reps):
            flush()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            call()
            b.record()
            torch.cuda.synchronize(dev)
            us.append(a.elapsed_time(b) * 1e3)
        t = sorted(us)[len(us) // 2]
        flops = 2.0 * n * h * h * co * ci * 9
        byts = 4.0 * (x.numel() + y.numel() + w.numel())
        out[tag] = {"us_per_launch_cold_l2": t, "fp32_grade_TFLOPs": flops / t / 1e6, "tensor_frac": flops / t / 1e6 / out["fp32_grade_tensor_roof_TFLOPs"],
                    "alg_bytes_per_launch": byts, "achieved_GBps": byts / t / 1e3, "hbm_frac": byts / t / 1e3 / pk["hbm_gbs"]}
    return out


def reference_eager_gpu(model_name, dev, steps=10, warmup=3):
    from oracle import torch_port
    model = build_model(model_name).to(dev).train()
    opt = torch.optim.Adam(list(model.parameters()), lr=1e-3, weight_decay=1e-5)
    x, y = synthetic_batch()
    x, y = x.to(dev), y.to(dev)
    for _ in range(warmup):
        torch_port.train_step(model, opt, x, y)
    torch.cuda.synchronize(dev)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        torch_port.train_step(model, opt, x, y)
    e1.record()
    torch.cuda.synchronize(dev)
    ms = e0.elapsed_time(e1) / steps
    return {"value": BS / (ms / 1e3), "unit": "img/s", "ms_per_step": ms, "steps": steps,
            "what": "oracle/torch_port.py (the reference's ATen op sequence) + torch.optim.Adam, stock PyTorch eager, "
                    "fp32 (TF32 off), batch resident, no L2 flush -- context, not the reference arm"}


def log(*a):
    if os.environ.get("BENCH_VERBOSE"):
        print(f"[bench {time.strftime('%H:%M:%S')}]", *a, file=sys.stderr, flush=True)


def main():
    if os.environ.get("BENCH_WATCHDOG"):
        import faulthandler
        faulthandler.dump_traceback_later(int(os.environ["BENCH_WATCHDOG"]), exit=True)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--model", default="gated", choices=["gated", "MedT", "logo"])
    ap.add_argument("--size", type=int, default=IMG, help="image side (128 = the BASELINE headline; 256 / 512 = configs 3 / 5)")
    ap.add_argument("--config", default=None, choices=sorted(WORKLOADS), help="shorthand for --model/--size of one BASELINE configuration")
    ap.add_argument("--configs", default=os.environ.get("BENCH_CONFIGS", "medt256,gated512,logo512"),
                    help="extra BASELINE configurations measured after the headline and mirrored under `also` ('' = none)")
    ap.add_argument("--single", action="store_true", help="only the --model network (skip the other headline model)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.config:
        args.model, args.size = WORKLOADS[args.config]
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_product(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
