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
def cpu_reference_steps(model_name, steps, warmup):
    """The reference's CPU path via its pinned restatement (oracle/torch_port.py)."""
    from oracle import torch_port
    model = build_model(model_name).train()
    opt = torch.optim.Adam(list(model.parameters()), lr=1e-3, weight_decay=1e-5)
    x, y = synthetic_batch()
    # use the host thread count that is FASTEST for the reference's op mix: its ~1.2k tiny ATen ops
    # per step get slower, not faster, when oversubscribed (128 threads: > 60 s/step on the GPU host)
    cores = os.cpu_count() or 1
    best, best_t = None, float("inf")
    for nt in [c for c in (8, 16, 32, 64, 128, 256) if c <= cores] or [cores]:
        torch.set_num_threads(nt)
        if best is None:
            torch_port.train_step(model, opt, x, y)            # first-touch warm-up
        t0 = time.perf_counter()
        torch_port.train_step(model, opt, x, y)
        dt = time.perf_counter() - t0
        if dt < best_t:
            best, best_t = nt, dt
        elif dt > 1.5 * best_t:
            break
    torch.set_num_threads(best)
    cpu_reference_steps.threads = best
    for _ in range(max(warmup - 1, 0)):
        torch_port.train_step(model, opt, x, y)
    t0 = time.perf_counter()
    for _ in range(steps):
        torch_port.train_step(model, opt, x, y)
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


def run_reference(args, rank):
    if rank != 0:
        return
    ips, ms = cpu_reference_steps(args.model, args.steps, min(args.warmup, 1))
    cores = os.cpu_count() or 1
    line = {"impl": "reference", "metric": METRIC, "value": ips, "unit": "img/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": min(args.warmup, 1), "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{'gatedaxialunet' if args.model == 'gated' else args.model} img_size=128 batch_size=4/GPU "
                                   "train step (fwd+CE+bwd+Adam)", "global_batch": BS, "parallelism": "cpu",
                       "impl_note": "reference CPU path: its ATen op sequence restated in oracle/torch_port.py (the reference is "
                                    "pure Python without setup.py, so it cannot be pip-installed onto the box)"},
            "cpu_baseline": {"value": ips, "unit": "img/s", "cores": cpu_reference_steps.threads, "host_cores": cores,
                             "kind": "port", "sample": f"{args.steps} full train steps of the bs=4 batch",
                             "cpu": cpu_model_string()},
            "e2e": {"value": ips, "unit": "img/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


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


# the 8 stages of one axial op as the ABI profiling hooks launch them (forward mask bits 0-3, backward 0-3)
STAGE_NAMES = ["k_proj", "k_attn2<stats>", "k_attn2<softmax+PV>", "k_out", "k_bwd_gather", "k_bwd2<T>", "k_bwd2<grads>",
               "k_bwd_dx+k_bwd_dw"]


def stage_profile(model, dev, flush):
    """Live per-kernel durations of the heaviest axial op (layer1.0 hight_block: C16, L64, B256):
    each of its 8 kernels launched alone through the ABI profiling hook, CUDA events on the
    launching stream, L2 flushed between launches."""
    from medt_b200 import _cabi
    from medt_b200.axial_op import _desc, _ptr
    lib_ = _cabi.load_library()
    mod = model.layer1[0].hight_block
    C_, L_ = mod.in_planes, mod.kernel_size
    x = torch.randn(BS, C_, L_, L_, device=dev)
    dout = torch.randn(BS, C_, L_, L_, device=dev)
    desc = _desc(x, 0, 1, mod._variant, True, 1, 1e-5, 0.1)
    bq, bs, bo = mod.bn_qkv, mod.bn_similarity, mod.bn_output
    rm = [t.clone() for t in (bq.running_mean, bq.running_var, bs.running_mean, bs.running_var, bo.running_mean, bo.running_var)]
    g = mod._variant == "gated"
    params = _cabi.MtAxialParams(_ptr(mod.qkv_transform.weight), _ptr(mod.relative), _ptr(mod.f_qr if g else None),
                                 _ptr(mod.f_kr if g else None), _ptr(mod.f_sv if g else None), _ptr(mod.f_sve if g else None),
                                 _ptr(bq.weight), _ptr(bq.bias), _ptr(rm[0]), _ptr(rm[1]),
                                 _ptr(bs.weight), _ptr(bs.bias), _ptr(rm[2]), _ptr(rm[3]),
                                 _ptr(bo.weight), _ptr(bo.bias), _ptr(rm[4]), _ptr(rm[5]))
    saved = torch.empty(lib_.mt_axial_saved_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
    ws = torch.empty(lib_.mt_axial_bwd_workspace_bytes(ctypes.byref(desc)), dtype=torch.uint8, device=dev)
    out = torch.empty_like(x)
    grads_t = [torch.empty_like(t) for t in (x, mod.qkv_transform.weight, mod.relative)] + [torch.empty(4, device=dev)] + \
              [torch.empty_like(t) for t in (bq.weight, bq.bias, bs.weight, bs.bias, bo.weight, bo.bias)]
    grads = _cabi.MtAxialGrads(*[_ptr(t) for t in grads_t])
    st = ctypes.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    _cabi.check(lib_.mt_axial_forward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(out), _ptr(saved), st), "fwd")
    _cabi.check(lib_.mt_axial_backward(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(dout), _ptr(saved),
                                       _ptr(ws), ctypes.byref(grads), st), "bwd")
    names = STAGE_NAMES
    res = {}
    for idx, name in enumerate(names):
        def launch():
            if idx < 4:
                lib_.mt_axial_forward_stages(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(out), _ptr(saved), st, 1 << idx)
            else:
                lib_.mt_axial_backward_stages(ctypes.byref(desc), ctypes.byref(params), _ptr(x), _ptr(dout), _ptr(saved),
                                              _ptr(ws), ctypes.byref(grads), st, 1 << (idx - 4))
        for _ in range(3):
            launch()
        res[name] = timed_steps(launch, 20, flush, dev) / 20 * 1e3          # us, cold L2
        res[name + "_warm"] = timed_steps(launch, 20, lambda: None, dev) / 20 * 1e3
    B = BS * L_
    shape = {"C": C_, "L": L_, "B": B}
    return res, shape


def run_product(args, rank, world, local_rank):
    import torch.distributed as dist
    from medt_b200 import _cabi
    from medt_b200.train_step import TrainStep
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    torch.backends.cudnn.benchmark = True          # fixed shapes: let cuDNN pick its fastest fp32 algorithms
    pk, pk_kind = peaks()
    flush = L2Flusher(dev)
    x_host, y_host = synthetic_batch()
    x_pin, y_pin = x_host.pin_memory(), y_host.pin_memory()

    def one_model(name):
        log("building", name)
        model = build_model(name).to(dev).train()
        n0 = _cabi.LAUNCHES["kernels"]
        ts = TrainStep(model, (BS, 3, IMG, IMG), lr=1e-3, weight_decay=1e-5, device=dev, use_graph=True)
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
        log("timed", name, ms / args.steps, "ms/step")
        return dict(model=model, ts=ts, ms=ms, ms_b2b=ms_b2b, ms_e2e=ms_e2e, launches=per_step_launches,
                    clocks=cs.summary(), loss=loss)

    main = one_model(args.model)
    other_name = "MedT" if args.model == "gated" else "gated"
    other = one_model(other_name) if not args.single else None

    if rank != 0:
        return
    K = args.steps
    ms_step = main["ms"] / K
    value = world * BS * K / (main["ms"] / 1e3)
    line = {"metric": METRIC, "value": value, "unit": "img/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic",
            "config": {"workload": f"{'gatedaxialunet' if args.model == 'gated' else args.model} img_size=128 "
                                   f"batch_size=4/GPU train step (fwd+CE+bwd+{'NCCL all-reduce+' if world > 1 else ''}Adam)",
                       "global_batch": BS * world, "parallelism": f"dp{world}", "l2": "flushed (256 MiB write) between timed steps",
                       "ms_per_step_back_to_back_no_flush": main["ms_b2b"] / K, "final_loss": main["loss"],
                       "cuda_graph": True, "cudnn_tf32": bool(torch.backends.cudnn.allow_tf32)},
            "e2e": {"value": world * BS * K / (main["ms_e2e"] / 1e3), "unit": "img/s",
                    "h2d_bytes_per_step": main["ts"].h2d_bytes_per_step(), "d2h_bytes_per_step": main["ts"].d2h_bytes_per_step()},
            "gpu_launches": main["launches"] * K, "gpu_launches_per_step": main["launches"],
            "clocks": main["clocks"]}
    if other is not None:
        line["also"] = {other_name: {"value": world * BS * K / (other["ms"] / 1e3), "unit": "img/s",
                                     "ms_per_step": other["ms"] / K,
                                     "e2e": world * BS * K / (other["ms_e2e"] / 1e3),
                                     "gpu_launches_per_step": other["launches"]}}
    # ---- roofline of the dominant hand-written kernel (heaviest op of the step), measured live
    log("stage profile")
    try:
        gm = main["model"] if args.model == "gated" else other["model"]
        st_us, shp = stage_profile(gm, dev, flush)
        names = STAGE_NAMES
        dom = max(names, key=lambda n: st_us[n])
        B, C, L = shp["B"], shp["C"], shp["L"]
        # algorithmic bytes per launch of each kernel (DESIGN.md "kernels"), fp32
        alg = dict(zip(STAGE_NAMES, [
            4 * (B * C * L + B * 2 * C * L),                       # x in, y_raw out
            4 * B * 2 * C * L,                                     # y_raw in
            4 * (B * 2 * C * L * 2 + B * 8 * L),                   # y_raw in, o_raw + lse out
            4 * (B * 2 * C * L + B * C * L),                       # o_raw in, out
            4 * (B * C * L * 2 + B * 2 * C * L),                   # dout in, g_s out, o_raw in
            4 * (B * 2 * C * L * 2 + B * C * L + B * 8 * L),       # y_raw, o_raw, g_s, lse in
            4 * (B * 2 * C * L * 3 + B * C * L + B * 8 * L),       # ... + dyh out
            4 * (B * 2 * C * L * 2 + B * C * L * 2)]))             # dyh, y_raw, x in, dx out
        traffic = None
        tpath = os.path.join(REPO, "profiles", "top_kernel_traffic.json")
        if os.path.exists(tpath):
            traffic = json.load(open(tpath)).get(dom)
        ach = alg[dom] / (st_us[dom] * 1e-6) / 1e9
        line["roofline"] = {"bound": "hbm", "kernel": dom, "op": f"gated layer1.0 hight_block C{C} L{L} B{B}",
                            "achieved": ach, "peak": pk["hbm_gbs"], "unit": "GB/s", "frac": ach / pk["hbm_gbs"],
                            "peak_source": f"MEASURED_PEAKS.json ({pk_kind})", "traffic": traffic,
                            "alg_bytes_per_launch": alg[dom], "us_per_launch_cold_l2": st_us[dom],
                            "us_per_launch_warm_l2": st_us[dom + "_warm"],
                            "exp_frac_of_4.5T_per_s": (8 * B * L * L / (st_us["k_attn2<softmax+PV>"] * 1e-6)) / 4.5e12,
                            "all_kernels_us_cold": {n: round(st_us[n], 2) for n in names},
                            "all_kernels_us_warm": {n: round(st_us[n + "_warm"], 2) for n in names}}
    except Exception as e:  # noqa: BLE001
        line["roofline"] = {"error": repr(e)}
    # ---- CPU baseline: the reference's path (oracle port) on this box's host cores, bounded sample
    log("cpu baseline, cores", os.cpu_count())
    if world == 1 and not args.no_cpu_baseline:
        ips, ms_cpu = cpu_reference_steps(args.model, 3, 1)
        line["cpu_baseline"] = {"value": ips, "unit": "img/s", "cores": cpu_reference_steps.threads,
                                "host_cores": os.cpu_count() or 1, "kind": "port",
                                "sample": "3 full train steps (after 1 warm-up) of the same bs=4 128x128 batch",
                                "ms_per_step": ms_cpu, "cpu": cpu_model_string()}
    # ---- still the baseline leg: the same checker (oracle/torch_port.py, the reference's op sequence) run by stock PyTorch
    # eager fp32 on THIS GPU -- SURVEY 8(d) "reference eager on the same B200".  Context next to the CPU number, measured
    # last and fully guarded so that it can never disturb the numbers above; never the reference arm.
    if world == 1 and not args.no_cpu_baseline:
        try:
            line["cpu_baseline"]["same_gpu_eager"] = reference_eager_gpu(args.model, dev)
        except Exception as e:  # noqa: BLE001
            line["cpu_baseline"]["same_gpu_eager"] = {"error": repr(e)[:200]}
    print(json.dumps(line), flush=True)


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
    ap.add_argument("--model", default="gated", choices=["gated", "MedT"])
    ap.add_argument("--single", action="store_true", help="only the --model network (skip the other headline model)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
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
