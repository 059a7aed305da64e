"""GPU tier, boundary row (SURVEY 8b / VERDICT r1 items 5-6): the UNMODIFIED reference scripts and checkpoints against the drop-in.

  * the reference's own train.py (baseline/_ref/train.py, byte-identical copy made by baseline/make_ref.py) is EXECUTED on
    the B200 through run_reference_script.py with the drop-in `lib`, on a synthetic PNG dataset, and -- same command line,
    same seeds -- with the reference's own `lib` (stock PyTorch eager on the same GPU); losses, checkpoints and written
    masks of the two runs are compared;
  * a checkpoint WRITTEN BY THE REFERENCE loads into the drop-in (and back, strict), eval logits agree with the reference's
    own eval forward, masks are bit-exact;
  * predict.py (the runnable equivalent of the reference's test.py) writes exactly the reference's masks.
baseline/_ref is git-ignored but travels with the gpurun snapshot; the tests skip (loudly) when it is absent."""
import json
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "medical-transformer_b200")
REF = os.path.join(REPO, "baseline", "_ref")
RUNNER = os.path.join(REPO, "tests", "ref_runner.py")
needs_ref = pytest.mark.skipif(not os.path.isfile(os.path.join(REF, "train.py")),
                               reason="baseline/_ref absent: run `python baseline/make_ref.py` where /root/reference exists")


def make_dataset(root, n, size, seed=11):
    """<root>/img/*.png (3-channel) and <root>/labelcol/*.png (0/255 masks): the layout utils.py:127-129 reads"""
    import cv2
    rs = np.random.RandomState(seed)
    os.makedirs(os.path.join(root, "img"), exist_ok=True)
    os.makedirs(os.path.join(root, "labelcol"), exist_ok=True)
    yy, xx = np.mgrid[0:size, 0:size]
    for i in range(n):
        cx, cy, r = rs.randint(size // 4, 3 * size // 4, 2).tolist() + [rs.randint(size // 8, size // 3)]
        blob = ((xx - cx) ** 2 + (yy - cy) ** 2 < r * r)
        img = (rs.rand(size, size, 3) * 90 + blob[:, :, None] * 120).astype(np.uint8)
        cv2.imwrite(os.path.join(root, "img", f"{i:03d}.png"), img)
        cv2.imwrite(os.path.join(root, "labelcol", f"{i:03d}.png"), (blob * 255).astype(np.uint8))


def run(cmd, env=None, cwd=None, timeout=900):
    e = dict(os.environ)
    e.update(env or {})
    r = subprocess.run(cmd, capture_output=True, text=True, env=e, cwd=cwd, timeout=timeout)
    assert r.returncode == 0, f"{' '.join(cmd)}\n--- stdout\n{r.stdout[-3000:]}\n--- stderr\n{r.stderr[-3000:]}"
    return r.stdout


def train_args(data, out, modelname, epochs=1):
    return ["--train_dataset", data, "--val_dataset", data, "--direc", out, "--batch_size", "4", "--epochs", str(epochs), "--save_freq", "1",
            "--modelname", modelname, "--learning_rate", "0.001", "--imgsize", "128", "--gray", "no"]


def stub_path():
    import importlib.util
    need = any(importlib.util.find_spec(m) is None for m in ("matplotlib", "skimage"))
    return os.path.join(PKG, "stubs") if need else ""


@needs_ref
@pytest.mark.parametrize("modelname", ["gatedaxialunet", "MedT"])
def test_unmodified_train_py_runs_on_the_dropin_and_tracks_the_reference(tmp_path, modelname):
    data = str(tmp_path / "data")
    make_dataset(data, 8, 128)
    report = str(tmp_path / "launches.json")
    # train.py builds the model BEFORE it seeds anything (train.py:95-121) and this PyTorch seeds its default generator from the OS,
    # so two runs start from different weights.  Both processes get the same start through a sitecustomize.py on PYTHONPATH
    # (test harness only; train.py itself stays byte-identical).
    site = tmp_path / "site"
    site.mkdir()
    (site / "sitecustomize.py").write_text("import torch\ntorch.manual_seed(20260923)\n")
    # --- the reference's train.py, drop-in `lib` (this repo)
    out_a = str(tmp_path / "dropin")
    so = run([sys.executable, os.path.join(PKG, "run_reference_script.py"), os.path.join(REF, "train.py")] + train_args(data, out_a, modelname),
             env={"MEDT_LAUNCH_REPORT": report, "CUDA_VISIBLE_DEVICES": "0", "PYTHONPATH": str(site)})
    rep = json.load(open(report))
    assert rep["kernels_launched"] > 0 and rep["library"].endswith("libmedt_b200.so"), rep     # the sm_100a kernels did the work
    loss_a = float(re.search(r"epoch \[0/1\], loss:([0-9.]+)", so).group(1))
    # --- the same script and command line with the reference's own `lib` (stock eager PyTorch on the same GPU)
    out_b = str(tmp_path / "reference")
    sp = stub_path()
    so_b = run([sys.executable, "train.py"] + train_args(data, out_b, modelname), cwd=REF,
               env={"PYTHONPATH": str(site) + (os.pathsep + sp if sp else ""), "CUDA_VISIBLE_DEVICES": "0", "NVIDIA_TF32_OVERRIDE": "0"})
    loss_b = float(re.search(r"epoch \[0/1\], loss:([0-9.]+)", so_b).group(1))
    assert re.search(r"Total_params: (\d+)", so).group(1) == re.search(r"Total_params: (\d+)", so_b).group(1)
    # two Adam steps from the same (default-seeded) init on the same shuffled batches: the epoch-mean loss is printed with 4
    # decimals; fp32 train-mode noise after one update is ~1e-3 (tests/golden/adam2_*.npz)
    assert abs(loss_a - loss_b) < 2e-2 * max(1.0, loss_b), (loss_a, loss_b)
    # checkpoints: same files, same schema, both load strictly into the other implementation
    for d in (out_a, out_b):
        assert os.path.isfile(d + "final_model.pth") and os.path.isfile(os.path.join(d, "0", modelname + ".pth"))
    sd_a, sd_b = torch.load(out_a + "final_model.pth", map_location="cpu"), torch.load(out_b + "final_model.pth", map_location="cpu")
    assert list(sd_a.keys()) == list(sd_b.keys())
    assert all(sd_a[k].shape == sd_b[k].shape and sd_a[k].dtype == sd_b[k].dtype for k in sd_a)
    step = max(float((sd_a[k].float() - sd_b[k].float()).abs().max()) for k in sd_a if sd_a[k].dtype.is_floating_point and "running" not in k)
    assert step <= 2 * 2 * 1e-3 * 1.05, step          # after 2 Adam steps no weight can differ by more than 2 x (2 lr)
    run([sys.executable, RUNNER, "load_strict", modelname, "128", out_a + "final_model.pth"], env={"CUDA_VISIBLE_DEVICES": ""})
    # validation masks written by train.py:176-213 (train-mode forward of single images): same files, nearly the same pixels
    import cv2
    names = sorted(os.listdir(os.path.join(out_a, "0")))
    assert names == sorted(os.listdir(os.path.join(out_b, "0"))) and len([n for n in names if n.endswith(".png")]) == 8
    diff = [float((cv2.imread(os.path.join(out_a, "0", n), 0) != cv2.imread(os.path.join(out_b, "0", n), 0)).mean()) for n in names if n.endswith(".png")]
    try:
        os.makedirs(os.path.join(REPO, "gpurun_out"), exist_ok=True)
        with open(os.path.join(REPO, "gpurun_out", f"reference_script_{modelname}.json"), "w") as f:
            json.dump({"loss_dropin": loss_a, "loss_reference_eager_same_gpu": loss_b, "launch_report": rep, "val_mask_mismatch_frac": diff,
                       "max_weight_difference_after_2_steps": step}, f, indent=1)
    except OSError:
        pass
    assert float(np.mean(diff)) < 0.02, diff


@needs_ref
@pytest.mark.parametrize("modelname,size", [("gatedaxialunet", 128), ("MedT", 128), ("logo", 128), ("axialunet", 128)])
def test_reference_written_checkpoint_round_trip(tmp_path, modelname, size):
    from medt_b200.infer import load_checkpoint
    sys.path.insert(0, PKG)
    import predict
    ckpt = str(tmp_path / "ref.pth")
    run([sys.executable, RUNNER, "write_ckpt", modelname, str(size), "77", ckpt], env={"CUDA_VISIBLE_DEVICES": ""})
    x = np.random.RandomState(5).rand(2, 3, size, size).astype(np.float32)
    np.save(str(tmp_path / "x.npy"), x)
    run([sys.executable, RUNNER, "eval_logits", modelname, str(size), ckpt, str(tmp_path / "x.npy"), str(tmp_path / "ref_logits.npy"), "cpu"],
        env={"CUDA_VISIBLE_DEVICES": ""})
    ref = np.load(str(tmp_path / "ref_logits.npy")).astype(np.float64)
    m = predict.build(modelname, size, 3).cuda()
    load_checkpoint(m, ckpt)                                  # strict
    m.eval()
    with torch.no_grad():
        got = m(torch.from_numpy(x).cuda()).cpu().numpy().astype(np.float64)
    rel = np.linalg.norm(got - ref) / np.linalg.norm(ref)
    assert rel < 1e-4, rel
    # bit-exact masks against the reference's fp32 CPU forward, except where its logit sits within fp32 rounding of the threshold
    close = np.abs(ref[:, 1] - 0.5) < 1e-4 * max(1.0, np.abs(ref).max())
    assert ((got[:, 1] >= 0.5) == (ref[:, 1] >= 0.5))[~close].all()
    assert ((got[:, 1] >= 0.5) != (ref[:, 1] >= 0.5)).mean() <= 2.0 / close.size
    # and back: a checkpoint written by the drop-in (with DataParallel's prefix, as train.py writes it on >1 GPU) loads strictly
    # into the reference, and unprefixed too
    back = str(tmp_path / "back.pth")
    torch.save(m.state_dict(), back)
    run([sys.executable, RUNNER, "load_strict", modelname, str(size), back], env={"CUDA_VISIBLE_DEVICES": ""})
    torch.save({"module." + k: v for k, v in m.state_dict().items()}, back)
    m2 = predict.build(modelname, size, 3).cuda()
    load_checkpoint(m2, back)
    assert all(torch.equal(a, b) for a, b in zip(m.state_dict().values(), m2.state_dict().values()))


@needs_ref
@pytest.mark.parametrize("graph", [False, True])
def test_predict_cli_writes_the_references_masks(tmp_path, graph):
    """predict.py == what test.py:86-146 would write if it ran (it dies at test.py:62): per image, eval logits of the reference
    from the same checkpoint, thresholded with the reference's numpy lines."""
    import cv2
    data = str(tmp_path / "data")
    make_dataset(data, 5, 128, seed=3)
    ckpt = str(tmp_path / "ref.pth")
    run([sys.executable, RUNNER, "write_ckpt", "gatedaxialunet", "128", "78", ckpt], env={"CUDA_VISIBLE_DEVICES": ""})
    out = str(tmp_path / "results")
    run([sys.executable, os.path.join(PKG, "predict.py"), "--loaddirec", ckpt, "--val_dataset", data, "--direc", out, "--batch_size", "1",
         "--modelname", "gatedaxialunet", "--imgsize", "128", "--gray", "no"] + (["--graph"] if graph else []))
    names = sorted(os.listdir(os.path.join(data, "img")))
    assert sorted(os.listdir(out)) == names
    x = np.stack([cv2.imread(os.path.join(data, "img", n)).transpose(2, 0, 1).astype(np.float32) / 255.0 for n in names])
    bad = 0
    for i, n in enumerate(names):                             # batch of 1, as test.py's valloader
        np.save(str(tmp_path / "x.npy"), x[i:i + 1])
        run([sys.executable, RUNNER, "eval_logits", "gatedaxialunet", "128", ckpt, str(tmp_path / "x.npy"), str(tmp_path / "l.npy"), "cpu"],
            env={"CUDA_VISIBLE_DEVICES": ""})
        tmp = np.load(str(tmp_path / "l.npy"))
        tmp[tmp >= 0.5] = 1
        tmp[tmp < 0.5] = 0
        yHaT = tmp.astype(int)
        yHaT[yHaT == 1] = 255
        want = yHaT[0, 1, :, :].astype(np.uint8)
        got = cv2.imread(os.path.join(out, n), 0)
        bad += int((got != want).sum())
    assert bad <= 2, bad                                       # eval mode: bit-exact up to a logit within rounding of 0.5
