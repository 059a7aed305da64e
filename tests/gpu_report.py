"""Developer aid (not a test): runs every golden op fixture on the GPU and prints the whole
error table, so one gpurun call shows every mismatch instead of the first."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "golden"), os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "medical-transformer_b200"), HERE]
import torch  # noqa: E402
from helpers import op_fixture_names  # noqa: E402
import test_gpu_axial as T  # noqa: E402

if __name__ == "__main__":
    nbad = 0
    for name in op_fixture_names():
        for mode in ("train", "eval"):
            try:
                errs = T.run_case(name, mode)
            except Exception as e:  # noqa: BLE001
                print(f"{name:28s} {mode:5s} EXC {type(e).__name__}: {e}")
                nbad += 1
                continue
            bad = {k: v for k, v in errs.items() if not (v[0] <= v[1])}
            nbad += len(bad)
            worst = max(errs.items(), key=lambda kv: kv[1][0] / max(kv[1][1], 1e-30))
            print(f"{name:28s} {mode:5s} {'FAIL' if bad else 'ok  '} worst {worst[0]}={worst[1][0]:.2e}")
            for k, v in bad.items():
                print(f"      {k:45s} {v[0]:.3e} (tol {v[1]:.0e})")
    print("TOTAL BAD", nbad)
    torch.cuda.synchronize()
