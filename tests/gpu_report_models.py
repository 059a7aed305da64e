"""Developer aid: prints the measured parity numbers of every model golden (TF32 on/off)."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [os.path.join(HERE, "golden"), os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), "medical-transformer_b200"), HERE]
import torch  # noqa: E402
import test_gpu_models as T  # noqa: E402
for tf32 in (True, False):
    torch.backends.cudnn.allow_tf32 = tf32
    print("=== cudnn.allow_tf32 =", tf32)
    for tag, factory, S, N, modes in T.CASES:
        for mode in modes:
            try:
                r = T.measure(tag, factory, S, N, mode)
                print(tag, mode, {k: (f"{v:.3e}" if isinstance(v, float) else v) for k, v in r.items()})
            except Exception as e:  # noqa: BLE001
                import traceback; traceback.print_exc()
                print(tag, mode, "EXC", e)
