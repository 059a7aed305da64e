#!/usr/bin/env python
"""Run an UNMODIFIED reference script (train.py / test.py) on the B200 path.

    python medical-transformer_b200/run_reference_script.py /path/to/Medical-Transformer/train.py \
        --train_dataset ... --val_dataset ... --modelname gatedaxialunet --imgsize 128 --batch_size 4 ...

`python /path/to/train.py` would put the reference directory first on sys.path, so its own
`lib` package would win.  This launcher instead puts THIS directory (the drop-in `lib`, `medt_b200`)
first, appends the reference directory (for the script's bare `import utils / metrics /
utils_gray`, train.py:22-24,74,77) and the empty `matplotlib` / `skimage` stubs if those import-only
packages are missing, then executes the script with runpy (which does not touch sys.path).
"""
import importlib.util
import os
import runpy
import sys


def setup_paths(script_path):
    here = os.path.dirname(os.path.abspath(__file__))
    ref_dir = os.path.dirname(os.path.abspath(script_path))
    sys.path[:] = [p for p in sys.path if os.path.abspath(p or ".") != ref_dir]
    sys.path.insert(0, here)
    sys.path.append(ref_dir)
    for mod in ("matplotlib", "skimage"):
        if importlib.util.find_spec(mod) is None:
            stubs = os.path.join(here, "stubs")
            if stubs not in sys.path:
                sys.path.append(stubs)
    return here, ref_dir


def main():
    if len(sys.argv) < 2:
        sys.exit(__doc__)
    script = sys.argv[1]
    setup_paths(script)
    sys.argv = [script] + sys.argv[2:]
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
