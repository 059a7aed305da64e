#!/usr/bin/env python
"""Recipe for the reference arm: put the UNMODIFIED reference (pure Python; it has no setup.py / pyproject, so
`pip install --target baseline/_ref /root/reference` fails with "not installable") under baseline/_ref/ by a plain copy of
its tree.  baseline/_ref/ is git-ignored (never committed) but NOT gpurun-ignored, so it travels to the GPU box with the
snapshot; `bench.py --impl reference`, `tests/test_gpu_reference_script.py` and nothing else read it.

    python baseline/make_ref.py [/root/reference]

Called by __graft_entry__.build() whenever /root/reference exists (i.e. in the development container)."""
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))


def make_ref(src="/root/reference"):
    dst = os.path.join(HERE, "_ref")
    if not os.path.isdir(src):
        return os.path.isdir(dst)
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    shutil.copytree(src, dst, ignore=shutil.ignore_patterns(".git", "__pycache__", "*.pyc"))
    return True


if __name__ == "__main__":
    ok = make_ref(*(sys.argv[1:2]))
    print("baseline/_ref", "ready" if ok else "unavailable (no reference tree here)")
