"""CPU tier: the experimental register-tiled wrapped-diagonal pass for d relative
(medical-transformer_b200/csrc/axial_diag_tile.cuh, MT_DEBUG_FLAGS bit 3) is a __host__ __device__ routine; this
compiles it with g++ and checks it against the plain double loop over (i, j) for every (d, RT, L) shape class the
kernel instantiates, including padded pitches and the degenerate one-block lengths."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(shutil.which("g++") is None, reason="needs g++")
def test_tiled_diagonal_pass_matches_plain_definition(tmp_path):
    exe = str(tmp_path / "diag_tile_check")
    src = os.path.join(HERE, "tests", "host", "diag_tile_check.cpp")
    subprocess.run(["g++", "-O1", "-std=c++17", "-o", exe, src], check=True, timeout=300)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert r.returncode == 0 and len(lines) >= 12 and all(l.startswith("OK") for l in lines), r.stdout + r.stderr
