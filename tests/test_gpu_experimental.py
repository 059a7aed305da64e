"""GPU tier: every alternate code path of the axial op that ships in libmedt_b200.so must pass the same op-level parity suite
(tests/test_gpu_axial.py: reference goldens + oracle at the baseline sizes) as the default path.  The selectors are read once
per process, hence the subprocesses.

  default                 single-sweep cooperative backward where the planner picks it, multi-sweep cooperative elsewhere
  MT_BWD_NOFUSE=1         multi-sweep cooperative backward everywhere (the round-2a kernel; serves d = 4 / 8 and short strips)
  MT_BWD_FUSED_ALL=1      single-sweep backward wherever it is applicable (d <= 4, power-of-two lengths), incl. 1 x 1 tiles
  MT_COOP_STREAM=1        cooperative kernels in streamed (non-resident) mode, as at 512 px
  MT_DEBUG_FLAGS=32       round-1 launch chain (what shapes the cooperative planners reject fall back to)
  MT_DEBUG_FLAGS=40/48/56 launch chain with the tiled wrapped-diagonal d-relative pass / split row-column CTAs / both"""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("env", [{"MT_BWD_NOFUSE": "1"}, {"MT_BWD_FUSED_ALL": "1"}, {"MT_COOP_STREAM": "1", "MT_BWD_FUSED_ALL": "1"},
                                 {"MT_DEBUG_FLAGS": "32"}, {"MT_DEBUG_FLAGS": "40"}, {"MT_DEBUG_FLAGS": "48"}, {"MT_DEBUG_FLAGS": "56"}],
                         ids=lambda e: ",".join(f"{k}={v}" for k, v in e.items()))
def test_alternate_paths_pass_the_axial_parity_suite(env):
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "tests", "test_gpu_axial.py"), "-q", "-x", "-m", "gpu"],
                       env=dict(os.environ, **env), capture_output=True, text=True, timeout=900, cwd=HERE)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
