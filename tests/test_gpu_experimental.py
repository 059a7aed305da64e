"""GPU tier, OPT-IN (MEDT_EXPERIMENTAL=1): the unmeasured k_bwd2 variants behind MT_DEBUG_FLAGS bits 3 / 4
(8 = register-tiled wrapped-diagonal d-relative pass, 16 = rows / columns in separate CTAs, 24 = both) must pass the
same op-level parity suite as the default kernels.  The flags are read once per process, hence the subprocesses.
Not part of the default run: these variants have not been on a GPU yet."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.skipif(os.environ.get("MEDT_EXPERIMENTAL") != "1", reason="opt-in: MEDT_EXPERIMENTAL=1")
@pytest.mark.parametrize("flags", [8, 16, 24])
def test_experimental_bwd_variants_pass_the_axial_parity_suite(flags):
    env = dict(os.environ, MT_DEBUG_FLAGS=str(flags))
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.join(HERE, "tests", "test_gpu_axial.py"), "-q", "-x", "-m", "gpu"],
                       env=env, capture_output=True, text=True, timeout=900, cwd=HERE)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-1000:]
