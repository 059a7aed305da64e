"""CPU tier: race check of the product kernels without a GPU -- the CPU-emulated library (tests/emu) under
ThreadSanitizer, one TSan fiber per GPU thread; __syncthreads, warp collectives, atomics and kernel boundaries are the
only happens-before edges, CTAs of one launch are unordered.  Any conflicting pair of accesses that nothing orders (a
missing barrier around shared-memory staging, two CTAs writing one global word, a finaliser reading what a sibling thread
updates) is a failure.  tests/emu/race_harness.cpp drives forward + backward of ten axial-op configurations (all layouts
and variants) and the shell kernels through the C ABI."""
import os
import shutil
import subprocess

import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def _has_tsan():
    if shutil.which("g++") is None:
        return False
    r = subprocess.run(["g++", "-print-file-name=libtsan.so"], capture_output=True, text=True)
    return os.path.isabs(r.stdout.strip()) and os.path.exists(r.stdout.strip())


@pytest.mark.skipif(not _has_tsan(), reason="needs g++ with libtsan")
def test_kernels_have_no_data_races_under_the_emulated_thread_model():
    r = subprocess.run([os.path.join(HERE, "emu", "run_tsan.sh")], capture_output=True, text=True, timeout=1500)
    log = "/tmp/medt_emu_tsan/log.txt"
    txt = open(log).read() if os.path.exists(log) else ""
    if "unexpected memory mapping" in txt or "FATAL: ThreadSanitizer" in txt:
        pytest.skip("ThreadSanitizer cannot run in this environment (address-space layout)")
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert "ALL DONE" in r.stdout and "race reports: 0" in r.stdout, r.stdout[-1500:] + txt[-3000:]
