"""pytest config: registers the `gpu` marker and puts the product package dir, the oracle
and the golden helpers on sys.path.  `-m "not gpu"` runs on a CPU-only box; `-m gpu`
needs a B200 and calls the product strictly through the C-ABI library."""
import os
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(REPO, "medical-transformer_b200")
for p in (os.path.join(REPO, "tests", "golden"), REPO, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
