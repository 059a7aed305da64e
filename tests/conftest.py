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


def pytest_collection_modifyitems(config, items):
    """`pytest tests` on a CPU-only box: the gpu-marked tests are skipped instead of failing at 'no NVIDIA driver'."""
    import pytest
    try:
        import torch
        has_cuda = torch.cuda.is_available()
    except Exception:  # noqa: BLE001
        has_cuda = False
    if has_cuda:
        return
    skip = pytest.mark.skip(reason="no CUDA device (gpu-marked tests need a B200)")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
