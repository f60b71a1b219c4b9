import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests need a CUDA device AND the built library; without either they skip (instead of failing from
    torch._C._cuda_init), so a CPU-only `pytest tests` reports the real state of the CPU suite."""
    import torch
    lib = os.path.join(ROOT, "volta_b200", "libvolta_b200.so")
    why = None
    if not torch.cuda.is_available():
        why = "needs a CUDA device"
    elif not os.path.exists(lib):
        why = "volta_b200/libvolta_b200.so is not built (python -m volta_b200.csrc.build)"
    if why is None:
        return
    skip = pytest.mark.skip(reason=why)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden_dir():
    return os.path.join(ROOT, "tests", "golden")
