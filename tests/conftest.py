import os
import sys
from pathlib import Path

import pytest

REPO = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(REPO))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def product_lib():
    """The in-tree C-ABI library, built on demand (nvcc cross-compiles without a GPU)."""
    from cadrays_b200 import build, _ffi
    build.build_library()
    return _ffi.load_library()


@pytest.fixture(scope="session")
def oracle_lib():
    from oracle import oracle_ffi
    oracle_ffi.build_oracle()
    return oracle_ffi.lib()


def _has_gpu() -> bool:
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if _has_gpu():
        return
    skip = pytest.mark.skip(reason="no CUDA device in this container")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)
