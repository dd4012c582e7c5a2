import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


@pytest.fixture(scope="session")
def weights():
    return load_golden("weights_seed0.npz")


@pytest.fixture(scope="session")
def built_lib():
    """libfluxgnn.so, built on demand (nvcc cross-compiles without a GPU)."""
    from gnn_plasma_flux_b200 import _lib, build
    if not os.path.exists(_lib.LIB_PATH):
        build.build_library()
    return _lib.lib()
