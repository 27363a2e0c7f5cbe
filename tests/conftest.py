"""pytest configuration: the ``gpu`` marker and shared fixtures."""
import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if REPO not in sys.path:
    sys.path.insert(0, REPO)
GOLDEN = os.path.join(REPO, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_sessionstart(session):
    """Build libdic_b200.so when it is missing or older than its sources (the .so is git-ignored; nvcc
    cross-compiles without a GPU).  A failed build is reported by the tests that need the library."""
    try:
        from deep_interpolation_clustering_b200 import build as _build
        _build.build()
    except Exception as e:          # noqa: BLE001
        sys.stderr.write(f"[conftest] could not build libdic_b200.so: {e}\n")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name + ".npz")) as f:
        return {k: f[k] for k in f.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden
