import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
for p in (ROOT, os.path.join(ROOT, "epipolarpose_b200")):
    if p not in sys.path:
        sys.path.insert(0, p)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a B200 (sm_100a) device")


def pytest_collection_modifyitems(config, items):
    import torch
    if torch.cuda.is_available():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return dict(np.load(os.path.join(GOLDEN, name + ".npz")))
    return load


def relerr(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


# Marker for GPU tests of kernels whose arithmetic is validated on the CPU (tests/harness, emulated
# ABI) but that have not had a hardware run yet: skipped unless EPB_RUN_HW_PENDING=1.  Unused when
# every test has been run on a B200 (the case at the end of round 1).
hw_pending = pytest.mark.skipif(os.environ.get("EPB_RUN_HW_PENDING") != "1",
                                reason="hardware run pending: set EPB_RUN_HW_PENDING=1 on a B200")
