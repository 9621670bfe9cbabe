import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    """`-m gpu` tests are skipped (not errored) on a machine without a CUDA device, so a plain `pytest tests`
    stays green on CPU CI.  On a GPU box nothing is skipped: a missing libdsea.so must fail loudly there."""
    reason = None
    try:
        import torch
        if not torch.cuda.is_available():
            reason = "no CUDA device"
    except Exception as exc:                     # pragma: no cover
        reason = f"torch unavailable: {exc}"
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    def load(name):
        return np.load(os.path.join(GOLDEN, name), allow_pickle=False)
    return load


def f32round(x: float) -> float:
    """g as the reference's drivers build it: through a float32 tensor (E0.py:95, chiF.py:66)."""
    return float(np.float64(np.float32(x)))
