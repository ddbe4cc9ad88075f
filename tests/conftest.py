import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as data:
        return {k: data[k] for k in data.files}


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_close_scaled(actual, expected, rel=1e-12):
    """North-star tolerance: max abs error <= rel * max|expected| (FP64)."""
    actual = np.asarray(actual)
    expected = np.asarray(expected)
    assert actual.shape == expected.shape, f"shape {actual.shape} != {expected.shape}"
    scale = max(float(np.abs(expected).max()), 1e-300)
    err = float(np.abs(actual - expected).max())
    assert err <= rel * scale, f"max abs error {err:.3e} > {rel:g} * max|ref| = {rel * scale:.3e}"
