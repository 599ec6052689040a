import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_case_names():
    with open(os.path.join(GOLDEN_DIR, "CASES.txt")) as fh:
        return [ln.strip() for ln in fh if ln.strip()]


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN_DIR


def load_case(name):
    import numpy as np
    return np.load(os.path.join(GOLDEN_DIR, f"case_{name}.npz"))


def rel_err(got, ref):
    """The tolerance north_star states: |a - b| <= tol * max(1, |b|) (values cross zero after the mean)."""
    import numpy as np
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    return float(np.max(np.abs(got - ref) / np.maximum(1.0, np.abs(ref)))) if ref.size else 0.0
