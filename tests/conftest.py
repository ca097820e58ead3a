import os
import sys

import numpy as np
import pytest

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(REPO, "tests", "golden")
for p in (REPO, GOLDEN, os.path.join(REPO, "tests")):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line(
        "markers", "gpu: needs a CUDA device (run on the B200 box)")
    config.addinivalue_line(
        "markers", "slow: takes more than a few seconds on CPU")


@pytest.fixture(scope="session")
def golden():
    """Loader for tests/golden/*.npz (made by tests/golden/make_golden.py
    from the unmodified reference)."""
    cache = {}

    def load(name):
        if name not in cache:
            cache[name] = np.load(os.path.join(GOLDEN, name + ".npz"))
        return cache[name]
    return load


@pytest.fixture()
def workdir(tmp_path, monkeypatch):
    """Run in a scratch directory: runtime-compiled custom kernels land in
    ./tmp of the current directory (as in the reference)."""
    monkeypatch.chdir(tmp_path)
    return tmp_path


def rel_l2(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.linalg.norm(b.ravel())
    return np.linalg.norm((a - b).ravel()) / (den if den > 0 else 1.0)
