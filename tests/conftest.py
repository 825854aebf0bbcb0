import os
import sys

import numpy as np
import pytest

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


@pytest.fixture(scope="session")
def golden():
    return load_golden


def assert_logodds_close(got, ref, tol=1e-9, what=""):
    """Parity criterion of SURVEY.md 8(d): identical -inf/NaN pattern and
    |delta| <= tol * max(1, |ref|) elsewhere."""
    got = np.asarray(got, dtype=float)
    ref = np.asarray(ref, dtype=float)
    assert got.shape == ref.shape, (what, got.shape, ref.shape)
    fin = np.isfinite(ref)
    assert np.array_equal(np.isfinite(got), fin), what + ": finite pattern differs"
    assert np.array_equal(got[~fin] == -np.inf, ref[~fin] == -np.inf), what + ": -inf pattern differs"
    if fin.any():
        err = np.abs(got[fin] - ref[fin]) / np.maximum(1.0, np.abs(ref[fin]))
        assert err.max() <= tol, "%s: max scaled err %.3e > %.1e" % (what, err.max(), tol)
        return float(err.max())
    return 0.0
