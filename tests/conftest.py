import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")


def load_golden(name):
    """-> dict of numpy arrays from tests/golden/<name>.npz (made by tests/golden/make_golden.py)."""
    with np.load(os.path.join(GOLDEN_DIR, name + ".npz")) as z:
        return {k: z[k] for k in z.files}


def rel_err(a, b):
    """(norm-relative, max-normalised) error of a against reference b (SURVEY 8c metric)."""
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    nb = np.linalg.norm(b)
    mb = np.abs(b).max() if b.size else 0.0
    d = a - b
    return (np.linalg.norm(d) / nb if nb > 0 else np.linalg.norm(d),
            np.abs(d).max() / mb if mb > 0 else (np.abs(d).max() if d.size else 0.0))


def assert_close(a, b, tol, what=""):
    assert np.shape(a) == np.shape(b), f"{what}: shape {np.shape(a)} vs {np.shape(b)}"
    e2, em = rel_err(a, b)
    assert e2 <= tol and em <= tol, f"{what}: rel-l2 {e2:.3e}, max-norm {em:.3e} > {tol:.1e}"


@pytest.fixture(scope="session")
def golden():
    return load_golden
