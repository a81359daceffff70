import bz2
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


@pytest.fixture(scope="session")
def oracle():
    from oracle import pyoracle
    pyoracle.build()
    return pyoracle


@pytest.fixture(scope="session")
def c16_path(tmp_path_factory):
    """data/problem-16-22106-pre.txt of the reference, shipped bz2-compressed under tests/golden/."""
    dst = tmp_path_factory.mktemp("bal") / "problem-16-22106-pre.txt"
    with bz2.open(os.path.join(GOLDEN, "problem-16-22106-pre.txt.bz2"), "rb") as f:
        dst.write_bytes(f.read())
    return str(dst)


@pytest.fixture(scope="session")
def c16(oracle, c16_path):
    """C16 after BALProblem::Normalize(), as the bundle_adjuster example feeds it to Ceres."""
    bal = oracle.BalProblem(c16_path)
    bal.normalize()
    return bal


@pytest.fixture(scope="session")
def c16_raw(oracle, c16_path):
    return oracle.BalProblem(c16_path)


def rel_err(a, b):
    a = np.asarray(a, dtype=float)
    b = np.asarray(b, dtype=float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))
