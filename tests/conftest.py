import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, os.path.join(ROOT, "tests"))


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_sessionstart(session):
    """The built library is git-ignored: in a fresh checkout with nvcc on the PATH, build it in-tree (what
    __graft_entry__.build() does) before the tests look for it.  Without nvcc the tests report it missing."""
    import shutil
    import linearmpchumanoid_b200 as lmh
    if not os.path.exists(lmh.LIB_PATH) and shutil.which("nvcc"):
        lmh.build_library()


@pytest.fixture(scope="session")
def oracle():
    """CPU restatement of the reference (oracle/liblmh_oracle.so): the CHECKER, never the thing under test."""
    import orc_py
    orc_py.build()
    return orc_py.oracle()


@pytest.fixture(scope="session")
def reflib():
    """The reference's own sources built against shim headers (oracle/_ref), if present or buildable."""
    import orc_py
    return orc_py.reference()


@pytest.fixture(scope="session")
def golden():
    return np.load(os.path.join(ROOT, "tests", "golden", "reference_vectors.npz"))


@pytest.fixture(scope="session")
def ik_golden():
    """postures of the reference's own Kinematics::compute for seeded targets (tests/golden/make_ik_golden.py)"""
    return np.load(os.path.join(ROOT, "tests", "golden", "ik_vectors.npz"))


@pytest.fixture(scope="session")
def qp_exact():
    return np.load(os.path.join(ROOT, "tests", "golden", "qp_exact.npz"))


@pytest.fixture(scope="session")
def hostemu():
    """Host build (NT = 1) of the device source lmh_model.cuh: a TEST TOOL for the kernel logic."""
    subprocess.run(["make", "-C", os.path.join(ROOT, "tests", "hostemu")], check=True, capture_output=True)
    import hostemu_py
    return hostemu_py


def scalederr(a, b):
    """max|a - b| / max(1, max|b|): an ABSOLUTE max-norm error for references below 1, a relative one above.  Used for the
    model intermediates, several of which are physically zero (a round-off sized reference has no relative error)."""
    a, b = np.asarray(a, float), np.asarray(b, float)
    return float(np.abs(a - b).max() / max(1.0, np.abs(b).max()))


def relerr(a, b):
    """norm-wise RELATIVE error max|a - b| / max|b| (the north star's "within 1e-6 relative"); an all-zero reference
    falls back to the absolute error"""
    a, b = np.asarray(a, float), np.asarray(b, float)
    scale = float(np.abs(b).max())
    return float(np.abs(a - b).max() / scale) if scale > 0.0 else float(np.abs(a - b).max())
