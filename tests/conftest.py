import glob
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG_PARENT = os.path.join(ROOT, "end-to-end-self-supervised-slam_b200")
for p in (ROOT, PKG_PARENT):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN_DIR = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


def golden_files():
    """Single-source goldens of tools/make_golden.py (the composite_* files of tools/make_golden_composite.py have their own fixture)."""
    return sorted(p for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")) if not os.path.basename(p).startswith("composite_"))


def composite_golden_files():
    return sorted(glob.glob(os.path.join(GOLDEN_DIR, "composite_*.npz")))


@pytest.fixture(params=golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def golden(request):
    return np.load(request.param)


@pytest.fixture(params=composite_golden_files(), ids=lambda p: os.path.basename(p)[:-4])
def composite_golden(request):
    return np.load(request.param)


def same_values(a, b):
    """Number of elements that differ as IEEE values (+0 == -0, NaN == NaN)."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    return int((~((a == b) | (np.isnan(a) & np.isnan(b)))).sum())


def rel_max(a, ref):
    """max|a - ref| / max|ref|  -- relative error in the max norm."""
    a, ref = np.asarray(a, dtype=np.float64), np.asarray(ref, dtype=np.float64)
    return float(np.abs(a - ref).max() / (np.abs(ref).max() + 1e-300))
