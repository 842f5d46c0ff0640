import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN_PATH = os.path.join(ROOT, "tests", "golden", "reference_vectors.json")

# Parity tolerance for floating-point results (BASELINE.json north_star):
# |dlogL| <= 1e-8 absolute, 1e-12 relative, whichever is larger for the value at hand.
LOGL_ATOL = 1e-8
LOGL_RTOL = 1e-12


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def has_cuda():
    try:
        import torch
        return torch.cuda.is_available()
    except Exception:
        return False


def pytest_collection_modifyitems(config, items):
    if has_cuda():
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    with open(GOLDEN_PATH) as f:
        return json.load(f)


@pytest.fixture(scope="session")
def libgsn():
    """libgsn.so built in-tree (compiled on demand: nvcc cross-compiles without a GPU)."""
    from gaussn_b200 import build, _engine
    build.build_lib()
    return _engine.load()


def assert_logl_close(got, want, what="", atol=LOGL_ATOL, rtol=LOGL_RTOL):
    got = np.asarray(got, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    nan_g, nan_w = np.isnan(got), np.isnan(want)
    assert np.array_equal(nan_g, nan_w), f"{what}: NaN pattern differs ({nan_g.sum()} vs {nan_w.sum()})"
    inf_g, inf_w = np.isinf(got), np.isinf(want)
    assert np.array_equal(inf_g, inf_w) and np.array_equal(got[inf_g], want[inf_w]), f"{what}: inf pattern differs"
    fin = ~(nan_w | inf_w)
    if fin.any():
        err = np.abs(got[fin] - want[fin])
        tol = np.maximum(atol, rtol * np.abs(want[fin]))
        worst = int(np.argmax(err / tol))
        assert np.all(err <= tol), (f"{what}: max |dlogL| = {err.max():.3e} (worst: got {got[fin][worst]!r}, "
                                    f"want {want[fin][worst]!r}, tol {tol[worst]:.1e})")
