import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def golden():
    return {name: np.load(os.path.join(GOLDEN, name + ".npz"))
            for name in ("angular", "radial", "pwd", "examples", "extras")}


def rel_err(x, ref, floor=1e-15):
    """SURVEY 7 tolerance: |x-ref| <= tol*|ref| + floor*colmax -> returns max of
    |x-ref| / (|ref| + floor/tol * colmax) style metric, evaluated by the caller."""
    x = np.asarray(x)
    ref = np.asarray(ref)
    return np.abs(x - ref)


def assert_close(x, ref, rtol=1e-10, floor=1e-15, name=""):
    """|x - ref| <= rtol*|ref| + floor*max|ref| (per last-axis column), NaN/Inf patterns equal."""
    x = np.asarray(x)
    ref = np.asarray(ref)
    assert x.shape == ref.shape, "%s shape %s vs %s" % (name, x.shape, ref.shape)
    bad_ref = ~np.isfinite(ref)
    bad_x = ~np.isfinite(x)
    assert np.array_equal(np.isnan(ref), np.isnan(x)), name + ": NaN pattern differs"
    inf_ref = np.isinf(ref.real) | np.isinf(ref.imag) if np.iscomplexobj(ref) else np.isinf(ref)
    inf_x = np.isinf(x.real) | np.isinf(x.imag) if np.iscomplexobj(x) else np.isinf(x)
    assert np.array_equal(inf_ref & ~np.isnan(ref), inf_x & ~np.isnan(x)), name + ": Inf pattern differs"
    ok = ~(bad_ref | bad_x)
    if not ok.any():
        return
    mag = np.where(ok, np.abs(np.where(ok, ref, 0)), 0.0)
    if ref.ndim >= 1 and ref.shape[-1] > 0:
        colmax = mag.reshape(-1, ref.shape[-1]).max(axis=0)
        colmax = np.broadcast_to(colmax, ref.shape)
    else:
        colmax = mag
    err = np.where(ok, np.abs(np.where(ok, x, 0) - np.where(ok, ref, 0)), 0.0)
    tol = rtol * mag + floor * colmax
    worst = np.max(err - tol)
    assert worst <= 0, "%s: max excess %.3e (max err %.3e, max ref %.3e)" % (
        name, worst, err.max(), mag.max())
