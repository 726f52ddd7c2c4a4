import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def _use_stubs_if_missing():
    """xarray / dask are not installed in this image: fall back to the minimal stand-ins under tests/stubs (README there)."""
    import importlib.util
    stubs = os.path.join(ROOT, "tests", "stubs")
    if (importlib.util.find_spec("xarray") is None or importlib.util.find_spec("dask") is None) and stubs not in sys.path:
        sys.path.append(stubs)        # appended: a real installation always wins


_use_stubs_if_missing()


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a real B200 (run with -m gpu on the GPU box)")


def _gpu_count():
    try:
        from xhistogram_b200 import _cabi
        return _cabi.device_count()
    except Exception:
        return 0


@pytest.fixture(scope="session")
def gpu_count():
    return _gpu_count()


@pytest.fixture(scope="session")
def golden():
    path = os.path.join(ROOT, "tests", "golden", "golden.npz")
    return np.load(path)


def golden_case(golden, name):
    """(h, [edges...]) of one golden case."""
    h = golden[f"{name}/h"]
    edges = []
    i = 0
    while f"{name}/edges{i}" in golden.files:
        edges.append(golden[f"{name}/edges{i}"])
        i += 1
    return h, edges, str(golden[f"{name}/digest"])


def assert_hist_equal(got, want, rtol=1e-6):
    """Parity bar: integer counts bit-exact; float results within ``rtol`` relative, NaNs in the same places."""
    got, want = np.asarray(got), np.asarray(want)
    assert got.shape == want.shape, (got.shape, want.shape)
    assert got.dtype == want.dtype, (got.dtype, want.dtype)
    if want.dtype.kind in "iu":
        assert np.array_equal(got, want)
        return
    assert np.array_equal(np.isnan(got), np.isnan(want))
    inf = np.isinf(want)
    assert np.array_equal(got[inf], want[inf])            # infinite sums (inf weights) must match in sign
    m = np.isfinite(want)
    scale = np.maximum(np.abs(want[m]), np.finfo(np.float64).tiny)
    err = np.abs(got[m] - want[m]) / scale
    assert err.size == 0 or err.max() <= rtol, float(err.max())
