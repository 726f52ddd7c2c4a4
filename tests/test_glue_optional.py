"""xarray / dask glue: runs only where those optional packages are installed (they are not in the build image;
the reference lists both as dependencies, xhistogram/setup.py:23).  The GPU is needed for the actual call."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_xarray_wrapper_matches_reference_behaviour():
    xr = pytest.importorskip("xarray")
    from xhistogram_b200.xarray import histogram

    dims = {"time": 5, "depth": 10, "lat": 45, "lon": 90}
    da = xr.DataArray(np.ones(list(dims.values())), dims=list(dims), name="ones")
    bins = np.array([0, 0.9, 1.1, 2])
    for d in (["lon"], ["lat", "lon"], ["depth", "lat", "lon"], list(dims)):
        h = histogram(da, bins=[bins], dim=d)                     # reference test_xarray.py:38-67
        other = [k for k in dims if k not in d]
        assert h.name == "histogram_ones" and h.dims[-1] == "ones_bin" and list(h.dims[:-1]) == other
        np.testing.assert_array_equal(h.sum(other).values, [0, da.size, 0])
        np.testing.assert_allclose(h["ones_bin"].values, 0.5 * (bins[1:] + bins[:-1]))
    hw = histogram(da, bins=[bins], weights=0.5 * da)              # reference test_xarray.py:99-135
    np.testing.assert_array_equal(hw.values, [0, 0.5 * da.size, 0])
    with pytest.raises(TypeError):
        histogram(np.ones(3), bins=[bins])                         # reference test_xarray.py:215-218


def test_dask_blockwise_path():
    dsa = pytest.importorskip("dask.array")
    from xhistogram_b200.core import histogram

    r = np.random.default_rng(0)
    a, b = r.standard_normal((10, 12)), r.standard_normal((10, 12))
    bins = [np.linspace(-4, 4, 9), np.linspace(-4, 4, 7)]
    for chunks in ((1, 12), (3, 5), (10, 4)):
        h, _ = histogram(dsa.from_array(a, chunks=chunks), dsa.from_array(b, chunks=chunks), bins=bins)
        want, _, _ = np.histogram2d(a.ravel(), b.ravel(), bins=bins)
        np.testing.assert_array_equal(h.compute(), want)           # reference test_chunking.py
    with pytest.raises(TypeError):
        histogram(dsa.from_array(a, chunks=(5, 6)), bins=10)        # dask inputs need explicit edges (core.py:377-381)
