"""The label glue (xhistogram_b200/xarray.py) and the dask branch of core.histogram, EXECUTED: restated from the reference's
xhistogram/test/test_xarray.py and test_chunking.py.  xarray and dask are not installed in this image, so unless they
are, the minimal stand-ins under tests/stubs are used (tests/conftest.py).  Every test runs twice: with the one native
call replaced by the oracle's block kernel (CPU suite) and, marked gpu, on the real CUDA path."""
from itertools import combinations

import numpy as np
import pytest

import xarray as xr

from tests.test_frontend_cpu import _oracle_desc_call
from xhistogram_b200 import core
from xhistogram_b200.xarray import histogram


@pytest.fixture(params=["oracle", pytest.param("cuda", marks=pytest.mark.gpu)])
def backend(request, monkeypatch):
    if request.param == "oracle":
        monkeypatch.setattr(core, "_desc_call", _oracle_desc_call)
        monkeypatch.setattr(core, "_minmax", lambda a: (float(np.min(a)), float(np.max(a))))
    return request.param


DIMS = {"time": 5, "depth": 10, "lat": 45, "lon": 90}
COORDS = {
    "time": ("time", np.arange("2000-01-01", "2000-01-06", dtype="datetime64[D]")),
    "depth": ("depth", np.arange(DIMS["depth"]) * 100.0 + 50),
    "lat": ("lat", np.arange(DIMS["lat"]) * 180 / DIMS["lat"] - 90 + 90 / DIMS["lat"]),
    "lon": ("lon", np.arange(DIMS["lon"]) * 360 / DIMS["lon"] + 180 / DIMS["lon"]),
}


@pytest.fixture(params=[("lon",), ("lat", "lon"), ("depth", "lat", "lon"), ("time", "depth", "lat", "lon")], ids=["1D", "2D", "3D", "4D"])
def ones(request):
    dims = request.param
    coords = {k: v for k, v in COORDS.items() if k in dims}
    return xr.DataArray(np.ones([DIMS[d] for d in dims], dtype="f8"), dims=dims, coords=coords, name="ones")


@pytest.mark.parametrize("ndims", [1, 2, 3, 4])
def test_histogram_ones(backend, ones, ndims):                                        # test_xarray.py:38-67
    if ones.ndim < ndims:
        pytest.skip("more dimension combinations than array dimensions")
    bins = np.array([0, 0.9, 1.1, 2])
    bins_c = 0.5 * (bins[1:] + bins[:-1])
    for d in combinations(ones.dims, ndims):
        h = histogram(ones, bins=[bins], dim=d)
        other_dims = [dim for dim in ones.dims if dim not in d]
        assert set(other_dims) <= set(h.dims)
        expected = xr.DataArray([0, ones.size, 0], dims=["ones_bin"], coords={"ones_bin": ("ones_bin", bins_c)}, name="histogram_ones")
        xr.testing.assert_identical(h.sum(other_dims), expected)


@pytest.mark.parametrize("ndims", [1, 2, 3, 4])
def test_histogram_ones_density(backend, ones, ndims):                                # test_xarray.py:70-94
    if ones.ndim < ndims:
        pytest.skip("more dimension combinations than array dimensions")
    bins = np.array([0, 0.9, 1.1, 2])
    for d in combinations(ones.dims, ndims):
        h = histogram(ones, bins=[bins], dim=d, density=True)
        np.testing.assert_allclose((h * 0.2).sum(dim="ones_bin").values, 1.0)


@pytest.mark.parametrize("ndims", [1, 2, 3])
def test_weights(backend, ones, ndims):                                               # test_xarray.py:99-135
    if ones.ndim < ndims:
        pytest.skip("more dimension combinations than array dimensions")
    bins = np.array([0, 0.9, 1.1, 2])
    bins_c = 0.5 * (bins[1:] + bins[:-1])
    weight_value = 0.5
    for n_combinations in range(ones.ndim):
        for weight_dims in combinations(ones.dims, n_combinations):
            weights = xr.full_like(ones.isel(**{dim: 0 for dim in weight_dims}), weight_value)
            for d in combinations(ones.dims, ndims):
                h = histogram(ones, weights=weights, bins=[bins], dim=d)
                other_dims = [dim for dim in ones.dims if dim not in d]
                expected = xr.DataArray([0, weight_value * ones.size, 0], dims=["ones_bin"], coords={"ones_bin": ("ones_bin", bins_c)},
                                        name="histogram_ones")
                xr.testing.assert_identical(h.sum(other_dims), expected)


def test_dims_and_coords(backend):                                                    # test_xarray.py:139-173 (issue #5)
    t, z, X, Y = np.arange(4), np.arange(10), np.arange(30), np.arange(30)
    r = np.random.RandomState(0)
    a1 = xr.DataArray(r.randint(0, 100, size=(4, 10, 30, 30)), coords=[t, z, X, Y], dims=["time", "depth", "X", "Y"], name="one")
    a2 = xr.DataArray(r.randint(0, 50, size=(4, 10, 30, 30)), coords=[t, z, X, Y], dims=["time", "depth", "X", "Y"], name="two")
    bins1, bins2 = np.linspace(0, 100, 50), np.linspace(0, 50, 25)
    result = histogram(a1, a2, dim=["X", "Y"], bins=[bins1, bins2])
    assert result.dims == ("time", "depth", "one_bin", "two_bin")
    assert result.time.identical(a1.time)
    assert result.depth.identical(a2.depth)
    want = np.stack([[np.histogram2d(a1.values[i, j].ravel(), a2.values[i, j].ravel(), bins=[bins1, bins2])[0] for j in range(10)] for i in range(4)])
    np.testing.assert_array_equal(result.values, want)


@pytest.mark.parametrize("number_of_inputs", [1, 2])
@pytest.mark.parametrize("keep_coords", [True, False])
@pytest.mark.parametrize("include_weights", [True, False])
def test_carry_coords(backend, keep_coords, number_of_inputs, include_weights):      # test_xarray.py:176-211
    data = np.random.RandomState(1).randint(0, 100, size=(40, 10, 10))
    da = xr.DataArray(data, coords=[np.arange(40), np.arange(10), np.arange(10)], dims=["time", "X", "Y"], name="one")
    weights = xr.full_like(da, 0.5) if include_weights else None
    da["lon"] = da.X ** 2 + da.Y ** 2
    assert "lon" in da.coords
    bins = np.linspace(0, 100, 10)
    result = histogram(*[da] * number_of_inputs, bins=[bins] * number_of_inputs, dim=["time"], weights=weights, keep_coords=keep_coords)
    assert ("lon" in result.coords) == keep_coords


def test_input_type_check(backend):                                                   # test_xarray.py:215-218 (issue #14)
    with pytest.raises(TypeError):
        histogram(np.arange(100))


# ---- dask: every chunking gives the numpy result (xhistogram/test/test_chunking.py) -------------------------------------
def example_dataarray(shape=(5, 20), seed=0, name="T"):
    return xr.DataArray(np.random.RandomState(seed).randn(*shape), dims=[f"dim_{i}" for i in range(len(shape))], name=name)


@pytest.mark.parametrize("weights", [False, True])
@pytest.mark.parametrize("chunksize", [1, 2, 3, 10])
@pytest.mark.parametrize("shape", [(10,), (10, 4)])
def test_chunked_weights(backend, chunksize, shape, weights):                         # test_chunking.py:8-30
    data_a = example_dataarray(shape).chunk((chunksize,))
    w = example_dataarray(shape, 1).chunk((chunksize,)) if weights else None
    bins_a = np.linspace(-4, 4, 7)
    h = histogram(data_a, bins=[bins_a], weights=w)
    assert h.shape == (6,)
    hist, _ = np.histogram(data_a.values, bins=bins_a, weights=None if w is None else w.values)
    np.testing.assert_allclose(hist, h.values)


@pytest.mark.parametrize("xchunksize", [1, 3, 10])
@pytest.mark.parametrize("ychunksize", [2, 3, 12])
class TestChunks2D:
    def test_2d_chunks(self, backend, xchunksize, ychunksize):                       # test_chunking.py:36-49
        data_a = example_dataarray(shape=(10, 12)).chunk((xchunksize, ychunksize))
        bins_a = np.linspace(-4, 4, 9)
        h = histogram(data_a, bins=[bins_a])
        assert h.shape == (8,)
        np.testing.assert_allclose(np.histogram(data_a.values, bins=bins_a)[0], h.values)

    @pytest.mark.parametrize("reduce_dim", ["dim_0", "dim_1"])
    def test_2d_chunks_broadcast_dim(self, backend, xchunksize, ychunksize, reduce_dim):   # test_chunking.py:51-80
        data_a = example_dataarray(shape=(10, 12)).chunk((xchunksize, ychunksize))
        dims = list(data_a.dims)
        broadcast_dim = [d for d in dims if d != reduce_dim][0]
        bins_a = np.linspace(-4, 4, 9)
        h = histogram(data_a, bins=[bins_a], dim=(reduce_dim,))
        assert h.shape == (data_a.sizes[broadcast_dim], 8)
        hist = np.apply_along_axis(lambda v: np.histogram(v, bins=bins_a)[0], dims.index(reduce_dim), data_a.values)
        got = h.values.T if reduce_dim == "dim_0" else h.values
        np.testing.assert_allclose(hist, got)

    def test_unaligned_data_chunks(self, backend, xchunksize, ychunksize):           # test_chunking.py:107-128
        data_a = example_dataarray(shape=(10, 12)).chunk((xchunksize, ychunksize))
        data_b = example_dataarray(shape=(10, 12), seed=2, name="S").chunk((xchunksize + 1, ychunksize + 1))   # (the reference names both "T")
        bins_a, bins_b = np.linspace(-4, 4, 9), np.linspace(-4, 4, 10)
        h = histogram(data_a, data_b, bins=[bins_a, bins_b])
        assert h.shape == (8, 9)
        hist, _, _ = np.histogram2d(data_a.values.ravel(), data_b.values.ravel(), bins=[bins_a, bins_b])
        np.testing.assert_allclose(hist, h.values)

    def test_unaligned_weights_chunks(self, backend, xchunksize, ychunksize):        # test_chunking.py:130-146
        data_a = example_dataarray(shape=(10, 12)).chunk((xchunksize, ychunksize))
        weights = example_dataarray(shape=(10, 12), seed=3).chunk((xchunksize + 1, ychunksize + 1))
        bins_a = np.linspace(-4, 4, 9)
        h = histogram(data_a, bins=[bins_a], weights=weights)
        np.testing.assert_allclose(np.histogram(data_a.values, bins=bins_a, weights=weights.values)[0], h.values)


def test_dask_inputs_need_explicit_edges(backend):                                   # test_core.py:276-313 / core.py:377-381
    import dask.array as dsa
    a = dsa.from_array(np.random.RandomState(5).randn(10, 12), chunks=(5, 6))
    with pytest.raises(TypeError):
        core.histogram(a, bins=10)
    h, _ = core.histogram(a, bins=np.linspace(-4, 4, 9), density=True)
    want = np.histogram(np.asarray(a), bins=np.linspace(-4, 4, 9), density=True)[0]
    np.testing.assert_allclose(np.asarray(h), want)


def test_dask_chunks_are_dealt_to_devices_in_turn(monkeypatch):                      # chunk -> device round-robin
    import dask.array as dsa
    seen = []

    def recording_desc_call(arrs, strides, w, wstride, bins, M, N, dtype, wdtype, mem, device, *rest, **kw):
        seen.append(device)
        return _oracle_desc_call(arrs, strides, w, wstride, bins, M, N, dtype, wdtype, mem, device, *rest, **kw)

    monkeypatch.setattr(core, "_desc_call", recording_desc_call)
    a = dsa.from_array(np.random.RandomState(6).randn(12, 10), chunks=(3, 5))            # 4 x 2 = 8 chunks
    bins = np.linspace(-4, 4, 9)
    core.set_chunk_devices([0, 1, 2, 3])
    try:
        h, _ = core.histogram(a, bins=bins)
    finally:
        core.set_chunk_devices(None)
    np.testing.assert_array_equal(np.asarray(h), np.histogram(np.asarray(a), bins=bins)[0])
    assert len(seen) == 8 and sorted(set(seen)) == [0, 1, 2, 3] and all(seen.count(d) == 2 for d in range(4))
