"""The drop-in, proven: the UNMODIFIED reference (``xhistogram.core`` from /root/reference or baseline/_ref) with its hot
path ``_bincount`` replaced by the ctypes stub of INTEGRATION.md (``integration/xhistogram_core_stub.py``), run through
the reference's own numpy test cases (xhistogram/test/test_core.py:25-228, 365-382, restated here) — needs a B200."""
import numpy as np
import pytest

from oracle import ref_loader

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref():
    if not ref_loader.reference_available():
        pytest.skip("reference package not found (neither /root/reference nor baseline/_ref)")
    from integration import xhistogram_core_stub as stub
    core = ref_loader.load_reference_core()
    original = stub.install(core)
    yield core
    core._bincount = original


bins_int = 10
bins_str = "auto"
bins_arr = np.linspace(-4, 4, 10)
range_ = (0, 1)


@pytest.mark.parametrize("density", [False, True])
@pytest.mark.parametrize("block_size", [None, 1, 2])
@pytest.mark.parametrize("axis", [1, None])
@pytest.mark.parametrize("bins", [10, np.linspace(-4, 4, 10), "auto"])
@pytest.mark.parametrize("range_", [None, (-4, 4)])
@pytest.mark.parametrize("add_nans", [False, True])
def test_histogram_results_1d(ref, block_size, density, axis, bins, range_, add_nans):          # test_core.py:25-69
    nrows, ncols = 5, 20
    np.random.seed(2)
    data = np.random.randn(nrows, ncols)
    if add_nans:
        N_nans = 20
        data.ravel()[np.random.choice(data.size, N_nans, replace=False)] = np.nan
    bins = np.linspace(-4, 4, 10)          # (the reference test overrides its own parameter the same way)
    h, bin_edges = ref.histogram(data, bins=bins, range=range_, axis=axis, block_size=block_size, density=density)
    expected_shape = (nrows, len(bin_edges[0]) - 1) if axis == 1 else (len(bin_edges[0]) - 1,)
    assert h.shape == expected_shape
    if axis:
        bins_np = np.histogram_bin_edges(data, bins=bins, range=range_)
        expected = np.stack([np.histogram(data[i], bins=bins_np, range=range_, density=density)[0] for i in range(nrows)])
    else:
        expected = np.histogram(data, bins=bins, range=range_, density=density)[0]
    np.testing.assert_allclose(h, expected)
    if density:
        widths = np.diff(bins)
        integral = np.sum(h * widths, axis)
        np.testing.assert_allclose(integral, 1.0)


@pytest.mark.parametrize("block_size", [None, 1, 2])
def test_histogram_results_1d_weighted(ref, block_size):                                         # test_core.py:72-80
    nrows, ncols = 5, 20
    data = np.random.RandomState(2).randn(nrows, ncols)
    bins = np.linspace(-4, 4, 10)
    h, _ = ref.histogram(data, bins=bins, axis=1, block_size=block_size)
    weights = 2 * np.ones_like(data)
    h_w, _ = ref.histogram(data, bins=bins, axis=1, weights=weights, block_size=block_size)
    np.testing.assert_array_equal(2 * h, h_w)


@pytest.mark.parametrize("block_size", [None, 1, 2, "auto"])
def test_histogram_results_1d_weighted_broadcasting(ref, block_size):                            # test_core.py:84-92
    nrows, ncols = 5, 20
    data = np.random.RandomState(3).randn(nrows, ncols)
    bins = np.linspace(-4, 4, 10)
    h, _ = ref.histogram(data, bins=bins, axis=1, block_size=block_size)
    weights = 2 * np.ones((1, ncols))
    h_w, _ = ref.histogram(data, bins=bins, axis=1, weights=weights, block_size=block_size)
    np.testing.assert_array_equal(2 * h, h_w)


@pytest.mark.parametrize("block_size", [None, 1, 2])
def test_histogram_right_edge(ref, block_size):                                                  # test_core.py:95-113
    nrows, ncols = 5, 20
    data = np.ones((nrows, ncols))
    bins = np.array([0, 0.5, 1])
    h, _ = ref.histogram(data, bins=bins, axis=1, block_size=block_size)
    assert h.shape == (nrows, len(bins) - 1)
    np.testing.assert_array_equal(h.sum(axis=1), ncols * np.ones(nrows))
    h_1d, _ = ref.histogram(data, bins=bins, block_size=block_size)
    assert h_1d.shape == (len(bins) - 1,)
    np.testing.assert_array_equal(h_1d, np.histogram(data, bins=bins)[0])


def test_histogram_results_2d(ref):                                                              # test_core.py:116-129
    nrows, ncols = 5, 20
    r = np.random.RandomState(4)
    data_a, data_b = r.randn(nrows, ncols), r.randn(nrows, ncols)
    nbins_a, nbins_b = 9, 10
    bins_a, bins_b = np.linspace(-4, 4, nbins_a + 1), np.linspace(-4, 4, nbins_b + 1)
    h, _ = ref.histogram(data_a, data_b, bins=[bins_a, bins_b])
    assert h.shape == (nbins_a, nbins_b)
    hist, _, _ = np.histogram2d(data_a.ravel(), data_b.ravel(), bins=[bins_a, bins_b])
    np.testing.assert_array_equal(hist, h)


def test_histogram_results_2d_broadcasting(ref):                                                 # test_core.py:132-157 (numpy variant)
    nrows, ncols = 5, 20
    r = np.random.RandomState(5)
    data_a, data_b = r.randn(ncols), r.randn(nrows, ncols)
    nbins_a, nbins_b = 9, 10
    bins_a, bins_b = np.linspace(-4, 4, nbins_a + 1), np.linspace(-4, 4, nbins_b + 1)
    h, _ = ref.histogram(data_a, data_b, bins=[bins_a, bins_b])
    assert h.shape == (nbins_a, nbins_b)
    hist, _, _ = np.histogram2d(np.broadcast_to(data_a, (nrows, ncols)).ravel(), data_b.ravel(), bins=[bins_a, bins_b])
    np.testing.assert_array_equal(hist, h)


@pytest.mark.parametrize("add_nans", [False, True])
def test_histogram_results_2d_density(ref, add_nans):                                            # test_core.py:160-187
    nrows, ncols = 5, 20
    r = np.random.RandomState(6)
    data_a, data_b = r.randn(nrows, ncols), r.randn(nrows, ncols)
    if add_nans:
        data_a.ravel()[r.choice(data_a.size, 20, replace=False)] = np.nan
        data_b.ravel()[r.choice(data_b.size, 20, replace=False)] = np.nan
    nbins_a, nbins_b = 9, 10
    bins_a, bins_b = np.linspace(-4, 4, nbins_a + 1), np.linspace(-4, 4, nbins_b + 1)
    h, _ = ref.histogram(data_a, data_b, bins=[bins_a, bins_b], density=True)
    assert h.shape == (nbins_a, nbins_b)
    hist, _, _ = np.histogram2d(data_a.ravel(), data_b.ravel(), bins=[bins_a, bins_b], density=True)
    np.testing.assert_allclose(hist, h)
    widths_a, widths_b = np.diff(bins_a), np.diff(bins_b)
    areas = np.outer(widths_a, widths_b)
    np.testing.assert_allclose(np.sum(h * areas), 1.0)


def test_histogram_results_3d_counts(ref):              # test_core.py:190-228 with density=False (Q2: the reference's own density for K >= 3 raises on numpy >= 1.24)
    nrows, ncols = 5, 20
    r = np.random.RandomState(7)
    a, b, c = (r.randn(nrows, ncols) for _ in range(3))
    bins = [np.linspace(-4, 4, 10), np.linspace(-4, 4, 11), np.linspace(-4, 4, 12)]
    h, _ = ref.histogram(a, b, c, bins=bins)
    want, _ = np.histogramdd(np.stack([a.ravel(), b.ravel(), c.ravel()]).T, bins=bins)
    np.testing.assert_array_equal(h, want)


def test_histogram_shape(ref):                                                                   # test_core.py:231-273 (numpy variant)
    from itertools import combinations
    shape = 10, 15, 12, 20
    b = np.random.RandomState(8).randn(*shape)
    bins = np.linspace(-4, 4, 27)
    c, _ = ref.histogram(b, bins=bins)
    assert c.shape == (len(bins) - 1,)
    for axis in [(0, 1, 2, 3), (0, 1, 3, 2), (3, 2, 1, 0), (3, 2, 0, 1)]:
        c, _ = ref.histogram(b, bins=bins, axis=axis)
        assert c.shape == (len(bins) - 1,)
        np.testing.assert_array_equal(c, np.histogram(b, bins=bins)[0])
    for axis in list(range(4)) + list(range(-1, -5, -1)):
        c, _ = ref.histogram(b, bins=bins, axis=axis)
        out_shape = list(shape); del out_shape[axis]; out_shape.append(len(bins) - 1)
        assert c.shape == tuple(out_shape)
    for nc in (2, 3):
        for axis in combinations(range(4), nc):
            c, _ = ref.histogram(b, bins=bins, axis=axis)
            out_shape = [shape[i] for i in range(4) if i not in axis] + [len(bins) - 1]
            assert c.shape == tuple(out_shape)
            moved = np.moveaxis(b, axis, tuple(range(-nc, 0))).reshape(tuple(out_shape[:-1]) + (-1,))
            want = np.apply_along_axis(lambda v: np.histogram(v, bins=bins)[0], -1, moved)
            np.testing.assert_array_equal(c, want)


@pytest.mark.parametrize("block_size", [None, 1, 2])
def test_histogram_results_datetime(ref, block_size):                                            # test_core.py:365-382 (numpy variant)
    import pandas as pd
    data = pd.date_range(start="2000-06-01", periods=5)
    bins = np.array([np.datetime64("1999-01-01"), np.datetime64("2000-01-01"), np.datetime64("2001-01-01")])
    h = ref.histogram(data, bins=bins, block_size=block_size)[0]
    expected = np.histogram(data, bins=bins)[0]
    np.testing.assert_allclose(h, expected)
