"""GPU tests of the round-2 additions: cached edge tables / probe verdicts, asynchronous device output, device-side
axis generality (xh_permute), strict device dtypes, large bin spaces in the leading-axis layout, K = 5..8 variables."""
import itertools

import numpy as np
import pytest

from oracle import hist_oracle as O
from tests.conftest import assert_hist_equal
from xhistogram_b200 import DeviceArray, _cabi, core

pytestmark = pytest.mark.gpu


def _xyw(n, seed=0, dtype=np.float32):
    r = np.random.default_rng(seed)
    return (r.standard_normal(n).astype(dtype), r.standard_normal(n).astype(dtype), r.random(n).astype(dtype))


def test_repeat_calls_use_cached_verdict_and_stay_exact():
    """Same buffers, same edges, many calls: the probe verdict is reused (no probe launch) and every result is identical."""
    n = 3_000_000
    x, y, w = _xyw(n, 1)
    e = np.linspace(-4, 4, 257)
    dx, dy, dw = (DeviceArray.from_numpy(a) for a in (x, y, w))
    want = O.block_bincount([x[None], y[None]], [e, e], w[None]).reshape(256, 256)
    first, _ = core.histogram(dx, dy, bins=[e, e], weights=dw)
    assert_hist_equal(first, want, rtol=1e-6)
    for _ in range(20):
        h, _ = core.histogram(dx, dy, bins=[e, e], weights=dw)
        assert np.array_equal(h, first)          # fixed point per CTA + float64 adds of identical partials in any order
    wc = O.block_bincount([x[None], y[None]], [e, e], None).reshape(256, 256)
    for _ in range(5):
        assert np.array_equal(core.histogram(dx, dy, bins=[e, e])[0], wc)


def test_cached_verdict_survives_data_changing_under_the_same_pointers():
    """The verdict is keyed by buffer addresses; new data in the same buffers must still give exact results (a sample
    outside the stale window / a weight outside the stale fixed-point form takes the slow exact path) and the slow-path
    counter makes the library probe again."""
    n = 2_000_000
    r = np.random.default_rng(5)
    e = np.linspace(-4, 4, 257)
    x, y, w = _xyw(n, 2)
    dx, dy, dw = (DeviceArray.from_numpy(a) for a in (x, y, w))
    for _ in range(3):
        core.histogram(dx, dy, bins=[e, e], weights=dw)
    lib = _cabi.lib()
    # same addresses, very different data: mass in a corner of the bin space, weights that are not multiples of 2^-24
    x2 = (r.random(n) * 1.5 + 2.4).astype(np.float32)
    y2 = (r.random(n) * 1.5 - 3.9).astype(np.float32)
    w2 = (r.standard_normal(n) * np.pi).astype(np.float32)
    for d, a in ((dx, x2), (dy, y2), (dw, w2)):
        _cabi.check(lib.xh_memcpy(0, d.ptr, a.ctypes.data, a.nbytes, _cabi.XH_DEVICE, _cabi.XH_HOST), "h2d")
    want = O.block_bincount([x2[None], y2[None]], [e, e], w2[None]).reshape(256, 256)
    for _ in range(24):
        h, _ = core.histogram(dx, dy, bins=[e, e], weights=dw)
        assert_hist_equal(h, want, rtol=1e-6)


def test_edges_changed_in_place_are_noticed():
    """The edge-table cache is keyed by edge CONTENT: mutating the caller's edge array between calls gives the new bins."""
    x = np.random.default_rng(3).standard_normal(200_000).astype(np.float32)
    e = np.linspace(-4, 4, 65)
    dx = DeviceArray.from_numpy(x)
    assert np.array_equal(core.histogram(dx, bins=e)[0], np.histogram(x, bins=e)[0])
    e[1:-1] += 0.013
    assert np.array_equal(core.histogram(dx, bins=e)[0], np.histogram(x, bins=e)[0])
    e *= 0.5
    assert np.array_equal(core.histogram(dx, bins=e)[0], np.histogram(x, bins=e)[0])


def test_many_different_edge_sets_evict_cleanly():
    x = np.random.default_rng(4).standard_normal(100_000).astype(np.float32)
    dx = DeviceArray.from_numpy(x)
    for i in range(40):                      # more than the cache holds
        e = np.linspace(-4, 4 + 0.01 * i, 33 + i)
        assert np.array_equal(core.histogram(dx, bins=e)[0], np.histogram(x, bins=e)[0])


@pytest.mark.parametrize("weighted,density", [(False, False), (True, False), (True, True), (False, True)])
def test_async_device_output(weighted, density):
    n = 1_500_000
    x, y, w = _xyw(n, 7)
    e = np.linspace(-4, 4, 101)
    dx, dy, dw = (DeviceArray.from_numpy(a) for a in (x, y, w))
    out = DeviceArray((100, 100), np.float64)
    want, _ = O.histogram(x, y, bins=[e, e], weights=w if weighted else None, density=density)
    for _ in range(3):                       # enqueue only; to_numpy() synchronises
        res, _ = core.histogram(dx, dy, bins=[e, e], weights=dw if weighted else None, density=density, out=out)
    got = res.to_numpy()
    if not weighted and not density:
        got = got.view(np.int64)
    assert_hist_equal(got, want, rtol=1e-6)


AXES_4D = [ax for r in (1, 2, 3) for ax in itertools.combinations(range(4), r)]


@pytest.mark.parametrize("axis", AXES_4D)
def test_device_inputs_every_axis_subset(axis):
    """Reference shape test (test_core.py:231-273) as a value test on device-resident inputs: every axis subset of a
    4-D array, including non-contiguous sets such as (1, 3) that need the device transpose."""
    r = np.random.default_rng(11)
    shape = (6, 5, 40, 33)
    a = r.standard_normal(shape).astype(np.float32)
    b = r.standard_normal(shape).astype(np.float32)
    w = r.random(shape).astype(np.float32)
    e1, e2 = np.linspace(-3, 3, 13), np.linspace(-2, 2, 8)
    da, db, dw = (DeviceArray.from_numpy(t) for t in (a, b, w))
    want, _ = O.histogram(a, b, bins=[e1, e2], axis=axis)
    got, _ = core.histogram(da, db, bins=[e1, e2], axis=axis)
    assert_hist_equal(got, want)
    wantw, _ = O.histogram(a, b, bins=[e1, e2], axis=axis, weights=w)
    gotw, _ = core.histogram(da, db, bins=[e1, e2], axis=axis, weights=dw)
    assert_hist_equal(gotw, wantw, rtol=1e-6)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("perm", [(1, 0), (0, 2, 1), (2, 0, 1), (2, 1, 0), (1, 3, 0, 2), (3, 2, 1, 0), (0, 1, 3, 2)])
def test_permute_matches_numpy(dtype, perm):
    import ctypes as C
    shape = (37, 50, 33, 9)[: len(perm)]
    a = np.random.default_rng(0).standard_normal(shape).astype(dtype)
    d = DeviceArray.from_numpy(a)
    out = DeviceArray(tuple(shape[i] for i in perm), dtype)
    shp = (C.c_int64 * len(perm))(*shape)
    prm = (C.c_int32 * len(perm))(*perm)
    _cabi.check(_cabi.lib().xh_permute(0, d.ptr, out.ptr, a.itemsize, len(perm), shp, prm), "xh_permute")
    assert np.array_equal(out.to_numpy(), np.ascontiguousarray(np.transpose(a, perm)))


class _Foreign:
    """A foreign __cuda_array_interface__ exporter over one of our buffers (dtype of our choosing)."""

    def __init__(self, dev, typestr, shape, stream=None):
        self._d = dev
        self._cai = {"shape": tuple(shape), "typestr": typestr, "data": (dev.ptr, False), "version": 3, "strides": None}
        if stream is not None:
            self._cai["stream"] = stream

    @property
    def __cuda_array_interface__(self):
        return self._cai


@pytest.mark.parametrize("typestr", ["<i4", "<i8", "<f2", "<u1"])
def test_device_arrays_of_other_dtypes_are_refused(typestr):
    d = DeviceArray((1024,), np.float64)
    f = _Foreign(d, typestr, (1024,))
    with pytest.raises(TypeError):
        core.histogram(f, bins=np.linspace(0, 1, 5))
    with pytest.raises(TypeError):
        core.histogram(f, bins=10)
    g = DeviceArray((1024,), np.float32)
    with pytest.raises(TypeError):
        core.histogram(g, bins=np.linspace(0, 1, 5), weights=_Foreign(d, typestr, (1024,)))


@pytest.mark.parametrize("stream", [None, 1, 2])
def test_foreign_cai_inputs_with_and_without_stream(stream):
    x = np.random.default_rng(9).standard_normal(300_000).astype(np.float32)
    d = DeviceArray.from_numpy(x)
    e = np.linspace(-3, 3, 50)
    h, _ = core.histogram(_Foreign(d, "<f4", x.shape, stream), bins=e)
    assert np.array_equal(h, np.histogram(x, bins=e)[0])


def test_leading_axis_with_bin_space_too_large_for_the_column_kernel():
    """(time, lat, lon) reduced over time with a 50 x 50 joint histogram used to raise NotImplementedError."""
    r = np.random.default_rng(21)
    a = r.standard_normal((300, 8, 40)).astype(np.float32)
    b = r.standard_normal((300, 8, 40)).astype(np.float32)
    e = np.linspace(-3, 3, 51)
    want, _ = O.histogram(a, b, bins=[e, e], axis=0)
    got, _ = core.histogram(a, b, bins=[e, e], axis=0)
    assert_hist_equal(got, want)
    da, db = DeviceArray.from_numpy(a), DeviceArray.from_numpy(b)
    got, _ = core.histogram(da, db, bins=[e, e], axis=0)
    assert_hist_equal(got, want)


@pytest.mark.parametrize("k", [5, 6, 8])
@pytest.mark.parametrize("weighted", [False, True])
def test_five_to_eight_variables(k, weighted):
    """k_hist<T, W, 0, 0> (runtime K): untested in round 1."""
    r = np.random.default_rng(100 + k)
    n = 200_000
    args = [r.standard_normal(n).astype(np.float32) for _ in range(k)]
    nb = {5: 6, 6: 5, 8: 3}[k]
    edges = [np.linspace(-2.5, 2.5, nb + 1) if i % 2 == 0 else np.sort(r.uniform(-2.5, 2.5, nb + 1)) for i in range(k)]
    w = r.standard_normal(n).astype(np.float64) if weighted else None
    want, _ = O.histogram(*args, bins=edges, weights=w)
    got, _ = core.histogram(*args, bins=edges, weights=w)
    assert_hist_equal(got, want, rtol=1e-6)
    if k == 5:
        rows = [a.reshape(4, -1) for a in args]
        want, _ = O.histogram(*rows, bins=edges, axis=1)
        got, _ = core.histogram(*rows, bins=edges, axis=1)
        assert_hist_equal(got, want)


def test_host_pipeline_reprobes_for_nonstationary_data():
    """Sorted input through the chunked host pipeline: every staged chunk has its mass elsewhere; results stay exact."""
    n = (1 << 23) * 2 + 12345
    r = np.random.default_rng(31)
    x = np.sort(r.standard_normal(n).astype(np.float32))
    y = r.standard_normal(n).astype(np.float32)
    w = r.random(n).astype(np.float32)
    e = np.linspace(-4, 4, 257)
    want, _ = O.histogram(x, y, bins=[e, e], weights=w, threads=8)
    got, _ = core.histogram(x, y, bins=[e, e], weights=w)
    assert_hist_equal(got, want, rtol=1e-6)
    wc, _ = O.histogram(x, y, bins=[e, e], threads=8)
    assert np.array_equal(core.histogram(x, y, bins=[e, e])[0], wc)
