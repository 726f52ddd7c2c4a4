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


def test_int_bins_on_host_arrays_cross_pcie_once():
    """bins=<int> without a range: the host arrays are uploaded once and both passes (min/max, histogram) read the device copy."""
    r = np.random.default_rng(41)
    x = r.standard_normal(2_000_003).astype(np.float32)
    y = (r.standard_normal(2_000_003) * 3).astype(np.float32)
    w = r.random(2_000_003).astype(np.float32)
    h, e = core.histogram(x, bins=50)
    hw, ew = np.histogram(x, bins=50)
    assert e[0].dtype == ew.dtype and np.array_equal(e[0], ew) and np.array_equal(h, hw)
    edges_y = np.linspace(-9, 9, 31)
    h2, e2 = core.histogram(x, y, bins=[40, edges_y], weights=w, density=True)
    want, ex, ey = np.histogram2d(x, y, bins=[np.histogram_bin_edges(x, 40), edges_y], weights=w, density=True)
    assert np.array_equal(e2[0], ex)
    assert_hist_equal(h2, want, rtol=1e-6)
    xr_ = x.reshape(-1, 1)[:2_000_000].reshape(1000, 2000)
    h3, e3 = core.histogram(xr_, bins=25, axis=1)
    want3 = np.stack([np.histogram(row, bins=np.histogram_bin_edges(xr_, 25))[0] for row in xr_])
    assert np.array_equal(h3, want3)


def test_nccl_minmax_and_file_bootstrap_single_rank(tmp_path, monkeypatch):
    from xhistogram_b200 import distributed as D
    monkeypatch.setenv("RANK", "0"); monkeypatch.setenv("WORLD_SIZE", "1"); monkeypatch.setenv("LOCAL_RANK", "0")
    comm = D.NcclCommunicator.from_env()
    try:
        assert comm.allreduce_minmax(-1.5, 2.5) == (-1.5, 2.5)
        x = np.random.default_rng(3).standard_normal(100_000).astype(np.float32)
        h, e = D.histogram(x, bins=20, comm=comm)
        hw, ew = np.histogram(x, bins=20)
        assert np.array_equal(e[0], ew) and np.array_equal(h, hw)
    finally:
        comm.close()


@pytest.mark.parametrize("where", ["host", "device", "device_one_pass"])
@pytest.mark.parametrize("case", ["flat_2d_fits", "rows", "flat_big_bins", "f64_three_weights"])
def test_list_of_weights_one_pass(where, case):
    """weights=[w1, w2, ...]: one call, one histogram per weight array; equals separate oracle calls.  Host inputs take the
    one-pass kernel (k_hist_mw), device-resident inputs one fused pass per weight array (faster there) unless
    XH_FLAG_ONE_PASS asks for k_hist_mw."""
    import contextlib

    r = np.random.default_rng(50)
    if case == "flat_2d_fits":
        shape, edges, wdt, nw, axis = (1_000_003,), [np.linspace(-3, 3, 101), np.linspace(-3, 3, 81)], np.float32, 2, None
    elif case == "rows":
        shape, edges, wdt, nw, axis = (37, 20_001), [np.linspace(-3, 3, 41)], np.float32, 3, 1
    elif case == "flat_big_bins":          # 2 x 256 x 256 float64 planes do not fit shared memory: global adds
        shape, edges, wdt, nw, axis = (400_000,), [np.linspace(-3, 3, 257), np.linspace(-3, 3, 257)], np.float32, 2, None
    else:
        shape, edges, wdt, nw, axis = (300_001,), [np.sort(r.uniform(-3, 3, 30))], np.float64, 3, None
    ddt = np.float64 if case == "f64_three_weights" else np.float32
    args = [r.standard_normal(shape).astype(ddt) for _ in edges]
    ws = [r.standard_normal(shape).astype(wdt) for _ in range(nw)]
    ws[0][...] = 1.0                                             # plane 0 = the counts, as floats
    if where != "host":
        a_in, w_in = [DeviceArray.from_numpy(a) for a in args], [DeviceArray.from_numpy(w) for w in ws]
    else:
        a_in, w_in = args, ws
    forced = core.debug_flags(_cabi.XH_FLAG_ONE_PASS) if where == "device_one_pass" else contextlib.nullcontext()
    with forced:
        _check_weight_list(a_in, w_in, args, ws, edges, axis, nw)


def _check_weight_list(a_in, w_in, args, ws, edges, axis, nw):
    h, _ = core.histogram(*a_in, bins=edges, axis=axis, weights=w_in)
    assert h.shape[0] == nw
    for q in range(nw):
        want, _ = O.histogram(*args, bins=edges, axis=axis, weights=ws[q])
        assert_hist_equal(h[q], want, rtol=1e-6)
    counts, _ = O.histogram(*args, bins=edges, axis=axis)
    assert np.array_equal(h[0], counts.astype(np.float64))
    hd, _ = core.histogram(*a_in, bins=edges, axis=axis, weights=w_in[:2], density=True)
    for q in range(2):
        want, _ = O.histogram(*args, bins=edges, axis=axis, weights=ws[q], density=True)
        assert_hist_equal(hd[q], want, rtol=1e-6)


def test_list_of_weights_weighted_mean_of_the_tutorial():
    """hist(w * a) / hist(w) in one pass (reference tutorial.ipynb:298-360 does two passes over the same samples)."""
    r = np.random.default_rng(51)
    a = r.standard_normal(500_000).astype(np.float32)
    t = (20 + 5 * r.standard_normal(500_000)).astype(np.float32)
    vol = r.random(500_000).astype(np.float32)
    e = np.linspace(-3, 3, 61)
    h, _ = core.histogram(a, bins=e, weights=[vol * t, vol])
    num, _ = O.histogram(a, bins=e, weights=vol * t)
    den, _ = O.histogram(a, bins=e, weights=vol)
    np.testing.assert_allclose(h[0] / h[1], num / den, rtol=1e-9)


def test_packed_counts_take_over_when_the_window_spills():
    """Counts over 256 x 256 bins with data spread over all of them: the windowed launch spills 12 %, the library notices
    (slow-path counter of the cached verdict) and later calls use the packed 16-bit histogram; every call is bit-exact."""
    n = 6_000_000
    r = np.random.default_rng(60)
    x, y = r.random(n).astype(np.float32), r.random(n).astype(np.float32)
    e = np.linspace(0, 1, 257)
    want, _ = O.histogram(x, y, bins=[e, e], threads=8)
    dx, dy = DeviceArray.from_numpy(x), DeviceArray.from_numpy(y)
    for _ in range(8):
        assert np.array_equal(core.histogram(dx, dy, bins=[e, e])[0], want)
    with core.debug_flags(_cabi.XH_FLAG_FORCE_PACKED):
        assert np.array_equal(core.histogram(dx, dy, bins=[e, e])[0], want)
        assert np.array_equal(core.histogram(x, y, bins=[e, e])[0], want)                 # host pipeline


@pytest.mark.parametrize("layout", ["flat", "rows"])
def test_packed_counts_guard_bit_carries(layout):
    """All samples in very few bins: the 16-bit fields reach 2^15 thousands of times (and both fields of one word are hot)."""
    n = 40_000_000
    x = np.zeros(n, dtype=np.float32)
    x[1::2] = 0.03                                   # two neighbouring bins that share a shared-memory word
    x[::1001] = 7.9
    y = np.full(n, 0.5, dtype=np.float32)
    e = np.linspace(-8, 8, 513)                      # 512 x 200 bins: does not fit as 4-byte bins
    e2 = np.linspace(0, 1, 201)
    if layout == "rows":
        x, y = x.reshape(4, -1), y.reshape(4, -1)
    axis = 1 if layout == "rows" else None
    want, _ = O.histogram(x, y, bins=[e, e2], axis=axis, threads=8)
    with core.debug_flags(_cabi.XH_FLAG_FORCE_PACKED):
        got, _ = core.histogram(x, y, bins=[e, e2], axis=axis)
    assert np.array_equal(got, want)


def test_call_plan_cache_is_revalidated():
    """Repeat calls on the same DeviceArrays / edge arrays reuse the filled descriptor; edits to the edges, new data in the
    buffers and new buffers must all be noticed."""
    r = np.random.default_rng(70)
    x = r.standard_normal(300_000).astype(np.float32)
    w = r.random(300_000).astype(np.float32)
    e = np.linspace(-4, 4, 65)
    dx, dw = DeviceArray.from_numpy(x), DeviceArray.from_numpy(w)
    for _ in range(3):
        h, be = core.histogram(dx, bins=[e], weights=dw, density=True)
        assert be[0] is e
        assert_hist_equal(h, O.histogram(x, bins=e, weights=w, density=True)[0], rtol=1e-6)
    e[1:-1] += 0.021                                    # same edge object, new content
    for _ in range(2):
        assert_hist_equal(core.histogram(dx, bins=[e], weights=dw, density=True)[0], O.histogram(x, bins=e, weights=w, density=True)[0], rtol=1e-6)
    x2 = (x * 0.5 + 1).astype(np.float32)               # same buffers, new data
    _cabi.check(_cabi.lib().xh_memcpy(0, dx.ptr, x2.ctypes.data, x2.nbytes, _cabi.XH_DEVICE, _cabi.XH_HOST), "h2d")
    for _ in range(2):
        assert_hist_equal(core.histogram(dx, bins=[e], weights=dw, density=True)[0], O.histogram(x2, bins=e, weights=w, density=True)[0], rtol=1e-6)
    for _ in range(3):                                  # counts and rows through the same cache
        assert np.array_equal(core.histogram(dx, bins=[e])[0], np.histogram(x2, bins=e)[0])
    dr = dx.reshape(300, 1000)
    for _ in range(3):
        got, _ = core.histogram(dr, bins=[e], axis=1)
        assert np.array_equal(got, np.stack([np.histogram(row, bins=e)[0] for row in x2.reshape(300, 1000)]))
    dx.free()
    with pytest.raises(Exception):
        core.histogram(dx, bins=[e], weights=dw, density=True)      # a freed array is not served from the cache


def test_kept_results_do_not_alias():
    """Results are returned in pinned blocks cut from slabs (device.ResultPool): a caller who keeps many of them must find
    every one intact after later calls wrote theirs (counts, weighted, density; host and device inputs)."""
    r = np.random.default_rng(77)
    e = np.linspace(-3, 3, 129)
    kept, want = [], []
    for i in range(14):
        x = (r.standard_normal(200_000) + 0.1 * i).astype(np.float32)
        y = r.standard_normal(200_000).astype(np.float32)
        w = r.random(200_000).astype(np.float32)
        kw = [dict(), dict(weights=w), dict(weights=w, density=True)][i % 3]
        if i % 2:
            dx, dy = DeviceArray.from_numpy(x), DeviceArray.from_numpy(y)
            dkw = {k: (DeviceArray.from_numpy(v) if k == "weights" else v) for k, v in kw.items()}
            kept.append(core.histogram(dx, dy, bins=[e, e], **dkw)[0])
        else:
            kept.append(core.histogram(x, y, bins=[e, e], **kw)[0])
        want.append(O.histogram(x, y, bins=[e, e], **kw)[0])
    addrs = [h.__array_interface__["data"][0] for h in kept]
    assert len(set(addrs)) == len(addrs)
    for h, wnt in zip(kept, want):
        assert_hist_equal(h, wnt, rtol=1e-6)
