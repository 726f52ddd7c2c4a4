"""Parity of the CUDA path (through the C-ABI) against the golden vectors and the oracle — needs a B200."""
import threading

import numpy as np
import pytest

from oracle import hist_oracle as O
from tests.conftest import assert_hist_equal, golden_case
from tests.golden.cases import CASES
from xhistogram_b200 import DeviceArray, _cabi, core

pytestmark = pytest.mark.gpu

PATHS = {
    "default": 0,
    "global_atomics": _cabi.XH_FLAG_FORCE_GLOBAL,
    "search_only": _cabi.XH_FLAG_FORCE_SEARCH,
    "windowed": _cabi.XH_FLAG_FORCE_WINDOW,
    "windowed_search": _cabi.XH_FLAG_FORCE_WINDOW | _cabi.XH_FLAG_FORCE_SEARCH,
    "no_fx32": _cabi.XH_FLAG_NO_FX32,
    "packed_counts": _cabi.XH_FLAG_FORCE_PACKED,
}


@pytest.mark.parametrize("path", sorted(PATHS))
@pytest.mark.parametrize("name", sorted(CASES))
def test_golden_cases(golden, name, path):
    args, kwargs = CASES[name]()
    h_ref, edges_ref, _ = golden_case(golden, name)
    with core.debug_flags(PATHS[path]):
        h, edges = core.histogram(*args, **kwargs)
    for e, er in zip(edges, edges_ref):
        assert e.dtype == er.dtype and np.array_equal(e, er)
    assert_hist_equal(h, h_ref, rtol=1e-6)      # counts bit-exact; float within 1e-6 relative (north_star)


def _random_problem(seed):
    r = np.random.default_rng(1000 + seed)
    k = int(r.integers(1, 5))
    dt = [np.float32, np.float64][int(r.integers(0, 2))]
    M = int(r.choice([1, 1, 2, 5, 33, 400]))
    N = int(r.choice([1, 3, 17, 1000, 4097, 50_001]))
    if M * N > 2_000_000:
        N = 2_000_000 // M
    scale = float(r.choice([1.0, 1e-3, 1e4]))
    args = [(r.standard_normal((M, N)) * scale).astype(dt) for _ in range(k)]
    edges = []
    for _ in range(k):
        nb = int(r.choice([1, 2, 7, 64, 300])) if k <= 2 else int(r.choice([1, 3, 12, 30]))
        kind = int(r.integers(0, 3))
        if kind == 0:
            e = np.linspace(-3 * scale, 3 * scale, nb + 1)
        elif kind == 1:
            e = np.sort(r.uniform(-3 * scale, 3 * scale, nb + 1))
        else:
            e = np.linspace(-3 * scale, 3 * scale, nb + 1).astype(np.float32).astype(np.float64)
        edges.append(e)
    wkind = int(r.integers(0, 3))
    w = None if wkind == 0 else r.standard_normal((M, N)).astype([np.float32, np.float64][wkind - 1])
    # plant exact edge values, NaN and infinities
    for a, e in zip(args, edges):
        flat = a.reshape(-1)
        idx = r.choice(flat.size, min(flat.size, e.size + 3), replace=False)
        vals = np.concatenate([e, [np.nan, np.inf, -np.inf]]).astype(dt)[: idx.size]
        flat[idx] = vals
    return args, edges, w


@pytest.mark.parametrize("seed", range(40))
def test_random_problems_match_oracle(seed):
    args, edges, w = _random_problem(seed)
    want = O.block_bincount(args, edges, w)
    for path in ("default", "windowed", "global_atomics", "packed_counts"):
        with core.debug_flags(PATHS[path]):
            h, _ = core.histogram(*args, bins=edges, axis=-1, weights=w)
        assert_hist_equal(h, want, rtol=1e-6)


@pytest.mark.parametrize("k", [1, 2, 3])
@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("layout", ["flat", "rows", "short_rows"])
def test_fx32_weights(k, dtype, layout):
    """fp32 weights that are multiples of 2**-24 take the one-limb accumulation (k_hist<W=3>); planted weights that
    do not fit it (negative, NaN, inf, huge, finer than the scale) must leave through the exact global path."""
    r = np.random.default_rng(77 + k)
    shape = {"flat": (1, 600_000), "rows": (7, 90_001), "short_rows": (3000, 60)}[layout]
    args = [r.standard_normal(shape).astype(dtype) for _ in range(k)]
    nb = {1: 1000, 2: 300, 3: 40}[k]
    edges = [np.linspace(-3, 3, nb + 1)] + [np.sort(r.uniform(-3, 3, nb // 2 + 1)) for _ in range(k - 1)]
    w = r.random(shape, dtype=np.float32)
    flat = w.reshape(-1)
    idx = r.choice(flat.size, 64, replace=False)
    flat[idx] = np.resize(np.array([-0.5, np.nan, np.inf, 3.0e6, 1e-30, 0.0, -0.0, 1.75, 2.5, 0.3, -np.inf, 2.0 ** -30],
                                   dtype=np.float32), 64)
    want = O.block_bincount(args, edges, w)
    got = {}
    for path in ("default", "windowed", "no_fx32", "global_atomics"):
        with core.debug_flags(PATHS[path]):
            got[path], _ = core.histogram(*args, bins=edges, axis=-1, weights=w)
        assert_hist_equal(got[path], want, rtol=1e-6)
    if layout == "flat":   # also from device memory
        h, _ = core.histogram(*[DeviceArray.from_numpy(a.reshape(-1)) for a in args], bins=edges,
                              weights=DeviceArray.from_numpy(w.reshape(-1)))
        assert_hist_equal(h, want[0], rtol=1e-6)


@pytest.mark.parametrize("case", ["1d_f32_edges", "2d_rows", "3d_mixed", "big_B", "short_rows", "empty_row"])
def test_density_on_device(case):
    """k_density reproduces core.py:444-462 (counts / bin areas / row sums, float32 products where numpy has them):
    bit-exact for counts (integer row sums are exact), 1e-6 for weighted sums."""
    import functools
    r = np.random.default_rng(11)
    if case == "1d_f32_edges":
        args, kw = [r.random(100_000).astype(np.float32)], dict(bins=50)                       # int bins on fp32 -> fp32 edges
    elif case == "2d_rows":
        args = [r.standard_normal((9, 20_000)).astype(np.float32) for _ in range(2)]
        kw = dict(bins=[np.linspace(-3, 3, 41).astype(np.float32), np.linspace(-3, 3, 31).astype(np.float32)], axis=1)
    elif case == "3d_mixed":
        args = [r.standard_normal(50_000) for _ in range(3)]
        kw = dict(bins=[np.linspace(-3, 3, 11).astype(np.float32), np.linspace(-3, 3, 13).astype(np.float32),
                        np.sort(r.uniform(-3, 3, 9))])
    elif case == "big_B":
        args = [r.standard_normal(400_000).astype(np.float32) for _ in range(2)]
        kw = dict(bins=[np.linspace(-4, 4, 301), np.linspace(-4, 4, 201)])
    elif case == "short_rows":
        args, kw = [r.standard_normal((5000, 40))], dict(bins=np.linspace(-2, 2, 9), axis=-1)
    else:
        x = r.standard_normal((3, 1000)); x[1] = 50.0                                          # a row with nothing in range
        args, kw = [x], dict(bins=np.linspace(-2, 2, 9), axis=-1)
    for weighted in (False, True):
        w = r.random(args[0].shape, dtype=np.float32) if weighted else None
        with np.errstate(all="ignore"):
            h, edges = core.histogram(*args, weights=w, density=True, **kw)
            c, _ = core.histogram(*args, weights=w, **kw)                                       # same kernels, no density
            areas = functools.reduce(np.multiply.outer, [np.diff(e) for e in edges])
            K = len(args)
            sums = c.sum(axis=tuple(range(-K, 0)))
            want = c / areas / np.reshape(sums, sums.shape + (1,) * K)
        assert h.dtype == np.float64
        if weighted:
            assert_hist_equal(h, want, rtol=1e-6)
        else:
            assert np.array_equal(h, want, equal_nan=True)


def test_fx32_limb_wraps_are_exact():
    """Every sample in ONE bin with weights just below 1: the u32 limb wraps every ~256 adds; the sum stays exact."""
    n = 3_000_000
    x = np.full(n, 0.25, dtype=np.float32)
    w = np.full(n, 1.0 - 2.0 ** -24, dtype=np.float32)
    w[::3] = 0.5
    e = np.linspace(0, 1, 11)
    h, _ = core.histogram(x, bins=e, weights=w)
    want = np.zeros(10)
    want[2] = w.astype(np.float64).sum()      # exact in float64
    assert np.array_equal(h, want)


@pytest.mark.parametrize("dtype", [np.float32, np.float64])
@pytest.mark.parametrize("weighted", [False, True])
def test_device_resident_inputs(dtype, weighted):
    n = 3_000_001
    x = DeviceArray.normal((n,), dtype, seed=1)
    y = DeviceArray.normal((n,), dtype, seed=2)
    w = DeviceArray.uniform((n,), dtype, seed=3) if weighted else None
    e = [np.linspace(-4, 4, 101), np.linspace(-4, 4, 65)]
    h, _ = core.histogram(x, y, bins=e, weights=w)
    want, _ = O.histogram(x.to_numpy(), y.to_numpy(), bins=e, weights=None if w is None else w.to_numpy())
    assert_hist_equal(h, want)
    # rows on the device: (30, 100000) reduce the trailing axis
    xr, yr = x.flat_slice(0, 3_000_000).reshape(30, 100_000), y.flat_slice(0, 3_000_000).reshape(30, 100_000)
    wr = w.flat_slice(0, 3_000_000).reshape(30, 100_000) if weighted else None
    h, _ = core.histogram(xr, yr, bins=e, weights=wr, axis=1)
    want, _ = O.histogram(xr.to_numpy(), yr.to_numpy(), bins=e, axis=1, weights=None if wr is None else wr.to_numpy())
    assert_hist_equal(h, want)


def test_int_bins_use_device_minmax():
    x = np.random.default_rng(5).random(1_000_000).astype(np.float32)       # BASELINE config 1
    h, edges = core.histogram(x, bins=100)
    want, want_edges = np.histogram(x, bins=100)
    assert edges[0].dtype == want_edges.dtype and np.array_equal(edges[0], want_edges)
    assert np.array_equal(h, want)
    xd = DeviceArray.from_numpy(x)
    h2, e2 = core.histogram(xd, bins=100)
    assert np.array_equal(h2, want) and np.array_equal(e2[0], want_edges)
    with pytest.raises(ValueError):
        core.histogram(np.array([1.0, np.nan], dtype=np.float32), bins=4)     # numpy: autodetected range not finite


def test_cfg3_shape_slab_and_invariants():
    """BASELINE config 3 at 2**27 samples: a slab against the oracle, the rest through invariants."""
    n = 1 << 27
    x, y = DeviceArray.normal((n,), np.float32, seed=3), DeviceArray.normal((n,), np.float32, seed=4)
    w = DeviceArray.uniform((n,), np.float32, seed=5)
    e = np.linspace(-4, 4, 257)
    hw, _ = core.histogram(x, y, bins=[e, e], weights=w)
    hc, _ = core.histogram(x, y, bins=[e, e])
    # slab parity (first 4M samples) against the oracle
    m = 1 << 22
    xs, ys, ws = x.flat_slice(0, m), y.flat_slice(0, m), w.flat_slice(0, m)
    hs, _ = core.histogram(xs, ys, bins=[e, e], weights=ws)
    want, _ = O.histogram(xs.to_numpy(), ys.to_numpy(), bins=[e, e], weights=ws.to_numpy(), threads=8)
    assert_hist_equal(hs, want)
    # partition invariance: halves add up (counts exactly, weighted within tolerance)
    h1, _ = core.histogram(x.flat_slice(0, n // 2), y.flat_slice(0, n // 2), bins=[e, e])
    h2, _ = core.histogram(x.flat_slice(n // 2, n), y.flat_slice(n // 2, n), bins=[e, e])
    assert np.array_equal(h1 + h2, hc)
    # every kernel path agrees on the counts
    for path in ("global_atomics", "search_only", "windowed"):
        with core.debug_flags(PATHS[path]):
            hp, _ = core.histogram(x, y, bins=[e, e])
        assert np.array_equal(hp, hc), path
    # fp32 weights in [0,1) are multiples of 2**-24: float64 sums are exact, hence order independent
    with core.debug_flags(PATHS["global_atomics"]):
        hg, _ = core.histogram(x, y, bins=[e, e], weights=w)
    assert np.array_equal(hg, hw)
    with core.debug_flags(PATHS["no_fx32"]):          # two-limb accumulation instead of the one-limb form
        h64, _ = core.histogram(x, y, bins=[e, e], weights=w)
    assert np.array_equal(h64, hw)
    # marginal of the joint histogram equals the 1-D histogram
    hx, _ = core.histogram(x, bins=e)
    inr = core.histogram(y, bins=np.array([-4.0, 4.0]))[0][0]
    assert hc.sum() <= min(hx.sum(), inr)
    # density integrates to one
    hd, _ = core.histogram(x, y, bins=[e, e], weights=w, density=True)
    np.testing.assert_allclose((hd * np.outer(np.diff(e), np.diff(e))).sum(), 1.0, rtol=1e-12)


def test_count_above_2_to_32_in_one_bin():
    n = (1 << 32) + 12345
    x = DeviceArray((n,), np.float32)
    _cabi.check(_cabi.lib().xh_memset(0, x.ptr, 0, x.nbytes))
    h, _ = core.histogram(x, bins=np.array([-1.0, -0.5, 0.5, 1.0]))
    assert h.dtype == np.int64 and h.tolist() == [0, n, 0]
    x.free()


def test_weights_two_gives_exactly_twice_the_counts():          # reference test_core.py:72-92
    r = np.random.default_rng(11)
    x = r.standard_normal((50, 20_000)).astype(np.float32)
    bins = np.linspace(-4, 4, 10)
    h, _ = core.histogram(x, bins=bins, axis=1)
    hw, _ = core.histogram(x, bins=bins, axis=1, weights=2 * np.ones((1, 20_000), dtype=np.float32))
    assert np.array_equal(2 * h, hw)


def test_concurrent_callers():
    r = np.random.default_rng(12)
    xs = [r.standard_normal(200_000).astype(np.float32) for _ in range(6)]
    bins = np.linspace(-4, 4, 50)
    res = [None] * len(xs)

    def work(i):
        res[i] = core.histogram(xs[i], bins=bins)[0]

    th = [threading.Thread(target=work, args=(i,)) for i in range(len(xs))]
    [t.start() for t in th]; [t.join() for t in th]
    for i, x in enumerate(xs):
        assert np.array_equal(res[i], np.histogram(x, bins=bins)[0])


def test_large_host_input_pipeline():
    r = np.random.default_rng(13)
    n = 20_000_003                       # several staged chunks, odd length
    x, y = r.standard_normal(n).astype(np.float32), r.standard_normal(n).astype(np.float32)
    e = np.linspace(-4, 4, 129)
    h, _ = core.histogram(x, y, bins=[e, e])
    want, _ = O.histogram(x, y, bins=[e, e], threads=8)
    assert np.array_equal(h, want)
    xr = x[:20_000_000].reshape(200, 100_000); yr = y[:20_000_000].reshape(200, 100_000)
    h, _ = core.histogram(xr, yr, bins=[e, e], axis=1)
    want, _ = O.histogram(xr, yr, bins=[e, e], axis=1, threads=8)
    assert np.array_equal(h, want)


def test_multi_gpu_in_process(gpu_count):
    if gpu_count < 2:
        pytest.skip("needs 2 GPUs")
    r = np.random.default_rng(14)
    x, y = r.standard_normal(4_000_001).astype(np.float32), r.standard_normal(4_000_001).astype(np.float32)
    w = r.random(4_000_001).astype(np.float32)
    e = np.linspace(-4, 4, 65)
    want, _ = O.histogram(x, y, bins=[e, e], weights=w)
    h, _ = core.histogram(x, y, bins=[e, e], weights=w, devices=list(range(gpu_count)))
    assert_hist_equal(h, want)
    xr = x[:4_000_000].reshape(40, 100_000)
    h, _ = core.histogram(xr, bins=e, axis=1, devices=list(range(gpu_count)))
    assert np.array_equal(h, O.histogram(xr, bins=e, axis=1)[0])


def test_fused_allreduce_single_rank():
    """distributed.histogram over an NCCL communicator of ONE rank: histogram kernels, ncclAllReduce, density and D2H
    run as one xh_hist call (XH_FLAG_ALLREDUCE | XH_FLAG_DENSITY); with one rank the result is the local one."""
    from xhistogram_b200 import distributed as D

    r = np.random.default_rng(21)
    x, y = r.standard_normal((6, 50_001)).astype(np.float32), r.standard_normal((6, 50_001)).astype(np.float32)
    w = r.random((6, 50_001), dtype=np.float32)
    e = [np.linspace(-4, 4, 65), np.linspace(-3, 3, 33)]
    with pytest.raises(RuntimeError, match="communicator"):          # the flag needs xh_comm_init_rank first
        core._bincount(x, y, weights=False, axis=[1], bins=e, _flags=_cabi.XH_FLAG_ALLREDUCE)
    comm = D.NcclCommunicator(0, 0, 1, D.NcclCommunicator.create_unique_id())
    try:
        for kw in (dict(), dict(weights=w), dict(weights=w, density=True), dict(density=True)):
            for axis in (1, None):
                h, _ = D.histogram(x, y, bins=e, axis=axis, comm=comm, sharded_axis=1, **kw)
                with np.errstate(all="ignore"):
                    want, _ = O.histogram(x, y, bins=e, axis=axis, **kw)
                assert h.shape == want.shape
                assert_hist_equal(h, want, rtol=1e-6)
        xd, yd, wd = (DeviceArray.from_numpy(a.reshape(-1)) for a in (x, y, w))
        h, _ = D.histogram(xd, yd, bins=e, weights=wd, density=True, comm=comm, sharded_axis=0)
        assert_hist_equal(h, O.histogram(x.reshape(-1), y.reshape(-1), bins=e, weights=w.reshape(-1), density=True)[0], rtol=1e-6)
    finally:
        comm.close()


@pytest.mark.parametrize("kind", ["wide_dynamic_range", "full_mantissa_f64", "negative_mixed", "all_zero", "huge", "tiny", "inf_nan_inside"])
def test_weight_accumulation_modes(kind):
    """Weights that exercise the fixed-point scale selection, its float64 fallback and the inexact-weight spill."""
    r = np.random.default_rng(31)
    n = 400_000
    x = r.standard_normal(n).astype(np.float32)
    y = r.standard_normal(n).astype(np.float32)
    if kind == "wide_dynamic_range":
        w = np.exp(r.uniform(-40, 40, n)).astype(np.float32)          # 35 decades: most weights inexact at any one scale
    elif kind == "full_mantissa_f64":
        w = r.random(n)                                                # 53-bit mantissas: probe must choose float64 adds
    elif kind == "negative_mixed":
        w = r.standard_normal(n).astype(np.float32)
    elif kind == "all_zero":
        w = np.zeros(n, np.float32)
    elif kind == "huge":
        w = (r.random(n) * 1e30).astype(np.float32)
    elif kind == "tiny":
        w = (r.random(n) * 1e-30).astype(np.float32)
    else:
        w = r.random(n).astype(np.float32)
        w[:50] = np.inf; w[50:100] = -np.inf; w[100:150] = np.nan
    for bins in ([np.linspace(-4, 4, 33)] * 2, [np.linspace(-4, 4, 257)] * 2):     # full and windowed shared histograms
        want = O.block_bincount([x[None], y[None]], bins, w[None])[0]
        with np.errstate(invalid="ignore"):
            h, _ = core.histogram(x, y, bins=bins, weights=w)
        assert h.shape == want.shape
        if kind == "inf_nan_inside":
            # +inf and -inf in one bin give NaN in any order; elsewhere the usual bar
            assert np.array_equal(np.isnan(h), np.isnan(want))
            m = np.isfinite(want)
            assert np.array_equal(np.isinf(h), np.isinf(want))
            np.testing.assert_allclose(h[m], want[m], rtol=1e-6, atol=0)
        elif kind == "negative_mixed":
            # cancellation: compare against the scale of the terms (the float64 sum order differs from numpy's)
            scale = O.block_bincount([x[None], y[None]], bins, np.abs(w)[None])[0]
            assert np.all(np.abs(h - want) <= 1e-9 * np.maximum(scale, 1e-300))
        else:
            assert_hist_equal(h, want, rtol=1e-6)


def test_row_regimes_device_resident():
    """Rows on the device in three regimes: many short rows, few long rows, odd (unaligned) row lengths."""
    for M, N, nb in ((5000, 64, 10), (3, 1_000_003, 200), (257, 4099, 64)):
        x = DeviceArray.normal((M, N), np.float32, seed=41)
        w = DeviceArray.uniform((M, N), np.float32, seed=42)
        e = np.linspace(-3, 3, nb + 1)
        xn, wn = x.to_numpy(), w.to_numpy()
        h, _ = core.histogram(x, bins=e, axis=1)
        assert np.array_equal(h, O.histogram(xn, bins=e, axis=1)[0])
        hw, _ = core.histogram(x, bins=e, axis=1, weights=w)
        assert_hist_equal(hw, O.histogram(xn, bins=e, axis=1, weights=wn)[0])
        x.free(); w.free()


def test_cabi_rejects_bad_requests():
    x = np.zeros(10, np.float32)
    with pytest.raises(ValueError, match="monotonically"):
        core._desc_call([x.reshape(1, -1)], [10], None, 0, [np.array([0.0, 2.0, 1.0])], 1, 10, _cabi.XH_F32, _cabi.XH_NONE,
                        _cabi.XH_HOST, 0, None, 0, None)
    with pytest.raises(ValueError):
        core._desc_call([x.reshape(1, -1)], [10], None, 0, [np.array([0.0, np.nan])], 1, 10, _cabi.XH_F32, _cabi.XH_NONE,
                        _cabi.XH_HOST, 0, None, 0, None)
    h, _ = core.histogram(x, bins=np.array([1.0]))       # a single edge means zero bins, as in numpy
    assert h.shape == (0,)


@pytest.mark.parametrize("shape,axis,nb", [((500, 40, 64), 0, 30), ((4, 300, 257), 1, 12), ((3000, 2000), 0, 100),
                                           ((50, 20, 1000), (0, 1), 20), ((9, 70, 33), (1,), 400)])
def test_column_layout_leading_axes(shape, axis, nb):
    """Leading / middle axes reduced: the column-layout kernel (device-resident and through the host pipeline)."""
    x = DeviceArray.normal(shape, np.float32, seed=51)
    w = DeviceArray.uniform(shape, np.float32, seed=52)
    xn, wn = x.to_numpy(), w.to_numpy()
    e = np.linspace(-3, 3, nb + 1)
    want = O.histogram(xn, bins=e, axis=axis)[0]
    want_w = O.histogram(xn, bins=e, axis=axis, weights=wn)[0]
    h, _ = core.histogram(x, bins=e, axis=axis)                       # device-resident
    assert np.array_equal(h, want)
    hw, _ = core.histogram(x, bins=e, axis=axis, weights=w)
    assert_hist_equal(hw, want_w)
    h, _ = core.histogram(xn, bins=e, axis=axis)                      # host arrays: staged slabs of the reduced axis
    assert np.array_equal(h, want)
    hw, _ = core.histogram(xn, bins=e, axis=axis, weights=wn)
    assert_hist_equal(hw, want_w)
    x.free(); w.free()


def test_column_layout_large_host_pipeline_and_joint():
    r = np.random.default_rng(53)
    x = r.standard_normal((600, 200, 256)).astype(np.float32)         # 30.7M samples: several staged slabs
    y = r.standard_normal((600, 200, 256)).astype(np.float32)
    e = [np.linspace(-3, 3, 9), np.sort(r.uniform(-3, 3, 8))]
    h, _ = core.histogram(x, y, bins=e, axis=0)
    want = O.histogram(x, y, bins=e, axis=0, threads=8)[0]
    assert np.array_equal(h, want)
