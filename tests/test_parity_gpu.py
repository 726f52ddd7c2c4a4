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
    for path in ("default", "windowed", "global_atomics"):
        with core.debug_flags(PATHS[path]):
            h, _ = core.histogram(*args, bins=edges, axis=-1, weights=w)
        assert_hist_equal(h, want, rtol=1e-6)


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
