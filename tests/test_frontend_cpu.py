"""Host logic of xhistogram_b200.core (axis/broadcast/edges/density/dtype policy) on CPU.

The one native call (`core._desc_call` -> xh_hist) is replaced by the oracle's block kernel so that
everything around it is exercised without a GPU.  The real CUDA path is covered by the `-m gpu`
tests, which run the same cases with nothing patched."""
import numpy as np
import pytest

from oracle import hist_oracle as O
from tests.conftest import assert_hist_equal, golden_case
from tests.golden.cases import CASES
from xhistogram_b200 import core


def _oracle_desc_call(arrs, strides, w, wstride, bins, M, N, dtype, wdtype, mem, device, devices, flags, timing,
                      out_device=None, n_inner=0, density_widths=None, infos=None, w_more=None):
    def full(a, stride):
        a = np.asarray(a)
        if n_inner > 1:   # column layout: (outer, N, inner) C-contiguous -> logical rows a*inner + m
            assert a.flags.c_contiguous and a.size == M * N
            return np.ascontiguousarray(a.reshape(M // n_inner, N, n_inner).transpose(0, 2, 1)).reshape(M, N)
        assert a.flags.c_contiguous and a.shape[1] == N
        if stride == 0:
            assert a.shape[0] == 1
            return np.broadcast_to(a, (M, N))
        assert stride == N and a.shape[0] == M
        return a
    data = [full(a, s) for a, s in zip(arrs, strides)]
    assert all(a.dtype == data[0].dtype for a in data)
    from xhistogram_b200 import _cabi
    if dtype == _cabi.XH_I64:
        assert data[0].dtype == np.int64 and all(np.asarray(b).dtype == np.int64 for b in bins)
        edges = [np.asarray(b) for b in bins]
    else:
        assert data[0].dtype == {_cabi.XH_F32: np.float32, _cabi.XH_F64: np.float64}[dtype]
        edges = [np.asarray(b, dtype=np.float64) for b in bins]
    ww = None if w is None else full(w, wstride)
    B = int(np.prod([len(b) - 1 for b in bins]))
    h = O.block_bincount(data, edges, ww).reshape(M, B)
    if w_more:                         # several weight arrays over the same samples: planes (n_weights, M, B), as xh_hist returns them
        assert density_widths is None
        planes = [h.astype(np.float64)] + [O.block_bincount(data, edges, full(wq, wstride)).reshape(M, B).astype(np.float64) for wq in w_more]
        return np.concatenate(planes, axis=0)
    if density_widths is not None:     # what k_density does on the device (core.py:444-462)
        import functools
        areas = functools.reduce(np.multiply.outer, density_widths).reshape(1, B)
        with np.errstate(all="ignore"):
            h = h / areas / h.sum(axis=1, keepdims=True)
    return h


@pytest.fixture
def patched(monkeypatch):
    monkeypatch.setattr(core, "_desc_call", _oracle_desc_call)
    monkeypatch.setattr(core, "_minmax", lambda a: (float(np.min(a)), float(np.max(a))))


@pytest.mark.parametrize("name", sorted(CASES))
def test_frontend_reproduces_golden(patched, golden, name):
    args, kwargs = CASES[name]()
    h_ref, edges_ref, _ = golden_case(golden, name)
    h, edges = core.histogram(*args, **kwargs)
    for e, er in zip(edges, edges_ref):
        assert e.dtype == er.dtype and np.array_equal(e, er)
    assert_hist_equal(h, h_ref, rtol=1e-12)


@pytest.mark.parametrize("block_size", [None, 1, 2, "auto", 5])
def test_block_size_is_accepted_and_irrelevant(patched, block_size):
    x = np.random.default_rng(0).standard_normal((5, 20))
    bins = np.linspace(-4, 4, 10)
    h, _ = core.histogram(x, bins=bins, axis=1, block_size=block_size)
    want = np.stack([np.histogram(x[i], bins=bins)[0] for i in range(5)])
    assert np.array_equal(h, want)


def test_shapes_for_every_axis_choice(patched):              # reference test_core.py:231-273
    from itertools import combinations
    b = np.random.default_rng(1).standard_normal((3, 4, 5, 6))
    bins = np.linspace(-4, 4, 27)
    assert core.histogram(b, bins=bins)[0].shape == (26,)
    for axis in [(0, 1, 2, 3), (0, 1, 3, 2), (3, 2, 1, 0), (3, 2, 0, 1)]:
        assert core.histogram(b, bins=bins, axis=axis)[0].shape == (26,)
    for axis in list(range(4)) + list(range(-1, -5, -1)):
        shape = list(b.shape); del shape[axis]
        assert core.histogram(b, bins=bins, axis=axis)[0].shape == tuple(shape) + (26,)
    for i, j in combinations(range(4), 2):
        keep = tuple(b.shape[k] for k in range(4) if k not in (i, j))
        assert core.histogram(b, bins=bins, axis=(i, j))[0].shape == keep + (26,)


def test_density_integrates_to_one_per_row(patched):          # reference test_core.py:66-69 and issue #51
    r = np.random.default_rng(2)
    x = r.standard_normal((4, 300)); x[1, :50] = np.nan
    bins = np.linspace(-4, 4, 10)
    h, _ = core.histogram(x, bins=bins, axis=1, density=True)
    np.testing.assert_allclose((h * np.diff(bins)).sum(axis=1), 1.0)


def test_three_variable_density(patched):
    r = np.random.default_rng(3)
    args = [r.standard_normal(300) for _ in range(3)]
    bins = [np.linspace(-4, 4, n) for n in (10, 11, 10)]
    h, _ = core.histogram(*args, bins=bins, density=True)
    want, _ = np.histogramdd(np.stack(args, -1), bins=bins, density=True)
    np.testing.assert_allclose(h, want)


def test_weight_row_broadcast_is_passed_with_stride_zero(monkeypatch):
    seen = {}

    def spy(arrs, strides, w, wstride, *rest, **kw):
        seen["wstride"], seen["wshape"] = wstride, w.shape
        return _oracle_desc_call(arrs, strides, w, wstride, *rest, **kw)

    monkeypatch.setattr(core, "_desc_call", spy)
    x = np.random.default_rng(4).standard_normal((5, 20))
    core.histogram(x, bins=np.linspace(-4, 4, 10), axis=1, weights=2 * np.ones((1, 20)))
    assert seen == {"wstride": 0, "wshape": (1, 20)}


@pytest.mark.parametrize("bins_in,n,ok", [
    (10, 1, True), ("auto", 2, True), (np.linspace(-4, 4, 10), 2, True), ([10], 1, True),
    ([10, "auto", np.linspace(0, 1, 3)], 3, True), ([np.linspace(0, 1, 3)], 2, False), (None, 1, False),
    ([np.linspace(0, 1, 3)] * 2, 1, False)])
def test_bins_formatting(bins_in, n, ok):                     # reference test_core.py:316-340
    if ok:
        assert len(core._ensure_correctly_formatted_bins(bins_in, n)) == n
    else:
        with pytest.raises((ValueError, TypeError)):
            core._ensure_correctly_formatted_bins(bins_in, n)


@pytest.mark.parametrize("range_in,n,expected", [
    ((0, 1), 1, [(0, 1)]), ((0, 1), 2, [(0, 1), (0, 1)]), ([(0, 1), (0, 1)], 2, [(0, 1), (0, 1)]),
    ([(0,)], 1, None), ([(0, 1)], 2, None), ([(0, 1), (0, 1)], 1, None)])
def test_range_formatting(range_in, n, expected):             # reference test_core.py:343-362
    if expected is not None:
        assert core._ensure_correctly_formatted_range(range_in, n) == expected
    else:
        with pytest.raises(ValueError):
            core._ensure_correctly_formatted_range(range_in, n)


def test_errors(patched):
    x = np.zeros((3, 4))
    with pytest.raises(ValueError, match="bins must be provided"):
        core.histogram(x)
    with pytest.raises(AssertionError):
        core.histogram(x, bins=3, range=(0, 1), axis=2)
    with pytest.raises(ValueError):
        core.histogram(x, np.zeros((5, 4)), bins=3, range=(0, 1))       # not broadcastable
    with pytest.raises(TypeError):
        core.histogram(x, bins="auto", weights=np.ones_like(x))         # numpy: estimators do not take weights
    with pytest.raises(ValueError):
        core.histogram(x, bins=np.array([0.0, 2.0, 1.0]))               # edges must increase
    with pytest.raises(TypeError):                                      # datetime data needs datetime edges
        core.histogram(np.array(["2000-01-01"], dtype="datetime64[D]"), bins=np.array([0.0, 1.0]))
    with pytest.raises(TypeError):
        core.histogram(np.array([1 + 2j]), bins=np.array([0.0, 2.0]))


def test_integer_and_bool_data(patched):
    r = np.random.default_rng(5)
    xi = r.integers(-5, 6, 400).astype(np.int32)
    bins = np.linspace(-5, 5, 11)
    assert np.array_equal(core.histogram(xi, bins=bins)[0], np.histogram(xi, bins=bins)[0])
    xb = r.integers(0, 2, 100).astype(bool)
    assert np.array_equal(core.histogram(xb, bins=2, range=(0, 1))[0], np.histogram(xb, bins=2, range=(0, 1))[0])
    big = np.array([2**60 + 1], dtype=np.int64)
    with pytest.raises(TypeError):
        core.histogram(big, bins=np.array([0.0, 2.0**61]))             # float edges: the float64 cast would be lossy
    assert core.histogram(big, bins=np.array([0, 2**61]))[0].tolist() == [1]   # integer edges: exact int64 path


def test_result_pool_slabs_and_recycling(monkeypatch):
    """device.ResultPool without CUDA (the page-locking call replaced by a plain allocator): a miss takes a slab of 4, 8,
    16, ... blocks in ONE allocation, blocks are disjoint, recycled when the result array dies, and the byte limit holds."""
    import ctypes as C
    import gc

    from xhistogram_b200 import device as dev_mod

    calls, keep = [], []

    class FakeLib:
        @staticmethod
        def xh_host_alloc(nbytes, pp):
            buf = C.create_string_buffer(nbytes)
            keep.append(buf)
            calls.append(nbytes)
            C.cast(pp, C.POINTER(C.c_void_p))[0] = C.addressof(buf)
            return 0

    monkeypatch.setattr(dev_mod._cabi, "lib", lambda: FakeLib)
    pool = dev_mod.ResultPool(limit=40 << 20, largest=8 << 20)
    cap = 1 << 19                                                   # size class of a 256 x 256 float64 result
    held = [pool.array((256, 256), np.float64) for _ in range(4)]
    assert calls == [4 * cap]                                       # one allocation served four results
    addrs = [a.__array_interface__["data"][0] for a in held]
    assert len(set(addrs)) == 4 and max(addrs) - min(addrs) == 3 * cap
    for i, a in enumerate(held):
        a[...] = i                                                  # disjoint, writable
    assert [float(a[0, 0]) for a in held] == [0.0, 1.0, 2.0, 3.0]
    held.append(pool.array((65536,), np.int64))                     # same class, pool empty: the next slab has 8 blocks
    assert calls == [4 * cap, 8 * cap]
    first = held[0].__array_interface__["data"][0]
    view = held[0][:10]                                             # a view keeps the block lent
    held[0] = None
    gc.collect()
    assert first not in pool.free[cap]
    del view
    gc.collect()
    assert first in pool.free[cap]
    n_alloc = len(calls)
    for _ in range(50):                                             # steady state of a loop that keeps the previous result
        held[1] = pool.array((256, 256), np.float64)
    assert len(calls) == n_alloc
    assert pool.array((3, 5), np.float64).shape == (3, 5) and calls[-1] == 4 * (1 << 16)      # smallest class: 64 KB blocks
    assert pool.array((2 << 20,), np.float64) is None               # above `largest`: the caller falls back to a pageable array
    assert pool.array((0, 7), np.float64) is None
    big = [pool.array((1 << 20,), np.float64) for _ in range(8)]    # 8 MB blocks against the 40 MB limit
    assert sum(b is not None for b in big) >= 1 and pool.total <= pool.limit and big[-1] is None


@pytest.mark.parametrize("axis", [None, 1, (0, 2), -1])
@pytest.mark.parametrize("density", [False, True])
def test_list_of_weights_host_logic(patched, axis, density):
    """weights=[w1, w2, w3] (host inputs): broadcasting of the weight arrays, axis handling, planes -> leading axis, density per
    plane — equal to one oracle call per weight array."""
    r = np.random.default_rng(8)
    x, y = r.standard_normal((3, 4, 50)), r.standard_normal((3, 4, 50))
    ws = [r.random((3, 4, 50)), r.random((1, 4, 50)), np.full((3, 4, 50), 2.0)]          # (one of them broadcasts along axis 0)
    bins = [np.linspace(-3, 3, 9), np.linspace(-3, 3, 6)]
    h, edges = core.histogram(x, y, bins=bins, axis=axis, weights=ws, density=density)
    assert h.shape[0] == 3 and all(np.array_equal(e, b) for e, b in zip(edges, bins))
    for q, wq in enumerate(ws):
        want, _ = O.histogram(x, y, bins=bins, axis=axis, weights=np.broadcast_to(wq, x.shape), density=density)
        assert h[q].shape == want.shape
        assert_hist_equal(h[q], want, rtol=1e-12)


def test_list_of_weights_argument_errors(patched):
    x = np.zeros(10)
    with pytest.raises(ValueError, match="2 to"):
        core.histogram(x, bins=4, range=(0, 1), weights=[np.ones(10)])
    with pytest.raises(ValueError, match="2 to"):
        core.histogram(x, bins=4, range=(0, 1), weights=[np.ones(10)] * 5)
    with pytest.raises(TypeError, match="out= or devices="):
        core.histogram(x, bins=4, range=(0, 1), weights=[np.ones(10)] * 2, devices=[0])
    with pytest.raises(TypeError, match="string bin"):
        core.histogram(x, bins="auto", weights=[np.ones(10)] * 2)
