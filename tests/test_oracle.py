"""Pin the oracle: golden vectors made by the unmodified reference, the live reference when it is
mounted, and numpy's own histogram functions (what the reference's tests assert against)."""
import numpy as np
import pytest

from oracle import hist_oracle as O
from oracle.ref_loader import load_reference_core, reference_available
from tests.conftest import assert_hist_equal, golden_case
from tests.golden.cases import CASES, input_digest


@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_golden(golden, name):
    args, kwargs = CASES[name]()
    h_ref, edges_ref, digest = golden_case(golden, name)
    assert input_digest(args, kwargs) == digest, "seeded inputs drifted from the ones the golden file was made with"
    h, edges = O.histogram(*args, **kwargs)
    assert len(edges) == len(edges_ref)
    for e, er in zip(edges, edges_ref):
        assert e.dtype == er.dtype and np.array_equal(e, er)
    # same numpy primitives in the same order: bit-identical, also for float sums
    assert h.dtype == h_ref.dtype and h.shape == h_ref.shape
    assert np.array_equal(h, h_ref, equal_nan=True)


@pytest.mark.skipif(not reference_available(), reason="/root/reference is only mounted in the build container")
@pytest.mark.parametrize("seed", range(6))
def test_oracle_matches_live_reference(seed):
    ref = load_reference_core()
    r = np.random.default_rng(100 + seed)
    shape = tuple(r.integers(2, 7, r.integers(1, 4)))
    k = int(r.integers(1, 4))
    dt = [np.float32, np.float64][seed % 2]
    args = [r.standard_normal(shape).astype(dt) for _ in range(k)]
    w = r.random(shape).astype(dt) if seed % 3 else None
    bins = [np.sort(r.uniform(-3, 3, int(r.integers(3, 9)))) for _ in range(k)]
    nd = len(shape)
    axis = None if seed % 2 else tuple(r.permutation(nd)[: int(r.integers(1, nd + 1))].tolist())
    dens = bool(seed % 2) and k < 3
    h_ref, _ = ref.histogram(*args, bins=bins, axis=axis, weights=w, density=dens, block_size=None)
    h, _ = O.histogram(*args, bins=bins, axis=axis, weights=w, density=dens)
    assert h.shape == h_ref.shape and h.dtype == h_ref.dtype
    assert np.array_equal(h, h_ref, equal_nan=True)


def test_oracle_matches_numpy_histogramdd_rows():
    r = np.random.default_rng(7)
    a, b = r.standard_normal((4, 500)), r.standard_normal((4, 500))
    w = r.random((4, 500))
    edges = [np.linspace(-3, 3, 8), np.sort(r.uniform(-3, 3, 6))]
    assert np.array_equal(O.block_bincount([a, b], edges), O.numpy_histogramdd_rows([a, b], edges))
    assert_hist_equal(O.block_bincount([a, b], edges, w), O.numpy_histogramdd_rows([a, b], edges, w), rtol=1e-12)


def test_oracle_three_var_density_matches_histogramdd():
    # the reference raises here on numpy >= 1.24 (core.py:454); its own test expects np.histogramdd
    r = np.random.default_rng(8)
    args = [r.standard_normal(400) for _ in range(3)]
    bins = [np.linspace(-4, 4, n) for n in (10, 11, 10)]
    h, _ = O.histogram(*args, bins=bins, density=True)
    want, _ = np.histogramdd(np.stack(args, -1), bins=bins, density=True)
    np.testing.assert_allclose(h, want)


def test_oracle_thread_slabs_equal_single_slab():
    r = np.random.default_rng(9)
    a, b = r.standard_normal((2, 4001)).astype(np.float32), r.standard_normal((2, 4001)).astype(np.float32)
    e = [np.linspace(-4, 4, 33)] * 2
    one = O.block_bincount([a, b], e)
    for t in (2, 3, 8):
        assert np.array_equal(O.block_bincount_threads([a, b], e, threads=t), one)


def test_oracle_property_against_numpy_and_the_live_reference():
    """Randomised (hypothesis): K = 1..3 variables, float32 / float64, non-uniform edges, samples planted exactly on edges,
    next to them, NaN and +-inf, optional weights, every axis choice — the oracle equals np.histogramdd per kept row (what the
    reference's own tests assert against) and, when the reference is mounted, the reference itself bit for bit."""
    hyp = pytest.importorskip("hypothesis")
    from hypothesis import given, settings, strategies as st

    ref = load_reference_core() if reference_available() else None

    @settings(max_examples=60, deadline=None, derandomize=True)
    @given(seed=st.integers(0, 2**31 - 1), k=st.integers(1, 3), f32=st.booleans(), weighted=st.booleans(), rows=st.integers(1, 4),
           n=st.integers(1, 200))
    def check(seed, k, f32, weighted, rows, n):
        r = np.random.default_rng(seed)
        dt = np.float32 if f32 else np.float64
        edges = [np.unique(np.round(r.uniform(-2, 2, int(r.integers(2, 9))), 2)) for _ in range(k)]
        edges = [e if len(e) >= 2 else np.array([-1.0, 1.0]) for e in edges]
        args = []
        for e in edges:
            a = r.normal(0, 1.5, (rows, n))
            m = r.random((rows, n))
            on = r.choice(e, (rows, n))
            a = np.where(m < 0.15, on, a)                                              # exactly on an edge
            a = np.where((m >= 0.15) & (m < 0.2), np.nextafter(on, np.inf), a)          # one ulp (float64) above
            a = np.where((m >= 0.2) & (m < 0.23), np.nan, a)
            a = np.where((m >= 0.23) & (m < 0.25), np.inf * np.sign(a), a)
            args.append(a.astype(dt))
        w = r.random((rows, n)).astype(dt) if weighted else None
        h, _ = O.histogram(*args, bins=edges, axis=1, weights=w)
        want = O.numpy_histogramdd_rows(args, edges, w)
        assert h.shape == (rows,) + tuple(len(e) - 1 for e in edges)
        if weighted:
            assert_hist_equal(h.reshape(rows, -1), want.reshape(rows, -1), rtol=1e-12)
        else:
            assert np.array_equal(h.reshape(rows, -1), want.reshape(rows, -1))
        if ref is not None:
            h_ref, _ = ref.histogram(*args, bins=edges, axis=1, weights=w, block_size=None)
            assert h.dtype == h_ref.dtype and np.array_equal(h, h_ref, equal_nan=True)
        flat, _ = O.histogram(*args, bins=edges, weights=w)                            # all axes reduced = sum of the rows
        if weighted:
            np.testing.assert_allclose(flat, h.sum(axis=0), rtol=1e-12, atol=1e-300)
        else:
            assert np.array_equal(flat, h.sum(axis=0))

    check()
