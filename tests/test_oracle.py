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
