"""Seeded input cases shared by the golden-vector generator and the parity tests.

Every case is a function returning ``(args, kwargs)`` for ``histogram(*args, **kwargs)``.  Inputs
are rebuilt from seeds (numpy Generator PCG64 streams are stable for a given numpy version; the
golden file additionally stores a SHA-256 of the rebuilt inputs so that drift is detected instead
of silently comparing different data).  The cases restate, in this repo's own words, the situations
the reference's tests pin (xhistogram/test/test_core.py) and add the ones SURVEY.md §8c lists as
unpinned there (fp32-vs-fp64 edge rounding, +-inf, duplicate edges, NaN weights, K=3/4, row
regimes).
"""
from __future__ import annotations

import hashlib

import numpy as np

CASES = {}


def case(fn):
    CASES[fn.__name__] = fn
    return fn


def _rng(seed):
    return np.random.default_rng(seed)


def input_digest(args, kwargs):
    h = hashlib.sha256()
    for a in list(args) + [kwargs.get("weights")]:
        if a is not None:
            a = np.ascontiguousarray(a)
            h.update(str(a.dtype).encode()); h.update(str(a.shape).encode()); h.update(a.tobytes())
    b = kwargs.get("bins")
    for e in (b if isinstance(b, (list, tuple)) else [b]):
        if isinstance(e, np.ndarray):
            h.update(np.ascontiguousarray(e).tobytes())
        else:
            h.update(repr(e).encode())
    return h.hexdigest()


# ---------------------------------------------------------------- reference-test situations
@case
def rows_1d_f64_linspace():                       # test_core.py:25-69 (axis=1)
    x = _rng(2).standard_normal((5, 20))
    return [x], dict(bins=np.linspace(-4, 4, 10), axis=1)


@case
def flat_1d_f64_linspace_nans_density():          # test_core.py:25-69 (axis=None, add_nans, density)
    r = _rng(3)
    x = r.standard_normal((5, 20))
    x.ravel()[r.choice(x.size, 20, replace=False)] = np.nan
    return [x], dict(bins=np.linspace(-4, 4, 10), range=(-4, 4), density=True)


@case
def rows_1d_int_bins_range():                     # bins=int + range (edges from histogram_bin_edges)
    x = _rng(4).standard_normal((6, 50))
    return [x], dict(bins=10, range=(-2.5, 3.0), axis=1)


@case
def rows_1d_weighted_const2():                    # test_core.py:72-80  (2*h == h_w exactly)
    x = _rng(5).standard_normal((5, 20))
    return [x], dict(bins=np.linspace(-4, 4, 10), axis=1, weights=2 * np.ones_like(x))


@case
def rows_1d_weighted_broadcast():                 # test_core.py:84-92  weights (1, ncols)
    x = _rng(6).standard_normal((5, 20))
    return [x], dict(bins=np.linspace(-4, 4, 10), axis=1, weights=2 * np.ones((1, 20)))


@case
def right_edge_rows():                            # test_core.py:95-113
    return [np.ones((5, 20))], dict(bins=np.array([0, 0.5, 1]), axis=1)


@case
def right_edge_flat():
    return [np.ones((5, 20))], dict(bins=np.array([0, 0.5, 1]))


@case
def joint_2d_f64():                               # test_core.py:116-129
    r = _rng(7)
    return [r.standard_normal((5, 20)), r.standard_normal((5, 20))], dict(
        bins=[np.linspace(-4, 4, 10), np.linspace(-4, 4, 11)])


@case
def joint_2d_broadcast_args():                    # test_core.py:132-157  (ncols,) vs (nrows, ncols)
    r = _rng(8)
    return [r.standard_normal(20), r.standard_normal((5, 20))], dict(
        bins=[np.linspace(-4, 4, 10), np.linspace(-4, 4, 11)])


@case
def joint_2d_density_nans():                      # test_core.py:160-187
    r = _rng(9)
    a, b = r.standard_normal((5, 20)), r.standard_normal((5, 20))
    a.ravel()[r.choice(a.size, 20, replace=False)] = np.nan
    b.ravel()[r.choice(b.size, 20, replace=False)] = np.nan
    return [a, b], dict(bins=[np.linspace(-4, 4, 10), np.linspace(-4, 4, 11)], density=True)


@case
def shape_4d_axis_pairs():                        # test_core.py:231-273 (values, not only shapes)
    x = _rng(10).standard_normal((4, 5, 6, 7))
    return [x], dict(bins=np.linspace(-4, 4, 27), axis=(1, 3))


@case
def shape_4d_axis_reversed_order():
    x = _rng(11).standard_normal((4, 5, 6, 7))
    return [x], dict(bins=np.linspace(-4, 4, 27), axis=(3, 0), weights=np.abs(x) + 0.25)


@case
def shape_4d_negative_axis():
    x = _rng(12).standard_normal((4, 5, 6, 7))
    return [x], dict(bins=np.linspace(-4, 4, 27), axis=-2)


# ---------------------------------------------------------------- situations the reference leaves unpinned
def _plant(x, values, r):
    idx = r.choice(x.size, len(values), replace=False)
    x.ravel()[idx] = values
    return x


@case
def f32_data_f64_edges_near_edges():              # rule R2: fp32 data vs float64 edges that are not fp32 numbers
    r = _rng(13)
    edges = np.linspace(-1.0, 1.0, 31)            # 1/15 steps: not representable in fp32
    x = r.uniform(-1.2, 1.2, 4000).astype(np.float32)
    e32 = edges.astype(np.float32)
    near = np.concatenate([e32, np.nextafter(e32, np.float32(np.inf)), np.nextafter(e32, np.float32(-np.inf))])
    return [_plant(x, near, r)], dict(bins=edges)


@case
def f32_data_int_bins():                          # bins=int on fp32 data -> fp32 edges (numpy 2.x)
    x = _rng(14).random(5000).astype(np.float32)
    return [x], dict(bins=100)


@case
def f32_joint_rows_128():                         # BASELINE cfg2 shape, scaled down
    r = _rng(15)
    x = r.standard_normal((12, 9000)).astype(np.float32)
    y = r.standard_normal((12, 9000)).astype(np.float32)
    e = np.linspace(-4, 4, 129)
    return [x, y], dict(bins=[e, e], axis=-1)


@case
def f32_joint_weighted_density_256():             # BASELINE cfg3 shape, scaled down
    r = _rng(16)
    n = 150_000
    x = r.standard_normal(n).astype(np.float32)
    y = r.standard_normal(n).astype(np.float32)
    w = r.random(n).astype(np.float32)
    e = np.linspace(-4, 4, 257)
    return [x, y], dict(bins=[e, e], weights=w, density=True)


@case
def f64_three_vars_nonuniform_weighted():         # BASELINE cfg5 shape, scaled down
    r = _rng(17)
    n = 60_000
    args = [r.standard_normal(n) for _ in range(3)]
    edges = []
    for m in (51, 61, 71):
        e = np.sort(r.uniform(-4, 4, m)); e[0], e[-1] = -4.0, 4.0
        edges.append(e)
    return args, dict(bins=edges, weights=r.random(n))


@case
def f64_four_vars():                              # test_chunking_hypotheses.py: 1..4 variables
    r = _rng(18)
    args = [r.standard_normal((3, 2000)) for _ in range(4)]
    return args, dict(bins=[np.linspace(-3, 3, n) for n in (5, 6, 7, 8)], axis=1)


@case
def inf_and_nan_samples():                        # rule R3
    r = _rng(19)
    x = r.standard_normal(500)
    x[:6] = [np.inf, -np.inf, np.nan, 4.0, -4.0, np.nextafter(4.0, 5.0)]
    return [x], dict(bins=np.linspace(-4, 4, 17))


@case
def infinite_outer_edges():
    r = _rng(20)
    x = r.standard_normal(500) * 3
    x[:3] = [np.inf, -np.inf, np.nan]
    return [x], dict(bins=np.array([-np.inf, -1.0, 0.0, 2.0, np.inf]))


@case
def duplicate_edges():
    r = _rng(21)
    x = r.integers(0, 5, 400).astype(np.float64)
    return [x], dict(bins=np.array([0.0, 1.0, 1.0, 2.0, 3.0, 3.0, 3.0, 4.0]))


@case
def nan_weights_in_and_out_of_range():            # rule R3: NaN weight poisons its bin only when in range
    r = _rng(22)
    x = r.uniform(-1, 2, 300)
    w = r.random(300)
    x[0], w[0] = 5.0, np.nan                      # out of range: no effect
    x[1], w[1] = 0.55, np.nan                     # in range: that bin becomes NaN
    return [x], dict(bins=np.linspace(0, 1, 11), weights=w)


@case
def negative_and_integer_like_weights():
    r = _rng(23)
    x = r.standard_normal((4, 300))
    w = r.integers(-3, 4, (4, 300)).astype(np.float64)
    return [x], dict(bins=np.linspace(-3, 3, 13), axis=1, weights=w)


@case
def f32_weights_on_f64_data():
    r = _rng(24)
    x = r.standard_normal(3000)
    return [x], dict(bins=np.linspace(-3, 3, 40), weights=r.random(3000).astype(np.float32))


@case
def f64_weights_on_f32_data_rows():
    r = _rng(25)
    x = r.standard_normal((7, 1001)).astype(np.float32)   # odd row length: unaligned rows
    return [x], dict(bins=np.linspace(-3, 3, 33), axis=1, weights=r.random((7, 1001)))


@case
def mixed_f32_f64_args():
    r = _rng(26)
    return [r.standard_normal(2000).astype(np.float32), r.standard_normal(2000)], dict(
        bins=[np.linspace(-2, 2, 9), np.linspace(-2, 2, 12)])


@case
def many_rows_tiny():                             # M >> N regime
    x = _rng(27).standard_normal((3000, 7))
    return [x], dict(bins=np.linspace(-2, 2, 6), axis=1)


@case
def single_sample_and_scalar_bins():
    return [np.array([0.25])], dict(bins=4, range=(0, 1))


@case
def large_offset_uniform_edges():                 # uniform fast path with heavy cancellation (e0 >> width)
    r = _rng(28)
    x = (1000.0 + r.random(5000) * 0.01).astype(np.float32)
    return [x], dict(bins=np.linspace(1000.0, 1000.01, 21))


@case
def log_spaced_edges_f32():
    r = _rng(29)
    x = np.exp(r.uniform(-9, 9, 6000)).astype(np.float32)
    return [x], dict(bins=np.logspace(-3, 3, 37))


# ---------------------------------------------------------------- exact integer path (int64 kernel)
@case
def datetime_ns_data_day_edges():                 # test_core.py:365-382 (datetime64 data and edges of different units)
    data = np.arange("2000-06-01", "2000-06-06", dtype="datetime64[D]").astype("datetime64[ns]")
    bins = np.array(["1999-01-01", "2000-01-01", "2001-01-01"], dtype="datetime64[D]")
    return [data], dict(bins=bins)


@case
def datetime_rows_with_nat():
    r = _rng(30)
    base = np.datetime64("2001-01-01T00:00:00", "s")
    data = (base + r.integers(-400 * 86400, 400 * 86400, (6, 500)).astype("timedelta64[s]"))
    data[2, 5] = np.datetime64("NaT")
    bins = np.arange("2000-01-01", "2002-02-01", dtype="datetime64[M]")
    return [data], dict(bins=bins, axis=1)


@case
def timedelta_flat():
    r = _rng(31)
    data = r.integers(-50, 150, 2000).astype("timedelta64[h]")
    return [data], dict(bins=np.arange(0, 101, 10).astype("timedelta64[h]"))


@case
def int64_beyond_2_53_integer_edges():            # not representable in float64: compared as integers
    r = _rng(32)
    big = 2**60
    data = big + r.integers(-1000, 1000, 3000)
    bins = big + np.arange(-1000, 1001, 100)
    return [data], dict(bins=bins)


@case
def int32_joint_integer_edges_weighted():
    r = _rng(33)
    a = r.integers(0, 50, (4, 800)).astype(np.int32)
    b = r.integers(-20, 20, (4, 800)).astype(np.int16)
    return [a, b], dict(bins=[np.arange(0, 51, 5), np.arange(-20, 21, 4)], axis=1, weights=r.random((4, 800)))


# ---------------------------------------------------------------- column layout (leading / middle axes reduced)
@case
def leading_axis_time_lat_lon():                  # reduce `time` of (time, lat, lon): kept axes trail
    r = _rng(34)
    x = r.standard_normal((40, 6, 8)).astype(np.float32)
    return [x], dict(bins=np.linspace(-3, 3, 13), axis=0)


@case
def leading_axes_weighted_joint():
    r = _rng(35)
    x = r.standard_normal((7, 9, 33)); y = r.standard_normal((7, 9, 33))
    return [x, y], dict(bins=[np.linspace(-3, 3, 7), np.linspace(-2, 2, 5)], axis=(1, 0), weights=r.random((7, 9, 33)))


@case
def middle_axis_nonuniform():
    r = _rng(36)
    x = r.standard_normal((3, 50, 40))
    return [x], dict(bins=np.sort(r.uniform(-3, 3, 12)), axis=1, weights=r.standard_normal((3, 50, 40)).astype(np.float32))
