"""Run the UNMODIFIED reference on all host cores — TEST / BENCHMARK INFRASTRUCTURE ONLY.

The reference parallelises through dask: ``blockwise(_bincount)`` over chunks, then ``.sum`` over the chunked axes
(xhistogram/core.py:418-439).  dask is not installed here, so its role is played by a thread pool: equal slabs of the
reduced axis, the reference's own ``_bincount`` (core.py:197-247, ``block_size=None`` — the default ``"auto"`` divides by
zero above 1e7 flat samples, core.py:114-117) on every slab concurrently (numpy releases the GIL inside
searchsorted / bincount), partial histograms summed, and the density taken exactly as the reference's front-end does
(core.py:444-462).  Every O(samples) instruction executed is the reference's.
"""
from __future__ import annotations

import functools
import operator
from concurrent.futures import ThreadPoolExecutor

import numpy as np


def reference_histogram_threads(ref_core, *args, bins, weights=None, density=False, threads=1):
    """Flat (``axis=None``) histogram of 1-D arrays through ``ref_core._bincount`` on ``threads`` column slabs."""
    n = args[0].size
    threads = max(1, min(int(threads), n if n else 1))
    bounds = [n * t // threads for t in range(threads + 1)]
    bins = [np.asarray(b) for b in bins]

    def work(t):
        sl = slice(bounds[t], bounds[t + 1])
        arrays = [a[sl] for a in args] + ([weights[sl]] if weights is not None else [])
        return ref_core._bincount(*arrays, weights=weights is not None, axis=None, bins=bins, density=density, block_size=None)

    if threads == 1:
        parts = [work(0)]
    else:
        with ThreadPoolExecutor(threads) as ex:
            parts = list(ex.map(work, range(threads)))
    counts = functools.reduce(operator.add, parts).squeeze(0)          # dask: bin_counts.sum(drop_axes), core.py:439
    if not density:
        return counts
    widths = [np.diff(b) for b in bins]                                # core.py:444-462
    areas = widths[0] if len(bins) == 1 else functools.reduce(np.multiply.outer, widths)
    return counts / areas / counts.sum()
