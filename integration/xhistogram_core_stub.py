"""The binding a maintainer of xgcm/xhistogram would add: a ctypes stub that replaces ``xhistogram.core._bincount``
(reference xhistogram/core.py:197-247) with one call into ``libxhist_b200.so`` (C-ABI: include/xhist_b200.h).

Nothing else of the reference changes: ``core.histogram`` (argument handling, bin edges, dask graph, density) and the xarray
wrapper keep running on top of it.  ``install(core_module)`` applies the stub to an imported reference module —
``tests/test_dropin_gpu.py`` does exactly that with the UNMODIFIED reference and runs the reference's own numpy test cases
through it.  This file uses nothing of the ``xhistogram_b200`` Python package: only the shared library.
"""
import ctypes as C
import os

import numpy as np

_XH_MAX_VARS = 8


class _XhDesc(C.Structure):                       # struct xh_desc, include/xhist_b200.h
    _fields_ = [("n_vars", C.c_int32), ("dtype", C.c_int32), ("w_dtype", C.c_int32), ("mem", C.c_int32),
                ("out_mem", C.c_int32), ("device", C.c_int32), ("flags", C.c_uint32), ("reserved", C.c_int32),
                ("n_rows", C.c_int64), ("n_cols", C.c_int64),
                ("data", C.c_void_p * _XH_MAX_VARS), ("row_stride", C.c_int64 * _XH_MAX_VARS),
                ("weights", C.c_void_p), ("w_row_stride", C.c_int64),
                ("edges", C.c_void_p * _XH_MAX_VARS), ("n_edges", C.c_int32 * _XH_MAX_VARS),
                ("out", C.c_void_p), ("stream", C.c_void_p), ("kernel_ms", C.c_void_p),
                ("iedges", C.c_void_p * _XH_MAX_VARS),                # int64 edges for dtype XH_I64 (datetime64, integers)
                ("n_inner", C.c_int64),                               # > 1: column layout (leading axes reduced)
                ("widths", C.c_void_p * _XH_MAX_VARS),                # XH_FLAG_DENSITY: np.diff(edges_k) as float64
                ("widths_f32", C.c_int32 * _XH_MAX_VARS),             # 1: numpy holds those widths as float32
                ("n_weights", C.c_int32), ("reserved2", C.c_int32),   # > 1: several weight arrays in one pass
                ("weights_more", C.c_void_p * 3)]


def _load(path=None):
    path = path or os.environ.get("XHIST_B200_LIB") or os.path.join(
        os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "xhistogram_b200", "libxhist_b200.so")
    xh = C.CDLL(path)
    assert xh.xh_desc_size() == C.sizeof(_XhDesc)                      # the C compiler agrees with this mirror
    xh.xh_hist.argtypes = [C.POINTER(_XhDesc)]
    xh.xh_hist.restype = C.c_int
    xh.xh_last_error.argtypes = [C.c_char_p, C.c_size_t]
    return xh


def make_bincount(xh, device=0):
    _range = range

    def _bincount(*all_arrays, weights=False, axis=None, bins=None, density=None, block_size=None):
        """Drop-in for core.py:197-247: same arguments, same return shape/dtype; the digitize ->
        ravel_multi_index -> bincount work of core.py:73-194 runs in one CUDA launch."""
        a0 = all_arrays[0]
        full = (axis is None) or (set(axis) == set(_range(a0.ndim)))
        kept = (1,) * a0.ndim if full else tuple(a0.shape[i] if i not in axis else 1 for i in _range(a0.ndim))

        def rows(a):                                   # core.py:211-227, unchanged
            if full:
                return np.ascontiguousarray(a).reshape(1, -1)
            c = np.moveaxis(a, axis, tuple(_range(-len(axis), 0)))
            return np.ascontiguousarray(c.reshape(int(np.prod(c.shape[:c.ndim - len(axis)])), -1))

        arrs = [rows(np.asarray(a)) for a in all_arrays]
        w = arrs.pop() if weights else None
        bins = [np.asarray(b) for b in bins]
        if arrs[0].dtype.kind in "mM":                                 # datetime64 / timedelta64: exact int64 ticks
            common = [np.result_type(a.dtype, b.dtype) for a, b in zip(arrs, bins)]
            arrs = [np.ascontiguousarray(a.astype(c).view(np.int64)) for a, c in zip(arrs, common)]
            edges = [np.ascontiguousarray(b.astype(c).view(np.int64)) for b, c in zip(bins, common)]
            xdt = 3
        else:
            dt = np.result_type(*[a.dtype for a in arrs], np.float32)  # float32 stays float32, else float64
            arrs = [np.ascontiguousarray(a, dtype=dt) for a in arrs]
            edges = [np.ascontiguousarray(b, dtype=np.float64) for b in bins]
            xdt = 1 if dt == np.float32 else 2
        if w is not None and w.dtype not in (np.float32, np.float64):
            w = w.astype(np.float64)                                   # np.bincount casts to double (core.py:81)
        if w is not None:
            w = np.ascontiguousarray(w)
        M, N = arrs[0].shape
        nb = [len(b) - 1 for b in bins]
        out = np.zeros((M, int(np.prod(nb))), dtype=np.int64 if w is None else np.float64)
        if M * N and out.size:
            d = _XhDesc(n_vars=len(arrs), dtype=xdt, w_dtype=0 if w is None else (1 if w.dtype == np.float32 else 2),
                        mem=0, out_mem=0, device=device, n_rows=M, n_cols=N, out=out.ctypes.data)
            for k, (a, e) in enumerate(zip(arrs, edges)):
                d.data[k], d.row_stride[k] = a.ctypes.data, N
                if xdt == 3:
                    d.iedges[k] = e.ctypes.data
                else:
                    d.edges[k] = e.ctypes.data
                d.n_edges[k] = e.size
            if w is not None:
                d.weights, d.w_row_stride = w.ctypes.data, N
            if xh.xh_hist(C.byref(d)) != 0:
                msg = C.create_string_buffer(512)
                xh.xh_last_error(msg, 512)
                raise RuntimeError(msg.value.decode())
        return out.reshape(kept + tuple(nb))

    return _bincount


def install(core_module, lib_path=None, device=0):
    """Replace ``core_module._bincount`` (the reference's hot path) with the GPU stub; returns the original."""
    original = core_module._bincount
    core_module._bincount = make_bincount(_load(lib_path), device)
    return original
