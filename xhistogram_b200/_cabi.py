"""ctypes binding of ``libxhist_b200.so`` (declared in ``include/xhist_b200.h``).

Thin by design: the structures and prototypes below are a 1:1 transcription of the header, and
every failure of the native library is raised as a Python exception carrying the library's own
message.  There is no fallback of any kind: if the shared library is missing, or no sm_100 GPU
is usable, the call raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

XH_MAX_VARS = 8
XH_MAX_WEIGHTS = 4
XH_NONE, XH_F32, XH_F64, XH_I64 = 0, 1, 2, 3
XH_HOST, XH_DEVICE = 0, 1
XH_FLAG_NO_ZERO = 1
XH_FLAG_FORCE_GLOBAL = 2
XH_FLAG_FORCE_SEARCH = 4
XH_FLAG_FORCE_WINDOW = 8
XH_FLAG_NO_FX32 = 16
XH_FLAG_DENSITY = 32
XH_FLAG_ALLREDUCE = 64
XH_FLAG_ASYNC = 128
XH_FLAG_OUT_PINNED = 256
XH_FLAG_FORCE_PACKED = 512
XH_FLAG_ONE_PASS = 1024
XH_NCCL_UNIQUE_ID_BYTES = 128

_ERRORS = {
    -1: ("XH_ERR_INVALID", ValueError),
    -2: ("XH_ERR_CUDA", RuntimeError),
    -3: ("XH_ERR_NO_DEVICE", RuntimeError),
    -4: ("XH_ERR_NOMEM", MemoryError),
    -5: ("XH_ERR_NCCL", RuntimeError),
    -6: ("XH_ERR_UNSUPPORTED", NotImplementedError),
}


class XhDesc(C.Structure):
    """``struct xh_desc`` of include/xhist_b200.h."""

    _fields_ = [
        ("n_vars", C.c_int32),
        ("dtype", C.c_int32),
        ("w_dtype", C.c_int32),
        ("mem", C.c_int32),
        ("out_mem", C.c_int32),
        ("device", C.c_int32),
        ("flags", C.c_uint32),
        ("reserved", C.c_int32),
        ("n_rows", C.c_int64),
        ("n_cols", C.c_int64),
        ("data", C.c_void_p * XH_MAX_VARS),
        ("row_stride", C.c_int64 * XH_MAX_VARS),
        ("weights", C.c_void_p),
        ("w_row_stride", C.c_int64),
        ("edges", C.c_void_p * XH_MAX_VARS),          # const double* [XH_MAX_VARS]
        ("n_edges", C.c_int32 * XH_MAX_VARS),
        ("out", C.c_void_p),
        ("stream", C.c_void_p),
        ("kernel_ms", C.POINTER(C.c_float)),
        ("iedges", C.c_void_p * XH_MAX_VARS),         # const int64_t* [XH_MAX_VARS]
        ("n_inner", C.c_int64),
        ("widths", C.c_void_p * XH_MAX_VARS),         # const double* [XH_MAX_VARS]
        ("widths_f32", C.c_int32 * XH_MAX_VARS),
        ("n_weights", C.c_int32),
        ("reserved2", C.c_int32),
        ("weights_more", C.c_void_p * (XH_MAX_WEIGHTS - 1)),
    ]


LIB_NAME = "libxhist_b200.so"
_lib = None
_lock = threading.Lock()

# name -> (restype, argtypes): every symbol include/xhist_b200.h declares
PROTOTYPES = {
    "xh_version": (C.c_int, []),
    "xh_desc_size": (C.c_int, []),
    "xh_device_count": (C.c_int, [C.POINTER(C.c_int)]),
    "xh_init": (C.c_int, [C.c_int]),
    "xh_shutdown": (C.c_int, []),
    "xh_last_error": (C.c_int, [C.c_char_p, C.c_size_t]),
    "xh_device_info": (C.c_int, [C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int64), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "xh_hist": (C.c_int, [C.POINTER(XhDesc)]),
    "xh_last_call_phases": (C.c_int, [C.POINTER(C.c_double)]),
    "xh_hist_multi": (C.c_int, [C.POINTER(XhDesc), C.POINTER(C.c_int32), C.c_int32]),
    "xh_minmax": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int64, C.POINTER(C.c_double), C.POINTER(C.c_double)]),
    "xh_malloc": (C.c_int, [C.c_int, C.c_size_t, C.POINTER(C.c_void_p)]),
    "xh_free": (C.c_int, [C.c_int, C.c_void_p]),
    "xh_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(C.c_void_p)]),
    "xh_host_free": (C.c_int, [C.c_void_p]),
    "xh_memcpy": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_size_t, C.c_int, C.c_int]),
    "xh_memset": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_size_t]),
    "xh_sync": (C.c_int, [C.c_int]),
    "xh_stream_wait": (C.c_int, [C.c_int, C.c_void_p]),
    "xh_permute": (C.c_int, [C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.POINTER(C.c_int64), C.POINTER(C.c_int32)]),
    "xh_fill_normal": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_int64]),
    "xh_fill_uniform": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_int64]),
    "xh_timer_start": (C.c_int, [C.c_int]),
    "xh_timer_stop": (C.c_int, [C.c_int, C.POINTER(C.c_float)]),
    "xh_flush_l2": (C.c_int, [C.c_int]),
    "xh_comm_unique_id": (C.c_int, [C.c_void_p]),
    "xh_comm_init_rank": (C.c_int, [C.c_int, C.c_void_p, C.c_int, C.c_int]),
    "xh_comm_allreduce": (C.c_int, [C.c_int, C.c_void_p, C.c_int64, C.c_int]),
    "xh_comm_destroy": (C.c_int, [C.c_int]),
}


def library_path() -> str:
    """In-tree shared library; ``XHIST_B200_LIB`` points at another build of it (A/B measurements of kernel variants)."""
    return os.environ.get("XHIST_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), LIB_NAME)


def lib():
    """Load (once) and return the native library; raises ImportError when it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        path = library_path()
        if not os.path.isfile(path):
            raise ImportError(
                f"{LIB_NAME} not found at {path}: build it with `make -C xhistogram_b200/csrc` "
                "(or `python -c 'import __graft_entry__ as g; g.build()'`). There is no CPU fallback."
            )
        handle = C.CDLL(path, mode=C.RTLD_GLOBAL)
        for name, (res, args) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = res
            fn.argtypes = args
        if handle.xh_desc_size() != C.sizeof(XhDesc):
            raise ImportError(f"{LIB_NAME} was built from a different include/xhist_b200.h (struct xh_desc is "
                              f"{handle.xh_desc_size()} bytes there, {C.sizeof(XhDesc)} here): rebuild it")
        _lib = handle
        return _lib


def last_error() -> str:
    buf = C.create_string_buffer(1024)
    lib().xh_last_error(buf, len(buf))
    return buf.value.decode("utf-8", "replace")


def check(rc: int, what: str = "") -> None:
    if rc == 0:
        return
    name, exc = _ERRORS.get(rc, (f"XH_ERR_{rc}", RuntimeError))
    raise exc(f"{what + ': ' if what else ''}{name}: {last_error()}")


def device_count() -> int:
    n = C.c_int(0)
    rc = lib().xh_device_count(C.byref(n))
    return n.value if rc == 0 else 0
