"""The C-ABI library loads and exports exactly what include/xhist_b200.h declares (no GPU needed)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from xhistogram_b200 import _cabi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "xhist_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"^\s*int\s+(xh_\w+)\s*\(", text, flags=re.M)))


def test_header_symbols_are_exported_and_bound():
    names = _declared_symbols()
    assert len(names) >= 20
    lib = _cabi.lib()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in the header but not exported"
        assert n in _cabi.PROTOTYPES, f"{n} has no ctypes prototype"
    assert sorted(_cabi.PROTOTYPES) == names


def test_version_and_desc_layout():
    assert _cabi.lib().xh_version() == 1
    # struct xh_desc: 8 int32 + 2 int64 + 8 ptr + 8 int64 + ptr + int64 + 8 ptr + 8 int32 + 3 ptr + 8 ptr + int64 + 8 ptr + 8 int32
    #                 + 2 int32 + 3 ptr (n_weights, reserved2, weights_more)
    assert C.sizeof(_cabi.XhDesc) == 32 + 16 + 64 + 64 + 8 + 8 + 64 + 32 + 24 + 64 + 8 + 64 + 32 + 8 + 24
    assert _cabi.lib().xh_desc_size() == C.sizeof(_cabi.XhDesc)      # the C compiler agrees with the ctypes mirror


def test_constants_match_the_header():
    """Every XH_FLAG_* / XH_ERR_* / dtype / mem constant of the Python binding has the value the header gives it."""
    text = open(os.path.join(ROOT, "include", "xhist_b200.h")).read()
    flags = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define\s+(XH_FLAG_\w+)\s+(\d+)u", text)}
    assert len(flags) >= 7
    for name, value in flags.items():
        assert getattr(_cabi, name) == value, name
    enums = {m.group(1): int(m.group(2)) for m in re.finditer(r"\b(XH_(?:OK|ERR_\w+|NONE|F32|F64|I64|HOST|DEVICE))\s*=\s*(-?\d+)", text)}
    for name in ("XH_NONE", "XH_F32", "XH_F64", "XH_I64", "XH_HOST", "XH_DEVICE"):
        assert getattr(_cabi, name) == enums[name], name
    for code, (name, _) in _cabi._ERRORS.items():
        assert enums[name] == code, name
    assert int(re.search(r"#define\s+XH_MAX_VARS\s+(\d+)", text).group(1)) == _cabi.XH_MAX_VARS


def test_density_flag_needs_widths_and_is_rejected_early():
    d = _cabi.XhDesc()
    e = (C.c_double * 3)(0.0, 1.0, 2.0)
    x = (C.c_float * 4)(0.1, 0.2, 1.5, 1.7)
    out = (C.c_double * 2)()
    d.n_vars, d.dtype, d.n_rows, d.n_cols = 1, _cabi.XH_F32, 1, 4
    d.data[0] = C.cast(x, C.c_void_p); d.row_stride[0] = 4
    d.edges[0] = C.cast(e, C.c_void_p); d.n_edges[0] = 3
    d.out = C.cast(out, C.c_void_p)
    d.flags = _cabi.XH_FLAG_DENSITY                       # no widths[]
    assert _cabi.lib().xh_hist(C.byref(d)) == -1 and "widths" in _cabi.last_error()
    d.flags = _cabi.XH_FLAG_DENSITY | _cabi.XH_FLAG_NO_ZERO
    assert _cabi.lib().xh_hist(C.byref(d)) == -1


def test_invalid_descriptor_is_rejected_without_touching_a_device():
    d = _cabi.XhDesc()
    d.n_vars = 0
    rc = _cabi.lib().xh_hist(C.byref(d))
    assert rc == -1
    assert "n_vars" in _cabi.last_error()
    with pytest.raises(ValueError):
        _cabi.check(rc, "xh_hist")


def test_no_cpu_fallback_without_a_device():
    if _cabi.device_count() > 0:
        pytest.skip("a GPU is present")
    from xhistogram_b200.core import histogram
    with pytest.raises(RuntimeError, match="XH_ERR_NO_DEVICE"):
        histogram(np.zeros(16), bins=np.linspace(0, 1, 3))
