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
    assert C.sizeof(_cabi.XhDesc) == 32 + 16 + 64 + 64 + 8 + 8 + 64 + 32 + 24 + 64 + 8 + 64 + 32
    assert _cabi.lib().xh_desc_size() == C.sizeof(_cabi.XhDesc)      # the C compiler agrees with the ctypes mirror


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
