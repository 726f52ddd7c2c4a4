"""B200-native replacement of the xhistogram hot path (digitize -> ravel_multi_index -> bincount).

Public surface mirrors the reference package (``__all__ = ["core", "xarray"]``,
xhistogram/__init__.py:6): ``xhistogram_b200.core.histogram`` and
``xhistogram_b200.xarray.histogram`` keep the reference signatures; the O(samples) work runs in
hand-written sm_100a CUDA behind a C-ABI (include/xhist_b200.h) loaded with ctypes.
"""
from . import core  # noqa: F401
from .device import DeviceArray, PinnedArray  # noqa: F401

__version__ = "0.1.0"
__all__ = ["core", "xarray", "DeviceArray", "PinnedArray"]


def __getattr__(name):
    # xarray is optional (as in the reference's docs); import the wrapper lazily
    if name == "xarray":
        import importlib

        return importlib.import_module(".xarray", __name__)
    raise AttributeError(name)
