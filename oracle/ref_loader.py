"""Load the UNMODIFIED reference ``xhistogram/core.py`` from a mounted checkout — TEST INFRASTRUCTURE ONLY.

The reference does an unconditional ``import dask`` (core.py:6) and dask is not installed in this
image, so a two-attribute stub is injected first: ``dask.is_dask_collection`` (the only attribute
the numpy path touches, core.py:339) returning False; ``import dask.array`` then raises
ImportError, which the reference handles itself (core.py:22-27, ``has_dask = False``).  Nothing
of the reference is copied: the file is executed from where it lies (``/root/reference``), which
exists only in the build container — never on the GPU box.  Used by
``tests/golden/make_golden.py`` (fixture generation) and ``tests/test_oracle.py`` (live
cross-check, skipped when the mount is absent).
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("XHIST_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "xhistogram", "core.py"))


def load_reference_core():
    """Return the reference's ``xhistogram.core`` module object (numpy path only)."""
    if not reference_available():
        raise FileNotFoundError(f"reference checkout not found under {REFERENCE_ROOT}")
    try:
        import dask  # noqa: F401  (a real dask, if ever present, is used as is)
    except ImportError:
        stub = types.ModuleType("dask")
        stub.is_dask_collection = lambda x: False
        sys.modules["dask"] = stub
    name = "_xhistogram_reference_core"
    if name in sys.modules:
        return sys.modules[name]
    spec = importlib.util.spec_from_file_location(name, os.path.join(REFERENCE_ROOT, "xhistogram", "core.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules[name] = mod
    spec.loader.exec_module(mod)
    return mod
