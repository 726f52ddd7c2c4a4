"""Load the UNMODIFIED reference ``xhistogram/core.py`` from a mounted checkout — TEST INFRASTRUCTURE ONLY.

The reference does an unconditional ``import dask`` (core.py:6) and dask is not installed in this
image, so a two-attribute stub is injected first: ``dask.is_dask_collection`` (the only attribute
the numpy path touches, core.py:339) returning False; ``import dask.array`` then raises
ImportError, which the reference handles itself (core.py:22-27, ``has_dask = False``).  Nothing
of the reference enters the repository history: the file is executed from where it lies — the
read-only mount ``/root/reference`` in the build container, or ``baseline/_ref/`` (git-ignored;
``__graft_entry__.build()`` places the reference package there, which is what lets it travel to the GPU box for
the CPU arm of ``bench.py`` and the drop-in test).  Used by ``tests/golden/make_golden.py`` (fixture generation),
``tests/test_oracle.py`` (live cross-check), ``tests/test_dropin_gpu.py`` and ``bench.py``'s CPU legs; all of
them skip / fall back to the numpy port when neither location exists.
"""
from __future__ import annotations

import importlib.util
import os
import sys
import types

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = [os.environ.get("XHIST_REFERENCE_ROOT"), "/root/reference", os.path.join(os.path.dirname(_HERE), "baseline", "_ref")]


def _find_root():
    for c in _CANDIDATES:
        if c and os.path.isfile(os.path.join(c, "xhistogram", "core.py")):
            return c
    return None


REFERENCE_ROOT = _find_root() or "/root/reference"


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
