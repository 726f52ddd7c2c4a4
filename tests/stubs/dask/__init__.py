"""Minimal stand-in for dask (see tests/stubs/README.md): just enough of ``dask.array`` for the blockwise + sum graph of
``core.histogram`` (reference xhistogram/core.py:403-439) to execute, chunk by chunk, eagerly."""
__stub__ = True


def is_dask_collection(x):
    from .array import Array
    return isinstance(x, Array)
