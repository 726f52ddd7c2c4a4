"""Generate tests/golden/golden.npz by running the UNMODIFIED reference on the seeded cases.

Run in the build container only (needs /root/reference):
    python tests/golden/make_golden.py
The reference is imported through oracle/ref_loader.py (dask stub, see there); every case is run
with ``block_size=None`` because the reference's default ``"auto"`` divides by zero for large flat
inputs (core.py:114-117) and block_size never changes results.  For each case the file stores the
reference histogram, its bin edges and a SHA-256 of the inputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

from oracle.ref_loader import load_reference_core  # noqa: E402
from tests.golden.cases import CASES, input_digest  # noqa: E402


def main():
    ref = load_reference_core()
    out = {}
    for name, fn in CASES.items():
        args, kwargs = fn()
        h, edges = ref.histogram(*args, block_size=None, **kwargs)
        out[f"{name}/h"] = np.asarray(h)
        for i, e in enumerate(edges):
            out[f"{name}/edges{i}"] = np.asarray(e)
        out[f"{name}/digest"] = np.array(input_digest(args, kwargs))
        print(f"{name:44s} h{tuple(h.shape)} {h.dtype}")
    np.savez_compressed(os.path.join(HERE, "golden.npz"), **out)
    print("wrote", os.path.join(HERE, "golden.npz"), os.path.getsize(os.path.join(HERE, "golden.npz")), "bytes")


if __name__ == "__main__":
    main()
