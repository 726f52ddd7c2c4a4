"""One config-5-shaped call (for ncu): 3 x fp64, non-uniform (50,60,70) bins, fp64 weights, n samples."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xhistogram_b200 import DeviceArray, core
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 100_000_000
r = np.random.default_rng(12)
edges = []
for m in (51, 61, 71):
    ee = np.sort(r.uniform(-4, 4, m)); ee[0], ee[-1] = -4.0, 4.0
    edges.append(ee)
xs = [DeviceArray.normal((n,), np.float64, seed=8 + i) for i in range(3)]
w = DeviceArray.uniform((n,), np.float64, seed=11)
t = {}
for _ in range(3):
    h = core._bincount(*xs, w, weights=True, axis=None, bins=edges, _timing=t)
    print("kernel_ms", t["kernel_ms"], "in-range weight", float(h.sum()))
