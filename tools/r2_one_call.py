"""A few identical config-3-shaped calls (for ncu launch lists): python tools/r2_one_call.py <n> [counts|weighted|generic|uniform] [calls]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xhistogram_b200 import DeviceArray, core
n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 125_000_000
kind = sys.argv[2] if len(sys.argv) > 2 else "weighted"
calls = int(sys.argv[3]) if len(sys.argv) > 3 else 4
e = np.linspace(-4, 4, 257)
if kind == "uniform":
    x = DeviceArray.uniform((n,), np.float32, seed=13); y = DeviceArray.uniform((n,), np.float32, seed=14)
    e = np.linspace(0, 1, 257)
else:
    x = DeviceArray.normal((n,), np.float32, seed=3); y = DeviceArray.normal((n,), np.float32, seed=4)
w = None if kind == "counts" else DeviceArray.uniform((n,), np.float32, seed=5)
t = {}
for _ in range(calls):
    h, _ = core.histogram(x, y, bins=[e, e], weights=w, density=w is not None)
print("ok", float(np.nansum(h)))
