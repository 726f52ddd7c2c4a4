"""Kernel time (CUDA events) of config-3-shaped calls, for A/B runs of library variants (XHIST_B200_LIB):
python tools/r2_ab.py <n> <weighted|counts|uniform|uniform_counts|rows> [calls]"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xhistogram_b200 import DeviceArray, core
n = int(float(sys.argv[1])); kind = sys.argv[2]; calls = int(sys.argv[3]) if len(sys.argv) > 3 else 8
e = np.linspace(-4, 4, 257)
axis = None
if kind.startswith("uniform"):
    x = DeviceArray.uniform((n,), np.float32, seed=13); y = DeviceArray.uniform((n,), np.float32, seed=14); e = np.linspace(0, 1, 257)
elif kind == "cfg4":
    x = DeviceArray.normal((1024, 720, 1440), np.float32, seed=6); y = DeviceArray.normal((1024, 720, 1440), np.float32, seed=7)
    e = np.linspace(-4, 4, 101); axis = [1, 2]; n = 1024 * 720 * 1440
elif kind == "rows":
    x = DeviceArray.normal((n // 100_000, 100_000), np.float32, seed=1); y = DeviceArray.normal((n // 100_000, 100_000), np.float32, seed=2)
    e = np.linspace(-4, 4, 129); axis = [1]
else:
    x = DeviceArray.normal((n,), np.float32, seed=3); y = DeviceArray.normal((n,), np.float32, seed=4)
w = DeviceArray.uniform((n,), np.float32, seed=5) if kind in ("weighted", "uniform") else None
arrays = [x, y] + ([w] if w is not None else [])
t = {}; ms = []
for _ in range(calls + 2):
    core._bincount(*arrays, weights=w is not None, axis=axis, bins=[e, e], _timing=t)
    ms.append(t["kernel_ms"])
print(f"{os.path.basename(os.environ.get('XHIST_B200_LIB', 'main')):12s} XH_PREFETCH={os.environ.get('XH_PREFETCH', 'auto'):5s} {kind:15s} n={n:.3g}  kernel_ms median {np.median(ms[2:]):.4f}  min {np.min(ms[2:]):.4f}")
