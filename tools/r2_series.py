"""Kernel time of N consecutive identical config-3 calls (is the time stable?): python tools/r2_series.py <n> <weighted|counts> <calls> [sleep_ms]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xhistogram_b200 import DeviceArray, core
n = int(float(sys.argv[1])); kind = sys.argv[2]; calls = int(sys.argv[3]); pause = float(sys.argv[4]) / 1e3 if len(sys.argv) > 4 else 0.0
e = np.linspace(-4, 4, 257)
x = DeviceArray.normal((n,), np.float32, seed=3); y = DeviceArray.normal((n,), np.float32, seed=4)
w = DeviceArray.uniform((n,), np.float32, seed=5) if kind == "weighted" else None
arrays = [x, y] + ([w] if w is not None else [])
t = {}; ms = []
for _ in range(calls):
    core._bincount(*arrays, weights=w is not None, axis=None, bins=[e, e], _timing=t)
    ms.append(t["kernel_ms"])
    if pause:
        time.sleep(pause)
print(kind, "pause", pause, " ".join(f"{v:.3f}" for v in ms))
