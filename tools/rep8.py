import sys, numpy as np
sys.path.insert(0, '.')
from xhistogram_b200 import DeviceArray, core
n = int(1e9)
x = DeviceArray.normal((n,), np.float32, seed=3); y = DeviceArray.normal((n,), np.float32, seed=4)
w = DeviceArray.uniform((n,), np.float32, seed=5)
e = np.linspace(-4, 4, 257)
t = {}
for i in range(8):
    core._bincount(x, y, w, weights=True, axis=None, bins=[e, e], _timing=t)
    print(i, t["kernel_ms"], flush=True)
