#!/bin/bash
# column-layout and weight-list kernels with / without the L2 prefetch (same box), then the bench line
cat > /tmp/cm.py <<'PY'
import os, sys, time
import numpy as np
sys.path.insert(0, os.getcwd())
from xhistogram_b200 import DeviceArray, core
T_, La, Lo = 1000, 512, 1024
x = DeviceArray.normal((T_, La, Lo), np.float32, seed=23); w = DeviceArray.uniform((T_, La, Lo), np.float32, seed=24)
t = {}
for name, wts, nb in (("cols counts 50 bins", None, 50), ("cols weighted 20 bins", w, 20)):
    e = np.linspace(-4, 4, nb + 1); ms = []
    arrays = [x] + ([wts] if wts is not None else [])
    for _ in range(6):
        core._bincount(*arrays, weights=wts is not None, axis=[0], bins=[e], _timing=t); ms.append(t["kernel_ms"])
    nbytes = T_ * La * Lo * 4 * len(arrays) + La * Lo * nb * 8
    print(f"XH_PREFETCH={os.environ.get('XH_PREFETCH','auto')} {name}: {min(ms[1:]):.4f} ms  frac {nbytes / (min(ms[1:]) * 1e-3) / 1e9 / 6551.4:.3f}")
x.free(); w.free()
n = int(5e8)
a = DeviceArray.normal((n,), np.float32, seed=3); b = DeviceArray.normal((n,), np.float32, seed=4)
w1 = DeviceArray.uniform((n,), np.float32, seed=5); w2 = DeviceArray.uniform((n,), np.float32, seed=15)
e1 = np.linspace(-4, 4, 101)
def wall(fn):
    fn(); ts = []
    for _ in range(4):
        t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
    return min(ts)
print(f"XH_PREFETCH={os.environ.get('XH_PREFETCH','auto')} weight list 5e8: one pass {wall(lambda: core.histogram(a, b, bins=[e1, e1], weights=[w1, w2])):.3f} ms, two calls {wall(lambda: (core.histogram(a, b, bins=[e1, e1], weights=w1), core.histogram(a, b, bins=[e1, e1], weights=w2))):.3f} ms")
PY
for pf in 0 1 0 1; do XH_PREFETCH=$pf python /tmp/cm.py; done
