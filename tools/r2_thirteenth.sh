#!/bin/bash
cat > /tmp/ab_r1.py <<'PY'
import os, sys
import numpy as np
sys.path.insert(0, sys.argv[1])
from xhistogram_b200 import DeviceArray, core
n = int(2.5e8)
for kind in ("uniform_counts", "uniform", "counts"):
    if kind.startswith("uniform"):
        x = DeviceArray.uniform((n,), np.float32, seed=13); y = DeviceArray.uniform((n,), np.float32, seed=14); e = np.linspace(0, 1, 257)
    else:
        x = DeviceArray.normal((n,), np.float32, seed=3); y = DeviceArray.normal((n,), np.float32, seed=4); e = np.linspace(-4, 4, 257)
    w = DeviceArray.uniform((n,), np.float32, seed=5) if kind == "uniform" else None
    arrays = [x, y] + ([w] if w is not None else [])
    t = {}; ms = []
    for _ in range(8):
        core._bincount(*arrays, weights=w is not None, axis=None, bins=[e, e], _timing=t); ms.append(t["kernel_ms"])
    print(sys.argv[1], kind, "kernel_ms median", round(float(np.median(ms[2:])), 4))
    for a in arrays: a.free()
PY
python /tmp/ab_r1.py $PWD/gpurun_variants/r1
python /tmp/ab_r1.py $PWD
