#!/usr/bin/env python
"""Where a config-3 call spends its time outside k_hist (round 2).

For the headline shape at 1e9 and at the 1/8 shard (1.25e8) it prints, per variant of the data,
the wall time of the public call (device-resident inputs), the CUDA-event time of the call's
kernels and their difference = the fixed per-call cost.  Variants: the bench data, generic fp32
weights (w * pi: not multiples of 2^-24), x/y uniform over the bin range (no window can hold the
mass), counts only.

    python tools/r2_overheads.py [--reps 20]
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from xhistogram_b200 import DeviceArray, _cabi, core  # noqa: E402


def run(name, args, w, bins, reps, density=True, rows=None):
    t = {}
    arrays = list(args) + ([w] if w is not None else [])
    widths = [np.diff(b) for b in bins] if density else None
    for _ in range(3):
        core.histogram(*args, bins=bins, weights=w, density=density)
    wall, kms = [], []
    for _ in range(reps):
        t0 = time.perf_counter()
        core.histogram(*args, bins=bins, weights=w, density=density)
        wall.append((time.perf_counter() - t0) * 1e3)
    ph = (_cabi.C.c_double * 4)()
    phases = []
    for _ in range(reps):
        t0 = time.perf_counter()
        core.histogram(*args, bins=bins, weights=w, density=density)
        tot = (time.perf_counter() - t0) * 1e6
        _cabi.lib().xh_last_call_phases(ph)
        phases.append([tot] + list(ph))
    for _ in range(reps):
        core._bincount(*arrays, weights=w is not None, axis=None, bins=bins, _timing=t, _density_widths=widths)
        kms.append(t["kernel_ms"])
    n = args[0].size
    rec = 4 * len(arrays)
    B = int(np.prod([len(b) - 1 for b in bins]))
    row = dict(case=name, n=n, wall_ms_med=float(np.median(wall)), wall_ms_min=float(np.min(wall)), kernel_ms_med=float(np.median(kms)),
               kernel_ms_min=float(np.min(kms)), fixed_ms=float(np.median(wall) - np.median(kms)),
               frac_kernel=(n * rec + B * 8) / (np.median(kms) * 1e-3) / 1e9 / 6551.4,
               phases_us_med=dict(zip(["python_total", "lib_tables_ready", "lib_enqueued", "lib_synced", "lib_return"],
                                      [float(v) for v in np.median(np.array(phases), axis=0)])))
    print(json.dumps(row), flush=True)
    if rows is not None:
        rows.append(row)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=20)
    a = ap.parse_args()
    e = np.linspace(-4, 4, 257)
    bins = [e, e]
    rows = []
    for n in (int(1e9), int(1.25e8), int(1e6)):
        x = DeviceArray.normal((n,), np.float32, seed=3)
        y = DeviceArray.normal((n,), np.float32, seed=4)
        w = DeviceArray.uniform((n,), np.float32, seed=5)
        run(f"cfg3 n={n:.3g}", [x, y], w, bins, a.reps, rows=rows)
        run(f"cfg3-counts n={n:.3g}", [x, y], None, bins, a.reps, density=False, rows=rows)
        if n >= int(1.25e8):
            # generic weights: w * pi on the host is too slow for 1e9; scale on the device through a float32 host slab
            m = min(n, 1 << 24)
            ws = (w.flat_slice(0, m).to_numpy() * np.float32(np.pi)).astype(np.float32)
            reps = n // m
            wg = DeviceArray((n,), np.float32)
            for i in range(reps):
                _cabi.check(_cabi.lib().xh_memcpy(0, wg.ptr + i * m * 4, ws.ctypes.data, m * 4, _cabi.XH_DEVICE, _cabi.XH_HOST), "h2d")
            rem = n - reps * m
            if rem:
                _cabi.check(_cabi.lib().xh_memcpy(0, wg.ptr + reps * m * 4, ws.ctypes.data, rem * 4, _cabi.XH_DEVICE, _cabi.XH_HOST), "h2d")
            run(f"cfg3 generic weights (w*pi) n={n:.3g}", [x, y], wg, bins, a.reps, rows=rows)
            wg.free()
            # x, y uniform over the whole bin range: U[0,1) data against edges linspace(0, 1, 257)
            ux = DeviceArray.uniform((n,), np.float32, seed=13)
            uy = DeviceArray.uniform((n,), np.float32, seed=14)
            eu = np.linspace(0, 1, 257)
            run(f"cfg3 uniform x,y n={n:.3g}", [ux, uy], w, [eu, eu], a.reps, rows=rows)
            run(f"cfg3-counts uniform x,y n={n:.3g}", [ux, uy], None, [eu, eu], a.reps, density=False, rows=rows)
            ux.free(); uy.free()
        x.free(); y.free(); w.free()
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "r2_overheads.json"), "w") as f:
        json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
