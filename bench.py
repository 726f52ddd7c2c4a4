#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 histogram hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples S]

Workload (BASELINE.json configs[2], the case the metric is quoted on): 2-D weighted histogram of
two fp32 arrays of 1e9 samples with fp32 weights, bins=(256, 256) on linspace(-4, 4, 257),
density=True.  One "step" = one pass of the hot path over the whole batch.  With N > 1 (launched
by torchrun, one rank per GPU) every rank holds its own 1e9-sample shard of the sample axis (weak
scaling); the per-rank partial histograms are summed with one ncclAllReduce inside the step.

Reported on ONE JSON line by rank 0:
  value      whole-job samples/s with the inputs resident in HBM (CUDA-event timed, max over ranks)
  e2e        same metric through the public API with PINNED HOST inputs (H2D inside the timed region)
  roofline   algorithmic bytes / measured kernel time against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the oracle port (numpy restatement of the reference path) on a bounded slab, host cores
`--impl reference` times that CPU port alone (the reference arm of the driver's ratio).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "samples/s, 2-var 1e9-sample fp32 weighted histogram (256x256 bins, density)"
UNIT = "samples/s"
NBINS = 256
EDGES = np.linspace(-4.0, 4.0, NBINS + 1)
SEEDS = (3, 4, 5)  # x, y, w


def workload_name(n):
    return f"cfg3: 2 x fp32 ({n:.3g},) + fp32 weights, bins=(256,256) linspace(-4,4,257), density=True, axis=None"


# ----------------------------------------------------------------------------------------------
# clocks: sample nvidia-smi while the timed regions run
# ----------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "50",
                                          "-i", str(self.gpu_index)], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(5)
        except Exception:
            self.proc.kill()
        rows = []
        try:
            for line in open(self.path):
                f = [s.strip() for s in line.split(",")]
                if len(f) >= 9:
                    rows.append(f)
            os.unlink(self.path)
        except Exception:
            pass
        sm, reasons, mx, pw = [], set(), None, []
        for f in rows:
            try:
                sm.append(float(f[1])); mx = float(f[2]); pw.append(float(f[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            # "under load": the samples whose power draw is in the upper half of what was seen
            thr = (max(pw) + min(pw)) / 2 if pw else 0
            loaded = [s for s, p in zip(sm, pw) if p >= thr] or sm
            out.update(sm_mhz=float(np.median(loaded)), sm_max_mhz=mx, reasons=sorted(reasons), samples=len(sm),
                       power_w_max=max(pw) if pw else None)
        return out


# ----------------------------------------------------------------------------------------------
# CPU arm: the oracle port (numpy restatement of the reference path) on a bounded slab
# ----------------------------------------------------------------------------------------------
def host_slab(n, seeds=SEEDS):
    r = [np.random.default_rng(s) for s in seeds]
    x = r[0].standard_normal(n, dtype=np.float32)
    y = r[1].standard_normal(n, dtype=np.float32)
    w = r[2].random(n, dtype=np.float32)
    return x, y, w


def cpu_port_throughput(x, y, w, threads):
    from oracle import hist_oracle as O

    t0 = time.perf_counter()
    h, _ = O.histogram(x, y, bins=[EDGES, EDGES], weights=w, density=True, threads=threads)
    return x.size / (time.perf_counter() - t0), h


def pick_cpu_sample(threads, target_s=12.0):
    """Probe with 4e6 samples, then size the slab for ~target_s seconds of CPU work (bounded by memory)."""
    x, y, w = host_slab(4_000_000)
    rate, _ = cpu_port_throughput(x, y, w, threads)
    n = int(min(max(rate * target_s, 4_000_000), 2.0e8))
    return n // 1000 * 1000


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    n = args.samples_cpu or pick_cpu_sample(threads, target_s=6.0)
    x, y, w = host_slab(n)
    for _ in range(args.warmup):
        cpu_port_throughput(x[: n // 8], y[: n // 8], w[: n // 8], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_port_throughput(x, y, w, threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} samples per step of the cfg3 workload (same distributions, host numpy RNG), {threads} column slabs in {threads} threads + sum"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(1e9), "timed_on": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------
_REAL_STDOUT = None


def quiet_stdout():
    """Route everything written to fd 1 (NCCL banners, library chatter) to stderr; the JSON line goes to the saved fd."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    out = _REAL_STDOUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    quiet_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--samples", type=float, default=1e9, help="samples per GPU (default: the 1e9 of the named config)")
    ap.add_argument("--samples-cpu", type=int, default=0, help="slab size of the CPU baseline (default: sized for ~10 s)")
    ap.add_argument("--e2e-steps", type=int, default=0, help="timed end-to-end steps (default: min(steps, 5))")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    from xhistogram_b200 import DeviceArray, PinnedArray, _cabi, core, distributed as D

    dev = local_rank
    lib = _cabi.lib()
    _cabi.check(lib.xh_init(dev), "xh_init")
    core.set_default_device(dev)

    dist = None
    comm = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        comm = D.NcclCommunicator.from_torch_distributed(dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        _cabi.check(lib.xh_sync(dev), "xh_sync")

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    n = int(args.samples)
    off = rank * n                                    # every rank gets its own slab of the counter-based streams
    x = DeviceArray.normal((n,), np.float32, seed=SEEDS[0], offset=off, device=dev)
    y = DeviceArray.normal((n,), np.float32, seed=SEEDS[1], offset=off, device=dev)
    w = DeviceArray.uniform((n,), np.float32, seed=SEEDS[2], offset=off, device=dev)
    bins = [EDGES, EDGES]

    timing = {}
    kernel_ms = []

    def step_device():
        """The user call with inputs resident in HBM: histogram (+ all-reduce of the partials) + density."""
        if comm is None:
            return core.histogram(x, y, bins=bins, weights=w, density=True)[0]     # density finished on the device
        h, _ = D.histogram(x, y, bins=bins, weights=w, density=True, comm=comm, sharded_axis=0)
        return h

    # ---- parity on a slab (outside any timed region): CUDA path vs the oracle on the same samples
    parity = "skipped"
    if rank == 0:
        from oracle import hist_oracle as O
        m = min(n, 1 << 22)
        xs, ys, ws = x.flat_slice(0, m), y.flat_slice(0, m), w.flat_slice(0, m)
        got, _ = core.histogram(xs, ys, bins=bins, weights=ws, density=True)
        want, _ = O.histogram(xs.to_numpy(), ys.to_numpy(), bins=bins, weights=ws.to_numpy(), density=True, threads=8)
        gc, _ = core.histogram(xs, ys, bins=bins)
        wc, _ = O.histogram(xs.to_numpy(), ys.to_numpy(), bins=bins, threads=8)
        rel = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        parity = {"slab_samples": m, "counts_bit_exact": bool(np.array_equal(gc, wc)), "weighted_density_max_rel_err": rel}
        if not parity["counts_bit_exact"] or rel > 1e-6:
            emit({"error": "parity check failed", "parity": parity})
            return 1

    # ---- multi-GPU invariant (outside the timed region): the all-reduced COUNT histogram must hold exactly the
    #      sum over ranks of the local in-range counts (int64, bit-exact through ncclAllReduce)
    if comm is not None:
        import torch
        hg, _ = D.histogram(x, y, bins=bins, comm=comm, sharded_axis=0)
        hl, _ = core.histogram(x, y, bins=bins)
        t = torch.tensor([int(hl.sum())], dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_reduce(t)
        ok = int(hg.sum()) == int(t.item()) and hg.dtype == np.int64 and bool((hg >= hl).all())
        if rank == 0:
            parity["allreduce_counts_exact"] = bool(ok)
        if not ok:
            if rank == 0:
                emit({"error": "all-reduced histogram does not match the per-rank totals", "parity": parity})
            return 1

    sampler = ClockSampler(dev)
    # ---- device-resident timing: W warm-ups, then exactly K steps between barriers, CUDA events on the library stream
    for _ in range(args.warmup):
        step_device()
    kernel_ms.clear()
    if comm is None:
        core._timing_sink = kernel_ms      # every timed step reports the device time of its kernels (library-stream events)
    barrier()
    if rank == 0:
        sampler.start()
    _cabi.check(lib.xh_timer_start(dev), "timer")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        h_last = step_device()
    ms = _cabi.C.c_float(0)
    _cabi.check(lib.xh_timer_stop(dev, _cabi.C.byref(ms)), "timer")
    wall_ms = (time.perf_counter() - t0) * 1e3
    core._timing_sink = None
    barrier()
    dev_ms = max(ms.value, 0.0)
    # events bracket the stream work; the host-side gaps between synchronous calls are inside them as well
    total_ms = max_over_ranks(max(dev_ms, wall_ms))
    ms_per_step = total_ms / args.steps
    value = world * n / (ms_per_step * 1e-3)

    # kernel time for the roofline: CUDA events on the library stream around the kernels of every TIMED step's call
    # (probe, k_hist and its idle sibling, [all-reduce], density), averaged over the K steps
    if not kernel_ms:
        for _ in range(3):
            core._bincount(x, y, w, weights=True, axis=None, bins=bins, _timing=timing, _density_widths=[np.diff(EDGES)] * 2)
            kernel_ms.append(timing["kernel_ms"])
    k_ms = float(np.mean(kernel_ms))

    # ---- end-to-end: pinned host inputs, H2D inside the timed region, result read back
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    hx, hy, hw = (PinnedArray((n,), np.float32) for _ in range(3))
    for src, dst in ((x, hx), (y, hy), (w, hw)):
        _cabi.check(lib.xh_memcpy(dev, dst.ptr, src.ptr, src.nbytes, _cabi.XH_HOST, _cabi.XH_DEVICE), "d2h")

    def step_e2e():
        if comm is None:
            return core.histogram(hx.array, hy.array, bins=bins, weights=hw.array, density=True)[0]
        return D.histogram(hx.array, hy.array, bins=bins, weights=hw.array, density=True, comm=comm, sharded_axis=0)[0]

    step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h_e2e = step_e2e()
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * n * e2e_steps / e2e_s
    e2e_ok = bool(np.allclose(h_e2e, h_last, rtol=1e-9, atol=0))

    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_hist) against the measured HBM peak
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    alg_bytes = n * 12 + NBINS * NBINS * 8
    achieved = alg_bytes / (k_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.isfile(tpath):
        try:
            t = json.load(open(tpath))
            if int(t.get("samples", 0)) == n:
                traffic = t.get("dram_bytes_per_launch")
        except Exception:
            pass
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "kernel": "k_hist<float, W=3 (fp32 weights, one-limb fixed point), K=2, MODE=1> (+ probe, sibling, density: all kernels of the call)", "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes,
                "peak_source": peak_src}

    cpu = None
    if not args.no_cpu:
        threads = os.cpu_count() or 1
        m = args.samples_cpu or pick_cpu_sample(threads)
        m = min(m, n)
        xs, ys, ws = (a.flat_slice(0, m).to_numpy() for a in (x, y, w))
        rate, _ = cpu_port_throughput(xs, ys, ws, threads)
        rate1, _ = cpu_port_throughput(xs[: m // 8], ys[: m // 8], ws[: m // 8], 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": "port",
               "sample": f"first {m} samples of the same device-generated workload, {threads} column slabs in {threads} threads + sum",
               "single_core_value": rate1}

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(n), "samples_per_gpu": n, "sharding": "sample axis, ncclAllReduce of partial histograms" if world > 1 else "single GPU",
                   "l2": "inputs (12 B/sample, 12 GB per GPU at 1e9) exceed the 126 MB L2; no flush needed between steps",
                   "accumulate": "float64 (np.bincount semantics), fp32 compare on round-up edges"},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * n * 12, "d2h_bytes_per_step": world * NBINS * NBINS * 8,
                "steps": e2e_steps, "matches_device_result": e2e_ok},
        "gpu_launches": args.steps * 5 * world,   # per step and rank: probe, the two sibling k_hist launches (one returns at once), 2 density kernels
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "parity": parity,
        "hbm_gbs_whole_step": alg_bytes / (ms_per_step * 1e-3) / 1e9,
    }
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
