#!/usr/bin/env python
"""bench.py — headline benchmark of the B200 histogram hot path.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--samples S]

Workload (BASELINE.json configs[2], the case the metric is quoted on): 2-D weighted histogram of
two fp32 arrays of 1e9 samples with fp32 weights, bins=(256, 256) on linspace(-4, 4, 257),
density=True.  One "step" = one pass of the hot path over the whole batch.  With N > 1 (launched
by torchrun, one rank per GPU) every rank holds its own 1e9-sample shard of the sample axis (weak
scaling); the per-rank partial histograms are summed inside the step (peer-memory reduction over
NVLink, or ncclAllReduce).

Reported on ONE JSON line by rank 0:
  value      whole-job samples/s with the inputs resident in HBM (CUDA-event timed, max over ranks) — weak scaling
  strong     N > 1: the SAME 1e9 samples in total (1e9/N per rank), same timing method, next to a single-GPU step over
             1e9 samples measured in the same run; also the same steps enqueued without a host round trip (out= in HBM)
  e2e        same metric through the public API with PINNED HOST inputs (H2D inside the timed region)
  roofline   algorithmic bytes / measured kernel time against MEASURED_PEAKS.json hbm_gbs
  configs    kernel-time roofline rows of the other BASELINE.json configurations and of adverse data for config 3
  parity     CUDA path vs the oracle: slab on rank 0; N > 1: the all-reduced result of every rank's first 2^22 samples
             against the oracle on their concatenation (counts bit-exact, weighted density <= 1e-6), kept-axis sharding
             with gather, and the single-process multi-device entry point
  cpu_baseline  the reference's own numpy path (baseline/_ref when present, else the oracle port) on a bounded slab
`--impl reference` times that CPU path alone (the reference arm of the driver's ratio).
"""
from __future__ import annotations

import argparse
import gc
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

    PERIOD_MS = os.environ.get("XH_BENCH_SMI_MS", "50")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", self.PERIOD_MS,
                                          "-i", str(self.gpu_index)], stdout=fd, stderr=subprocess.DEVNULL)
            os.close(fd)
        except Exception:
            self.proc = None

    def wait_ready(self, timeout=5.0):
        """Block until nvidia-smi has written its first sample: its start-up (driver / NVML initialisation, ~0.1-0.5 s) must
        not overlap a timed region — it was seen to add 0.2-0.7 ms to every step it overlapped."""
        if self.proc is None:
            return
        t0 = time.time()
        while time.time() - t0 < timeout:
            try:
                if os.path.getsize(self.path) > 0:
                    return
            except OSError:
                pass
            time.sleep(0.02)

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
# CPU arm: the reference's own numpy path on all host cores (baseline/_ref, placed by __graft_entry__.build()), or the
# oracle port (numpy restatement) when the reference package is not there
# ----------------------------------------------------------------------------------------------
def host_slab(n, seeds=SEEDS):
    r = [np.random.default_rng(s) for s in seeds]
    x = r[0].standard_normal(n, dtype=np.float32)
    y = r[1].standard_normal(n, dtype=np.float32)
    w = r[2].random(n, dtype=np.float32)
    return x, y, w


_CPU_IMPL = None


def cpu_impl():
    """("reference", fn) with the unmodified reference's _bincount on thread slabs, else ("port", fn) with the oracle."""
    global _CPU_IMPL
    if _CPU_IMPL is None:
        try:
            from oracle import ref_loader, ref_runner
            if not ref_loader.reference_available():
                raise FileNotFoundError
            rc = ref_loader.load_reference_core()

            def run(x, y, w, threads):
                return ref_runner.reference_histogram_threads(rc, x, y, bins=[EDGES, EDGES], weights=w, density=True, threads=threads)
            _CPU_IMPL = ("reference", run, f"unmodified xhistogram.core._bincount from {os.path.relpath(ref_loader.REFERENCE_ROOT, ROOT)} "
                         "(block_size=None), dask's blockwise + sum played by a thread pool over column slabs")
        except Exception:
            from oracle import hist_oracle as O

            def run(x, y, w, threads):
                return O.histogram(x, y, bins=[EDGES, EDGES], weights=w, density=True, threads=threads)[0]
            _CPU_IMPL = ("port", run, "oracle/hist_oracle.py (numpy restatement of the reference path), column slabs in threads + sum")
    return _CPU_IMPL


def cpu_throughput(x, y, w, threads):
    t0 = time.perf_counter()
    h = cpu_impl()[1](x, y, w, threads)
    return x.size / (time.perf_counter() - t0), h


def pick_cpu_sample(threads, target_s=12.0):
    """Probe with 4e6 samples, then size the slab for ~target_s seconds of CPU work (bounded by memory)."""
    x, y, w = host_slab(4_000_000)
    rate, _ = cpu_throughput(x, y, w, threads)
    n = int(min(max(rate * target_s, 4_000_000), 2.0e8))
    return n // 1000 * 1000


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    threads = os.cpu_count() or 1
    kind, _, how = cpu_impl()
    n = args.samples_cpu or pick_cpu_sample(threads, target_s=6.0)
    x, y, w = host_slab(n)
    for _ in range(args.warmup):
        cpu_throughput(x[: n // 8], y[: n // 8], w[: n // 8], threads)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        cpu_throughput(x, y, w, threads)
    dt = time.perf_counter() - t0
    value = n * args.steps / dt
    sample = f"{n} samples per step of the cfg3 workload (same distributions, host numpy RNG), {threads} column slabs in {threads} threads + sum; {how}"
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_name(1e9), "timed_on": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": sample},
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


def pin_to_gpu_numa_node(dev):
    """Run this rank (and allocate its pinned buffers) on the NUMA node its GPU hangs off.  Returns what was done."""
    try:
        bus = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(dev)],
                             capture_output=True, text=True, timeout=20).stdout.strip().lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        node = int(open(f"/sys/bus/pci/devices/{bus}/numa_node").read())
        if node < 0:
            # no NUMA information (one node, or a virtualised topology): give every rank its own block of cores, so that the
            # ranks' Python threads neither migrate nor share a core
            world, lr = int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))
            cores = sorted(os.sched_getaffinity(0))
            if world > 1 and len(cores) >= 2 * world:
                per = len(cores) // world
                os.sched_setaffinity(0, set(cores[lr * per:(lr + 1) * per]))
                return {"numa_node": node, "pinned": False, "cores_for_this_rank": per}
            return {"numa_node": node, "pinned": False}
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            a, _, b = part.partition("-")
            cpus.update(range(int(a), int(b or a) + 1))
        cpus &= os.sched_getaffinity(0)
        if cpus:
            os.sched_setaffinity(0, cpus)
        return {"numa_node": node, "pinned": bool(cpus), "cpus": len(cpus)}
    except Exception as e:        # no sysfs / no permission: run unpinned
        return {"numa_node": None, "pinned": False, "why": type(e).__name__}


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
    ap.add_argument("--no-configs", action="store_true", help="skip the rows of the other configurations")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference_arm(args)

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        args.gpus = world

    dev = local_rank
    numa = pin_to_gpu_numa_node(dev)

    from xhistogram_b200 import DeviceArray, PinnedArray, _cabi, core, distributed as D

    lib = _cabi.lib()
    _cabi.check(lib.xh_init(dev), "xh_init")
    core.set_default_device(dev)

    dist = None
    comm = None
    host_group = None
    if world > 1:
        import torch
        import torch.distributed as dist

        torch.cuda.set_device(dev)
        dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
        host_group = dist.new_group(backend="gloo")          # host-side barriers that keep no GPU spinning
        comm = D.NcclCommunicator.from_env(dev)               # the library's own bootstrap (file store); torch.distributed above is
                                                              # only this script's timing plumbing (barrier, max over ranks)

    def barrier():
        if dist is not None:
            dist.barrier()
        _cabi.check(lib.xh_sync(dev), "xh_sync")

    def host_barrier():
        if dist is not None:
            dist.barrier(group=host_group)

    def max_over_ranks(v):
        if dist is None:
            return v
        import torch
        t = torch.tensor([v], dtype=torch.float64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    def all_ok(ok):
        if dist is None:
            return bool(ok)
        import torch
        t = torch.tensor([1 if ok else 0], dtype=torch.int64, device=f"cuda:{dev}")
        dist.all_reduce(t, op=dist.ReduceOp.MIN)
        return bool(t.item())

    n = int(args.samples)
    off = rank * n                                    # every rank gets its own slab of the counter-based streams
    x = DeviceArray.normal((n,), np.float32, seed=SEEDS[0], offset=off, device=dev)
    y = DeviceArray.normal((n,), np.float32, seed=SEEDS[1], offset=off, device=dev)
    w = DeviceArray.uniform((n,), np.float32, seed=SEEDS[2], offset=off, device=dev)
    bins = [EDGES, EDGES]

    def timed_steps(fn, steps, warmup, after_warmup=None):
        """W warm-ups, then exactly K steps between barriers; CUDA events on the library stream and the host clock,
        max over ranks.  Returns (ms per step, last result)."""
        last = None
        for _ in range(warmup):
            last = fn()        # (kept while the next call runs, as in the timed loop: the result pool reaches its steady-state size —
                               # page-locking one more 512 KB block costs 4-15 ms — during the warm-up, not in timed step 2)
        if after_warmup is not None:
            after_warmup()
        gc.collect()
        gc.disable()          # (as timeit does: a generation-2 collection inside a 5-step region costs several steps)
        try:
            barrier()
            _cabi.check(lib.xh_timer_start(dev), "timer")
            t0 = time.perf_counter()
            trace = [] if os.environ.get("XH_BENCH_TRACE") else None
            for _ in range(steps):
                if trace is not None:
                    ts = time.perf_counter()
                last = fn()
                if trace is not None:
                    trace.append(round((time.perf_counter() - ts) * 1e3, 3))
            if trace is not None:
                print(f"[trace rank {rank}] {getattr(fn, '__name__', 'step')}: {trace}", file=sys.stderr, flush=True)
            ms = _cabi.C.c_float(0)
            _cabi.check(lib.xh_timer_stop(dev, _cabi.C.byref(ms)), "timer")      # synchronises the library stream
            wall_ms = (time.perf_counter() - t0) * 1e3
        finally:
            gc.enable()
        barrier()
        # events bracket the stream work; the host-side gaps between synchronous calls are inside them as well
        return max_over_ranks(max(ms.value, wall_ms)) / steps, last

    def step_device():
        """The user call with inputs resident in HBM: histogram (+ reduction of the partials over NVLink) + density."""
        if comm is None:
            return core.histogram(x, y, bins=bins, weights=w, density=True)[0]     # density finished on the device
        return D.histogram(x, y, bins=bins, weights=w, density=True, comm=comm, sharded_axis=0)[0]

    # ---- parity on a slab (outside any timed region): CUDA path vs the oracle on the same samples
    from oracle import hist_oracle as O
    parity = {}
    if rank == 0:
        m = min(n, 1 << 22)
        xs, ys, ws = x.flat_slice(0, m), y.flat_slice(0, m), w.flat_slice(0, m)
        got, _ = core.histogram(xs, ys, bins=bins, weights=ws, density=True)
        want, _ = O.histogram(xs.to_numpy(), ys.to_numpy(), bins=bins, weights=ws.to_numpy(), density=True, threads=8)
        got_c, _ = core.histogram(xs, ys, bins=bins)
        want_c, _ = O.histogram(xs.to_numpy(), ys.to_numpy(), bins=bins, threads=8)
        rel = float(np.max(np.abs(got - want)) / np.max(np.abs(want)))
        parity = {"slab_samples": m, "counts_bit_exact": bool(np.array_equal(got_c, want_c)), "weighted_density_max_rel_err": rel}
    ok = (not parity) or (parity["counts_bit_exact"] and parity["weighted_density_max_rel_err"] <= 1e-6)
    if not all_ok(ok):
        if rank == 0:
            emit({"error": "parity check failed", "parity": parity})
        return 1

    # ---- multi-rank parity (outside the timed regions), bin by bin against the oracle
    if comm is not None:
        mr = multi_rank_parity(rank, world, dev, n, x, y, w, comm, bins, D, core, O, DeviceArray, _cabi, host_barrier)
        okr = all_ok(mr.pop("_ok"))
        if rank == 0:
            parity.update(mr)
        if not okr:
            if rank == 0:
                emit({"error": "multi-rank parity check failed", "parity": parity})
            return 1

    sampler = ClockSampler(dev)
    # ---- device-resident timing (weak scaling: n samples per rank)
    kernel_ms = []
    if rank == 0:
        sampler.start()
        sampler.wait_ready()
    def sink_on():
        if comm is None:
            core._timing_sink = kernel_ms  # every TIMED step reports the device time of its kernels (library-stream events)
    ms_per_step, h_last = timed_steps(step_device, args.steps, args.warmup, sink_on)
    core._timing_sink = None
    value = world * n / (ms_per_step * 1e-3)

    # kernel time for the roofline: CUDA events on the library stream around the kernels of every TIMED step's call
    # (k_hist, density; the probe only on the first call of a buffer set), averaged over the K steps
    timing = {}
    if not kernel_ms:
        for _ in range(3):
            core._bincount(x, y, w, weights=True, axis=None, bins=bins, _timing=timing, _density_widths=[np.diff(EDGES)] * 2)
            kernel_ms.append(timing["kernel_ms"])
    k_ms = float(np.mean(kernel_ms))

    # ---- strong scaling: the same 1e9 samples in total, 1/N per rank, same timing method
    strong = None
    if comm is not None:
        ns = n // world
        xs_, ys_, ws_ = x.flat_slice(0, ns), y.flat_slice(0, ns), w.flat_slice(0, ns)

        def step_strong():
            return D.histogram(xs_, ys_, bins=bins, weights=ws_, density=True, comm=comm, sharded_axis=0)[0]

        def step_single():          # one GPU over all n samples, no collective: the N = 1 step, measured in this run
            return core.histogram(x, y, bins=bins, weights=w, density=True)[0]

        ms_strong, _ = timed_steps(step_strong, args.steps, args.warmup)
        ms_single, _ = timed_steps(step_single, args.steps, args.warmup)
        outd = DeviceArray((NBINS, NBINS), np.float64, device=dev)

        def step_strong_enqueue():  # result stays in HBM (out=): no host round trip between steps
            D.histogram(xs_, ys_, bins=bins, weights=ws_, density=True, comm=comm, sharded_axis=0, out=outd)

        ms_enq, _ = timed_steps(step_strong_enqueue, args.steps, args.warmup)
        h_enq = outd.to_numpy()
        h_sync = step_strong()
        strong = {"samples_total": n, "samples_per_gpu": ns, "ms_per_step": ms_strong, "value": n / (ms_strong * 1e-3),
                  "single_gpu_ms_per_step_same_run": ms_single, "speedup_vs_single_gpu_same_run": ms_single / ms_strong,
                  "enqueue_only": {"ms_per_step": ms_enq, "value": n / (ms_enq * 1e-3), "speedup_vs_single_gpu_same_run": ms_single / ms_enq,
                                   "what": "the same steps with out= (result kept in HBM, no host synchronisation between steps); one sync after the K steps",
                                   "matches_synchronous_result": bool(np.array_equal(h_enq, h_sync))},
                  "timing": "barrier + sync, K steps, CUDA events on the library stream and host clock, max over ranks; every step returns the finished float64 density to the host"}

    # ---- end-to-end: pinned host inputs, H2D inside the timed region, result read back
    e2e_steps = args.e2e_steps or min(args.steps, 5)
    hx, hy, hw = (PinnedArray((n,), np.float32) for _ in range(3))
    for src, dst in ((x, hx), (y, hy), (w, hw)):
        _cabi.check(lib.xh_memcpy(dev, dst.ptr, src.ptr, src.nbytes, _cabi.XH_HOST, _cabi.XH_DEVICE), "d2h")

    def step_e2e():
        if comm is None:
            return core.histogram(hx.array, hy.array, bins=bins, weights=hw.array, density=True)[0]
        return D.histogram(hx.array, hy.array, bins=bins, weights=hw.array, density=True, comm=comm, sharded_axis=0)[0]

    h_e2e = step_e2e()
    h_e2e = step_e2e()            # (the second warm-up call runs while the first result is held, as every timed step does)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        h_e2e = step_e2e()
    mine_s = time.perf_counter() - t0
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    clocks = sampler.stop() if rank == 0 else None
    e2e_value = world * n * e2e_steps / e2e_s
    e2e_ok = bool(np.allclose(h_e2e, h_last, rtol=1e-9, atol=0))
    h2d_gbs_min = -max_over_ranks(-(n * 12 * e2e_steps / mine_s / 1e9))      # slowest rank's own H2D rate
    for a in (hx, hy, hw):
        a.free()

    # ---- the other configurations and adverse data for config 3 (kernel-time roofline rows, 3 repetitions each)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (of measured)"
    else:
        peak, peak_src = 6650.0, "B200_PROFILING.md fallback (of fallback)"
    cpu_sample = None
    if rank == 0 and not args.no_cpu:      # host copy of the CPU leg's slab (config_rows releases the device arrays)
        threads = os.cpu_count() or 1
        m = min(args.samples_cpu or pick_cpu_sample(threads), n)
        cpu_sample = tuple(a.flat_slice(0, m).to_numpy() for a in (x, y, w))
    configs = None
    if not args.no_configs:
        configs = config_rows(rank, world, dev, comm, peak, x, y, w, n, D, core, DeviceArray, PinnedArray, _cabi, timed_steps, max_over_ranks)
    # ---- sustained: the same call back to back for about a second (single GPU).  The K timed steps above last ~50 ms; this
    #      kernel keeps the B200 above its 1000 W power limit (instantaneous draw 1.05-1.2 kW), so after ~25 back-to-back calls
    #      the driver lowers the SM clock (sw_power_cap, 1965 -> 1600-1760 MHz) and the kernel time settles ~13 % higher.
    sustained = None
    if comm is None and not args.no_configs:
        # (config_rows released the headline arrays to make room: the counter-based streams regenerate them; last leg of the run,
        #  so that the power state it leaves behind disturbs nothing else)
        x = DeviceArray.normal((n,), np.float32, seed=SEEDS[0], offset=off, device=dev)
        y = DeviceArray.normal((n,), np.float32, seed=SEEDS[1], offset=off, device=dev)
        w = DeviceArray.uniform((n,), np.float32, seed=SEEDS[2], offset=off, device=dev)
        for _ in range(3):
            core._bincount(x, y, w, weights=True, axis=None, bins=bins, _timing=timing, _density_widths=[np.diff(EDGES)] * 2)
        time.sleep(1.0)
        sus_ms = []
        s2 = ClockSampler(dev)
        s2.start()
        s2.wait_ready()
        t_end = time.perf_counter() + 1.0
        while time.perf_counter() < t_end:
            core._bincount(x, y, w, weights=True, axis=None, bins=bins, _timing=timing, _density_widths=[np.diff(EDGES)] * 2)
            sus_ms.append(timing["kernel_ms"])
        c2 = s2.stop()
        tail = sus_ms[len(sus_ms) // 2:]
        sustained = {"calls": len(sus_ms), "kernel_ms_first_10": float(np.mean(sus_ms[:10])), "kernel_ms_second_half": float(np.mean(tail)),
                     "frac_second_half": (n * 12 + NBINS * NBINS * 8) / (float(np.mean(tail)) * 1e-3) / 1e9 / peak,
                     "clocks": c2, "what": "about 1 s of back-to-back device-resident calls; kernel time by CUDA events per call"}


    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return 0

    # ---- roofline of the dominant kernel (k_hist) against the measured HBM peak
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
                "kernel": "k_hist<float, W=3 (fp32 weights, one-limb fixed point), K=2, MODE=1> (+ density: all kernels of the call)", "kernel_ms": k_ms, "algorithmic_bytes": alg_bytes,
                "peak_source": peak_src}

    cpu = None
    if cpu_sample is not None:
        threads = os.cpu_count() or 1
        xs, ys, ws = cpu_sample
        m = xs.size
        rate, _ = cpu_throughput(xs, ys, ws, threads)
        rate1, _ = cpu_throughput(xs[: m // 8], ys[: m // 8], ws[: m // 8], 1)
        cpu = {"value": rate, "unit": UNIT, "cores": threads, "kind": cpu_impl()[0],
               "sample": f"first {m} samples of the same device-generated workload, {threads} column slabs in {threads} threads + sum; {cpu_impl()[2]}",
               "single_core_value": rate1}

    # launches per step and rank inside the timed region: k_hist (the verdict of the probe is cached after the first call of a
    # buffer set: no probe, no idle sibling), the zero-fill of the partial (a memset node), [k_peer_allreduce], 2 density kernels
    launches_per_step = 3 + (1 if world > 1 else 0)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": workload_name(n), "samples_per_gpu": n,
                   "sharding": "sample axis, partial histograms summed through peer memory over NVLink (k_peer_allreduce) inside the step" if world > 1 else "single GPU",
                   "l2": "inputs (12 B/sample, 12 GB per GPU at 1e9) exceed the 126 MB L2; no flush needed between steps",
                   "accumulate": "float64 (np.bincount semantics), fp32 compare on round-up edges", "host_numa": numa},
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": world * n * 12, "d2h_bytes_per_step": world * NBINS * NBINS * 8,
                "steps": e2e_steps, "matches_device_result": e2e_ok, "h2d_gb_per_s_slowest_gpu": h2d_gbs_min},
        "gpu_launches": args.steps * launches_per_step * world,
        "roofline": roofline,
        "cpu_baseline": cpu,
        "clocks": clocks,
        "parity": parity,
        "hbm_gbs_whole_step": alg_bytes / (ms_per_step * 1e-3) / 1e9,
    }
    if strong is not None:
        line["strong"] = strong
    if sustained is not None:
        line["sustained"] = sustained
    if configs is not None:
        line["configs"] = configs
    emit(line)
    if dist is not None:
        dist.destroy_process_group()
    return 0


def multi_rank_parity(rank, world, dev, n, x, y, w, comm, bins, D, core, O, DeviceArray, _cabi, host_barrier):
    """All-reduced results against the oracle on the concatenated shards, bin by bin (reference role: dask's blockwise +
    sum over chunks, core.py:418-439; spec: xhistogram/test/test_chunking.py:33-101)."""
    out = {}
    ok = True
    m = min(n, 1 << 22)
    xs, ys, ws = x.flat_slice(0, m), y.flat_slice(0, m), w.flat_slice(0, m)
    hc, _ = D.histogram(xs, ys, bins=bins, comm=comm, sharded_axis=0)                                   # counts
    hd, _ = D.histogram(xs, ys, bins=bins, weights=ws, density=True, comm=comm, sharded_axis=0)         # weighted density
    hh, _ = D.histogram(xs.to_numpy(), ys.to_numpy(), bins=bins, weights=ws.to_numpy(), comm=comm, sharded_axis=0)   # host shards
    if rank == 0:
        # every rank's first m samples, regenerated from the counter-based streams (element i depends on (seed, offset + i))
        parts = []
        for r in range(world):
            px = DeviceArray.normal((m,), np.float32, seed=SEEDS[0], offset=r * n, device=dev)
            py = DeviceArray.normal((m,), np.float32, seed=SEEDS[1], offset=r * n, device=dev)
            pw = DeviceArray.uniform((m,), np.float32, seed=SEEDS[2], offset=r * n, device=dev)
            parts.append((px.to_numpy(), py.to_numpy(), pw.to_numpy()))
            for q in (px, py, pw):
                q.free()
        cx, cy, cw = (np.concatenate([p[i] for p in parts]) for i in range(3))
        wc, _ = O.histogram(cx, cy, bins=bins, threads=8)
        wd, _ = O.histogram(cx, cy, bins=bins, weights=cw, density=True, threads=8)
        ww, _ = O.histogram(cx, cy, bins=bins, weights=cw, threads=8)
        out["allreduce_counts_bit_exact_vs_oracle"] = bool(hc.dtype == np.int64 and np.array_equal(hc, wc))
        out["allreduce_weighted_density_max_rel_err"] = float(np.max(np.abs(hd - wd)) / np.max(np.abs(wd)))
        out["allreduce_host_shards_weighted_max_rel_err"] = float(np.max(np.abs(hh - ww)) / np.max(np.abs(ww)))
        out["allreduce_samples"] = int(world * m)
        ok = ok and out["allreduce_counts_bit_exact_vs_oracle"] and out["allreduce_weighted_density_max_rel_err"] <= 1e-6 \
            and out["allreduce_host_shards_weighted_max_rel_err"] <= 1e-6
    # kept axis sharded (config-4 shape in small: (time, lat, lon), reduce lat/lon, time split over the ranks), gathered
    T, La, Lo = 4 * world + 1, 48, 96
    r = np.random.default_rng(40)
    fa = r.standard_normal((T, La, Lo)).astype(np.float32)
    fb = r.standard_normal((T, La, Lo)).astype(np.float32)
    e4 = [np.linspace(-4, 4, 101)] * 2
    t0, t1 = D.shard_bounds(T, world, rank)
    tc = D.TorchCommunicator(device=f"cuda:{dev}")
    hk, _ = D.histogram(fa[t0:t1], fb[t0:t1], bins=e4, axis=(1, 2), comm=tc, sharded_axis=0, gather=True)
    hk2, _ = D.histogram(np.ascontiguousarray(fa[:, :, rank::world]), np.ascontiguousarray(fb[:, :, rank::world]), bins=e4, axis=(1, 2),
                         comm=comm, sharded_axis=2)         # the same result with a REDUCED axis sharded (lon), rows kept: fused all-reduce of (T, 100, 100)
    if rank == 0:
        wk, _ = O.histogram(fa, fb, bins=e4, axis=(1, 2))
        out["kept_axis_sharded_gather_bit_exact"] = bool(np.array_equal(hk, wk))
        out["reduced_axis_sharded_rows_kept_bit_exact"] = bool(np.array_equal(hk2, wk))
        ok = ok and out["kept_axis_sharded_gather_bit_exact"] and out["reduced_axis_sharded_rows_kept_bit_exact"]
    # single-process multi-device entry point (xh_hist_multi): rank 0 drives devices 0..1 while the others wait on the host
    host_barrier()
    if rank == 0 and _cabi.device_count() >= 2:
        mm = 3_000_001
        a, b, c = host_slab(mm, seeds=(21, 22, 23))
        hm, _ = core.histogram(a, b, bins=bins, weights=c, devices=[0, 1])                              # columns sharded + NCCL
        wm, _ = O.histogram(a, b, bins=bins, weights=c, threads=8)
        a7, b7 = a[:2_800_000].reshape(7, -1), b[:2_800_000].reshape(7, -1)
        hr, _ = core.histogram(a7, b7, bins=bins, axis=1, devices=[0, 1])                               # rows sharded
        wr, _ = O.histogram(a7, b7, bins=bins, axis=1)
        out["xh_hist_multi_columns_max_rel_err"] = float(np.max(np.abs(hm - wm)) / np.max(np.abs(wm)))
        out["xh_hist_multi_rows_bit_exact"] = bool(np.array_equal(hr, wr))
        ok = ok and out["xh_hist_multi_columns_max_rel_err"] <= 1e-6 and out["xh_hist_multi_rows_bit_exact"]
    host_barrier()
    out["_ok"] = ok
    return out


def config_rows(rank, world, dev, comm, peak, x, y, w, n, D, core, DeviceArray, PinnedArray, _cabi, timed_steps, max_over_ranks):
    """Kernel-time roofline rows of BASELINE.json's other configurations (algorithmic bytes as in SURVEY.md §8d: inputs read
    once + histogram written once) and of adverse data for config 3.  N = 1: whole configurations on one GPU; N > 1: the
    per-rank shard of the N-GPU case, config 5 with its reduction inside the timed step, config 4 with rows (time) sharded
    and no collective."""
    rows = []
    lib = _cabi.lib()

    def kernel_row(name, arrays, wts, bins, axis, nbytes, samples, note="", reps=3):
        t = {}
        every = list(arrays) + ([wts] if wts is not None else [])
        ms = []
        for _ in range(reps + 1):
            core._bincount(*every, weights=wts is not None, axis=axis, bins=bins, _timing=t)
            ms.append(t["kernel_ms"])
        best = max_over_ranks(float(np.min(ms[1:])))
        gbs = nbytes / (best * 1e-3) / 1e9
        rows.append({"config": name, "samples_per_gpu": samples, "algorithmic_bytes_per_gpu": nbytes, "kernel_ms": best,
                     "gb_per_s": gbs, "frac": gbs / peak, "timed": f"CUDA events around the kernels of one call, best of {reps}, max over ranks", "note": note})

    # config 3 on adverse data (single GPU rows; the data of the headline run is the favourable case)
    e = np.linspace(-4.0, 4.0, NBINS + 1)
    if world == 1:
        m = 1 << 24
        ws_host = (w.flat_slice(0, m).to_numpy() * np.float32(np.pi)).astype(np.float32)      # not multiples of 2^-24
        wg = DeviceArray((n,), np.float32, device=dev)
        for i in range(0, n, m):
            c = min(m, n - i)
            _cabi.check(lib.xh_memcpy(dev, wg.ptr + i * 4, ws_host.ctypes.data, c * 4, _cabi.XH_DEVICE, _cabi.XH_HOST), "h2d")
        kernel_row("cfg3, generic fp32 weights (w * pi: two-limb fixed point / float64 path)", [x, y], wg, [e, e], None, n * 12 + NBINS * NBINS * 8, n)
        wg.free()
        ux = DeviceArray.uniform((n,), np.float32, seed=13, device=dev)
        uy = DeviceArray.uniform((n,), np.float32, seed=14, device=dev)
        eu = np.linspace(0.0, 1.0, NBINS + 1)
        kernel_row("cfg3, x and y uniform over the whole bin range (no window can hold the mass)", [ux, uy], w, [eu, eu], None, n * 12 + NBINS * NBINS * 8, n)
        kernel_row("cfg3 counts, x and y uniform over the whole bin range", [ux, uy], None, [eu, eu], None, n * 8 + NBINS * NBINS * 8, n)
        ux.free(); uy.free()
        kernel_row("cfg3 counts (no weights)", [x, y], None, [e, e], None, n * 8 + NBINS * NBINS * 8, n)
        # sorted x through the HOST pipeline: every staged chunk has its mass elsewhere (the window is chosen per chunk)
        ms_ = min(n, 1 << 26)
        hxs = np.sort(x.flat_slice(0, ms_).to_numpy())
        hys, hws = y.flat_slice(0, ms_).to_numpy(), w.flat_slice(0, ms_).to_numpy()
        px, py, pw = (PinnedArray((ms_,), np.float32) for _ in range(3))
        px.array[:], py.array[:], pw.array[:] = hxs, hys, hws
        t0 = time.perf_counter()
        for _ in range(2):
            core.histogram(px.array, py.array, bins=[e, e], weights=pw.array, density=True)
        dt = (time.perf_counter() - t0) / 2
        rows.append({"config": "cfg3, x sorted, pinned host inputs through the chunked H2D pipeline (window re-chosen per staged chunk)",
                     "samples_per_gpu": ms_, "wall_ms": dt * 1e3, "h2d_gb_per_s": ms_ * 12 / dt / 1e9,
                     "note": "PCIe-bound: compare with e2e.value * 12 B of the unsorted run"})
        for a in (px, py, pw):
            a.free()
        # several weightings of the same samples (the weighted mean of the reference's tutorial: hist(w * a) / hist(w)): ONE pass with
        # weights=[w1, w2] against two separate calls; 100 x 100 bins (the two float64 planes fit shared memory); wall time of the
        # synchronous calls, device-resident inputs
        e1 = np.linspace(-4.0, 4.0, 101)
        w2 = DeviceArray.uniform((n,), np.float32, seed=15, device=dev)

        def wall(fn, reps=3):
            fn()
            ts = []
            for _ in range(reps):
                t0 = time.perf_counter(); fn(); ts.append((time.perf_counter() - t0) * 1e3)
            return float(np.min(ts))
        t_list = wall(lambda: core.histogram(x, y, bins=[e1, e1], weights=[w, w2]))       # the library's choice on device data
        with core.debug_flags(_cabi.XH_FLAG_ONE_PASS):
            t_one = wall(lambda: core.histogram(x, y, bins=[e1, e1], weights=[w, w2]))    # k_hist_mw forced
            h1f, _ = core.histogram(x, y, bins=[e1, e1], weights=[w, w2])
        t_two = wall(lambda: (core.histogram(x, y, bins=[e1, e1], weights=w), core.histogram(x, y, bins=[e1, e1], weights=w2)))
        h1, _ = core.histogram(x, y, bins=[e1, e1], weights=[w, w2])
        ha, _ = core.histogram(x, y, bins=[e1, e1], weights=w)
        hb, _ = core.histogram(x, y, bins=[e1, e1], weights=w2)
        # the same from pinned host memory (what a numpy user has): the samples cross PCIe once instead of twice
        mh = min(n, 1 << 27)
        ph = [PinnedArray((mh,), np.float32) for _ in range(4)]
        for src, dst in zip((x, y, w, w2), ph):
            _cabi.check(lib.xh_memcpy(dev, dst.ptr, src.ptr, mh * 4, _cabi.XH_HOST, _cabi.XH_DEVICE), "d2h")
        hx_, hy_, hw1_, hw2_ = (q.array for q in ph)
        th_one = wall(lambda: core.histogram(hx_, hy_, bins=[e1, e1], weights=[hw1_, hw2_]), reps=2)
        th_two = wall(lambda: (core.histogram(hx_, hy_, bins=[e1, e1], weights=hw1_), core.histogram(hx_, hy_, bins=[e1, e1], weights=hw2_)), reps=2)
        for q in ph:
            q.free()
        def rel(h):
            return float(max(np.max(np.abs(h[0] - ha)) / np.max(np.abs(ha)), np.max(np.abs(h[1] - hb)) / np.max(np.abs(hb))))
        rows.append({"config": "cfg3 shape with TWO weight arrays, 100x100 bins: weights=[w1, w2] in one call vs two calls",
                     "host_inputs": {"samples": mh, "one_pass_ms": th_one, "two_calls_ms": th_two, "speedup": th_two / th_one,
                                     "note": "pinned host arrays, one-pass kernel: 16 B/sample over PCIe instead of 24"},
                     "samples_per_gpu": n, "list_call_ms": t_list, "one_pass_kernel_ms": t_one, "two_calls_ms": t_two,
                     "algorithmic_bytes_per_sample": 16, "bytes_read_per_sample_list_call": 24,
                     "gb_per_s": (n * 16) / (t_list * 1e-3) / 1e9, "frac": (n * 16) / (t_list * 1e-3) / 1e9 / peak,
                     "max_rel_diff_vs_separate_calls": rel(h1), "max_rel_diff_one_pass_kernel": rel(h1f),
                     "note": "device-resident inputs: the list call runs one fused exact fixed-point pass per weight array (24 B/sample) because the "
                             "one-pass kernel (16 B/sample, XH_FLAG_ONE_PASS) is bound by its float64 shared adds, not by the reads it saves; "
                             "host inputs take the one-pass kernel (PCIe-bound)"})
        w2.free()
    x.free(); y.free(); w.free()

    # config 2: two fp32 (1e4, 1e5), 128 x 128, axis=-1 (one GPU)
    if world == 1:
        M, N = 10_000, 100_000
        a = DeviceArray.normal((M, N), np.float32, seed=1, device=dev); b = DeviceArray.normal((M, N), np.float32, seed=2, device=dev)
        e2 = np.linspace(-4, 4, 129)
        kernel_row("cfg2: 2 x fp32 (1e4,1e5), 128x128, axis=-1", [a, b], None, [e2, e2], [1], M * N * 8 + M * 128 * 128 * 8, M * N)
        a.free(); b.free()

    # config 4: (8192, 720, 1440) x 2, 100 x 100, dim=(lat, lon); time sharded over 8 GPUs: 1024 time steps per GPU, no collective
    M, N = 1024, 720 * 1440
    a = DeviceArray.normal((M, 720, 1440), np.float32, seed=6, offset=rank * M * N, device=dev)
    b = DeviceArray.normal((M, 720, 1440), np.float32, seed=7, offset=rank * M * N, device=dev)
    e4 = np.linspace(-4, 4, 101)
    kernel_row("cfg4 shard: 2 x fp32 (1024,720,1440) per GPU, 100x100, axis=(1,2) (time sharded: disjoint output slabs, no collective)",
               [a, b], None, [e4, e4], [1, 2], M * N * 8 + M * 100 * 100 * 8, M * N, note="1/8 of the 8-GPU case on every rank")
    a.free(); b.free()

    # config 5: three fp64 (4e8,), non-uniform (50,60,70) bins, fp64 weights; N > 1: 4e8 / N per rank + reduction inside the step
    r = np.random.default_rng(12)
    e5 = []
    for k in (51, 61, 71):
        ee = np.sort(r.uniform(-4, 4, k)); ee[0], ee[-1] = -4.0, 4.0
        e5.append(ee)
    n5 = int(4e8) // world
    xs = [DeviceArray.normal((n5,), np.float64, seed=8 + i, offset=rank * n5, device=dev) for i in range(3)]
    w5 = DeviceArray.uniform((n5,), np.float64, seed=11, offset=rank * n5, device=dev)
    nb5 = 50 * 60 * 70
    if world == 1:
        kernel_row("cfg5: 3 x fp64 (4e8,), fp64 weights, non-uniform (50,60,70)", xs, w5, e5, None, n5 * 32 + nb5 * 8, n5)
        n8 = n5 // 8
        kernel_row("cfg5 1/8 shard: 3 x fp64 (5e7,), fp64 weights, non-uniform (50,60,70)", [q.flat_slice(0, n8) for q in xs], w5.flat_slice(0, n8), e5, None,
                   n8 * 32 + nb5 * 8, n8)
    else:
        def step5():
            return D.histogram(*xs, bins=e5, weights=w5, comm=comm, sharded_axis=0)[0]
        ms5, _ = timed_steps(step5, 10, 3)
        nbytes = n5 * 32 + nb5 * 8
        rows.append({"config": f"cfg5: 3 x fp64 (4e8,) over {world} GPUs ({n5} per rank), fp64 weights, non-uniform (50,60,70), partials reduced inside the step",
                     "samples_per_gpu": n5, "algorithmic_bytes_per_gpu": nbytes, "ms_per_step": ms5, "value_samples_per_s": 4e8 / (ms5 * 1e-3),
                     "gb_per_s_per_gpu": nbytes / (ms5 * 1e-3) / 1e9, "frac": nbytes / (ms5 * 1e-3) / 1e9 / peak,
                     "timed": "whole synchronous step (kernels + reduction + D2H), 10 steps after 3 warm-ups, max over ranks"})
    for q in xs + [w5]:
        q.free()
    return rows


if __name__ == "__main__":
    sys.exit(main())
