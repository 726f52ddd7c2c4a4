"""Debug: config-5-shaped step under torchrun, with and without the collective; prints wall per step and the library's phase clock."""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, torch.distributed as dist
from xhistogram_b200 import DeviceArray, _cabi, core, distributed as D

rank, world, dev = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=torch.device("cuda", dev))
core.set_default_device(dev)
comm = D.NcclCommunicator.from_env(dev)
lib = _cabi.lib()
r = np.random.default_rng(12)
e5 = []
for k in (51, 61, 71):
    ee = np.sort(r.uniform(-4, 4, k)); ee[0], ee[-1] = -4.0, 4.0
    e5.append(ee)
n5 = int(float(sys.argv[1])) if len(sys.argv) > 1 else 50_000_000
xs = [DeviceArray.normal((n5,), np.float64, seed=8 + i, offset=rank * n5, device=dev) for i in range(3)]
w5 = DeviceArray.uniform((n5,), np.float64, seed=11, offset=rank * n5, device=dev)
ph = (_cabi.C.c_double * 4)()
TAG = "nccl" if os.environ.get("XH_NO_P2P") else "p2p"
os.makedirs("gpurun_out", exist_ok=True)
sys.stdout = open(f"gpurun_out/r2k_{TAG}_rank{rank}.log", "w")

def timed(fn, name, reps=8):
    for _ in range(3):
        fn()
    dist.barrier(); lib.xh_sync(dev)
    rows = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); t = (time.perf_counter() - t0) * 1e3
        lib.xh_last_call_phases(ph); rows.append([t] + [p / 1e3 for p in ph])
    m = np.median(np.array(rows), axis=0)
    print(f"{TAG} rank {rank} {name}: wall {m[0]:.3f} ms  lib: tables {m[1]:.3f} enqueued {m[2]:.3f} synced {m[3]:.3f} return {m[4]:.3f}", flush=True)

timed(lambda: core.histogram(*xs, bins=e5, weights=w5), "local only")
timed(lambda: D.histogram(*xs, bins=e5, weights=w5, comm=comm, sharded_axis=0), "with reduction")
e = np.linspace(-4, 4, 257)
x = DeviceArray.normal((n5,), np.float32, seed=3, offset=rank * n5, device=dev); y = DeviceArray.normal((n5,), np.float32, seed=4, offset=rank * n5, device=dev)
w = DeviceArray.uniform((n5,), np.float32, seed=5, offset=rank * n5, device=dev)
timed(lambda: core.histogram(x, y, bins=[e, e], weights=w), "cfg3-shape local only")
timed(lambda: D.histogram(x, y, bins=[e, e], weights=w, comm=comm, sharded_axis=0), "cfg3-shape with reduction")
dist.barrier()
dist.destroy_process_group()
