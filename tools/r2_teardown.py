#!/usr/bin/env python
"""2+ ranks: a few all-reduced histograms through the peer-memory kernel, then NcclCommunicator.close() on every rank
(xh_comm_destroy: barrier, release of the IPC-mapped buffers, ncclCommDestroy), with rank 1 arriving late on purpose.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29531 tools/r2_teardown.py
"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xhistogram_b200 import DeviceArray, core, distributed as D  # noqa: E402

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = int(os.environ["LOCAL_RANK"])
core.set_default_device(dev)
comm = D.NcclCommunicator.from_env(dev)
n = 4_000_001
e = [np.linspace(-4, 4, 257)] * 2
x = DeviceArray.normal((n,), np.float32, seed=1, offset=rank * n, device=dev)
y = DeviceArray.normal((n,), np.float32, seed=2, offset=rank * n, device=dev)
w = DeviceArray.uniform((n,), np.float32, seed=3, offset=rank * n, device=dev)
for i in range(5):
    h, _ = D.histogram(x, y, bins=e, weights=w, comm=comm, sharded_axis=0)
    c, _ = D.histogram(x, y, bins=e, comm=comm, sharded_axis=0)
assert int(c.sum()) <= world * n and int(c.sum()) > 0.99 * world * n, c.sum()
if rank == 1:
    time.sleep(1.5)            # the others must wait in close(), not free memory this rank's last kernel could still read
t0 = time.perf_counter()
comm.close()
print(f"rank {rank}: closed after {time.perf_counter() - t0:.3f} s, counts {int(c.sum())}, weighted {float(h.sum()):.6g}", flush=True)
