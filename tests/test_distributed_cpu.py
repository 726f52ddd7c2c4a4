"""World-size-2 gloo test of the N>1 path's host logic (sharding + reduce / gather) on CPU.

Each rank's local GPU call is replaced by the oracle block kernel (as in test_frontend_cpu.py);
the partial-histogram reduction and the kept-axis gather run over a real torch.distributed gloo
group on 127.0.0.1."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    import torch.distributed as dist
    from oracle import hist_oracle as O
    from tests.test_frontend_cpu import _oracle_desc_call
    from xhistogram_b200 import core, distributed as D

    core._desc_call = _oracle_desc_call
    core._minmax = lambda a: (float(np.min(a)), float(np.max(a)))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        comm = D.TorchCommunicator()
        r = np.random.default_rng(77)
        x = r.standard_normal((6, 4001)).astype(np.float32)
        y = r.standard_normal((6, 4001)).astype(np.float32)
        w = r.random((6, 4001))
        e = [np.linspace(-4, 4, 33), np.linspace(-3, 3, 17)]
        ok = True
        # (a) sharded axis is reduced: columns split, partial histograms all-reduced; counts bit-exact
        c0, c1 = D.shard_bounds(4001, world, rank)
        h, _ = D.histogram(x[:, c0:c1], y[:, c0:c1], bins=e, axis=1, comm=comm, sharded_axis=1)
        ok &= np.array_equal(h, O.histogram(x, y, bins=e, axis=1)[0])
        # weighted + density, flat: reduce then normalise
        h, _ = D.histogram(x[:, c0:c1], y[:, c0:c1], bins=e, weights=w[:, c0:c1], density=True, comm=comm, sharded_axis=1)
        want = O.histogram(x, y, bins=e, weights=w, density=True)[0]
        ok &= bool(np.max(np.abs(h - want)) <= 1e-9 * np.max(np.abs(want)))
        # integer bins: global min/max across ranks before numpy's edge formula
        h, ed = D.histogram(x[:, c0:c1], bins=20, comm=comm, sharded_axis=1)
        hw, edw = np.histogram(x, bins=20)
        ok &= np.array_equal(ed[0], edw) and np.array_equal(h, hw)
        # (b) sharded axis is kept: rows split, no reduction, optional gather (uneven: 6 rows -> 3+3, 5 rows -> 2+3)
        for rows in (6, 5):
            r0, r1 = D.shard_bounds(rows, world, rank)
            h, _ = D.histogram(x[r0:r1], y[r0:r1], bins=e, axis=1, comm=comm, sharded_axis=0, gather=True)
            ok &= np.array_equal(h, O.histogram(x[:rows], y[:rows], bins=e, axis=1)[0])
            hl, _ = D.histogram(x[r0:r1], y[r0:r1], bins=e, axis=1, comm=comm, sharded_axis=0)
            ok &= np.array_equal(hl, O.histogram(x[r0:r1], y[r0:r1], bins=e, axis=1)[0])
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(180)
def test_two_rank_gloo_shard_reduce_gather():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    results = [q.get(timeout=150) for _ in procs]
    [p.join(30) for p in procs]
    assert sorted(results) == [(0, True), (1, True)]


def test_shard_bounds_cover_everything():
    from xhistogram_b200.distributed import shard_bounds
    for n in (0, 1, 7, 1000):
        for world in (1, 2, 3, 8):
            b = [shard_bounds(n, world, r) for r in range(world)]
            assert b[0][0] == 0 and b[-1][1] == n
            assert all(b[i][1] == b[i + 1][0] for i in range(world - 1))


def test_file_bootstrap_of_the_nccl_id(tmp_path, monkeypatch):
    """NcclCommunicator.from_env: rank 0 publishes the id through a file, the other ranks read exactly those 128 bytes
    (the NCCL calls themselves are stubbed: no GPU here)."""
    from xhistogram_b200 import distributed as D

    got = {}

    def fake_init(self, device, rank, world, unique_id):
        self.device, self.rank, self.world = device, rank, world
        got[rank] = bytes(unique_id)

    uid = bytes(range(128))
    monkeypatch.setattr(D.NcclCommunicator, "__init__", fake_init)
    monkeypatch.setattr(D.NcclCommunicator, "create_unique_id", staticmethod(lambda: uid))
    path = str(tmp_path / "id")
    monkeypatch.setenv("XHIST_NCCL_ID_FILE", path)
    monkeypatch.setenv("WORLD_SIZE", "3")

    def from_env_as(rank, timeout=20.0):
        monkeypatch.setenv("RANK", str(rank))
        monkeypatch.setenv("LOCAL_RANK", str(rank))
        return D.NcclCommunicator.from_env(timeout=timeout)

    c0 = from_env_as(0)                             # publishes; with __init__ stubbed it returns (and removes the file) at once
    assert c0.rank == 0 and got[0] == uid and not os.path.exists(path)
    with open(path, "wb") as f:                     # what a still-joining rank 0 leaves for the others
        f.write(uid)
    c1 = from_env_as(1)
    assert c1.rank == 1 and c1.world == 3 and c1.device == 1 and got[1] == uid
    with open(path, "wb") as f:                     # a truncated file is not accepted
        f.write(uid[:50])
    with pytest.raises(TimeoutError):
        from_env_as(2, timeout=0.3)
