"""One-rank-per-GPU histograms: shard, histogram locally on the GPU, combine the partials.

This replaces the role dask plays in the reference — ``blockwise(_bincount)`` over chunks followed
by ``.sum`` over the chunked axes (xhistogram/core.py:403-439):

* the sharded axis is REDUCED (flat 1e9-sample histograms): every rank builds a partial
  histogram of its slab and the partials are summed with one all-reduce (NCCL inside
  ``libxhist_b200.so``: ``xh_comm_allreduce``; int64 counts stay bit-exact, float64 sums differ
  only by add order);
* the sharded axis is KEPT (e.g. ``time`` in the (time, lat, lon) case): rows are independent,
  every rank owns a disjoint slab of the output and no reduction is needed (an optional
  all-gather assembles the full array).

``Communicator`` is the small interface the algorithm needs.  ``NcclCommunicator`` is the product
implementation (library-owned NCCL communicator, bootstrapped by broadcasting the 128-byte
unique id through any out-of-band channel, e.g. ``torch.distributed``);
``TorchCommunicator`` runs the same algorithm over ``torch.distributed`` tensors and is what the
world-size-2 ``gloo`` CPU tests use.
"""
from __future__ import annotations

import ctypes as C
import functools

import numpy as np

from . import _cabi
from . import core as _core

_range = range


def shard_bounds(n: int, world: int, rank: int):
    """Contiguous, balanced [start, stop) of ``n`` items for ``rank`` of ``world``."""
    return n * rank // world, n * (rank + 1) // world


class Communicator:
    rank = 0
    world = 1

    def allreduce_sum(self, a: np.ndarray) -> np.ndarray:  # int64 or float64
        return a

    def allreduce_minmax(self, mn: float, mx: float):
        return mn, mx

    def allgather_rows(self, a: np.ndarray, counts) -> np.ndarray:  # concatenate along axis 0
        return a


class TorchCommunicator(Communicator):
    """``torch.distributed`` plumbing (any backend).  CPU tensors for gloo, CUDA tensors for nccl."""

    def __init__(self, device=None):
        import torch.distributed as dist

        self._dist = dist
        self.rank, self.world = dist.get_rank(), dist.get_world_size()
        self._device = device

    def _t(self, a):
        import torch

        t = torch.from_numpy(np.ascontiguousarray(a))
        return t.to(self._device) if self._device is not None else t

    def allreduce_sum(self, a):
        t = self._t(a)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return t.cpu().numpy()

    def allreduce_minmax(self, mn, mx):
        t = self._t(np.array([-mn, mx], dtype=np.float64))  # NaN propagates through MAX on both backends
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        v = t.cpu().numpy()
        return -float(v[0]), float(v[1])

    def allgather_rows(self, a, counts):
        import torch

        t = self._t(a)
        outs = [torch.empty((int(c),) + tuple(a.shape[1:]), dtype=t.dtype, device=t.device) for c in counts]
        self._dist.all_gather(outs, t) if len(set(int(c) for c in counts)) == 1 else self._uneven(outs, t)
        return torch.cat(outs).cpu().numpy()

    def _uneven(self, outs, t):
        for r, o in enumerate(outs):
            if r == self.rank:
                o.copy_(t)
            self._dist.broadcast(o, src=r)


class NcclCommunicator(Communicator):
    """NCCL communicator owned by ``libxhist_b200.so`` (one rank per GPU, NVLink/NVSwitch)."""

    def __init__(self, device: int, rank: int, world: int, unique_id: bytes):
        self.device, self.rank, self.world = int(device), int(rank), int(world)
        buf = C.create_string_buffer(unique_id, _cabi.XH_NCCL_UNIQUE_ID_BYTES)
        _cabi.check(_cabi.lib().xh_comm_init_rank(self.device, buf, self.world, self.rank), "xh_comm_init_rank")

    @staticmethod
    def create_unique_id() -> bytes:
        buf = C.create_string_buffer(_cabi.XH_NCCL_UNIQUE_ID_BYTES)
        _cabi.check(_cabi.lib().xh_comm_unique_id(buf), "xh_comm_unique_id")
        return buf.raw

    @classmethod
    def from_torch_distributed(cls, device: int):
        """Bootstrap over an initialised ``torch.distributed`` group (used only to ship the id)."""
        import torch.distributed as dist

        rank, world = dist.get_rank(), dist.get_world_size()
        box = [cls.create_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(box, src=0)
        return cls(device, rank, world, box[0])

    @classmethod
    def from_env(cls, device: int = None, timeout: float = 120.0):
        """Bootstrap without any other communication library (one node): ranks read ``RANK`` / ``WORLD_SIZE`` /
        ``LOCAL_RANK`` (set by torchrun, mpirun wrappers, ...); rank 0 creates the NCCL id and publishes it through a file
        that the other ranks wait for.  The file name is unique per launch (``XHIST_NCCL_ID_FILE``, or the launcher's pid —
        the common parent of all ranks — and ``MASTER_PORT``)."""
        import os
        import tempfile
        import time

        rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        device = int(os.environ.get("LOCAL_RANK", "0")) if device is None else device
        if world == 1:
            return cls(device, 0, 1, cls.create_unique_id())
        path = os.environ.get("XHIST_NCCL_ID_FILE") or os.path.join(
            tempfile.gettempdir(), f"xhist_b200_nccl_{os.getppid()}_{os.environ.get('MASTER_PORT', '0')}.id")
        if rank == 0:
            uid = cls.create_unique_id()
            tmp = path + f".{os.getpid()}"
            with open(tmp, "wb") as f:
                f.write(uid)
            os.replace(tmp, path)                      # atomic: a reader sees nothing or all 128 bytes
        else:
            t0 = time.time()
            while True:
                try:
                    with open(path, "rb") as f:
                        uid = f.read()
                    if len(uid) == _cabi.XH_NCCL_UNIQUE_ID_BYTES:
                        break
                except FileNotFoundError:
                    pass
                if time.time() - t0 > timeout:
                    raise TimeoutError(f"rank {rank}: no NCCL id at {path} after {timeout} s")
                time.sleep(0.01)
        comm = cls(device, rank, world, uid)           # ncclCommInitRank returns once every rank has joined
        if rank == 0:
            try:
                os.unlink(path)
            except OSError:
                pass
        return comm

    def allreduce_device(self, ptr: int, count: int, is_f64: bool):
        _cabi.check(_cabi.lib().xh_comm_allreduce(self.device, ptr, count, 1 if is_f64 else 0), "xh_comm_allreduce")

    def allreduce_sum(self, a):
        from .device import DeviceArray

        is_f64 = a.dtype == np.float64
        d = DeviceArray.from_numpy(a.view(np.float64) if not is_f64 else a, device=self.device)
        self.allreduce_device(d.ptr, a.size, is_f64)
        out = d.to_numpy()
        d.free()
        return out if is_f64 else out.view(np.int64)

    def allreduce_minmax(self, mn, mx):
        # every rank deposits (min, max) in its own slot of a zero vector; the sum is an all-gather (NaN stays NaN)
        v = np.zeros(2 * self.world, dtype=np.float64)
        v[2 * self.rank], v[2 * self.rank + 1] = mn, mx
        v = self.allreduce_sum(v)
        if np.isnan(v).any():
            return float("nan"), float("nan")
        return float(v[0::2].min()), float(v[1::2].max())

    def close(self):
        _cabi.check(_cabi.lib().xh_comm_destroy(self.device), "xh_comm_destroy")


def histogram(*local_args, bins=None, range=None, axis=None, weights=None, density=False, block_size="auto",
              comm: Communicator = None, sharded_axis=0, gather=False, out=None):
    """SPMD histogram: every rank passes ITS shard of the arrays (split along ``sharded_axis``).

    Same arguments as ``core.histogram`` otherwise.  Returns ``(h, edges)``; when the sharded axis
    is reduced ``h`` is the global histogram on every rank, when it is kept ``h`` is this rank's
    slab (or the assembled array with ``gather=True``).  ``out`` (device-resident shards, reduced sharded axis, NCCL
    communicator): a ``DeviceArray`` that receives the global histogram — the call only enqueues (kernels, reduction
    over NVLink, density) and the result stays in HBM, valid in stream order.
    """
    comm = comm or Communicator()
    # repeat call on the same device-resident shards and edge arrays: reuse the filled descriptor (core._plan_*)
    plan_key = None
    if (isinstance(comm, NcclCommunicator) and out is None and range is None and _core._debug_flags == 0 and _core._timing_sink is None
            and isinstance(bins, (list, tuple)) and len(bins) == len(local_args) and all(type(b) is np.ndarray for b in bins)
            and all(type(a) is _core.DeviceArray for a in list(local_args) + ([weights] if weights is not None else []))):
        every = list(local_args) + ([weights] if weights is not None else [])
        plan_key = ("dist", comm.device, tuple(id(a) for a in every), weights is not None, tuple(id(b) for b in bins),
                    None if axis is None else tuple(np.atleast_1d(axis).tolist()), bool(density), int(sharded_axis))
        plan = _core._plan_lookup(plan_key, every, bins)
        if plan is not None:
            return _core._plan_run(plan), list(bins)
    a0 = local_args[0]
    nd = np.ndim(a0) if not _core.is_device_array(a0) else len(_core.as_device_view(a0)[1])
    sharded_axis = sharded_axis if sharded_axis >= 0 else nd + sharded_axis
    if axis is None:
        red = list(_range(nd))
    else:
        red = [int(a) if a >= 0 else nd + int(a) for a in np.atleast_1d(axis)]
    n = len(local_args)
    bins_l = _core._ensure_correctly_formatted_bins(bins, n)
    range_l = _core._ensure_correctly_formatted_range(range, n)
    edges = []
    for a, b, r in zip(local_args, bins_l, range_l):
        if isinstance(b, str):
            raise TypeError("distributed histograms need explicit bin edges or an integer bin count")
        if np.ndim(b) == 0 and r is None:
            mn, mx = comm.allreduce_minmax(*_core._minmax(a))      # global range, then numpy's own edge formula
            dt = _core.as_device_view(a)[2] if _core.is_device_array(a) else np.asarray(a).dtype
            edges.append(np.histogram_bin_edges(np.array([mn, mx], dtype=dt), bins=b))
        else:
            edges.append(_core._resolve_edges(a if _core.is_device_array(a) else np.asarray(a), b, r, None))
    if sharded_axis in red:
        if isinstance(comm, NcclCommunicator):
            # ONE native call: histogram kernels -> ncclAllReduce of the partials in HBM -> density -> D2H of the result
            h = _fused_allreduce(local_args, weights, edges, axis, nd, comm, density, out)
            if plan_key is not None and h.size and all(e is b for e, b in zip(edges, bins)):
                shape = _core.as_device_view(local_args[0])[1]
                full = axis is None or set(red) == set(_range(nd))
                mn = _core.M_N_of_rows(shape, nd, full, red)
                on_device = (not density) or all(np.asarray(e).dtype in (np.float32, np.float64) for e in edges)
                if mn is not None and on_device:
                    every = list(local_args) + ([weights] if weights is not None else [])
                    infos = [_core._edge_info(e) for e in edges]
                    _core._plan_store(plan_key, every, edges, infos, [_core.as_device_view(a) for a in every], len(local_args), mn[0], mn[1],
                                      h.shape, weights is not None, bool(density), _cabi.XH_FLAG_ALLREDUCE)
            return h, edges
        if out is not None:
            raise TypeError("out= needs an NcclCommunicator")
        h, _ = _core.histogram(*local_args, bins=edges, axis=axis, weights=weights, density=False, block_size=block_size)
        h = comm.allreduce_sum(np.ascontiguousarray(h))         # partial histograms -> global (core.py:439)
    else:
        if out is not None:
            raise TypeError("out= is for a reduced sharded axis")
        h, _ = _core.histogram(*local_args, bins=edges, axis=axis, weights=weights, density=False, block_size=block_size)
        if gather:
            kept = [i for i in _range(nd) if i not in red]
            pos = kept.index(sharded_axis)
            hm = np.ascontiguousarray(np.moveaxis(h, pos, 0))
            counts = comm.allreduce_sum(np.eye(comm.world, dtype=np.int64)[comm.rank] * hm.shape[0])
            h = np.moveaxis(comm.allgather_rows(hm, counts), 0, pos)
    if density:                                                     # after the reduce, on O(bins) data (core.py:444-462)
        areas = functools.reduce(np.multiply.outer, [np.diff(e) for e in edges])
        bin_axes = tuple(_range(-n, 0))
        h = h / areas / h.sum(axis=bin_axes, keepdims=True)
    return h, edges


def _fused_allreduce(local_args, weights, edges, axis, nd, comm, density, out=None):
    """Shard (host or device resident) -> partial histogram in HBM -> ncclAllReduce in place -> density on the device
    -> one D2H of the finished result, all inside one ``xh_hist`` call (XH_FLAG_ALLREDUCE [| XH_FLAG_DENSITY])."""
    ax = None if axis is None else [int(a) if a >= 0 else nd + int(a) for a in np.atleast_1d(axis)]
    arrays = list(local_args) + ([weights] if weights is not None else [])
    if not _core.is_device_array(local_args[0]):
        arrays = list(np.broadcast_arrays(*[np.asarray(a) for a in arrays]))
    on_device = density and all(np.asarray(e).dtype in (np.float32, np.float64) for e in edges)
    widths = [np.diff(e) for e in edges] if on_device else None
    if out is not None:
        if not _core.is_device_array(local_args[0]) or (density and not on_device):
            raise TypeError("out= needs device-resident shards (and float bin edges for density=True)")
        _core._bincount(*arrays, weights=weights is not None, axis=ax, bins=edges, _flags=_cabi.XH_FLAG_ALLREDUCE | _cabi.XH_FLAG_ASYNC,
                        _density_widths=widths, _out_device=out.reshape(-1))
        return out
    h = _core._bincount(*arrays, weights=weights is not None, axis=ax, bins=edges, _flags=_cabi.XH_FLAG_ALLREDUCE,
                        _density_widths=widths, _devices=[comm.device])
    if ax is not None:
        h = h.squeeze(tuple(ax))
    else:
        h = h.reshape(h.shape[len(h.shape) - len(edges):])
    if density and not on_device:
        areas = functools.reduce(np.multiply.outer, [np.diff(e) for e in edges])
        h = h / areas / h.sum(axis=tuple(_range(-len(edges), 0)), keepdims=True)
    return h
