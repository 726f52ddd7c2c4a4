"""Device-resident arrays for the device path of ``xhistogram_b200.core.histogram``.

No CuPy / PyTorch: buffers come from ``xh_malloc`` in the native library.  ``DeviceArray``
exposes ``__cuda_array_interface__`` so other CUDA libraries can view it, and the histogram
front-end accepts any object that exposes that interface (C-contiguous fp32/fp64).
"""
from __future__ import annotations

import ctypes as C
import math

import numpy as np

from . import _cabi

_DT = {np.dtype(np.float32): _cabi.XH_F32, np.dtype(np.float64): _cabi.XH_F64}


def xh_dtype(dtype) -> int:
    try:
        return _DT[np.dtype(dtype)]
    except KeyError:
        raise TypeError(f"device arrays must be float32 or float64, got {np.dtype(dtype)}") from None


class DeviceArray:
    """A C-contiguous fp32/fp64 array in the memory of one GPU."""

    def __init__(self, shape, dtype=np.float32, device=0, _ptr=None, _owner=None):
        self.shape = tuple(int(s) for s in (shape if np.iterable(shape) else (shape,)))
        self.dtype = np.dtype(dtype)
        xh_dtype(self.dtype)
        self.device = int(device)
        self.size = int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1
        self.nbytes = self.size * self.dtype.itemsize
        self._owner = _owner
        if _ptr is None:
            p = C.c_void_p()
            _cabi.check(_cabi.lib().xh_malloc(self.device, self.nbytes, C.byref(p)), "xh_malloc")
            self.ptr = p.value
            self._owned = True
        else:
            self.ptr = int(_ptr)
            self._owned = False

    # -- construction ----------------------------------------------------------------------
    @classmethod
    def from_numpy(cls, a, device=0):
        a = np.ascontiguousarray(a)
        out = cls(a.shape, a.dtype, device)
        _cabi.check(_cabi.lib().xh_memcpy(device, out.ptr, a.ctypes.data, a.nbytes, _cabi.XH_DEVICE, _cabi.XH_HOST), "xh_memcpy")
        return out

    @classmethod
    def normal(cls, shape, dtype=np.float32, seed=0, offset=0, device=0):
        out = cls(shape, dtype, device)
        _cabi.check(_cabi.lib().xh_fill_normal(device, out.ptr, xh_dtype(dtype), out.size, seed, offset), "xh_fill_normal")
        return out

    @classmethod
    def uniform(cls, shape, dtype=np.float32, seed=0, offset=0, device=0):
        out = cls(shape, dtype, device)
        _cabi.check(_cabi.lib().xh_fill_uniform(device, out.ptr, xh_dtype(dtype), out.size, seed, offset), "xh_fill_uniform")
        return out

    # -- views / transfers -------------------------------------------------------------------
    @property
    def ndim(self):
        return len(self.shape)

    def reshape(self, *shape):
        shape = shape[0] if len(shape) == 1 and np.iterable(shape[0]) else shape
        shape = tuple(int(s) for s in shape)
        if -1 in shape:
            known = int(np.prod([s for s in shape if s != -1], dtype=np.int64))
            shape = tuple(self.size // known if s == -1 else s for s in shape)
        if int(np.prod(shape, dtype=np.int64)) != self.size:
            raise ValueError(f"cannot reshape {self.shape} into {shape}")
        return DeviceArray(shape, self.dtype, self.device, _ptr=self.ptr, _owner=self)

    def flat_slice(self, start, stop):
        """View of elements [start, stop) of the flattened array."""
        start, stop = int(start), int(stop)
        if not (0 <= start <= stop <= self.size):
            raise IndexError("slice out of range")
        return DeviceArray((stop - start,), self.dtype, self.device, _ptr=self.ptr + start * self.dtype.itemsize, _owner=self)

    def to_numpy(self):
        out = np.empty(self.shape, self.dtype)
        if self.nbytes:
            _cabi.check(_cabi.lib().xh_memcpy(self.device, out.ctypes.data, self.ptr, self.nbytes, _cabi.XH_HOST, _cabi.XH_DEVICE), "xh_memcpy")
        return out

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, False), "version": 3, "strides": None}

    def free(self):
        if self._owned and self.ptr:
            _cabi.lib().xh_free(self.device, self.ptr)
            self.ptr = 0
            self._owned = False

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass

    def __repr__(self):
        return f"DeviceArray(shape={self.shape}, dtype={self.dtype}, device={self.device})"


class PinnedArray:
    """Page-locked host memory wrapped as a numpy array (``.array``) for fast host->device copies."""

    def __init__(self, shape, dtype=np.float32):
        self.dtype = np.dtype(dtype)
        self.shape = tuple(int(s) for s in (shape if np.iterable(shape) else (shape,)))
        n = int(np.prod(self.shape, dtype=np.int64))
        self.nbytes = n * self.dtype.itemsize
        p = C.c_void_p()
        _cabi.check(_cabi.lib().xh_host_alloc(max(self.nbytes, 1), C.byref(p)), "xh_host_alloc")
        self.ptr = p.value
        buf = (C.c_char * max(self.nbytes, 1)).from_address(self.ptr)
        self.array = np.frombuffer(buf, dtype=self.dtype, count=n).reshape(self.shape)

    def free(self):
        if self.ptr:
            self.array = None
            _cabi.lib().xh_host_free(self.ptr)
            self.ptr = 0

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class _ResultBlock:
    """One pinned, device-mapped host block lent to a result array: numpy keeps this object as the array's base, and when
    the array (and every view of it) is gone the block goes back to the pool."""

    __slots__ = ("ptr", "cap", "__array_interface__", "_pool", "__weakref__")

    def __init__(self, pool, ptr, cap, shape, typestr):
        self._pool, self.ptr, self.cap = pool, ptr, cap
        self.__array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3}

    def __del__(self):
        try:
            self._pool._give(self.ptr, self.cap)
        except Exception:
            pass


class ResultPool:
    """Small results (a 256 x 256 float64 histogram is 512 KB) are returned in page-locked host memory that the GPU writes
    directly (the density kernel stores into it; plain histograms are DMA-ed into it): no staging copy on the host.
    Blocks are recycled by size class; at most ``limit`` bytes are lent out or cached — beyond that, and for larger
    results, ordinary pageable arrays are used."""

    def __init__(self, limit=256 << 20, largest=8 << 20):
        self.limit, self.largest = limit, largest
        self.free = {}
        self.total = 0
        self._slab = {}          # size class -> blocks in its next slab

    def array(self, shape, dtype):
        dtype = np.dtype(dtype)
        nbytes = math.prod(shape) * dtype.itemsize
        if nbytes == 0 or nbytes > self.largest:
            return None
        cap = max(1 << 16, 1 << (nbytes - 1).bit_length())
        lst = self.free.get(cap)
        if lst:
            ptr = lst.pop()
        else:
            # Page-locking is slow (4-15 ms per allocation, whatever its size), so a miss takes a SLAB of blocks in one
            # allocation — 4 blocks the first time, twice as many on every later miss of the class: a caller who keeps the
            # previous result while asking for the next one never pays it in steady state, and one who keeps every result
            # (a list of histograms) pays it O(log n) times.  Slabs are never given back (process lifetime).
            n = min(self._slab.get(cap, 4), (self.limit - self.total) // cap)
            if n < 1:
                return None
            p = C.c_void_p()
            if _cabi.lib().xh_host_alloc(n * cap, C.byref(p)) != 0:
                return None
            self.total += n * cap
            self._slab[cap] = min(2 * n, 64)
            ptr = p.value
            self.free.setdefault(cap, []).extend(ptr + i * cap for i in range(n - 1, 0, -1))
        return np.asarray(_ResultBlock(self, ptr, cap, tuple(shape), dtype.str))

    def _give(self, ptr, cap):
        self.free.setdefault(cap, []).append(ptr)


result_pool = ResultPool()


def is_device_array(a) -> bool:
    return isinstance(a, DeviceArray) or (hasattr(a, "__cuda_array_interface__") and not isinstance(a, np.ndarray))


def as_device_view(a):
    """(ptr, shape, dtype, device) of a DeviceArray or any C-contiguous ``__cuda_array_interface__`` object."""
    if isinstance(a, DeviceArray):
        return a.ptr, a.shape, a.dtype, a.device
    cai = a.__cuda_array_interface__
    if cai.get("strides") is not None:
        shape, item = tuple(cai["shape"]), np.dtype(cai["typestr"]).itemsize
        expect = tuple(int(np.prod(shape[i + 1:], dtype=np.int64)) * item for i in range(len(shape)))
        if tuple(cai["strides"]) != expect:
            raise ValueError("device inputs must be C-contiguous")
    dev = getattr(getattr(a, "device", None), "index", None)
    if dev is None:
        dev = getattr(getattr(a, "device", None), "id", 0) or 0
    return int(cai["data"][0]), tuple(cai["shape"]), np.dtype(cai["typestr"]), int(dev)
