"""Minimal stand-in for xarray (see tests/stubs/README.md): DataArray with named dims, coords, attrs and the handful of
methods the histogram wrapper and its tests use.  Eager, numpy-backed, no indexes beyond exact alignment."""
import numpy as np

__all__ = ["DataArray", "align", "full_like", "testing"]
__stub__ = True


class _Coords(dict):
    pass


class DataArray:
    def __init__(self, data, coords=None, dims=None, name=None, attrs=None):
        self.data = data if hasattr(data, "shape") and not isinstance(data, (list, tuple)) else np.asarray(data)
        nd = len(self.data.shape)
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(nd))
        self.dims = (dims,) if isinstance(dims, str) else tuple(dims)
        assert len(self.dims) == nd, (self.dims, self.data.shape)
        self.name = name
        self.attrs = dict(attrs or {})
        self.coords = _Coords()
        if coords is not None:
            if isinstance(coords, (list, tuple)):
                coords = dict(zip(self.dims, coords))
            for k, v in coords.items():
                self._set_coord(k, v)

    def _set_coord(self, k, v):
        if isinstance(v, DataArray):
            c = DataArray(v.data, dims=v.dims, name=k, attrs=v.attrs)
        elif isinstance(v, tuple):
            cd, cv = v[0], v[1]
            c = DataArray(np.asarray(cv), dims=cd, name=k, attrs=v[2] if len(v) > 2 else None)
        else:
            c = DataArray(np.asarray(v), dims=(k,), name=k)
        for d, n in zip(c.dims, c.data.shape):
            assert d in self.dims and n == self.shape[self.dims.index(d)], f"coordinate {k} does not fit"
        self.coords[k] = c

    # -- basics
    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return len(self.dims)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def values(self):
        return np.asarray(self.data.compute() if hasattr(self.data, "compute") else self.data)

    def __getitem__(self, key):
        if isinstance(key, str):
            return self.coords[key]
        raise NotImplementedError

    def __setitem__(self, key, value):
        self._set_coord(key, value)

    def __getattr__(self, item):
        coords = self.__dict__.get("coords")
        if coords is not None and item in coords:
            return coords[item]
        raise AttributeError(item)

    @property
    def sizes(self):
        return dict(zip(self.dims, self.shape))

    def chunk(self, chunks):
        import dask.array as dsa
        return self._like(dsa.from_array(self.values, chunks=chunks), self.dims)

    def get_axis_num(self, d):
        return self.dims.index(d)

    def _like(self, data, dims, keep_coords=True):
        out = DataArray(data, dims=dims, name=self.name, attrs=self.attrs)
        if keep_coords:
            for k, c in self.coords.items():
                if set(c.dims) <= set(dims):
                    out.coords[k] = c
        return out

    def reset_coords(self, drop=False):
        assert drop
        out = self._like(self.data, self.dims, keep_coords=False)
        for k, c in self.coords.items():
            if k in self.dims:                 # dimension coordinates stay
                out.coords[k] = c
        return out

    def expand_dims(self, dims):
        out = self
        for k, n in dims.items():
            assert n == 1
            out = out._like(out.data[None] if not hasattr(out.data, "_expand") else out.data._expand(), (k,) + out.dims)
        return out

    def transpose(self, *dims):
        perm = [self.dims.index(d) for d in dims]
        data = self.data.transpose(perm) if hasattr(self.data, "transpose") else np.transpose(self.data, perm)
        return self._like(data, tuple(dims))

    def isel(self, **sel):
        idx = tuple(sel.get(d, slice(None)) for d in self.dims)
        dims = tuple(d for d in self.dims if d not in sel)
        out = DataArray(self.data[idx], dims=dims, name=self.name, attrs=self.attrs)
        for k, c in self.coords.items():
            if set(c.dims) <= set(dims):
                out.coords[k] = c
        return out

    def sum(self, dim=None):
        dims = [dim] if isinstance(dim, str) else list(dim if dim is not None else self.dims)
        axes = tuple(self.dims.index(d) for d in dims)
        keep = tuple(d for d in self.dims if d not in dims)
        return self._like(np.asarray(self.values).sum(axis=axes) if axes else self.values, keep)

    def identical(self, other):
        return _identical(self, other)

    def _binary(self, other, op):
        o = other.values if isinstance(other, DataArray) else other
        return self._like(op(self.values, o), self.dims)

    def __mul__(self, other):
        return self._binary(other, np.multiply)

    def __add__(self, other):
        if isinstance(other, DataArray) and other.dims != self.dims:      # outer broadcast of 1-D coordinates (da.X**2 + da.Y**2)
            dims = self.dims + tuple(d for d in other.dims if d not in self.dims)
            a = self.values.reshape(self.shape + (1,) * (len(dims) - self.ndim))
            b = other.values.reshape((1,) * (len(dims) - other.ndim) + other.shape)
            return DataArray(a + b, dims=dims)
        return self._binary(other, np.add)

    def __pow__(self, p):
        return self._like(self.values ** p, self.dims)

    def __repr__(self):
        return f"<stub DataArray {self.name} {dict(zip(self.dims, self.shape))}>"


def _identical(a, b):
    if a.name != b.name or a.dims != b.dims or a.attrs != b.attrs:
        return False
    av, bv = np.asarray(a.values), np.asarray(b.values)
    if av.shape != bv.shape or not np.array_equal(av, bv):
        return False
    if set(a.coords) != set(b.coords):
        return False
    return all(a.coords[k].dims == b.coords[k].dims and np.array_equal(a.coords[k].values, b.coords[k].values)
               and a.coords[k].attrs == b.coords[k].attrs for k in a.coords)


def align(*arrays, join="exact"):
    assert join == "exact"
    seen = {}
    for a in arrays:
        for d, n in zip(a.dims, a.shape):
            if d in seen and seen[d][0] != n:
                raise ValueError(f"indexes along dimension {d!r} are not equal")
            c = a.coords.get(d)
            if d in seen and c is not None and seen[d][1] is not None and not np.array_equal(seen[d][1].values, c.values):
                raise ValueError(f"indexes along dimension {d!r} are not equal")
            if d not in seen or seen[d][1] is None:
                seen[d] = (n, c)
    return tuple(arrays)


def full_like(a, value):
    return a._like(np.full(a.shape, value, dtype=np.result_type(a.values.dtype, type(value))), a.dims)


class testing:
    @staticmethod
    def assert_identical(a, b):
        assert _identical(a, b), (a, b, a.values, b.values, dict(a.coords), dict(b.coords))
