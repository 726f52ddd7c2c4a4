"""dask.array stand-in: chunked arrays evaluated eagerly, with the chunk bookkeeping of ``blockwise`` kept faithful —
chunks of the inputs are unified along shared axes (unaligned chunks between arguments are re-cut at the union of their
boundaries, as dask's ``unify_chunks`` does), the function is called once per block of the unified grid, every block
result occupies one cell along adjusted axes, and ``.sum`` over those axes adds the per-block partial results."""
import itertools

import numpy as np

__stub__ = True


def _normalize_chunks(chunks, shape):
    if isinstance(chunks, int):
        chunks = (chunks,) * len(shape)
    chunks = tuple(chunks) + (None,) * (len(shape) - len(tuple(chunks)))
    out = []
    for c, n in zip(chunks, shape):
        if c is None or c == -1:
            out.append((n,) if n else (0,))
        elif isinstance(c, int):
            full, rem = divmod(n, c)
            out.append((c,) * full + ((rem,) if rem else ()))
        else:
            assert sum(c) == n
            out.append(tuple(c))
    return tuple(out)


class Array:
    def __init__(self, a, chunks):
        self._a = np.asarray(a)
        self.chunks = _normalize_chunks(chunks, self._a.shape)
        self.calls = 0

    shape = property(lambda self: self._a.shape)
    ndim = property(lambda self: self._a.ndim)
    dtype = property(lambda self: self._a.dtype)
    size = property(lambda self: self._a.size)

    def compute(self):
        return self._a

    def __array__(self, dtype=None, copy=None):
        return self._a if dtype is None else self._a.astype(dtype)

    def rechunk(self, chunks):
        return Array(self._a, chunks)

    def transpose(self, axes):
        return Array(self._a.transpose(axes), tuple(self.chunks[i] for i in axes))

    def __getitem__(self, idx):
        if idx is None or (isinstance(idx, tuple) and idx[0] is None and all(i == slice(None) for i in idx[1:])):
            return Array(self._a[None], ((1,),) + self.chunks)
        b = self._a[idx]
        return Array(b, tuple((n,) for n in b.shape))

    def _expand(self):
        return Array(self._a[None], ((1,),) + self.chunks)

    def sum(self, axis=None):
        r = self._a.sum(axis=axis)
        return Array(r, tuple((n,) for n in np.shape(r)))

    def squeeze(self, axis=None):
        r = self._a.squeeze(axis)
        return Array(r, tuple((n,) for n in r.shape))

    def reshape(self, *shape):
        r = self._a.reshape(*shape)
        return Array(r, tuple((n,) for n in r.shape))

    def _bin(self, other, op):
        o = other._a if isinstance(other, Array) else other
        r = op(self._a, o)
        return Array(r, tuple((n,) for n in np.shape(r)))

    def __truediv__(self, o):
        return self._bin(o, np.divide)

    def __mul__(self, o):
        return self._bin(o, np.multiply)


def from_array(a, chunks):
    return Array(a, chunks)


def asarray(a, chunks=None):
    return a if isinstance(a, Array) else Array(a, chunks if chunks is not None else -1)


def reshape(a, shape):
    return a.reshape(shape)


def broadcast_arrays(*arrays):
    raw = [a._a if isinstance(a, Array) else np.asarray(a) for a in arrays]
    out = np.broadcast_arrays(*raw)
    res = []
    for a, b in zip(arrays, out):
        if isinstance(a, Array) and a.shape == b.shape:
            res.append(Array(b, a.chunks))
        elif isinstance(a, Array):
            # broadcast axes get one chunk, existing axes keep theirs
            lead = b.ndim - a.ndim
            ch = tuple((n,) for n in b.shape[:lead]) + tuple(c if sum(c) == n else (n,) for c, n in zip(a.chunks, b.shape[lead:]))
            res.append(Array(b, ch))
        else:
            res.append(Array(b, -1))
    return res


def _bounds(chunks):
    return list(np.cumsum((0,) + tuple(chunks)))


def blockwise(func, out_ind, *args, new_axes=None, adjust_chunks=None, meta=None, **kwargs):
    arrays, inds = args[0::2], args[1::2]
    new_axes, adjust_chunks = new_axes or {}, adjust_chunks or {}
    # unify the chunks of every index over all arguments (union of the boundaries)
    cuts = {}
    for a, ind in zip(arrays, inds):
        for ax, i in enumerate(ind):
            cuts.setdefault(i, set()).update(_bounds(a.chunks[ax]))
    cuts = {i: sorted(c) for i, c in cuts.items()}
    in_inds = [i for i in out_ind if i not in new_axes]
    grid = [range(len(cuts[i]) - 1) for i in in_inds]
    out_shape_blocks = {}
    blocks = {}
    for pos in itertools.product(*grid):
        where = dict(zip(in_inds, pos))
        block_args = []
        for a, ind in zip(arrays, inds):
            sl = tuple(slice(cuts[i][where[i]], cuts[i][where[i] + 1]) for i in ind)
            block_args.append(a._a[sl])
        blocks[pos] = np.asarray(func(*block_args, **kwargs))
        out_shape_blocks[pos] = blocks[pos].shape
    # assemble: concatenate the block results along every input axis in turn (adjusted axes have extent 1 per block)
    def assemble(prefix, depth):
        if depth == len(in_inds):
            return blocks[tuple(prefix)]
        parts = [assemble(prefix + [k], depth + 1) for k in grid[depth]]
        return np.concatenate(parts, axis=depth)
    full = assemble([], 0) if blocks else np.zeros((0,) * len(out_ind))
    for ax, i in enumerate(in_inds):
        if i in adjust_chunks:
            assert full.shape[ax] == len(cuts[i]) - 1, "adjust_chunks: every block must collapse to extent 1"
    out = Array(full, tuple((1,) * full.shape[ax] if in_inds[ax] in adjust_chunks else tuple(np.diff(cuts[in_inds[ax]]))
                             for ax in range(len(in_inds))) + tuple((n,) for n in full.shape[len(in_inds):]))
    out.calls = len(blocks)
    return out

import types as _types
core = _types.SimpleNamespace(Array=Array)
