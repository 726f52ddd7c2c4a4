"""Xarray API: label-handling glue of ``xhistogram.xarray.histogram``, kept unchanged in behaviour.

Mirrors reference xhistogram/xarray.py:13-201: type/name checks, ``reset_coords``,
``xr.align(join="exact")``, manual broadcast via ``expand_dims`` + ``transpose``, ``dim`` -> ``axis``,
then the numpy-level ``core.histogram`` (whose hot path runs on the GPU) and reconstruction of a
``DataArray`` with ``<name>_bin`` centre coordinates.  Nothing here is O(samples) except xarray's
own transposition of its inputs.  The reference's unreachable ``args.transposed`` typo branch
(xarray.py:146-149) is not reproduced: arrays are always brought to the common dim order.
"""
from __future__ import annotations

from .core import histogram as _histogram

_range = range


def histogram(*args, bins=None, range=None, dim=None, weights=None, density=False, block_size="auto",
              keep_coords=False, bin_dim_suffix="_bin"):
    """Histogram applied along specified dimensions (signature of ``xhistogram.xarray.histogram``)."""
    import xarray as xr

    args = list(args)
    n_args = len(args)
    for a in args:
        if not isinstance(a, xr.DataArray):
            raise TypeError(
                "xhistogram.xarray.histogram accepts only xarray.DataArray "
                f"objects but a {type(a).__name__} was provided"
            )
    for a in args:
        assert a.name is not None, "all arrays must have a name"

    if not keep_coords:                                    # xarray.py:120-123
        args = [da.reset_coords(drop=True) for da in args]
    if weights is not None:
        args += [weights.reset_coords(drop=True)]
    args = list(xr.align(*args, join="exact"))             # xarray.py:126
    a0 = args[0]
    a_coords = a0.coords

    dims_ordered = list(dict.fromkeys(d for a in args for d in a.dims))   # first-seen union, xarray.py:135-136
    aligned = []
    for a in args:
        missing = [d for d in dims_ordered if d not in a.dims]
        a = a.expand_dims({k: 1 for k in missing})
        aligned.append(a.transpose(*dims_ordered))
    data = [a.data for a in aligned]
    weights_data = data.pop() if weights is not None else None

    if dim is not None:                                    # xarray.py:157-162
        dims_to_keep = [d for d in dims_ordered if d not in dim]
        axis = [aligned[0].get_axis_num(d) for d in dim]
    else:
        dims_to_keep = []
        axis = None

    h_data, edges = _histogram(*data, weights=weights_data, bins=bins, range=range, axis=axis,
                               density=density, block_size=block_size)

    new_dims = [a.name + bin_dim_suffix for a in args[:n_args]]          # xarray.py:175-183
    output_dims = dims_to_keep + new_dims
    centres = [0.5 * (e[:-1] + e[1:]) for e in edges]
    coords = {name: a0[name] for name in dims_to_keep if name in a_coords}
    coords.update({name: ((name,), c, a.attrs) for name, c, a in zip(new_dims, centres, args)})
    if keep_coords:                                        # xarray.py:192-195
        for c in a_coords:
            if c not in coords and set(a0[c].dims).issubset(output_dims):
                coords[c] = a0[c]
    name = "_".join(["histogram"] + [a.name for a in args[:n_args]])
    return xr.DataArray(h_data, dims=output_dims, coords=coords, name=name)
