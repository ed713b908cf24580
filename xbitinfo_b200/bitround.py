"""``xr_bitround``: round to ``keepbits`` mantissa bits with the sm_100a bitround kernel.

Host-side mirror of ``xbitinfo/bitround.py:7-101``.  ``_cuda_bitround`` takes the place of
``_np_bitround`` (numcodecs ``BitRound`` encode+decode on a copy); the keepbits adapter,
the per-variable recursion, the ``apply_ufunc(dask="parallelized")`` chunk mapping and the
``_QuantizeBitRoundNumberOfSignificantDigits`` attribute are kept.
"""
from __future__ import annotations

import numpy as np

from . import _labelled, _xr, engine

_MANTISSA_BITS = {"float16": 10, "float32": 23, "float64": 52}


def _cuda_bitround(data, keepbits):
    """Bitround for arrays: new array, input untouched  [bitround.py:7-12].  Error
    behaviour of numcodecs.BitRound: negative keepbits or more than the mantissa holds ->
    ValueError, non-float data -> TypeError; keepbits == mantissa bits returns the data
    unchanged."""
    keepbits = int(keepbits)
    if keepbits < 0:
        raise ValueError("keepbits must be zero or positive")
    dtype = _dtype_of(data)
    if dtype.kind != "f" or dtype.itemsize > 8:
        raise TypeError("Only float arrays (16-64bit) can be bit-rounded")
    if keepbits > _MANTISSA_BITS[dtype.name]:
        raise ValueError("keepbits too large for given dtype")
    return engine.bitround(data, keepbits)


def _dtype_of(data) -> np.dtype:
    if type(data).__module__.startswith("torch"):
        return np.dtype(str(data.dtype).replace("torch.", ""))
    return np.asarray(data).dtype


_ONE_INFLEVEL = "Information content is only allowed for one 'inflevel' here. Please make a selection."
_ONE_DIM = ("Information content is only allowed along one dimension here. Please select one `dim`. "
            "To find the maximum keepbits, simply use `keepbits.max(dim='dim')`")


def _require_single(labelled, coord, message, optional=False):
    """AssertionError unless coordinate ``coord`` of a keepbits object holds at most one entry
    (the reference asserts the same, so callers that catch AssertionError keep working)."""
    if optional and coord not in labelled.coords:
        return
    if not labelled.coords[coord].shape <= (1,):
        raise AssertionError(message)


def _lookup(name, table, wrap=lambda v: v):
    if name not in table.keys():
        raise ValueError(f"name {name} not for in keepbits: {table.keys()}")
    return wrap(table[name])


def _keepbits_interface(da, keepbits):
    """Resolve any of the accepted keepbits forms to the int that applies to ``da``.

    Same contract as the reference adapter (xbitinfo/bitround.py:15-66): a plain integer is taken as
    is; a mapping ``{name: int}`` and a keepbits Dataset are indexed by ``da.name`` (missing ->
    ValueError); a keepbits DataArray must carry the same name (else KeyError); Dataset / DataArray
    forms may hold one ``inflevel`` and one ``dim`` only (AssertionError); anything else -> TypeError.
    """
    if not _xr.is_dataarray(da):
        raise AssertionError("expected a DataArray")
    if isinstance(keepbits, (int, np.integer)) and not isinstance(keepbits, bool):
        return int(keepbits)
    if isinstance(keepbits, dict):
        return _lookup(da.name, keepbits)
    if _xr.is_dataset(keepbits):
        _require_single(keepbits, "inflevel", _ONE_INFLEVEL)
        _require_single(keepbits, "dim", _ONE_DIM, optional=True)
        return _lookup(da.name, keepbits, int)
    if _xr.is_dataarray(keepbits):
        _require_single(keepbits, "inflevel", _ONE_INFLEVEL)
        _require_single(keepbits, "dim", _ONE_DIM)
        if da.name != keepbits.name:
            raise KeyError(f"no keepbits found for variable {da.name}")
        return int(keepbits)
    raise TypeError(f"type {type(keepbits)} is not a valid type for keepbits.")


def _round_variable(da, keep):
    """One DataArray -> rounded DataArray carrying the CF quantisation attribute (bitround.py:99-100)."""
    if _xr.is_xarray(da):  # pragma: no cover - xarray is absent from the build image
        rounded = _xr.xarray_module().apply_ufunc(_cuda_bitround, da, keep, dask="parallelized", keep_attrs=True)
    else:
        rounded = da._replace(_cuda_bitround(da.data, keep))
    rounded.attrs["_QuantizeBitRoundNumberOfSignificantDigits"] = keep
    return rounded


def xr_bitround(da, keepbits):
    """Bitround a Dataset or DataArray to the keepbits :py:func:`get_keepbits` returned (or an int /
    ``{name: int}``); a new object comes back, the input is untouched.  Public twin of the reference's
    ``xr_bitround`` (xbitinfo/bitround.py:69-101) with K3 in place of numcodecs' BitRound."""
    if _xr.is_dataset(da):
        out = da.copy()
        for name in da.data_vars:
            out[name] = _round_variable(da[name], _keepbits_interface(da[name], keepbits))
        return out
    return _round_variable(da, _keepbits_interface(da, keepbits))


def bitround_along_dim(ds, info_per_bit, dim, inflevels=[1.0, 0.9999, 0.99, 0.975, 0.95], keepbits=None):
    """Apply bitrounding on slices along ``dim`` based on ``inflevels`` (figure helper,
    [bitround.py:138-207]): the axis is cut into ``len(inflevels)`` slices (the last one takes
    the remainder), slice i is rounded to ``get_keepbits(info_per_bit, inflevels[i])`` (left
    untouched for inflevel 1), and the slices are concatenated again."""
    from .keepbits import get_keepbits

    new_ds = []
    if inflevels is not None and keepbits is not None:
        raise ValueError("Either inflevel or keepbits should be None")
    elif inflevels is not None:
        size = ds[dim].size if _xr.is_xarray(ds) else ds.sizes[dim]
        stride = size // len(inflevels)
        for i, inf in enumerate(inflevels):  # last slice might be a bit larger
            ds_slice = ds.isel({dim: slice(stride * i, stride * (i + 1) if i != len(inflevels) - 1 else None)})
            keepbits_slice = get_keepbits(info_per_bit, inf)
            new_ds.append(xr_bitround(ds_slice, keepbits_slice) if inf != 1 else ds_slice)
    elif keepbits is not None:
        new_ds = [xr_bitround(ds, keepbits)]
    else:
        raise ValueError("Either inflevel or keepbits should NOT be None")
    if _xr.is_xarray(ds):  # pragma: no cover - xarray is absent from the build image
        return _xr.xarray_module().concat(new_ds, dim)
    return _labelled.concat(new_ds, dim)
