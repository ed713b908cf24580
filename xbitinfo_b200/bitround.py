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


def _keepbits_interface(da, keepbits):
    """Common interface to the allowed keepbits types  [bitround.py:15-66]."""
    assert _xr.is_dataarray(da)
    if isinstance(keepbits, (int, np.integer)) and not isinstance(keepbits, bool):
        keep = int(keepbits)
    elif isinstance(keepbits, dict):
        v = da.name
        if v in keepbits.keys():
            keep = keepbits[v]
        else:
            raise ValueError(f"name {v} not for in keepbits: {keepbits.keys()}")
    elif _xr.is_dataset(keepbits):
        assert keepbits.coords["inflevel"].shape <= (
            1,
        ), "Information content is only allowed for one 'inflevel' here. Please make a selection."
        if "dim" in keepbits.coords:
            assert keepbits.coords["dim"].shape <= (
                1,
            ), "Information content is only allowed along one dimension here. Please select one `dim`. To find the maximum keepbits, simply use `keepbits.max(dim='dim')`"
        v = da.name
        if v in keepbits.keys():
            keep = int(keepbits[v])
        else:
            raise ValueError(f"name {v} not for in keepbits: {keepbits.keys()}")
    elif _xr.is_dataarray(keepbits):
        assert keepbits.coords["inflevel"].shape <= (
            1,
        ), "Information content is only allowed for one 'inflevel' here. Please make a selection."
        assert keepbits.coords["dim"].shape <= (
            1,
        ), "Information content is only allowed along one dimension here. Please select one `dim`. To find the maximum keepbits, simply use `keepbits.max(dim='dim')`"
        v = da.name
        if v == keepbits.name:
            keep = int(keepbits)
        else:
            raise KeyError(f"no keepbits found for variable {v}")
    else:
        raise TypeError(f"type {type(keepbits)} is not a valid type for keepbits.")
    return keep


def xr_bitround(da, keepbits):
    """Apply bitrounding based on keepbits from :py:func:`get_keepbits` to a Dataset or
    DataArray  [bitround.py:69-101]."""
    if _xr.is_dataset(da):
        da_bitrounded = da.copy()
        for v in da.data_vars:
            da_bitrounded[v] = xr_bitround(da[v], keepbits)
        return da_bitrounded

    assert _xr.is_dataarray(da)
    keep = _keepbits_interface(da, keepbits)
    if _xr.is_xarray(da):  # pragma: no cover - xarray is absent from the build image
        xr = _xr.xarray_module()
        out = xr.apply_ufunc(_cuda_bitround, da, keep, dask="parallelized", keep_attrs=True)
    else:
        out = da._replace(_cuda_bitround(da.data, keep))
    out.attrs["_QuantizeBitRoundNumberOfSignificantDigits"] = keep
    return out


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
