"""xarray is optional: use it when the caller does, fall back to the in-repo containers."""
from __future__ import annotations

from . import _labelled

try:  # pragma: no cover - xarray is not installed in the build image
    import xarray as _xarray
except ImportError:  # pragma: no cover
    _xarray = None


def xarray_module():
    return _xarray


def is_xarray(obj) -> bool:
    return _xarray is not None and isinstance(obj, (_xarray.Dataset, _xarray.DataArray))


def is_dataset(obj) -> bool:
    if isinstance(obj, _labelled.Dataset):
        return True
    return _xarray is not None and isinstance(obj, _xarray.Dataset)


def is_dataarray(obj) -> bool:
    if isinstance(obj, _labelled.DataArray):
        return True
    return _xarray is not None and isinstance(obj, _xarray.DataArray)


def namespace_for(obj):
    """Module-like object providing Dataset / DataArray constructors matching ``obj``."""
    if is_xarray(obj):
        return _xarray
    return _labelled


def raw_data(da):
    """The array behind a DataArray without forcing it to the host: a torch tensor stays a
    torch tensor, a lazy array (dask, zarr, h5py ...) stays lazy and is streamed slab by slab
    (``streaming.py``), anything else becomes a numpy array."""
    from .streaming import is_lazy_array

    if isinstance(da, _labelled.DataArray):
        return da.data
    data = getattr(da, "data", None)
    if data is not None and (type(data).__module__.startswith("torch") or is_lazy_array(data)):
        return data
    return da.values
