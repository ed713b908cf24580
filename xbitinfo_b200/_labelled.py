"""Minimal labelled-array containers used when ``xarray`` is not installed.

The public API of xbitinfo takes and returns ``xarray.Dataset`` / ``xarray.DataArray``
objects.  xarray is an optional dependency of this package (it is absent from the build
image): when it is importable and the caller passes xarray objects,
xarray objects come back; otherwise these small stand-ins carry the same information
(``dims``, ``coords``, ``attrs``, ``encoding``, ``name``, ``values``) and implement the few
members the hot path touches (``xbitinfo/xbitinfo.py:81-115, 337-441``,
``xbitinfo/bitround.py:15-101``).  They are containers, not a re-implementation of xarray.
"""
from __future__ import annotations

from collections import OrderedDict

import numpy as np


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _is_lazy(data) -> bool:
    from .streaming import is_lazy_array

    return is_lazy_array(data)


class DataArray:
    def __init__(self, data, dims=None, coords=None, name=None, attrs=None, encoding=None):
        if not _is_torch_tensor(data) and not isinstance(data, np.ndarray) and not _is_lazy(data):
            data = np.asarray(data)
        self.data = data
        ndim = len(data.shape)
        if dims is None:
            dims = tuple(f"dim_{i}" for i in range(ndim))
        if isinstance(dims, str):
            dims = (dims,)
        dims = tuple(dims)
        if len(dims) != ndim:
            raise ValueError(f"different number of dimensions on data and dims: {ndim} vs {len(dims)}")
        self.dims = dims
        self.coords = OrderedDict()
        for k, v in (coords or {}).items():
            self.coords[k] = v if isinstance(v, DataArray) else _coord(k, v, dims)
        self.name = name
        self.attrs = dict(attrs or {})
        self.encoding = dict(encoding or {})

    # --- array-ish -----------------------------------------------------------------
    @property
    def values(self):
        if _is_torch_tensor(self.data):
            return self.data.detach().cpu().numpy()
        if _is_lazy(self.data):
            return np.asarray(self.data[...])
        return self.data

    @property
    def dtype(self):
        if _is_torch_tensor(self.data):
            import torch  # noqa: F401

            return np.dtype(str(self.data.dtype).replace("torch.", ""))
        return self.data.dtype

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64))

    @property
    def sizes(self):
        return OrderedDict(zip(self.dims, self.shape))

    def __array__(self, dtype=None, copy=None):
        v = self.values
        return v.astype(dtype) if dtype is not None else v

    def __int__(self):
        return int(np.asarray(self.values).reshape(()))

    def __float__(self):
        return float(np.asarray(self.values).reshape(()))

    def item(self):
        return np.asarray(self.values).item()

    def get_axis_num(self, dim):
        try:
            return self.dims.index(dim)
        except ValueError:
            raise ValueError(f"{dim!r} not found in array dimensions {self.dims!r}") from None

    def astype(self, dtype):
        if _is_torch_tensor(self.data):
            import torch

            data = self.data.to(getattr(torch, np.dtype(dtype).name))
        else:
            data = self.data.astype(dtype)
        return self._replace(data)

    def copy(self, deep=True):
        data = self.data.clone() if _is_torch_tensor(self.data) else (self.data.copy() if deep else self.data)
        return self._replace(data)

    def _replace(self, data, dims=None, coords=None):
        out = DataArray.__new__(DataArray)
        out.data = data
        out.dims = self.dims if dims is None else tuple(dims)
        out.coords = OrderedDict(self.coords if coords is None else coords)
        out.name = self.name
        out.attrs = dict(self.attrs)
        out.encoding = dict(self.encoding)
        return out

    # --- selection ---------------------------------------------------------------------
    def isel(self, **indexers):
        data, dims = self.values, list(self.dims)
        coords = OrderedDict(self.coords)
        for d, idx in indexers.items():
            ax = dims.index(d)
            data = np.take(data, idx, axis=ax) if not isinstance(idx, slice) else data[(slice(None),) * ax + (idx,)]
            if d in coords:
                c = coords[d]
                cv = c.values[idx]
                coords[d] = DataArray(cv, dims=(d,) if np.ndim(cv) else (), name=d, attrs=c.attrs)
            if not isinstance(idx, slice) and np.ndim(idx) == 0:
                dims.pop(ax)
        return self._replace(data, dims=dims, coords=coords)

    def _reduce(self, fn, dim):
        """Reduce over one named dimension (``kb.max(dim="dim")`` of the reference's pipeline test)."""
        if dim is None:
            return self._replace(np.asarray(fn(self.values)), dims=(), coords={})
        ax = self.get_axis_num(dim)
        dims = [d for d in self.dims if d != dim]
        coords = OrderedDict((k, c) for k, c in self.coords.items() if dim not in getattr(c, "dims", ()) and k != dim)
        return self._replace(fn(self.values, axis=ax), dims=dims, coords=coords)

    def max(self, dim=None):
        return self._reduce(np.max, dim)

    def min(self, dim=None):
        return self._reduce(np.min, dim)

    def sel(self, **indexers):
        pos = {}
        for d, label in indexers.items():
            labels = list(np.asarray(self.coords[d].values).reshape(-1))
            if np.ndim(label) == 0:
                pos[d] = labels.index(label)
            else:
                pos[d] = [labels.index(v) for v in label]
        return self.isel(**pos)

    def __repr__(self):
        return f"<xbitinfo_b200.DataArray {self.name!r} {dict(self.sizes)} {self.dtype}>\n{self.values!r}"


def _coord(name, value, parent_dims):
    v = np.asarray(value)
    if v.ndim == 0:
        return DataArray(v, dims=(), name=name)
    return DataArray(v, dims=(name,), name=name)


class Dataset:
    def __init__(self, data_vars=None, coords=None, attrs=None):
        self._vars = OrderedDict()
        self.coords = OrderedDict()
        for k, v in (coords or {}).items():
            self.coords[k] = v if isinstance(v, DataArray) else _coord(k, v, ())
        self.attrs = dict(attrs or {})
        for k, v in (data_vars or {}).items():
            self[k] = v

    @property
    def data_vars(self):
        return self._vars

    @property
    def dims(self):
        out = OrderedDict()
        for v in self._vars.values():
            for d, n in zip(v.dims, v.shape):
                out.setdefault(d, n)
        return out

    sizes = dims

    def keys(self):
        return self._vars.keys()

    def __iter__(self):
        return iter(self._vars)

    def __contains__(self, k):
        return k in self._vars

    def __len__(self):
        return len(self._vars)

    def __getitem__(self, key):
        if isinstance(key, (list, tuple)):
            return Dataset({k: self._vars[k] for k in key}, coords=self.coords, attrs=self.attrs)
        if key in self._vars:
            return self._vars[key]
        return self.coords[key]

    def __setitem__(self, key, value):
        if isinstance(value, tuple):
            dims, data = value[0], value[1]
            value = DataArray(data, dims=dims, attrs=value[2] if len(value) > 2 else None)
        if not isinstance(value, DataArray):
            raise TypeError("Dataset values must be DataArray or (dims, data) tuples")
        if value.name != key:
            value = value._replace(value.data)
            value.name = key
        self._vars[key] = value
        for ck, cv in value.coords.items():
            self.coords.setdefault(ck, cv)

    def copy(self, deep=False):
        out = Dataset(coords=self.coords, attrs=self.attrs)
        for k, v in self._vars.items():
            out._vars[k] = v.copy(deep=deep)
        return out

    def astype(self, dtype):
        out = Dataset(coords=self.coords, attrs=self.attrs)
        for k, v in self._vars.items():
            out._vars[k] = v.astype(dtype)
        return out

    def sel(self, **indexers):
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            use = {d: i for d, i in indexers.items() if d in v.dims}
            out[k] = v.sel(**use) if use else v
        return out

    def max(self, dim=None):
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            out[k] = v.max(dim=dim) if (dim is None or dim in v.dims) else v
        return out

    def min(self, dim=None):
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            out[k] = v.min(dim=dim) if (dim is None or dim in v.dims) else v
        return out

    def isel(self, indexers=None, **kw):
        indexers = dict(indexers or {}, **kw)
        out = Dataset(attrs=self.attrs)
        for k, v in self._vars.items():
            use = {d: i for d, i in indexers.items() if d in v.dims}
            out[k] = v.isel(**use) if use else v
        return out

    def __repr__(self):
        lines = [f"<xbitinfo_b200.Dataset dims={dict(self.dims)}>"]
        for k, v in self._vars.items():
            lines.append(f"    {k} {v.dims} {v.dtype}")
        return "\n".join(lines)


def concat(objs, dim):
    """Concatenate DataArrays or Datasets along an existing dimension (numpy data only)."""
    first = objs[0]
    if isinstance(first, Dataset):
        out = Dataset(attrs=first.attrs)
        for k in first.data_vars:
            out[k] = concat([o[k] for o in objs], dim)
        return out
    ax = first.get_axis_num(dim)
    data = np.concatenate([np.asarray(o.values) for o in objs], axis=ax)
    coords = OrderedDict(first.coords)
    if dim in coords:
        cv = np.concatenate([np.atleast_1d(o.coords[dim].values) for o in objs])
        coords[dim] = DataArray(cv, dims=(dim,), name=dim, attrs=first.coords[dim].attrs)
    return first._replace(data, coords=coords)
