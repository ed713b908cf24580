"""``get_bitinformation`` with the B200 backend behind the ``implementation=`` argument.

Host-side mirror of the bit-information part of ``xbitinfo/xbitinfo.py`` (reference
lines in brackets).  Names, arguments, return structure and error behaviour follow the
reference so its tests read the same against this module; the arithmetic is done by
``engine.bitinformation`` (CUDA kernels through the C ABI).  There is no CPU fallback.

Differences that are deliberate and documented in DESIGN.md:
* ``implementation`` defaults to ``"cuda"`` (the reference defaults to ``"julia"`` [147]);
  ``"julia"`` / ``"python"`` name the reference's own backends, which live in the
  reference package, not here.
* The keyword arguments documented at [171-176] are honoured by the cuda backend:
  ``masked_value`` (default NaN for floats, no mask for integers), ``set_zero_insignificant``
  (default True) and ``confidence`` (default 0.99).  ``masked_value`` is compared with the
  raw values, before the signed-exponent transform (the reference's julia path compares
  after it, which makes a NaN mask a no-op: SURVEY.md quirk B3).
"""
from __future__ import annotations

import json
import logging
import os
import warnings

import numpy as np

from . import _xr, engine
from ._version import __version__

IMPLEMENTATIONS = ("cuda",)
_KNOWN_KWARGS = ("masked_value", "set_zero_insignificant", "confidence")


# ---------------------------------------------------------------------------------------
# bit layout helpers  [xbitinfo.py:46-78]
# ---------------------------------------------------------------------------------------
def bit_partitioning(dtype):
    """(n_bits, n_sign, n_exponent, n_mantissa) of a dtype  [46-67]."""
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        info = np.finfo(dtype)
        parts = (info.bits, 1, info.nexp, info.nmant)
    elif dtype.kind == "i":
        bits = np.iinfo(dtype).bits
        parts = (bits, 1, 0, bits - 1)
    elif dtype.kind == "u":
        bits = np.iinfo(dtype).bits
        parts = (bits, 0, 0, bits)
    else:
        raise ValueError(f"dtype {dtype} neither known nor implemented.")
    assert parts[1] + parts[2] + parts[3] == parts[0], "The components of the datatype could not be safely inferred."
    return parts


def get_bit_coords(dtype):
    """Coordinate labels of the bits: sign, exponent bits, mantissa bits  [70-78]."""
    _, n_sign, n_exponent, n_mantissa = bit_partitioning(dtype)
    return n_sign * ["±"] + [f"e{i}" for i in range(1, n_exponent + 1)] + [f"m{i}" for i in range(1, n_mantissa + 1)]


def _backend_version_attr(implementation="cuda"):
    return f"implementation='{implementation}'"


def dict_to_dataset(info_per_bit, like=None):
    """Result dictionary {var: {"bitinfo","dim","axis","dtype"}} -> Dataset  [81-115].
    ``like`` selects the container family (xarray if it is an xarray object)."""
    ns = _xr.namespace_for(like)
    dsb = ns.Dataset()
    for v, rec in info_per_bit.items():
        dtype = np.dtype(rec["dtype"])
        dim_name = f"bit{dtype}"
        da = ns.DataArray(
            np.asarray(rec["bitinfo"], dtype="float64"),
            dims=[dim_name],
            coords={dim_name: get_bit_coords(dtype), "dim": rec["dim"]},
            name=v,
            attrs={"long_name": f"{v} bitwise information", "units": 1},
        )
        dsb[v] = da
    _decorate(dsb)
    return dsb


def _decorate(dsb):
    dsb.attrs = {
        "xbitinfo_description": "bitinformation calculated by xbitinfo_b200.get_bitinformation (sm_100a CUDA backend)",
        "python_repository": "https://github.com/observingClouds/xbitinfo",
        "julia_repository": "https://github.com/milankl/BitInformation.jl",
        "reference_paper": "http://www.nature.com/articles/s43588-021-00156-2",
        "xbitinfo_version": __version__,
        "BitInformation.jl_version": _backend_version_attr(),
    }
    for c in dsb.coords:
        if "bit" in c:
            dsb.coords[c].attrs = {
                "description": "name of the bits: '±' refers to the sign bit, 'e' to the exponents bits and 'm' to the mantissa bits."
            }
    if "dim" in dsb.coords:
        dsb.coords["dim"].attrs = {
            "description": "dimension of the source dataset along which the bitwise information has been analysed."
        }


# ---------------------------------------------------------------------------------------
# argument checks  [118-138]
# ---------------------------------------------------------------------------------------
def _check_bitinfo_kwargs(implementation=None, axis=None, dim=None, kwargs=None):
    kwargs = kwargs or {}
    if implementation == "julia":
        raise ImportError('The julia backend lives in the reference package; use implementation="cuda".')
    if axis is not None and dim is not None:
        raise ValueError("Please provide either `axis` or `dim` but not both.")
    if axis:
        if not isinstance(axis, int):
            raise ValueError(f"Please provide `axis` as `int`, found {type(axis)}.")
    if dim:
        if not isinstance(dim, str) and not isinstance(dim, list):
            raise ValueError(f"Please provide `dim` as `str` or `list`, found {type(dim)}.")
    if "mask" in kwargs:
        raise ValueError("`xbitinfo` does not wrap the mask argument. Mask your xr.Dataset with NaNs instead.")


# ---------------------------------------------------------------------------------------
# public entry  [141-238]
# ---------------------------------------------------------------------------------------
def get_bitinformation(ds, dim=None, axis=None, label=None, overwrite=False, implementation="cuda", **kwargs):
    """Information content of every bit of every variable of ``ds`` along ``dim``/``axis``.

    Same contract as ``xbitinfo.get_bitinformation`` [141-216]: returns a Dataset with one
    variable per input variable on a ``bit<dtype>`` dimension (plus a ``dim`` dimension when
    several dimensions are analysed); ``label`` caches the result dictionary as JSON.
    ``implementation="cuda"`` runs the sm_100a kernels of this package.
    """
    if overwrite is False and label is not None:
        try:
            return load_bitinformation(label, like=ds)
        except FileNotFoundError:
            logging.info(f"No bitinformation could be found for {label}. Please set `overwrite=True` for recalculation...")
    else:
        _check_bitinfo_kwargs(implementation, axis, dim, kwargs)
    return _get_bitinformation(ds, dim=dim, axis=axis, label=label, overwrite=overwrite,
                               implementation=implementation, **kwargs)


def _get_bitinformation(ds, dim=None, axis=None, label=None, overwrite=False, implementation="cuda", **kwargs):
    """[241-286]"""
    if (dim is None and axis is None) or (isinstance(dim, list) and axis is None):
        return _get_bitinformation_along_dims(ds, dim=dim, label=label, overwrite=overwrite,
                                              implementation=implementation, **kwargs)
    info_per_bit = _get_bitinformation_along_axis(ds, implementation, axis, dim, **kwargs)
    if label is not None:
        out_fn = label + ".json"
        if not os.path.exists(out_fn) or overwrite:
            save_bitinformation(info_per_bit, out_fn)
    info_per_bit = dict_to_dataset(info_per_bit, like=ds)
    _copy_source_attrs(ds, info_per_bit)
    return info_per_bit


def _copy_source_attrs(ds, info_per_bit):
    for var in info_per_bit.data_vars:  # keep attrs from input with source_ prefix  [283-285]
        for a in ds[var].attrs.keys():
            info_per_bit[var].attrs["source_" + a] = ds[var].attrs[a]


def _quantized_variable_is_scaled(ds, var) -> bool:
    """[289-306]"""
    enc = getattr(ds[var], "encoding", {}) or {}
    if not any(["add_offset" in enc, "scale_factor" in enc]):
        return False
    loaded_dtype = ds[var].dtype
    storage_dtype = enc.get("dtype", None)
    assert storage_dtype is not None, f"Variable {var} is likely quantized, but does not have a storage dtype"
    return loaded_dtype != storage_dtype


def _resolve_kwargs(da_dtype, kwargs):
    """Defaults of the documented keyword arguments [171-176, 444-464] for the cuda backend."""
    unknown = [k for k in kwargs if k not in _KNOWN_KWARGS]
    if unknown:
        raise TypeError(f"get_bitinformation got unexpected keyword argument(s) {unknown}")
    kind = np.dtype(da_dtype).kind
    if "masked_value" not in kwargs:
        if kind in "iu":
            logging.warning("No masked value given for integer type variable. Assuming no mask to apply.")
            masked_value = None
        elif kind == "f":
            masked_value = float("nan")
        else:
            raise ValueError(f"Dtype kind ({kind}) not supported.")
    else:
        masked_value = kwargs["masked_value"]
        if isinstance(masked_value, str) and masked_value == "nothing":
            masked_value = None
    set_zero = bool(kwargs.get("set_zero_insignificant", True))
    confidence = float(kwargs.get("confidence", 0.99))
    return masked_value, set_zero, confidence


def _cuda_get_bitinformation(ds, var, axis, dim, kwargs={}, precomputed=None):
    """The backend function of the drop-in boundary: same signature and result dictionary
    as ``_py_get_bitinformation`` / ``_jl_get_bitinformation`` [309-370]; ``None`` when the
    variable does not have the dimension.  ``precomputed``: ``{(var, dim): bitinfo}`` from
    :func:`_cuda_precompute_dims` (the all-dimensions call analyses a variable once for all of them)."""
    da = ds[var]
    masked_value, set_zero, confidence = _resolve_kwargs(da.dtype, kwargs)
    if axis is not None:
        dim = da.dims[axis]
    if isinstance(dim, str):
        try:
            axis = da.get_axis_num(dim)
        except ValueError:
            logging.info(f"Variable {var} does not have dimension {dim}. Skipping.")
            return None
    if precomputed is not None and (var, dim) in precomputed:
        bitinfo = precomputed[(var, dim)]
    else:
        logging.info("Calling cuda implementation now")
        bitinfo = engine.bitinformation(_xr.raw_data(da), axis, masked_value=masked_value,
                                        set_zero_insignificant=set_zero, confidence=confidence)
    return {"bitinfo": bitinfo, "dim": dim, "axis": int(axis), "dtype": str(np.dtype(da.dtype))}


def _cuda_precompute_dims(ds, dims, kwargs):
    """``{(var, dim): bitinfo}`` for every variable and every one of ``dims`` it has, each variable
    crossing PCIe once for all its dimensions (``engine.bitinformation_multi``) instead of once per
    dimension as the reference's loop over dimensions would [387-399]."""
    out = {}
    for var in ds.data_vars:
        da = ds[var]
        mine = [d for d in dims if d in da.dims]
        if len(mine) < 2:
            continue
        masked_value, set_zero, confidence = _resolve_kwargs(da.dtype, kwargs)
        axes = [da.get_axis_num(d) for d in mine]
        infos = engine.bitinformation_multi(_xr.raw_data(da), axes, masked_value=masked_value,
                                            set_zero_insignificant=set_zero, confidence=confidence)
        for d, info in zip(mine, infos):
            out[(var, d)] = info
    return out


def _get_bitinformation_along_axis(ds, implementation, axis, dim, precomputed=None, **kwargs):
    """Per-variable loop and implementation dispatch  [409-441]."""
    info_per_bit = {}
    for var in ds.data_vars:
        if _quantized_variable_is_scaled(ds, var):
            loaded_dtype = ds[var].dtype
            quantized_storage_dtype = ds[var].encoding["dtype"]
            warnings.warn(
                f"Variable {var} is quantized as {quantized_storage_dtype}, but loaded as {loaded_dtype}. Consider reopening using `mask_and_scale=False` to get sensible results",
                category=UserWarning,
            )
        if implementation == "cuda":
            info_per_bit_var = _cuda_get_bitinformation(ds, var, axis, dim, kwargs, precomputed)
            if info_per_bit_var is None:
                continue
            info_per_bit[var] = info_per_bit_var
        elif implementation in ("julia", "python"):
            raise ImportError(
                f'implementation="{implementation}" is provided by the reference package xbitinfo, not by xbitinfo_b200; '
                'use implementation="cuda".'
            )
        else:
            raise ValueError(
                f"Implementation of bitinformation algorithm {implementation} is unknown. Please choose a different one."
            )
    return info_per_bit


def _get_bitinformation_along_dims(ds, dim=None, label=None, overwrite=False, implementation="cuda", **kwargs):
    """One analysis per dimension, gathered on a new ``dim`` dimension  [373-406].  The
    reference merges the per-dimension datasets with an outer join, which orders the ``dim``
    coordinate alphabetically and fills variables that lack a dimension with NaN."""
    if dim is None:
        dim = list(ds.dims)
    precomputed = None
    if implementation == "cuda" and len(dim) > 1:
        # silence the per-call default-mask warning here; the per-dimension loop below issues it as before
        level = logging.getLogger().level
        logging.getLogger().setLevel(max(level, logging.ERROR))
        try:
            precomputed = _cuda_precompute_dims(ds, dim, kwargs)
        finally:
            logging.getLogger().setLevel(level)
    per_dim = {}
    for d in dim:
        logging.info(f"Get bitinformation along dimension {d}")
        if label is not None:
            label = "_".join([label, d])  # (sic) the suffix accumulates, as in the reference [392-393]
        info = _get_bitinformation_along_axis(ds, implementation, None, d, precomputed=precomputed, **kwargs)
        if label is not None:
            out_fn = label + ".json"
            if not os.path.exists(out_fn) or overwrite:
                save_bitinformation(info, out_fn)
        per_dim[d] = info
    return _merge_dims(ds, per_dim)


def _merge_dims(ds, per_dim):
    ns = _xr.namespace_for(ds)
    dims_sorted = sorted(per_dim.keys())
    variables = []
    for d in per_dim:
        for v in per_dim[d]:
            if v not in variables:
                variables.append(v)
    out = ns.Dataset()
    for v in variables:
        rec0 = next(per_dim[d][v] for d in per_dim if v in per_dim[d])
        dtype = np.dtype(rec0["dtype"])
        nbits = bit_partitioning(dtype)[0]
        dim_name = f"bit{dtype}"
        stack = np.full((len(dims_sorted), nbits), np.nan, dtype="float64")
        for i, d in enumerate(dims_sorted):
            if v in per_dim[d]:
                stack[i] = per_dim[d][v]["bitinfo"]
        if len(dims_sorted) == 1:  # .squeeze() in the reference [403-405]
            da = ns.DataArray(stack[0], dims=[dim_name],
                              coords={dim_name: get_bit_coords(dtype), "dim": dims_sorted[0]}, name=v)
        else:
            da = ns.DataArray(stack, dims=["dim", dim_name],
                              coords={"dim": dims_sorted, dim_name: get_bit_coords(dtype)}, name=v)
        da.attrs = {"long_name": f"{v} bitwise information", "units": 1}
        out[v] = da
    _decorate(out)
    _copy_source_attrs(ds, out)
    return out


# ---------------------------------------------------------------------------------------
# JSON label cache: the on-disk format is the reference's, so caches written by either package load in the
# other (encoder: xbitinfo/xbitinfo.py:947-957; load: 467-476; save: 594-599)
# ---------------------------------------------------------------------------------------
_JSON_CONVERSIONS = (
    ((np.ndarray, np.number), lambda o: o.tolist()),
    (complex, lambda o: [o.real, o.imag]),
    (set, list),
    (bytes, bytes.decode),
)


class JsonCustomEncoder(json.JSONEncoder):
    """numpy arrays / scalars, complex numbers, sets and bytes as plain JSON -- the conversions of the
    reference's encoder of the same name (xbitinfo/xbitinfo.py:947-957)."""

    def default(self, obj):
        for types, convert in _JSON_CONVERSIONS:
            if isinstance(obj, types):
                return convert(obj)
        return super().default(obj)


def load_bitinformation(label, like=None):
    """Dataset from the cache file ``<label>.json``; FileNotFoundError when there is none
    (xbitinfo/xbitinfo.py:467-476)."""
    path = f"{label}.json"
    if not os.path.exists(path):
        raise FileNotFoundError(f"No bitinformation could be found at {path}")
    logging.debug(f"Load bitinformation from {path}")
    with open(path) as f:
        return dict_to_dataset(json.load(f), like=like)


def save_bitinformation(info_per_bit, out_fn, overwrite=False):
    """Write the per-variable result dicts to ``out_fn`` as JSON (xbitinfo/xbitinfo.py:594-599; like the
    reference, an existing file is always replaced)."""
    logging.debug(f"Save bitinformation to {out_fn}")
    with open(out_fn, "w") as f:
        json.dump(info_per_bit, f, cls=JsonCustomEncoder)
