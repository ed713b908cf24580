"""``get_keepbits``: number of mantissa bits to keep for a given information level.

Host-side mirror of ``xbitinfo/xbitinfo.py:602-730`` (and the optional "Gradient"
artificial-information filter, ``479-591``).  O(nbits) work per variable, done in numpy on
the host; the result must match the reference exactly (integers).
"""
from __future__ import annotations

import numpy as np

from . import _xr
from .bitinfo import bit_partitioning

_BITDIMS = [
    "bitfloat16", "bitfloat32", "bitfloat64",
    "bitint16", "bitint32", "bitint64",
    "bituint16", "bituint32", "bituint64",
    # narrower integers are accepted by the cuda backend as well
    "bitint8", "bituint8",
]


def _cdf_from_info_per_bit(info: np.ndarray) -> np.ndarray:
    """CDF over the last axis  [720-730]: information not above 1.5 x the maximum of the
    last four bits is treated as rounding noise (dropped), then a NaN-skipping cumulative
    sum normalised by its last value."""
    info = np.asarray(info, dtype="float64")
    with np.errstate(invalid="ignore", divide="ignore"):
        tail_max = _nanmax_keepnan(info[..., -4:])
        cleaned = np.where(info > tail_max[..., None] * 1.5, info, np.nan)
        csum = np.cumsum(np.where(np.isnan(cleaned), 0.0, cleaned), axis=-1)
        return csum / csum[..., -1:]


def _nanmax_keepnan(a):
    allnan = np.isnan(a).all(axis=-1)
    out = np.nanmax(np.where(np.isnan(a), -np.inf, a), axis=-1)
    return np.where(allnan, np.nan, out)


def get_cdf_without_artificial_information(info: np.ndarray, dtype, threshold, tolerance) -> np.ndarray:
    """The "Gradient" filter on a 1-D information vector  [560-590]: cut at the first
    mantissa bit whose CDF increment is below ``tolerance`` once ``threshold`` of the total
    information has been accumulated; normalise the CDF up to there, 1 beyond."""
    info = np.asarray(info, dtype="float64")
    cdf = _cdf_from_info_per_bit(info).copy()
    _, n_sign, n_exponent, _ = bit_partitioning(dtype)
    lead = n_sign + n_exponent
    total = 0.0
    for v in info:  # python sum() order, as in the reference [563]
        total = total + float(v)
    running = 0.0
    for v in info[:lead]:
        running = running + float(v)
    gradient = np.diff(cdf)
    infbits = None
    for i in range(lead, len(gradient) - 1):
        running = running + float(info[i])
        if gradient[i] < tolerance and running >= threshold * total:
            infbits = i
            break
    if infbits is None:
        # the reference leaves `infbits` unbound here (quirk B8) and fails with the same error
        raise UnboundLocalError("no bit satisfies the Gradient filter; cannot cut the CDF")
    for i in range(0, infbits + 1):
        cdf[i] = cdf[i] / cdf[infbits]
    cdf[infbits + 1:] = 1
    return cdf


def get_keepbits(info_per_bit, inflevel=0.99, information_filter=None, **kwargs):
    """Number of mantissa bits to keep per variable (and ``dim``) for each ``inflevel``.

    Same contract as ``xbitinfo.get_keepbits`` [602-717]: returns a Dataset of int64 with an
    ``inflevel`` dimension (after ``dim`` when the information was analysed along several
    dimensions); ``inflevel == 1.0`` keeps all mantissa bits; with
    ``information_filter="Gradient"`` the keywords ``threshold`` and ``tolerance`` are
    required, as in the reference [690-691].
    """
    if not isinstance(inflevel, list):
        inflevel = [inflevel]
    levels = np.asarray(inflevel, dtype="float64")
    if (levels < 0).any() or (levels > 1.0).any():
        raise ValueError("Please provide `inflevel` from interval [0.,1.]")
    ns = _xr.namespace_for(info_per_bit)
    out = ns.Dataset()
    dim_coord = None
    if "dim" in info_per_bit.coords:
        dim_coord = info_per_bit.coords["dim"]
    for bitdim in _BITDIMS:
        bit_vars = [v for v in info_per_bit.data_vars if bitdim in info_per_bit[v].dims]
        if not bit_vars:
            continue
        data_type = np.dtype(bitdim.replace("bit", ""))
        n_bits, _, _, n_mant = bit_partitioning(data_type)
        non_mantissa = n_bits - n_mant
        for v in bit_vars:
            da = info_per_bit[v]
            info = np.asarray(da.values, dtype="float64")
            bit_axis = da.dims.index(bitdim)
            info = np.moveaxis(info, bit_axis, -1)
            other_dims = [d for d in da.dims if d != bitdim]
            if information_filter == "Gradient":
                flat = info.reshape(-1, n_bits)
                cdf = np.stack([
                    get_cdf_without_artificial_information(row, data_type, kwargs["threshold"], kwargs["tolerance"])
                    for row in flat
                ]).reshape(info.shape)
            else:
                cdf = _cdf_from_info_per_bit(info)
            with np.errstate(invalid="ignore"):
                above = cdf[..., None, :] > levels[:, None]  # (..., inflevel, bit)
            keep = above.argmax(axis=-1) + 1 - non_mantissa
            keep = np.where(levels == 1.0, n_bits - non_mantissa, keep).astype("int64")
            coords = {"inflevel": levels}
            if dim_coord is not None:
                coords["dim"] = dim_coord.values if hasattr(dim_coord, "values") else dim_coord
            out[v] = ns.DataArray(keep, dims=other_dims + ["inflevel"], coords=coords, name=v)
    return out
