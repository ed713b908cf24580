"""CPU oracle for get_keepbits -- TEST INFRASTRUCTURE ONLY (see bitinfo_oracle.py header).

Plain-array restatement of ``/root/reference/xbitinfo/xbitinfo.py``:
``bit_partitioning`` (46-67), ``_cdf_from_info_per_bit`` (720-730),
``get_cdf_without_artificial_information`` (479-591) and the keepbits rule of
``get_keepbits`` (696-712).  Written as explicit Python loops on purpose (the product
code in ``xbitinfo_b200/keepbits.py`` is vectorised numpy), so the two do not share an
implementation.  PINNED by the reference's self-contained golden vector
``tests/test_get_keepbits.py:35-138`` (19 without filter, 7 with the Gradient filter).
"""
from __future__ import annotations

import math

import numpy as np


def bit_partitioning(dtype):
    """(n_bits, n_sign, n_exponent, n_mantissa)  -- ``xbitinfo.py:46-67``."""
    dtype = np.dtype(dtype)
    if dtype.kind == "f":
        fi = np.finfo(dtype)
        return fi.bits, 1, fi.nexp, fi.nmant
    if dtype.kind == "i":
        n = np.iinfo(dtype).bits
        return n, 1, 0, n - 1
    if dtype.kind == "u":
        n = np.iinfo(dtype).bits
        return n, 0, 0, n
    raise ValueError(f"dtype {dtype} neither known nor implemented.")


def cdf_from_info(info) -> list:
    """``xbitinfo.py:720-730``: drop (-> NaN) every bit whose information is not above
    1.5 x max(last four bits); NaN-skipping cumulative sum divided by its last entry."""
    info = [float(v) for v in info]
    tail = [v for v in info[-4:] if not math.isnan(v)]
    thr = (max(tail) if tail else float("nan")) * 1.5
    run, acc = [], np.float64(0.0)
    for v in info:
        if v > thr:  # False for NaN, like xarray's where()
            acc = acc + np.float64(v)
        run.append(acc)
    total = run[-1]
    with np.errstate(divide="ignore", invalid="ignore"):
        return [float(np.float64(r) / total) for r in run]


def cdf_gradient_filtered(info, dtype, threshold: float, tolerance: float) -> list:
    """``xbitinfo.py:560-590``: cut the CDF at the first mantissa bit whose CDF increment
    drops below ``tolerance`` once ``threshold`` of the total information has been seen."""
    info = [float(v) for v in info]
    cdf = cdf_from_info(info)
    _, n_sign, n_exp, _ = bit_partitioning(dtype)
    lead = n_sign + n_exp
    total = 0.0
    for v in info:  # python sum(), line 563
        total = total + v
    running = 0.0
    for v in info[:lead]:  # line 570
        running = running + v
    grad = [cdf[i + 1] - cdf[i] for i in range(len(cdf) - 1)]
    cut = None
    for i in range(lead, len(grad) - 1):
        running = running + info[i]
        if grad[i] < tolerance and running >= threshold * total:
            cut = i
            break
    if cut is None:
        raise UnboundLocalError("no bit satisfies the Gradient filter (reference quirk B8)")
    # lines 586-590; note cdf[cut] is overwritten with 1.0 at i == cut, as in the
    # reference, so earlier entries are divided by the ORIGINAL value only while i < cut
    for i in range(0, cut + 1):
        cdf[i] = cdf[i] / cdf[cut]
    for i in range(cut + 1, len(cdf)):
        cdf[i] = 1.0
    return cdf


def keepbits_from_info(info, dtype, inflevel: float = 0.99, information_filter=None,
                       threshold=None, tolerance=None) -> int:
    """``xbitinfo.py:686-712`` for one variable / one inflevel."""
    if inflevel < 0 or inflevel > 1.0:
        raise ValueError("Please provide `inflevel` from interval [0.,1.]")
    n_bits, _, _, n_mant = bit_partitioning(dtype)
    non_mantissa = n_bits - n_mant
    if inflevel == 1.0:
        return n_bits - non_mantissa
    if information_filter == "Gradient":
        cdf = cdf_gradient_filtered(info, dtype, threshold, tolerance)
    else:
        cdf = cdf_from_info(info)
    first = 0
    for i, c in enumerate(cdf):
        if c > inflevel:  # NaN compares False
            first = i
            break
    return first + 1 - non_mantissa
