"""CPU oracle for xr_bitround's arithmetic -- TEST INFRASTRUCTURE ONLY.

The reference delegates the arithmetic to a third-party package that is absent from
/root/reference and from this image: ``numcodecs.bitround.BitRound`` (pinned
``numcodecs>=0.10.0,<0.16.4`` in ``pyproject.toml:14``; call site
``xbitinfo/bitround.py:7-12``: encode() then decode() on a copy of the data).

PARITY UNPINNED against numcodecs itself.  What is restated here is the algorithm
numcodecs documents for BitRound (round to nearest, ties to even, on the integer view
of the IEEE bits; M. Kloewer's BitInformation.jl ``round!``): with ``maskbits =
mantissa_bits - keepbits``, add ``((b >> maskbits) & 1) + (2**(maskbits-1) - 1)`` to the
integer view ``b`` and clear the low ``maskbits`` bits.  The reference's own tests only
check closeness (``tests/test_bitround.py:13-45``) and equality with Julia's ``round!``
(``67-78``, needs Julia).  The oracle is cross-checked in tests/ against an independent
formulation (exact rational round-half-even of the significand) on every finite
float16 and on random float32/float64.
"""
from __future__ import annotations

import numpy as np

MANTISSA_BITS = {"float16": 10, "float32": 23, "float64": 52}


def bitround(data: np.ndarray, keepbits: int) -> np.ndarray:
    """New array with ``keepbits`` explicit mantissa bits kept (RNE on the bit pattern)."""
    if keepbits < 0:
        raise ValueError("keepbits must be zero or positive")
    data = np.asarray(data)
    if data.dtype.kind != "f" or data.dtype.itemsize > 8:
        raise TypeError("Only float arrays (16-64bit) can be bit-rounded")
    mbits = MANTISSA_BITS[data.dtype.name]
    if keepbits == mbits:
        return data.copy()
    if keepbits > mbits:
        raise ValueError("keepbits too large for given dtype")
    width = data.dtype.itemsize * 8
    utype = np.dtype(f"u{data.dtype.itemsize}")
    b = data.copy().view(utype)
    drop = mbits - keepbits
    full = (1 << width) - 1
    keepmask = utype.type((full >> drop) << drop)
    half_minus_one = utype.type((1 << (drop - 1)) - 1)
    tie = (b >> utype.type(drop)) & utype.type(1)
    with np.errstate(over="ignore"):
        b = b + tie + half_minus_one  # wraps like the signed add in the codec
    b = b & keepmask
    return b.view(data.dtype)


def bitround_float_rne(data: np.ndarray, keepbits: int) -> np.ndarray:
    """Independent check (no integer tricks): do the rounding in float64 arithmetic.
    For a finite float16/float32 value x with unbiased exponent e (clamped to the
    subnormal exponent of its dtype) the kept grid spacing is q = 2**(e - keepbits);
    the result is rint(x / q) * q (``np.rint`` rounds half to even; every step is exact
    in float64 for these two dtypes), cast back (overflow -> inf, as in the bit
    algorithm).  Valid for finite inputs and keepbits >= 1."""
    data = np.asarray(data)
    assert data.dtype.name in ("float16", "float32")
    emin = {"float16": -14, "float32": -126}[data.dtype.name]
    x = data.astype(np.float64)
    _, e = np.frexp(x)  # x = m * 2**e with 0.5 <= |m| < 1  ->  unbiased exponent e-1
    e = np.maximum(e - 1, emin)
    q = np.ldexp(1.0, e - keepbits)
    with np.errstate(over="ignore"):
        r = (np.rint(x / q) * q).astype(data.dtype)
    # rint(-0.3) = -0.0 keeps the sign, as the bit algorithm does
    return r
