"""CPU oracle for the bit-information hot path -- TEST INFRASTRUCTURE ONLY.

This file restates, in plain numpy, the algorithm of the reference's python
implementation (``/root/reference/xbitinfo/_py_bitinfo.py`` and the calling glue in
``xbitinfo/xbitinfo.py``).  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product
package ``xbitinfo_b200`` never does (it fails loudly when the CUDA library is missing).

Parity status
-------------
* ``signed_exponent``, pair counts and mutual information are PINNED: they reproduce,
  bit for bit, the outputs of the reference's own ``_py_bitinfo.py`` executed in the
  build container (``tests/golden/make_golden.py`` -> ``tests/golden/py_bitinfo_golden.npz``)
  and the reference's doctest known answers (``_py_bitinfo.py:12-15,28-31,57-66``).
* ``masked_value`` exclusion and ``set_zero_insignificant``/``confidence`` are
  "parity unpinned": the python path of the reference rejects them
  (``xbitinfo.py:338-346``) and the Julia package that defines them
  (BitInformation.jl v0.6.3, pinned at ``julia_helpers.py:113-117``) is neither vendored nor
  runnable here.  Their published algorithm is restated in ``paircount_masked`` and
  ``free_entropy_threshold`` below (see DESIGN.md "Semantics defined only in Julia").

Two evaluation routes are provided for the counts:
* ``paircount_unpackbits`` -- op-for-op the reference route (shift, cast to u1,
  ``np.unpackbits``, ``(a<<1)|b``, compare with 0..3, sum).  ~90 B of temporaries per
  pair and byte plane: use for small arrays and as the timed CPU baseline.
* ``paircount_planes`` -- closed form per bit plane from three population counts
  (n1x, nx1, n11); identical integers, feasible up to ~2**26 elements.
"""
from __future__ import annotations

import math

import numpy as np

FLOAT_LAYOUT = {  # dtype name -> (exponent bits, explicit mantissa bits)
    "float16": (5, 10),
    "float32": (8, 23),
    "float64": (11, 52),
}


# --------------------------------------------------------------------------------------
# exponent helpers  (reference: _py_bitinfo.py:6-19, 22-39)
# --------------------------------------------------------------------------------------
def exponent_bias(dtype) -> int:
    """2**(E-1) - 1 with E the number of exponent bits (``_py_bitinfo.py:6-19``)."""
    fi = np.finfo(dtype)
    return 2 ** (fi.bits - fi.nmant - 2) - 1


def exponent_mask(dtype) -> int:
    """Bit mask of the exponent field (``_py_bitinfo.py:22-39``)."""
    e, m = FLOAT_LAYOUT[np.dtype(dtype).name]
    return ((1 << e) - 1) << m


# --------------------------------------------------------------------------------------
# signed exponent  (reference: _py_bitinfo.py:42-94)
# --------------------------------------------------------------------------------------
def signed_exponent(A: np.ndarray) -> np.ndarray:
    """Biased -> sign-magnitude exponent, following ``_py_bitinfo.py:68-94`` step by step
    (sign+mantissa kept, e = field - bias, |e| modulo the field range, sign flag in the
    field's top bit).  Returns the unsigned view the caller makes at ``xbitinfo.py:355``."""
    A = np.asarray(A)
    ebits, mbits = FLOAT_LAYOUT[A.dtype.name]
    width = A.dtype.itemsize * 8
    utype = np.dtype(f"u{A.dtype.itemsize}")
    raw = A.view(utype).astype(np.uint64)  # widen: keeps every step exact for all widths
    top = 1 << (width - 1)
    keep = np.uint64(top | ((1 << mbits) - 1))  # sign bit and mantissa (line 73)
    field = np.uint64(exponent_mask(A.dtype))  # line 74
    flag = np.uint64(top >> 1)  # line 75
    e = ((raw & field) >> np.uint64(mbits)).astype(np.int64) - exponent_bias(A.dtype)  # line 85
    span = (int(np.iinfo(utype).max) >> mbits) + 1  # lines 86-87
    mag = (np.abs(e) % span).astype(np.uint64)
    neg = np.where(e < 0, flag, np.uint64(0))  # line 88
    out = (raw & keep) | neg | (mag << np.uint64(mbits))  # lines 92-93
    return out.astype(utype)


def signed_exponent_closed_form(A: np.ndarray) -> np.ndarray:
    """Branch-free form the CUDA kernels use: field F -> F-bias if F >= bias else ~F.
    Must equal :func:`signed_exponent` on every bit pattern (tested exhaustively for f16)."""
    A = np.asarray(A)
    ebits, mbits = FLOAT_LAYOUT[A.dtype.name]
    utype = np.dtype(f"u{A.dtype.itemsize}")
    raw = A.view(utype)
    field = utype.type(exponent_mask(A.dtype))
    one = utype.type(exponent_bias(A.dtype) << mbits)
    return np.where((raw & field) >= one, raw - one, raw ^ field).astype(utype)


def to_analysis_words(x: np.ndarray) -> np.ndarray:
    """``xbitinfo.py:349-355``: floats get the signed-exponent transform, every dtype is
    then reinterpreted as the unsigned integer of the same width."""
    x = np.asarray(x)
    if x.dtype.kind == "f":
        return signed_exponent(x)
    if x.dtype.kind in "iu":
        return x.astype(f"u{x.dtype.itemsize}")
    raise ValueError(f"dtype {x.dtype} not supported")


# --------------------------------------------------------------------------------------
# pair selection  (reference: _py_bitinfo.py:155-160)
# --------------------------------------------------------------------------------------
def adjacent_views(X: np.ndarray, axis: int):
    lead = [slice(None)] * X.ndim
    lag = [slice(None)] * X.ndim
    lead[axis] = slice(0, -1)
    lag[axis] = slice(1, None)
    return X[tuple(lead)], X[tuple(lag)]


# --------------------------------------------------------------------------------------
# counting, route 1: the reference's unpackbits route (_py_bitinfo.py:97-140)
# --------------------------------------------------------------------------------------
def _paircount_byteplane(a_u1: np.ndarray, b_u1: np.ndarray) -> np.ndarray:
    """(8,2,2) joint histogram of one byte plane (``_py_bitinfo.py:97-125``)."""
    ua = np.unpackbits(a_u1.reshape(-1))
    ub = np.unpackbits(b_u1.reshape(-1))
    code = ((ua << 1) | ub).reshape(-1, 8)
    onehot = code[..., None] == np.arange(4, dtype="u1")
    return onehot.sum(axis=0).reshape(8, 2, 2)


def paircount_unpackbits(a: np.ndarray, b: np.ndarray) -> np.ndarray:
    """(nbits,2,2) int64 counts, most significant bit first (``_py_bitinfo.py:128-140``)."""
    assert a.dtype.kind == "u" and b.dtype.kind == "u"
    nbytes = a.dtype.itemsize
    a, b = np.broadcast_arrays(a, b)
    planes = []
    for i in range(nbytes):
        s = (nbytes - 1 - i) * 8
        planes.append(_paircount_byteplane((a >> s).astype("u1"), (b >> s).astype("u1")))
    return np.concatenate(planes, axis=0).astype(np.int64)


# --------------------------------------------------------------------------------------
# counting, route 2: closed form from plane population counts (same integers)
# --------------------------------------------------------------------------------------
def paircount_planes(a: np.ndarray, b: np.ndarray, valid: np.ndarray | None = None) -> np.ndarray:
    """Same (nbits,2,2) table as :func:`paircount_unpackbits`, from
    n1x = #(a_k=1), nx1 = #(b_k=1), n11 = #(a_k=b_k=1) per bit k and the pair count N:
    [[N-n1x-nx1+n11, nx1-n11], [n1x-n11, n11]].  ``valid`` (bool, same shape) restricts
    the pair set (used for the masked_value semantics)."""
    a = np.ascontiguousarray(a).reshape(-1)
    b = np.ascontiguousarray(b).reshape(-1)
    if valid is not None:
        v = np.ascontiguousarray(valid).reshape(-1)
        a, b = a[v], b[v]
    nbits = a.dtype.itemsize * 8
    N = a.size
    ab = a & b
    out = np.zeros((nbits, 2, 2), dtype=np.int64)
    one = a.dtype.type(1)
    for k in range(nbits):
        sh = a.dtype.type(nbits - 1 - k)
        n1x = int(np.count_nonzero((a >> sh) & one))
        nx1 = int(np.count_nonzero((b >> sh) & one))
        n11 = int(np.count_nonzero((ab >> sh) & one))
        out[k] = [[N - n1x - nx1 + n11, nx1 - n11], [n1x - n11, n11]]
    return out


# --------------------------------------------------------------------------------------
# mutual information  (reference: _py_bitinfo.py:143-152)
# --------------------------------------------------------------------------------------
def mutual_information_from_counts(counts: np.ndarray, size: int) -> np.ndarray:
    """Per-bit mutual information in bits from the (nbits,2,2) table, same numpy
    masked-array expression, operand order and reduction order as ``_py_bitinfo.py:147-151``.
    ``size == 0`` (no pairs) gives NaN for every bit, like the reference."""
    with np.errstate(divide="ignore", invalid="ignore"):
        p = counts.astype("float") / size
        p = np.ma.masked_equal(p, 0)
        pr = p.sum(axis=-1)[..., np.newaxis]
        ps = p.sum(axis=-2)[..., np.newaxis, :]
        mi = (p * np.ma.log(p / (pr * ps))).sum(axis=(-1, -2)) / np.log(2)
    return np.ma.filled(np.ma.asarray(mi), np.nan).astype(np.float64)


def mutual_information_scalar(c00: int, c01: int, c10: int, c11: int) -> float:
    """The same value written as the scalar recipe the device epilogue (K2) follows:
    p=c/N; pr=p_a0+p_a1; ps=p_0b+p_1b; t=p*log(p/(pr*ps)) for p>0;
    (((t00+t01)+t10)+t11)/log(2).  Uses numpy's float64 ``log`` so it can be compared
    bit-for-bit with :func:`mutual_information_from_counts`."""
    n = c00 + c01 + c10 + c11
    if n == 0:
        return float("nan")
    f = np.float64
    p = [[f(c00) / f(n), f(c01) / f(n)], [f(c10) / f(n), f(c11) / f(n)]]
    pr = [p[0][0] + p[0][1], p[1][0] + p[1][1]]
    ps = [p[0][0] + p[1][0], p[0][1] + p[1][1]]
    acc = None
    for a_ in range(2):
        for b_ in range(2):
            if p[a_][b_] == 0:
                continue
            t = p[a_][b_] * np.log(p[a_][b_] / (pr[a_] * ps[b_]))
            acc = t if acc is None else acc + t
    if acc is None:
        return float("nan")
    return float(acc / np.log(2))


# --------------------------------------------------------------------------------------
# masked pairs + significance filter (BitInformation.jl v0.6.3 semantics; parity unpinned)
# --------------------------------------------------------------------------------------
def masked_elements(x: np.ndarray, masked_value) -> np.ndarray | None:
    """Element mask on the RAW values.  ``None``/"nothing" -> no mask; NaN -> every NaN
    payload; any other value -> bit-identity with that value cast to x.dtype."""
    if masked_value is None or (isinstance(masked_value, str) and masked_value == "nothing"):
        return None
    x = np.asarray(x)
    if x.dtype.kind == "f" and isinstance(masked_value, float) and math.isnan(masked_value):
        return np.isnan(x)
    mv = np.array(masked_value).astype(x.dtype)
    ut = f"u{x.dtype.itemsize}"
    return x.view(ut) == mv.view(ut)


def binom_confidence_z(confidence: float) -> float:
    """Two-sided normal quantile z = Phi^-1(1 - (1-c)/2), evaluated with the stdlib
    (``statistics.NormalDist``) so the oracle does not share code with the product."""
    from statistics import NormalDist

    return NormalDist().inv_cdf(1.0 - (1.0 - confidence) / 2.0)


def free_entropy_threshold(npairs: int, confidence: float = 0.99) -> float:
    """Hfree = 1 - H2(p), p = min(1, 0.5 + z/(2*sqrt(N))) (BitInformation.jl
    ``binom_confidence`` / ``binom_free_entropy``)."""
    z = binom_confidence_z(confidence)
    p = min(1.0, 0.5 + z / (2.0 * math.sqrt(npairs)))
    if p >= 1.0:
        return 1.0
    h = -(p * math.log2(p) + (1.0 - p) * math.log2(1.0 - p))
    return 1.0 - h


def set_zero_insignificant(mi: np.ndarray, npairs: int, confidence: float = 0.99) -> np.ndarray:
    thr = free_entropy_threshold(npairs, confidence)
    out = np.array(mi, dtype=np.float64, copy=True)
    out[out <= thr] = 0.0
    return out


# --------------------------------------------------------------------------------------
# whole path for one variable
# --------------------------------------------------------------------------------------
def bitinformation(
    x: np.ndarray,
    axis: int,
    masked_value=None,
    set_zero_insignificant_flag: bool = False,
    confidence: float = 0.99,
    route: str = "planes",
):
    """counts (nbits,2,2) int64, number of pairs, information per bit (float64).

    With ``masked_value=None`` and ``set_zero_insignificant_flag=False`` this is exactly
    the reference's python path (``xbitinfo.py:337-370`` + ``_py_bitinfo.py``)."""
    x = np.asarray(x)
    axis = axis % x.ndim if x.ndim else 0
    X = to_analysis_words(x)
    a, b = adjacent_views(X, axis)
    mask = masked_elements(x, masked_value)
    if mask is None:
        npairs = a.size
        counts = paircount_unpackbits(a, b) if route == "unpackbits" else paircount_planes(a, b)
    else:
        ma, mb = adjacent_views(mask, axis)
        valid = ~(ma | mb)
        npairs = int(np.count_nonzero(valid))
        counts = paircount_planes(a, b, valid)
    mi = mutual_information_from_counts(counts, npairs)
    if set_zero_insignificant_flag and npairs > 0:
        mi = set_zero_insignificant(mi, npairs, confidence)
    return counts, npairs, mi
