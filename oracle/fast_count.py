"""ctypes front end of oracle/fast_count.c -- TEST INFRASTRUCTURE ONLY.

Builds the C restatement with gcc on first use (``oracle/_build/libfastcount.so``; ``__graft_entry__.build()``
builds it ahead of time so that it travels to the GPU box) and exposes ``count_pairs(x, axis, masked_value)``
with the return layout of ``bitinfo_oracle.bitinformation`` (int64 ``(nbits, 2, 2)``)."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(_HERE, "fast_count.c")
LIB = os.path.join(_HERE, "_build", "libfastcount.so")
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    gcc = shutil.which("gcc")
    if gcc is None:
        raise RuntimeError("gcc not found: cannot build oracle/fast_count.c")
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    cmd = [gcc, "-O3", "-march=x86-64-v2", "-mpopcnt", "-fopenmp", "-shared", "-fPIC", "-std=c99", "-Wall", "-Wextra", SRC, "-o", LIB]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("gcc failed on oracle/fast_count.c:\n" + r.stderr)
    return LIB


def lib():
    global _lib
    if _lib is None:
        L = C.CDLL(build())
        L.fast_count_pairs.restype = C.c_int
        L.fast_count_pairs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_uint64,
                                       C.c_void_p, C.c_int]
        _lib = L
    return _lib


def count_pairs(x: np.ndarray, axis: int, masked_value=None, threads: int = 0) -> np.ndarray:
    """(nbits, 2, 2) int64 pair counts of ``x`` along ``axis`` (all host cores by default)."""
    x = np.ascontiguousarray(x)
    if x.ndim == 0:
        raise ValueError("cannot analyse a 0-d array")
    axis %= x.ndim
    outer = int(np.prod(x.shape[:axis], dtype=np.int64))
    n = int(x.shape[axis])
    inner = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    is_float = x.dtype.kind == "f"
    mode, bits = 0, 0
    if masked_value is not None:
        if is_float and isinstance(masked_value, float) and masked_value != masked_value:
            mode = 1
        else:
            mode = 2
            bits = int(np.array(masked_value).astype(x.dtype).view(f"u{x.dtype.itemsize}"))
    nbits = 8 * x.dtype.itemsize
    counts = np.zeros(nbits * 4, dtype=np.uint64)
    rc = lib().fast_count_pairs(x.ctypes.data_as(C.c_void_p), x.dtype.itemsize, int(is_float), outer, n, inner, mode, bits,
                                counts.ctypes.data_as(C.c_void_p), int(threads))
    if rc != 0:
        raise ValueError("fast_count_pairs rejected its arguments")
    return counts.reshape(nbits, 2, 2).astype(np.int64)
