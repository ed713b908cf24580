#!/usr/bin/env python
"""Generate the golden vectors under tests/golden/ by executing the REFERENCE's own
``xbitinfo/_py_bitinfo.py`` (read from /root/reference, never copied).

The reference package cannot be imported in this image (xarray, dask, numcodecs, zarr and
matplotlib are missing), but the hot-path module ``_py_bitinfo.py``
only needs three things from dask.array: ``da.ma.masked_equal`` and, on its array
arguments, ``.flatten().map_blocks(np.unpackbits, ...)``.  This script installs a
20-line stand-in for ``dask.array`` (numpy arrays with a ``map_blocks`` method that
applies the function to the single in-memory block, which is exactly what dask does for a
one-chunk array) and then loads the reference file by path.  Every array below is
therefore produced by the reference's own arithmetic, op for op.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

Outputs (committed): tests/golden/py_bitinfo_golden.npz
"""
import importlib.util
import os
import sys
import types

import numpy as np

REF = "/root/reference/xbitinfo/_py_bitinfo.py"
HERE = os.path.dirname(os.path.abspath(__file__))


class _OneBlock(np.ndarray):
    """numpy array that answers the two dask-array calls the reference makes."""

    def map_blocks(self, func, *args, drop_axis=None, meta=None, chunks=None, **kw):
        return np.asarray(func(np.asarray(self), *args, **kw)).view(_OneBlock)

    def compute(self):
        return np.asarray(self)

    def __array_function__(self, func, types, args, kwargs):
        # dask arrays stay dask arrays through np.broadcast_arrays (_py_bitinfo.py:133)
        if func is np.broadcast_arrays:
            kwargs = dict(kwargs, subok=True)
        return super().__array_function__(func, types, args, kwargs)


def _install_dask_shim():
    dask = types.ModuleType("dask")
    darr = types.ModuleType("dask.array")
    dma = types.ModuleType("dask.array.ma")
    dma.masked_equal = np.ma.masked_equal
    darr.ma = dma
    darr.array = lambda x: np.asarray(x).view(_OneBlock)
    dask.array = darr
    sys.modules["dask"] = dask
    sys.modules["dask.array"] = darr
    sys.modules["dask.array.ma"] = dma


def load_reference_module():
    _install_dask_shim()
    spec = importlib.util.spec_from_file_location("ref_py_bitinfo", REF)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def smooth_field(rng, shape, dtype, axis=-1, scale=1.0, base=280.0):
    """Correlated synthetic data: base + cumulative sum of noise along ``axis``."""
    x = base + np.cumsum(rng.standard_normal(shape) * scale, axis=axis)
    return x.astype(dtype)


def reference_bitinformation(pb, x, axis):
    """What xbitinfo.xbitinfo._py_get_bitinformation (xbitinfo.py:347-366) does, using the
    reference module ``pb`` for every arithmetic step."""
    X = np.asarray(x).view(_OneBlock)
    if X.dtype in (np.float16, np.float32, np.float64):
        X = pb.signed_exponent(X)
    X = np.asarray(X).astype(f"u{x.dtype.itemsize}").view(_OneBlock)
    sa = tuple(slice(0, -1) if i == axis else slice(None) for i in range(X.ndim))
    sb = tuple(slice(1, None) if i == axis else slice(None) for i in range(X.ndim))
    a, b = X[sa], X[sb]
    counts = np.asarray(pb.bitpaircount(a, b))
    mi = pb.mutual_information(a, b)
    mi = np.ma.filled(np.ma.asarray(mi), np.nan).astype("float64")
    return np.asarray(X), counts.astype("int64"), mi


def main():
    pb = load_reference_module()
    rng = np.random.default_rng(20261019)
    out = {}

    # --- signed_exponent: special values + random bit patterns, all three float widths
    for dt, ut in (("float16", "uint16"), ("float32", "uint32"), ("float64", "uint64")):
        nb = np.dtype(ut).itemsize * 8
        if dt == "float16":
            bits = np.arange(1 << 16, dtype=ut)  # every float16 pattern
        else:
            bits = rng.integers(0, 1 << 63, size=4096, dtype="uint64").astype(ut)
            if dt == "float64":
                bits = bits | (rng.integers(0, 2, size=4096, dtype="uint64") << np.uint64(63))
            special = np.array(
                [0.0, -0.0, 1.0, -1.0, 0.03125, 2.0, 0.5, np.inf, -np.inf, np.nan, -999.0,
                 np.finfo(dt).tiny, np.finfo(dt).max, np.finfo(dt).smallest_subnormal, 1.5, 3.0],
                dtype=dt,
            ).view(ut)
            bits = np.concatenate([special, bits])
        A = bits.view(dt).view(_OneBlock)
        B = np.asarray(pb.signed_exponent(A)).astype(ut)
        out[f"sexp_{dt}_in"] = bits
        out[f"sexp_{dt}_out"] = B
        assert B.dtype.itemsize * 8 == nb

    # --- bitinformation cases: (name, array, axis)
    cases = []
    cases.append(("f32_3d_last", smooth_field(rng, (3, 17, 40), "float32", axis=-1, scale=0.7), 2))
    cases.append(("f32_3d_mid", smooth_field(rng, (5, 23, 16), "float32", axis=1, scale=0.3), 1))
    cases.append(("f32_3d_first", smooth_field(rng, (19, 6, 12), "float32", axis=0, scale=2.0), 0))
    cases.append(("f64_2d_last", smooth_field(rng, (11, 50), "float64", axis=-1, scale=0.05, base=1.0e3), 1))
    cases.append(("f64_2d_first", smooth_field(rng, (37, 9), "float64", axis=0, scale=5.0, base=-3.0), 0))
    cases.append(("f16_2d_last", smooth_field(rng, (13, 64), "float16", axis=-1, scale=0.1, base=2.0), 1))
    cases.append(("f16_3d_first", smooth_field(rng, (21, 4, 10), "float16", axis=0, scale=0.5, base=10.0), 0))
    cases.append(("f32_1d", smooth_field(rng, (777,), "float32", axis=0, scale=1.0), 0))
    # mixed-sign small numbers incl. zeros, subnormals, inf and nan (python path has no mask)
    odd = rng.standard_normal((6, 33)).astype("float32") * np.float32(1e-3)
    odd[0, :4] = [0.0, -0.0, np.inf, -np.inf]
    odd[1, 5] = np.nan
    odd[2, 7:9] = np.finfo("float32").smallest_subnormal
    cases.append(("f32_specials", odd, 1))
    cases.append(("i16_2d", (smooth_field(rng, (9, 41), "float64", axis=-1, scale=30.0, base=0.0)).astype("int16"), 1))
    cases.append(("u8_2d", rng.integers(0, 256, size=(7, 29), dtype="uint8"), 1))
    cases.append(("i32_3d", (smooth_field(rng, (4, 5, 31), "float64", axis=-1, scale=1e5, base=0.0)).astype("int32"), 2))
    cases.append(("i64_2d", (smooth_field(rng, (5, 27), "float64", axis=0, scale=1e12, base=0.0)).astype("int64"), 0))
    cases.append(("f32_const", np.full((4, 9), 1.25, dtype="float32"), 1))
    cases.append(("f32_n1", smooth_field(rng, (5, 1), "float32", axis=0), 1))  # no pairs along the axis
    # one eraint_uvz-shaped variable of BASELINE.json configs[0], reduced in the two outer dims
    cases.append(("f32_eraint_like", smooth_field(rng, (1, 2, 241, 480), "float32", axis=-1, scale=3.0, base=5.0e4), 3))

    names = []
    for name, x, axis in cases:
        _, counts, mi = reference_bitinformation(pb, x, axis)
        out[f"case_{name}_x"] = x
        out[f"case_{name}_axis"] = np.int64(axis)
        out[f"case_{name}_counts"] = counts
        out[f"case_{name}_mi"] = mi
        names.append(name)
        print(f"{name:18s} shape={x.shape} axis={axis} dtype={x.dtype} mi[:4]={mi[:4]}")
    out["case_names"] = np.array(names)

    # --- mutual_information rounding pins: random 2x2 tables straight into the formula
    #     (reference lines 147-151) via a fake pair of arrays is not possible without data,
    #     so pin it through bitpaircount on tiny random uint8 arrays of varying length.
    tabs_a, tabs_b, tabs_mi = [], [], []
    for i in range(64):
        nn = int(rng.integers(2, 300))
        a = rng.integers(0, 256, size=nn, dtype="uint8").view(_OneBlock)
        b = rng.integers(0, 256, size=nn, dtype="uint8").view(_OneBlock)
        mi = np.ma.filled(np.ma.asarray(pb.mutual_information(a, b)), np.nan)
        tabs_a.append(np.pad(np.asarray(a), (0, 300 - nn)))
        tabs_b.append(np.pad(np.asarray(b), (0, 300 - nn)))
        tabs_mi.append(np.concatenate([[nn], mi]))
    out["mi_pairs_a"] = np.stack(tabs_a)
    out["mi_pairs_b"] = np.stack(tabs_b)
    out["mi_pairs_n_and_mi"] = np.stack(tabs_mi)

    path = os.path.join(HERE, "py_bitinfo_golden.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
