"""The public API with implementation="cuda" against the CPU oracle -- written after the
reference's own tests (tests/test_get_bitinformation.py, test_get_keepbits.py,
test_bitround.py), on synthetic stand-ins for the tutorial datasets (no network here)."""
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

import xbitinfo_b200 as xb
from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from oracle import keepbits_oracle as ko
from xbitinfo_b200 import _native as nat
from xbitinfo_b200 import engine
from xbitinfo_b200.device import PairCounter, bitround_

from conftest import smooth_field
from test_counts_gpu import assert_mi_close

pytestmark = pytest.mark.gpu


def rasm_like(dtype="float64"):
    rng = np.random.default_rng(0)
    t = smooth_field(rng, (12, 40, 56), "float64", axis=0, scale=2.0, base=5.0)
    t[:, rng.random((40, 56)) < 0.3] = np.nan
    return xb.Dataset({"Tair": xb.DataArray(t.astype(dtype), dims=("time", "y", "x"), name="Tair", attrs={"units": "C"})})


def eraint_like():
    """BASELINE.json configs[0]: float32 z/u/v of shape (2, 3, 241, 480)."""
    out = xb.Dataset()
    for i, (name, base, scale) in enumerate([("z", 5.0e4, 30.0), ("u", 3.0, 0.8), ("v", -1.0, 0.6)]):
        rng = np.random.default_rng(i)
        x = smooth_field(rng, (2, 3, 241, 480), "float32", axis=-1, scale=scale, base=base)
        out[name] = xb.DataArray(x, dims=("time", "level", "latitude", "longitude"), name=name)
    return out


def test_returns_dataset_and_is_sensitive_to_dim():
    ds = rasm_like()
    b0 = xb.get_bitinformation(ds, axis=0, implementation="cuda")
    b2 = xb.get_bitinformation(ds, axis=2, implementation="cuda")
    assert isinstance(b0, xb.Dataset)
    assert not np.array_equal(b0["Tair"].values, b2["Tair"].values)
    bx = xb.get_bitinformation(ds, dim="x")
    assert np.array_equal(bx["Tair"].values, b2["Tair"].values)


@pytest.mark.parametrize("dtype", ["float64", "float32", "float16", "int16"])
def test_dtype_gives_number_of_bits(dtype):
    ds = rasm_like()
    if dtype == "int16":
        ds = xb.Dataset({"Tair": xb.DataArray(np.nan_to_num(ds["Tair"].values).astype(dtype), dims=("time", "y", "x"))})
    else:
        ds = ds.astype(dtype)
    bi = xb.get_bitinformation(ds, dim="x")
    assert len(bi.coords["bit" + dtype].values) == np.dtype(dtype).itemsize * 8


def test_masked_value_semantics():
    ds = rasm_like()
    x = ds["Tair"].values
    default = xb.get_bitinformation(ds, dim="x", set_zero_insignificant=False)
    nomask = xb.get_bitinformation(ds, dim="x", masked_value="nothing", set_zero_insignificant=False)
    nomask2 = xb.get_bitinformation(ds, dim="x", masked_value=None, set_zero_insignificant=False)
    filled = xb.Dataset({"Tair": xb.DataArray(np.where(np.isnan(x), -999.0, x), dims=("time", "y", "x"))})
    fill = xb.get_bitinformation(filled, dim="x", masked_value=-999.0, set_zero_insignificant=False)
    assert np.array_equal(nomask["Tair"].values, nomask2["Tair"].values)
    # the default masks NaN on the raw values (documented deviation from reference quirk B3)
    assert not np.array_equal(default["Tair"].values, nomask["Tair"].values)
    # masking NaN or masking the fill value removes the same pairs: identical sign/mantissa information
    assert_mi_close(default["Tair"].values, fill["Tair"].values)
    _, _, ref = bo.bitinformation(x, 2, masked_value=float("nan"))
    assert_mi_close(default["Tair"].values, ref)


def test_set_zero_insignificant_and_confidence():
    ds = eraint_like()
    on = xb.get_bitinformation(ds, dim="longitude")  # default True, 0.99
    off = xb.get_bitinformation(ds, dim="longitude", set_zero_insignificant=False)
    c99 = xb.get_bitinformation(ds, dim="longitude", confidence=0.99)
    c50 = xb.get_bitinformation(ds, dim="longitude", confidence=0.5)
    assert np.array_equal(on["u"].values, c99["u"].values)
    assert not np.array_equal(on["u"].values, off["u"].values)
    assert not np.array_equal(c99["u"].values, c50["u"].values)
    assert ((on["u"].values == 0) | (on["u"].values == off["u"].values)).all()


def test_config0_implementations_agree_and_pipeline_is_exact():
    """configs[0]: get_bitinformation(dim='longitude') + get_keepbits(0.99) + xr_bitround,
    cuda vs the oracle of the reference's python path: information within 1e-12 (same
    settings as the reference's own cross-implementation test), keepbits and rounded
    arrays identical."""
    ds = eraint_like()
    bi = xb.get_bitinformation(ds, dim="longitude", implementation="cuda", set_zero_insignificant=False,
                               masked_value=None)
    keep = xb.get_keepbits(bi, 0.99)
    rounded = xb.xr_bitround(ds, keep)
    for v in ds.data_vars:
        x = ds[v].values
        _, _, ref = bo.bitinformation(x, 3, route="unpackbits")
        assert_mi_close(bi[v].values, ref)
        kref = ko.keepbits_from_info(ref, "float32", 0.99)
        assert int(keep[v].values[0]) == kref
        assert rounded[v].attrs["_QuantizeBitRoundNumberOfSignificantDigits"] == kref
        assert np.array_equal(rounded[v].values.view("u4"), bro.bitround(x, kref).view("u4"))
        assert not np.shares_memory(rounded[v].values, x)
        assert (rounded[v].values != x).any()


def test_all_dims_mixed_variables():
    ds = rasm_like()
    ds["Tair_mean"] = xb.DataArray(np.nanmean(ds["Tair"].values, axis=0), dims=("y", "x"))
    ds["Tair32"] = ds["Tair"].astype("float32")
    ds["Tair16"] = ds["Tair"].astype("float16")
    bi = xb.get_bitinformation(ds, set_zero_insignificant=False)
    assert list(bi.coords["dim"].values) == ["time", "x", "y"]
    for bitdim in ("bitfloat16", "bitfloat32", "bitfloat64"):
        assert bitdim in bi.dims
    assert np.isnan(bi["Tair_mean"].sel(dim="time").values).all()
    for d, axis in (("time", 0), ("y", 1), ("x", 2)):
        _, _, ref = bo.bitinformation(ds["Tair16"].values, axis, masked_value=float("nan"))
        assert_mi_close(bi["Tair16"].sel(dim=d).values, ref)


@pytest.mark.parametrize("dtype,keeps", [("float16", range(0, 11)), ("float32", range(0, 24)), ("float64", range(0, 53, 3))])
def test_bitround_bit_exact(dtype, keeps):
    rng = np.random.default_rng(2)
    x = (rng.standard_normal(100003) * np.exp(rng.uniform(-6, 6, 100003))).astype(dtype)
    x[:7] = [0.0, -0.0, np.inf, -np.inf, np.nan, np.finfo(dtype).max, np.finfo(dtype).tiny]
    ut = f"u{np.dtype(dtype).itemsize}"
    for k in keeps:
        got = engine.bitround(x, k)
        assert np.array_equal(got.view(ut), bro.bitround(x, k).view(ut)), (dtype, k)
    # every float16 bit pattern
    if dtype == "float16":
        allp = np.arange(1 << 16, dtype="uint16").view("float16")
        for k in (0, 3, 9):
            assert np.array_equal(engine.bitround(allp, k).view("u2"), bro.bitround(allp, k).view("u2"))


def test_xr_bitround_interface():
    ds = eraint_like().astype("float32")
    for keepbits in (6, dict.fromkeys(ds.data_vars, 6)):
        out = xb.xr_bitround(ds, keepbits)
        for v in ds.data_vars:
            np.testing.assert_allclose(out[v].values, ds[v].values, atol=0.01, rtol=0.01)
            assert out[v].attrs["_QuantizeBitRoundNumberOfSignificantDigits"] == 6
            assert (out[v].values != ds[v].values).any()
    da = xb.xr_bitround(ds["u"], 6)
    assert isinstance(da, xb.DataArray)
    with pytest.raises(ValueError):
        xb.xr_bitround(ds["u"], 24)
    with pytest.raises(ValueError):
        xb.xr_bitround(ds["u"], -1)
    with pytest.raises(TypeError):
        xb.xr_bitround(xb.DataArray(np.arange(5), dims=("x",), name="i"), 3)


def test_device_tensors_unaligned_and_slab_accumulation():
    rng = np.random.default_rng(8)
    x = smooth_field(rng, (64, 96, 200), "float32", axis=-1, scale=0.2)
    ref = {ax: bo.bitinformation(x, ax)[0] for ax in range(3)}
    xt = torch.from_numpy(x).cuda()
    pc = PairCounter(xt.dtype, xt.device)
    for ax in range(3):
        pc.reset(); pc.add(xt, ax)
        assert np.array_equal(pc.counts.cpu().numpy(), ref[ax])
    # unaligned base pointers: flat views starting 1..3 elements into the buffer
    flat = torch.from_numpy(x.reshape(-1)).cuda()
    for off in (1, 2, 3):
        v = flat[off: off + 300000]
        pc.reset(); pc.add(v, 0)
        assert np.array_equal(pc.counts.cpu().numpy(), bo.bitinformation(x.reshape(-1)[off: off + 300000], 0)[0])
    # slabs along a non-analysed axis add up; slabs along the analysed axis need the halo row
    pc.reset()
    for lo in range(0, 64, 16):
        pc.add(xt[lo: lo + 16], 2)
    assert np.array_equal(pc.counts.cpu().numpy(), ref[2])
    pc.reset()
    for lo in range(0, 64, 16):
        hi = min(lo + 16 + 1, 64)  # halo: first row of the next slab
        pc.add(xt[lo:hi], 0)
    assert np.array_equal(pc.counts.cpu().numpy(), ref[0])
    mi = pc.information().cpu().numpy()
    assert_mi_close(mi, bo.mutual_information_from_counts(ref[0], int(ref[0][0].sum())))
    # torch tensors through the public API stay on the device
    ds = xb.Dataset({"t": xb.DataArray(xt, dims=("a", "b", "c"), name="t")})
    bi = xb.get_bitinformation(ds, dim="c", masked_value=None, set_zero_insignificant=False)
    assert_mi_close(bi["t"].values, bo.bitinformation(x, 2)[2])
    y = bitround_(xt.clone(), 9)
    assert np.array_equal(y.cpu().numpy().view("u4"), bro.bitround(x, 9).view("u4"))


def test_host_path_streams_large_arrays_in_slabs(monkeypatch):
    """Force tiny staging buffers so every slab plan of xbi_bitinformation_host runs:
    whole blocks, column strips and row strips with halo."""
    import os
    import subprocess
    import sys

    code = r'''
import sys, numpy as np
sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from conftest import smooth_field
from oracle import bitinfo_oracle as bo
from xbitinfo_b200 import engine
rng = np.random.default_rng(5)
x = smooth_field(rng, (6, 700, 384), "float32", axis=1, scale=0.3)     # 6.4 MB, staging is 1 MB
for ax in range(3):
    mi, c = engine.bitinformation(x, ax, return_counts=True)
    assert np.array_equal(c, bo.bitinformation(x, ax)[0]), ax
y = smooth_field(rng, (3000000,), "float32", axis=0, scale=0.3)          # 12 MB 1-D: row strips with halo
mi, c = engine.bitinformation(y, 0, return_counts=True)
assert np.array_equal(c, bo.bitinformation(y, 0)[0])
z = engine.bitround(y, 5)
from oracle import bitround_oracle as bro
assert np.array_equal(z.view("u4"), bro.bitround(y, 5).view("u4"))
print("ok")
'''
    env = dict(os.environ, XBI_STAGING_MB="1")
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("env_extra", [{}, {"XBI_HOST_STAGING": "0"}, {"XBI_HOST_THREADS": "3"}])
def test_pageable_and_pinned_host_buffers_give_identical_results(env_extra):
    """Pageable numpy memory goes through the pinned ring and the copy threads (csrc/xbi_hostpool.cu),
    page-locked memory straight to the DMA engine; contiguous slabs, pitched column strips (4 MB staging
    buffer) and bitround into a fresh / the same / a pinned array must all agree with the oracle."""
    code = r'''
import sys
import numpy as np, torch
sys.path.insert(0, "tests")
from conftest import smooth_field
from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from xbitinfo_b200 import engine
rng = np.random.default_rng(21)
x = smooth_field(rng, (3, 2000, 1024), "float32", axis=1, scale=0.3)     # 24.6 MB, staging is 4 MB
pinned = torch.empty(x.shape, dtype=torch.float32, pin_memory=True).numpy()
pinned[...] = x
for ax in range(3):
    want = bo.bitinformation(x, ax)[0]
    for arr in (x, pinned):
        mi, c = engine.bitinformation(arr, ax, return_counts=True)
        assert np.array_equal(c, want), ax
y = smooth_field(rng, (11_000_003,), "float32", axis=0, scale=0.3)       # 44 MB: several pinned slots, ragged end
want = bro.bitround(y, 6).view("u4")
assert np.array_equal(engine.bitround(y, 6).view("u4"), want)
yp = torch.empty(y.shape, dtype=torch.float32, pin_memory=True).numpy()
yp[...] = y
assert np.array_equal(engine.bitround(yp, 6).view("u4"), want)
# the C entry point allows src == dst
import ctypes as C
from xbitinfo_b200 import _native as nat
z = y.copy()
nat.check(nat.lib().xbi_bitround_host(z.ctypes.data_as(C.c_void_p), z.ctypes.data_as(C.c_void_p), nat.F32, z.size, 6, 0))
assert np.array_equal(z.view("u4"), want)
out_pinned = torch.empty(y.shape, dtype=torch.float32, pin_memory=True).numpy()
nat.check(nat.lib().xbi_bitround_host(y.ctypes.data_as(C.c_void_p), out_pinned.ctypes.data_as(C.c_void_p), nat.F32, y.size, 6, 0))
assert np.array_equal(out_pinned.view("u4"), want)
d = smooth_field(rng, (5_000_001,), "float64", axis=0, scale=0.3)
assert np.array_equal(engine.bitround(d, 20).view("u8"), bro.bitround(d, 20).view("u8"))
print("ok")
'''
    env = dict(os.environ, XBI_STAGING_MB="4", **env_extra)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.parametrize("staging_mb", ["1", "256"])
def test_all_axes_in_one_pass_over_the_host_array(staging_mb):
    """xbi_bitinformation_host_multi (every slab on the device once for all the analysed axes; one halo row
    for the first dimension) gives exactly the per-axis results and the oracle's; 1 MB staging = many slabs."""
    code = r'''
import sys
import numpy as np
sys.path.insert(0, "tests")
from conftest import smooth_field
from oracle import bitinfo_oracle as bo
from xbitinfo_b200 import engine
rng = np.random.default_rng(31)
cases = [
    (smooth_field(rng, (41, 30, 50, 24), "float32", axis=1, scale=0.3), None),
    (smooth_field(rng, (300, 65, 33), "float64", axis=0, scale=0.3), float("nan")),
    (smooth_field(rng, (2, 700, 512), "float16", axis=1, scale=0.05, base=3.0), None),   # one row > half the 1 MB buffer
    (rng.integers(-900, 900, size=(57, 120, 40), dtype="int16"), 7),
    (smooth_field(rng, (100000,), "float32", axis=0, scale=0.3), None),
]
cases[1][0][rng.random(cases[1][0].shape) < 0.1] = np.nan
for x, mv in cases:
    axes = list(range(x.ndim))[::-1]
    got = engine.bitinformation_multi(x, axes + axes[:1], masked_value=mv, set_zero_insignificant=True, return_counts=True)
    assert len(got) == x.ndim + 1
    for a, (mi, c) in zip(axes + axes[:1], got):
        want_c, _, want_mi = bo.bitinformation(x, a, masked_value=mv, set_zero_insignificant_flag=True)
        assert np.array_equal(c, want_c), (x.shape, a)
        mi1, c1 = engine.bitinformation(x, a, masked_value=mv, set_zero_insignificant=True, return_counts=True)
        assert np.array_equal(c, c1) and np.array_equal(mi, mi1, equal_nan=True), (x.shape, a)
print("ok")
'''
    env = dict(os.environ, XBI_STAGING_MB=staging_mb)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True,
                       cwd=os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    assert r.returncode == 0 and "ok" in r.stdout, r.stdout + r.stderr


def test_get_bitinformation_all_dims_equals_dim_by_dim():
    ds = eraint_like()
    every = xb.get_bitinformation(ds)
    for d in every.coords["dim"].values:
        one = xb.get_bitinformation(ds, dim=str(d))
        for v in ds.data_vars:
            assert np.array_equal(every[v].sel(dim=d).values, one[v].values, equal_nan=True), (v, d)


def test_host_entry_points_from_several_threads():
    """``xr.apply_ufunc(..., dask="parallelized")`` calls the bitround boundary from several worker threads
    (xbitinfo/bitround.py:99): the host entry points keep their staging state per thread and share the copy
    pool; concurrent calls must give the single-threaded results."""
    from concurrent.futures import ThreadPoolExecutor

    rng = np.random.default_rng(41)
    blocks = [smooth_field(rng, (700_001 + 50_000 * i,), "float32", axis=0, scale=0.3) for i in range(8)]
    fields = [smooth_field(rng, (9 + i, 300, 200), "float32", axis=1, scale=0.3) for i in range(8)]
    want_round = [bro.bitround(b, 3 + i).view("u4") for i, b in enumerate(blocks)]
    want_counts = [bo.bitinformation(f, 1)[0] for f in fields]

    def work(i):
        r = engine.bitround(blocks[i], 3 + i).view("u4")
        _, c = engine.bitinformation(fields[i], 1, return_counts=True)
        multi = engine.bitinformation_multi(fields[i], [1, 2], return_counts=True)
        return np.array_equal(r, want_round[i]) and np.array_equal(c, want_counts[i]) and np.array_equal(multi[0][1], want_counts[i])

    with ThreadPoolExecutor(max_workers=4) as ex:
        for _ in range(3):
            assert all(ex.map(work, range(8)))


def test_launch_counters_move():
    before = nat.kernel_launches()
    engine.bitinformation(np.zeros((8, 64), dtype="float32"), 1)
    assert nat.kernel_launches() > before


def test_streamed_from_memmap_equals_in_memory(tmp_path):
    """Row f1: slab streaming from a file-backed array (np.memmap), incl. the halo row when
    the sliced axis is the analysed one, equals the in-memory result exactly."""
    from xbitinfo_b200.streaming import bitinformation_streamed, slab_plan

    rng = np.random.default_rng(17)
    x = smooth_field(rng, (301, 48, 64), "float32", axis=0, scale=0.4)
    x[rng.random(x.shape) < 0.05] = np.nan
    path = tmp_path / "field.dat"
    x.tofile(path)
    mm = np.memmap(path, dtype="float32", mode="r", shape=x.shape)
    assert len(slab_plan(x.shape, 0, 4, 200_000)) > 10
    for axis in (0, 1, 2):
        for mv in (None, float("nan")):
            mi_ref, c_ref = engine.bitinformation(x, axis, masked_value=mv, return_counts=True)
            mi, c = bitinformation_streamed(mm, axis, slab_bytes=200_000, masked_value=mv, return_counts=True)
            assert np.array_equal(c, c_ref), (axis, mv)
            assert_mi_close(mi, mi_ref)
    y = x[:, 0, 0].copy()  # 1-D
    _, c = bitinformation_streamed(y, 0, slab_bytes=64, return_counts=True)
    assert np.array_equal(c, bo.bitinformation(y, 0)[0])


def test_bitround_along_dim():
    from xbitinfo_b200.bitround import bitround_along_dim

    ds = eraint_like()
    info = xb.get_bitinformation(ds, dim="longitude")
    out = bitround_along_dim(ds, info, dim="longitude", inflevels=[1.0, 0.9999, 0.99, 0.975])
    assert out["u"].shape == ds["u"].shape and out["u"].dtype == ds["u"].dtype
    assert np.array_equal(out["u"].values[..., :120], ds["u"].values[..., :120])  # inflevel 1: untouched
    assert (out["u"].values[..., 120:] != ds["u"].values[..., 120:]).any()
    assert abs(float(np.mean(ds["u"].values - out["u"].values))) < 0.05
    out2 = bitround_along_dim(ds, info, dim="longitude", inflevels=None, keepbits=2)
    assert out2["u"].shape == ds["u"].shape
    with pytest.raises(ValueError):
        bitround_along_dim(ds, info, dim="latitude", keepbits=2, inflevels=[1.0, 0.99])
    with pytest.raises(ValueError):
        bitround_along_dim(ds, info, dim="latitude", inflevels=None)


class _LazyChunked:
    """A stand-in for a dask / zarr array: sliceable, has ``chunks``, counts what is read."""

    def __init__(self, arr, chunk_rows):
        self._a = arr
        self.shape, self.dtype = arr.shape, arr.dtype
        self.chunks = (chunk_rows,) + arr.shape[1:]
        self.rows_read = 0
        self.reads = []

    def __getitem__(self, key):
        out = self._a[key]
        if isinstance(key, slice):
            self.rows_read += out.shape[0]
            self.reads.append((key.start, key.stop))
        return out


def test_lazy_arrays_are_streamed_on_chunk_boundaries(monkeypatch):
    """f1: a lazy variable goes through get_bitinformation without being materialised: slabs
    along the first axis, cut on the source's chunk boundaries, each row read once (plus one
    halo row per slab when the first axis is the analysed one)."""
    from xbitinfo_b200 import streaming

    rng = np.random.default_rng(8)
    x = smooth_field(rng, (50, 24, 36), "float32", axis=2, scale=0.4, base=280.0)
    x[:, rng.random((24, 36)) < 0.2] = np.nan
    orig = streaming.bitinformation_streamed
    monkeypatch.setattr(streaming, "bitinformation_streamed",
                        lambda src, axis, **kw: orig(src, axis, slab_bytes=13 * 24 * 36 * 4, **kw))
    for dim, axis in (("x", 2), ("time", 0)):
        lazy = _LazyChunked(x, chunk_rows=6)
        ds = xb.Dataset({"T": xb.DataArray(lazy, dims=("time", "y", "x"), name="T")})
        got = xb.get_bitinformation(ds, dim=dim, implementation="cuda")
        ref = xb.get_bitinformation(xb.Dataset({"T": xb.DataArray(x, dims=("time", "y", "x"), name="T")}), dim=dim)
        assert np.array_equal(got["T"].values, ref["T"].values)
        if axis != 0:
            assert lazy.rows_read == 50
            assert all(lo % 6 == 0 for lo, _ in lazy.reads)  # 12-row slabs: two whole chunks each
        else:
            assert lazy.rows_read == 50 + len(lazy.reads) - 1  # one halo row per slab but the last


def test_lazy_arrays_are_read_once_for_all_dimensions():
    """``get_bitinformation(ds)`` over every dimension of a lazy variable reads (decodes) the source once:
    slabs on its chunk boundaries, one halo row per slab because the first dimension is analysed too."""
    from xbitinfo_b200.streaming import bitinformation_streamed_multi

    rng = np.random.default_rng(18)
    x = smooth_field(rng, (50, 24, 36), "float32", axis=2, scale=0.4, base=280.0)
    x[:, rng.random((24, 36)) < 0.2] = np.nan
    lazy = _LazyChunked(x, chunk_rows=6)
    got = bitinformation_streamed_multi(lazy, [0, 1, 2], slab_bytes=13 * 24 * 36 * 4, masked_value=float("nan"),
                                        return_counts=True)
    assert lazy.rows_read == 50 + len(lazy.reads) - 1 and len(lazy.reads) >= 4
    assert all(lo % 6 == 0 for lo, _ in lazy.reads)
    for a, (mi, c) in enumerate(got):
        want_c, _, want_mi = bo.bitinformation(x, a, masked_value=float("nan"))
        assert np.array_equal(c, want_c), a
        assert_mi_close(mi, want_mi)
    # without the first axis no halo is read; a subset in any order
    lazy = _LazyChunked(x, chunk_rows=7)
    got = bitinformation_streamed_multi(lazy, [2, 1], slab_bytes=10 * 24 * 36 * 4, return_counts=True)
    assert lazy.rows_read == 50
    assert np.array_equal(got[0][1], bo.bitinformation(x, 2)[0]) and np.array_equal(got[1][1], bo.bitinformation(x, 1)[0])
    # one-row slabs, 2-d, first axis of length 1 (no pairs along it)
    y = smooth_field(rng, (9, 40), "float64", axis=1, scale=0.3)
    got = bitinformation_streamed_multi(_LazyChunked(y, chunk_rows=4), [0, 1], slab_bytes=1, return_counts=True)
    assert np.array_equal(got[0][1], bo.bitinformation(y, 0)[0]) and np.array_equal(got[1][1], bo.bitinformation(y, 1)[0])
    z = y[:1]
    got = bitinformation_streamed_multi(_LazyChunked(z, chunk_rows=1), [0, 1], return_counts=True)
    assert int(np.abs(got[0][1]).sum()) == 0 and np.array_equal(got[1][1], bo.bitinformation(z, 1)[0])
    # the public API takes this route for dim=None
    lazy = _LazyChunked(x, chunk_rows=6)
    ds = xb.Dataset({"T": xb.DataArray(lazy, dims=("time", "y", "x"), name="T")})
    every = xb.get_bitinformation(ds, implementation="cuda")
    ref = xb.get_bitinformation(xb.Dataset({"T": xb.DataArray(x, dims=("time", "y", "x"), name="T")}))
    assert np.array_equal(every["T"].values, ref["T"].values, equal_nan=True)
    assert lazy.rows_read <= 50 + len(lazy.reads)


def test_bitround_streamed_between_two_files(tmp_path):
    """Out-of-core rounding: file-backed input and output (np.memmap), slab by slab."""
    from xbitinfo_b200.streaming import bitround_streamed

    rng = np.random.default_rng(19)
    x = smooth_field(rng, (101, 40, 30), "float32", axis=0, scale=0.4)
    x.tofile(tmp_path / "in.dat")
    src = np.memmap(tmp_path / "in.dat", dtype="float32", mode="r", shape=x.shape)
    dst = np.memmap(tmp_path / "out.dat", dtype="float32", mode="w+", shape=x.shape)
    lazy = _LazyChunked(src, chunk_rows=8)
    bitround_streamed(lazy, 7, dst, slab_bytes=20 * 40 * 30 * 4)
    dst.flush()
    assert lazy.rows_read == 101 and all(lo % 8 == 0 for lo, _ in lazy.reads) and len(lazy.reads) == 7
    got = np.fromfile(tmp_path / "out.dat", dtype="float32").reshape(x.shape)
    assert np.array_equal(got.view("u4"), bro.bitround(x, 7).view("u4"))
    y = smooth_field(rng, (1000,), "float64", axis=0)
    out = np.empty_like(y)
    assert bitround_streamed(y, 30, out, slab_bytes=800) is out
    assert np.array_equal(out.view("u8"), bro.bitround(y, 30).view("u8"))
    with pytest.raises(ValueError):
        bitround_streamed(y, 30, np.empty(999))
    with pytest.raises(TypeError):
        bitround_streamed(np.arange(10), 3, np.empty(10, dtype="int64"))


def test_bitround_by_chunks_equals_the_per_block_loop_of_the_chunking_notebook():
    from xbitinfo_b200.chunked import bitround_by_chunks, slices_from_chunks, normalize_chunks

    rng = np.random.default_rng(9)
    x = smooth_field(rng, (120, 25, 53), "float32", axis=1, scale=0.3, base=280.0)
    x[:, :10] *= 1e-3  # a region with a different bit spectrum
    chunks = {1: 5, 2: 10}
    rounded, keep = bitround_by_chunks(x, chunks, axis=1, inflevel=0.99, masked_value=None, set_zero_insignificant=False)
    grid = normalize_chunks(x.shape, chunks)
    assert keep.shape == (1, 5, 6) and grid[2][-1] == 3
    want = np.empty_like(x)
    for idx, sl in zip(np.ndindex(*keep.shape), slices_from_chunks(grid)):
        _, _, mi = bo.bitinformation(x[sl], 1)
        k = max(0, ko.keepbits_from_info(mi, "float32", 0.99))
        assert keep[idx] == k, idx
        want[sl] = bro.bitround(np.ascontiguousarray(x[sl]), k)
    assert np.array_equal(rounded.view("u4"), want.view("u4"))
    assert len(np.unique(keep)) > 1
    # device tensors stay on the device
    t = torch.from_numpy(x).cuda()
    r2, k2 = bitround_by_chunks(t, chunks, axis=1, inflevel=0.99, masked_value=None, set_zero_insignificant=False)
    assert r2.is_cuda and np.array_equal(r2.cpu().numpy().view("u4"), want.view("u4")) and np.array_equal(k2, keep)
