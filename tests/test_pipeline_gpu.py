"""End to end, after the reference's tests/test_bitinformation_pipeline.py (test_full,
test_full_max_keepbits): get_bitinformation -> get_keepbits -> xr_bitround -> shuffle ->
entropy coder, on synthetic stand-ins for the tutorial datasets the reference downloads
(no network here).  The reference writes netCDF with zlib + shuffle at complevel 9
(xbitinfo/save_compressed.py:44-80); here the byte shuffle is K4 on the device and zlib runs on
the host, which is the same byte stream libhdf5 would deflate for a one-chunk variable."""
import zlib

import numpy as np
import pytest
import torch

import xbitinfo_b200 as xb
from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from oracle import keepbits_oracle as ko
from xbitinfo_b200 import shuffle as xs

from conftest import smooth_field

pytestmark = pytest.mark.gpu


def _dataset(name):
    rng = np.random.default_rng(abs(hash(name)) % 2**32)
    if name == "air_temperature":  # float32 (time, lat, lon), values around 280 K
        x = smooth_field(rng, (240, 25, 53), "float32", axis=2, scale=0.6, base=280.0)
        return xb.Dataset({"air": xb.DataArray(x, dims=("time", "lat", "lon"), name="air")})
    if name == "rasm":  # float64 with a NaN land mask
        x = smooth_field(rng, (36, 60, 80), "float64", axis=2, scale=0.8, base=5.0)
        x[:, rng.random((60, 80)) < 0.35] = np.nan
        return xb.Dataset({"Tair": xb.DataArray(x, dims=("time", "y", "x"), name="Tair")})
    if name == "eraint_uvz":  # configs[0]: three float32 variables on (2, 3, 241, 480)
        ds = xb.Dataset()
        for v, base, scale in (("z", 5.0e4, 30.0), ("u", 3.0, 0.8), ("v", -1.0, 0.6)):
            ds[v] = xb.DataArray(smooth_field(rng, (2, 3, 241, 480), "float32", axis=3, scale=scale, base=base),
                                 dims=("time", "level", "latitude", "longitude"), name=v)
        return ds
    if name == "era52mt":  # a long time series on a small grid
        x = smooth_field(rng, (1500, 9, 12), "float32", axis=0, scale=0.4, base=270.0)
        return xb.Dataset({"t2m": xb.DataArray(x, dims=("time", "latitude", "longitude"), name="t2m")})
    if name == "icon_grid_demo":  # unstructured: (time, ncells)
        x = smooth_field(rng, (4, 20480), "float32", axis=1, scale=0.05, base=1.0e5)
        return xb.Dataset({"pres": xb.DataArray(x, dims=("time", "ncells"), name="pres")})
    raise KeyError(name)


def _deflated_size(arr):
    """bytes after byte shuffle (K4, on the device) + zlib level 9"""
    t = torch.from_numpy(np.ascontiguousarray(arr)).cuda()
    return len(zlib.compress(xs.byte_shuffle(t).cpu().numpy().tobytes(), 9))


@pytest.mark.parametrize("name,dim,axis", [
    ("air_temperature", "lon", None),
    ("rasm", "x", None),
    ("eraint_uvz", "longitude", None),
    ("era52mt", "time", None),
    ("icon_grid_demo", "ncells", None),
    ("air_temperature", None, -1),
])
def test_full(name, dim, axis):
    ds = _dataset(name)
    bitinfo = xb.get_bitinformation(ds, dim=dim, axis=axis, implementation="cuda")
    keepbits = xb.get_keepbits(bitinfo)
    ds_bitrounded = xb.xr_bitround(ds, keepbits)
    for v in ds.data_vars:
        x = ds[v].values
        ax = ds[v].get_axis_num(dim) if dim is not None else axis % x.ndim
        # every stage against the oracle (default keywords of the cuda backend)
        _, _, mi = bo.bitinformation(x, ax, masked_value=float("nan"), set_zero_insignificant_flag=True, confidence=0.99)
        scale = np.nanmax(np.abs(mi))
        np.testing.assert_allclose(bitinfo[v].values, mi, rtol=1e-12, atol=1e-12 * scale)
        k = ko.keepbits_from_info(mi, str(x.dtype), 0.99)
        assert int(keepbits[v].values.reshape(-1)[0]) == k
        u = f"u{x.dtype.itemsize}"
        assert np.array_equal(ds_bitrounded[v].values.view(u), bro.bitround(x, k).view(u))
        assert ds_bitrounded[v].attrs["_QuantizeBitRoundNumberOfSignificantDigits"] == k
        # size reduction, as the reference checks on disk
        ori, comp, rounded = x.nbytes, _deflated_size(x), _deflated_size(ds_bitrounded[v].values)
        assert comp < ori
        assert rounded < comp


def test_full_max_keepbits():
    """bitinformation along every dimension -> keepbits -> the maximum over dims -> bitround."""
    ds = _dataset("air_temperature")
    bi = xb.get_bitinformation(ds, implementation="cuda")
    kb = xb.get_keepbits(bi)
    assert kb["air"].dims == ("dim", "inflevel") and kb["air"].shape == (3, 1)
    kb_max = kb.max(dim="dim")
    want = int(kb["air"].values.max())
    assert int(kb_max["air"].values.reshape(-1)[0]) == want
    rounded = xb.xr_bitround(ds, kb_max)
    x = ds["air"].values
    assert np.array_equal(rounded["air"].values.view("u4"), bro.bitround(x, want).view("u4"))
    planes = xs.BitShuffle(4).encode(rounded["air"].values).reshape(32, -1)
    assert not planes[: 23 - want].any()  # the dropped mantissa bits are empty planes
