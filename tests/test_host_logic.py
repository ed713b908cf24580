"""Host-side logic of the API mirror, on CPU.  Where a test needs numbers from the CUDA
backend it monkeypatches ``engine.bitinformation`` with the oracle (tests may use the
oracle; the product never does)."""
import json
import os

import numpy as np
import pytest

import xbitinfo_b200 as xb
from oracle import bitinfo_oracle as bo
from oracle import keepbits_oracle as ko
from xbitinfo_b200 import bitinfo as xbi_bitinfo
from xbitinfo_b200 import distributed, engine, keepbits as xbi_keepbits

from conftest import smooth_field
from test_oracle import RH2_INFO


def fake_engine(data, axis, masked_value=None, set_zero_insignificant=False, confidence=0.99, **kw):
    _, _, mi = bo.bitinformation(np.asarray(data), axis, masked_value=masked_value,
                                 set_zero_insignificant_flag=set_zero_insignificant, confidence=confidence)
    return mi


@pytest.fixture
def oracle_backend(monkeypatch):
    monkeypatch.setattr(engine, "bitinformation", fake_engine)
    monkeypatch.setattr(engine, "bitinformation_multi", lambda data, axes, **kw: [fake_engine(data, a, **kw) for a in axes])


def rasm_like(seed=0):
    """float64, (time, y, x), NaN-masked 'ocean' -- the shape family of xarray's rasm."""
    rng = np.random.default_rng(seed)
    t = smooth_field(rng, (6, 20, 28), "float64", axis=0, scale=2.0, base=5.0)
    t[:, rng.random((20, 28)) < 0.3] = np.nan
    da = xb.DataArray(t, dims=("time", "y", "x"), name="Tair", attrs={"units": "C", "long_name": "Surface air temperature"})
    return xb.Dataset({"Tair": da})


def test_bit_coords_and_partitioning():
    assert xb.get_bit_coords("float32") == ["±"] + [f"e{i}" for i in range(1, 9)] + [f"m{i}" for i in range(1, 24)]
    assert xb.get_bit_coords("uint8") == [f"m{i}" for i in range(1, 9)]
    assert xb.get_bit_coords("int16")[0] == "±" and len(xb.get_bit_coords("int16")) == 16
    assert xb.bit_partitioning(np.dtype("float64")) == (64, 1, 11, 52)
    with pytest.raises(ValueError):
        xb.bit_partitioning(np.dtype("complex64"))


def test_keepbits_golden_vector_product():
    da = xb.DataArray(RH2_INFO, dims=["bitfloat32"], coords={"bitfloat32": xb.get_bit_coords("float32"), "dim": "x"})
    ds = xb.Dataset({"RH2": da})
    assert xb.get_keepbits(ds, inflevel=[0.99], information_filter=None, threshold=0.7, tolerance=0.001)["RH2"].values == 19
    assert xb.get_keepbits(ds, inflevel=[0.99], information_filter="Gradient", threshold=0.7, tolerance=0.001)["RH2"].values == 7
    k = xb.get_keepbits(ds, inflevel=[1.0, 0.5, 0.99])
    assert k["RH2"].dims == ("inflevel",) and list(k["RH2"].values) == [23, ko.keepbits_from_info(RH2_INFO, "float32", 0.5), 19]
    assert list(k.coords["inflevel"].values) == [1.0, 0.5, 0.99]
    with pytest.raises(ValueError):
        xb.get_keepbits(ds, inflevel=1.5)


@pytest.mark.parametrize("dtype", ["float16", "float32", "float64", "int16", "uint32"])
def test_keepbits_equals_oracle_on_random_information(dtype):
    rng = np.random.default_rng(3)
    nbits = np.dtype(dtype).itemsize * 8
    for _ in range(200):
        info = np.abs(rng.standard_normal(nbits)) * np.exp(-np.arange(nbits) * rng.uniform(0.05, 0.6))
        info[rng.random(nbits) < 0.2] = 0.0
        da = xb.DataArray(info, dims=[f"bit{dtype}"], coords={f"bit{dtype}": xb.get_bit_coords(dtype), "dim": "x"})
        ds = xb.Dataset({"v": da})
        for lvl in (0.9, 0.99, 0.999999, 1.0):
            assert int(xb.get_keepbits(ds, lvl)["v"].values[0]) == ko.keepbits_from_info(info, dtype, lvl)


def test_cdf_matches_oracle_bitwise():
    rng = np.random.default_rng(4)
    info = np.abs(rng.standard_normal((5, 32))) * np.exp(-0.4 * np.arange(32))
    info[4] = 1e-3  # nothing above the noise threshold: the CDF is NaN everywhere
    got = xbi_keepbits._cdf_from_info_per_bit(info)
    assert not np.isnan(got[:4]).any() and np.isnan(got[4]).all()
    for r in range(5):
        assert np.array_equal(got[r], np.array(ko.cdf_from_info(info[r])), equal_nan=True)


def test_get_bitinformation_structure(oracle_backend):
    ds = rasm_like()
    bi = xb.get_bitinformation(ds, axis=0)
    assert isinstance(bi, xb.Dataset)
    assert bi["Tair"].dims == ("bitfloat64",) and bi["Tair"].shape == (64,)
    assert list(bi.coords["bitfloat64"].values) == xb.get_bit_coords("float64")
    assert bi.coords["dim"].item() == "time"
    assert bi["Tair"].attrs["units"] == 1 and bi["Tair"].attrs["source_units"] == "C"
    assert bi.attrs["xbitinfo_version"] == xb.__version__
    bix = xb.get_bitinformation(ds, dim="x")
    bi2 = xb.get_bitinformation(ds, axis=2)
    assert np.array_equal(bix["Tair"].values, bi2["Tair"].values, equal_nan=True)
    assert not np.array_equal(bi["Tair"].values, bi2["Tair"].values)


def test_get_bitinformation_all_dims_and_missing_dims(oracle_backend):
    ds = rasm_like()
    ds["Tair_mean"] = xb.DataArray(np.nanmean(ds["Tair"].values, axis=0), dims=("y", "x"))
    ds["Tair32"] = ds["Tair"].astype("float32")
    bi = xb.get_bitinformation(ds)
    assert list(bi.coords["dim"].values) == ["time", "x", "y"]  # outer join orders alphabetically
    assert bi["Tair"].dims == ("dim", "bitfloat64") and bi["Tair32"].dims == ("dim", "bitfloat32")
    assert np.isnan(bi["Tair_mean"].sel(dim="time").values).all()
    assert not np.isnan(bi["Tair_mean"].sel(dim="x").values).any()
    bl = xb.get_bitinformation(ds, dim=["x", "y"])
    assert list(bl.coords["dim"].values) == ["x", "y"]
    b1 = xb.get_bitinformation(ds, dim=["x"])
    assert b1["Tair"].dims == ("bitfloat64",) and b1.coords["dim"].item() == "x"
    keep = xb.get_keepbits(bi, inflevel=[0.99, 1.0])
    assert keep["Tair"].dims == ("dim", "inflevel") and keep["Tair"].dtype == np.int64
    assert (keep["Tair"].sel(inflevel=1.0).values == 52).all()


def test_kwargs_defaults_and_errors(oracle_backend, monkeypatch):
    seen = {}

    def spy(data, axis, **kw):
        seen.update(kw)
        return fake_engine(data, axis, **kw)

    monkeypatch.setattr(engine, "bitinformation", spy)
    ds = rasm_like()
    xb.get_bitinformation(ds, dim="x")
    assert np.isnan(seen["masked_value"]) and seen["set_zero_insignificant"] is True and seen["confidence"] == 0.99
    xb.get_bitinformation(ds, dim="x", masked_value="nothing", set_zero_insignificant=False, confidence=0.5)
    assert seen["masked_value"] is None and seen["set_zero_insignificant"] is False and seen["confidence"] == 0.5
    xb.get_bitinformation(ds.astype("int16"), dim="x")
    assert seen["masked_value"] is None
    with pytest.raises(ValueError):
        xb.get_bitinformation(ds, dim="x", axis=2)
    with pytest.raises(ValueError):
        xb.get_bitinformation(ds, dim="x", mask=np.ones(3))
    with pytest.raises(ValueError):
        xb.get_bitinformation(ds, dim=3)
    with pytest.raises(ValueError, match="is unknown"):
        xb.get_bitinformation(ds, dim="x", implementation="fortran")
    with pytest.raises(ImportError):
        xb.get_bitinformation(ds, dim="x", implementation="julia")
    with pytest.raises(TypeError):
        xb.get_bitinformation(ds, dim="x", bogus=1)


def test_label_json_cache(oracle_backend, tmp_path, monkeypatch):
    ds = rasm_like()
    label = str(tmp_path / "rasm")
    first = xb.get_bitinformation(ds, dim="x", label=label)
    assert os.path.exists(label + ".json")
    rec = json.load(open(label + ".json"))
    assert set(rec["Tair"]) == {"bitinfo", "dim", "axis", "dtype"} and rec["Tair"]["axis"] == 2
    monkeypatch.setattr(engine, "bitinformation", lambda *a, **k: (_ for _ in ()).throw(AssertionError("recomputed")))
    again = xb.get_bitinformation(ds, dim="x", label=label)
    assert np.array_equal(first["Tair"].values, again["Tair"].values, equal_nan=True)


def test_quantized_warning(oracle_backend):
    ds = rasm_like()
    ds["Tair"].encoding.update({"scale_factor": 0.01, "add_offset": 0.0, "dtype": np.dtype("int16")})
    with pytest.warns(UserWarning, match="quantized"):
        xb.get_bitinformation(ds, dim="x")


def test_keepbits_interface_types():
    from xbitinfo_b200.bitround import _keepbits_interface

    da = xb.DataArray(np.zeros(3, dtype="float32"), dims=("x",), name="air")
    assert _keepbits_interface(da, 7) == 7
    assert _keepbits_interface(da, {"air": 5}) == 5
    with pytest.raises(ValueError):
        _keepbits_interface(da, {"other": 5})
    kb = xb.Dataset({"air": xb.DataArray(np.array([9]), dims=("inflevel",), coords={"inflevel": [0.99], "dim": "x"})})
    assert _keepbits_interface(da, kb) == 9
    assert _keepbits_interface(da, kb["air"]) == 9
    with pytest.raises(TypeError):
        _keepbits_interface(da, 7.5)


def test_shard_planner():
    # every index is owned exactly once, halo only when the analysed axis is cut
    for shape, axis in [((10, 4, 6), 2), ((3, 100), 1), ((50,), 0), ((8, 8), 0), ((1, 9, 5), 1)]:
        for world in (1, 2, 3, 4, 8):
            plans = [distributed.plan_shard(shape, axis, world, r) for r in range(world)]
            ax = plans[0].split_axis
            assert all(p.split_axis == ax for p in plans)
            owned = sum(p.hi - p.lo for p in plans)
            assert owned == shape[ax] and plans[0].lo == 0 and plans[-1].hi == shape[ax]
            for a, b in zip(plans, plans[1:]):
                assert a.hi == b.lo
            if ax % len(shape) == axis % len(shape):
                assert all(p.halo == (1 if (p.hi < shape[ax] and p.hi > p.lo) else 0) for p in plans)
            else:
                assert all(p.halo == 0 for p in plans)


def test_sharded_counts_add_up_to_global_counts():
    rng = np.random.default_rng(9)
    x = smooth_field(rng, (7, 12, 20), "float32", axis=1, scale=0.4)
    for axis in range(3):
        ref, _, _ = bo.bitinformation(x, axis)
        for world in (2, 3, 5):
            total = np.zeros_like(ref)
            for r in range(world):
                p = distributed.plan_shard(x.shape, axis, world, r)
                sl = [slice(None)] * x.ndim
                sl[p.split_axis] = p.read_slice
                slab = x[tuple(sl)]
                if slab.shape[axis] >= 2:
                    total += bo.bitinformation(slab, axis)[0]
            assert np.array_equal(total, ref), (axis, world)


# ---------------------------------------------------------------------------------------
# host logic of the "next" rows (streaming front end, per-chunk mode, multi-file flow): CPU only
# ---------------------------------------------------------------------------------------
def test_slab_plan_respects_the_sources_chunks():
    from xbitinfo_b200.streaming import _chunk_edges_axis0, is_lazy_array, slab_plan

    edges = [0, 6, 12, 18, 24, 30, 36, 42, 48, 50]
    row = 24 * 36 * 4
    # analysed axis is not the sliced one: slabs end on chunk boundaries, every row exactly once
    plan = slab_plan((50, 24, 36), 2, 4, 13 * row, edges)
    assert plan == [(0, 12, 0), (12, 24, 0), (24, 36, 0), (36, 48, 0), (48, 50, 0)]
    # a chunk larger than a slab is cut
    assert slab_plan((50, 24, 36), 1, 4, 4 * row, [0, 25, 50])[:3] == [(0, 4, 0), (4, 8, 0), (8, 12, 0)]
    # the sliced axis is the analysed one: one halo row per slab, the pairs tile [0, n-1)
    plan0 = slab_plan((50, 24, 36), 0, 4, 13 * row, edges)
    assert all(h == 1 for _, _, h in plan0) and plan0[0][0] == 0 and plan0[-1][1] == 49
    assert all(a[1] == b[0] for a, b in zip(plan0, plan0[1:]))

    class Lazy:
        shape, dtype = (50, 24, 36), np.dtype("float32")

        def __init__(self, chunks):
            self.chunks = chunks

        def __getitem__(self, k):
            raise AssertionError("not read")

    assert is_lazy_array(Lazy(None)) and not is_lazy_array(np.zeros(3)) and not is_lazy_array(3.0)
    assert _chunk_edges_axis0(Lazy(((6,) * 8 + (2,), (24,), (36,)))) == edges  # dask style
    assert _chunk_edges_axis0(Lazy((6, 24, 36))) == edges                      # zarr / h5py style
    assert _chunk_edges_axis0(Lazy(None)) is None


def test_chunk_helpers_of_the_chunking_notebook():
    from xbitinfo_b200.chunked import keepbits_of_information, normalize_chunks, slices_from_chunks

    # the doctest of docs/chunking.ipynb cell 8
    assert slices_from_chunks(((2, 2), (3, 3, 3))) == [
        (slice(0, 2), slice(0, 3)), (slice(0, 2), slice(3, 6)), (slice(0, 2), slice(6, 9)),
        (slice(2, 4), slice(0, 3)), (slice(2, 4), slice(3, 6)), (slice(2, 4), slice(6, 9))]
    assert normalize_chunks((2920, 25, 53), {1: 5, 2: 10}) == ((2920,), (5,) * 5, (10,) * 5 + (3,))
    assert normalize_chunks((4, 9), ((2, 2), 3)) == ((2, 2), (3, 3, 3))
    with pytest.raises(ValueError):
        normalize_chunks((4,), ((3, 2),))
    rng = np.random.default_rng(0)
    for dtype in ("float16", "float32", "float64"):
        nbits = np.dtype(dtype).itemsize * 8
        info = np.sort(rng.random(nbits))[::-1] * (rng.random(nbits) > 0.2)
        for lvl in (0.9, 0.99, 1.0):
            assert keepbits_of_information(info, dtype, lvl) == ko.keepbits_from_info(info, dtype, lvl)


def test_flow_path_selection_and_container_round_trip(tmp_path):
    from xbitinfo_b200 import flow

    paths = [f"f{i}.npz" for i in range(7)]
    assert flow.select_analysis_paths(paths, "first") == ["f0.npz"]
    assert flow.select_analysis_paths(paths, "first_last") == ["f0.npz", "f6.npz"]
    assert flow.select_analysis_paths(paths, "all") == paths
    assert flow.select_analysis_paths(paths, 3) == ["f0.npz", "f3.npz", "f6.npz"]
    for bad in ("last", "middle", 2.5, True):
        with pytest.raises(ValueError):
            flow.select_analysis_paths(paths, bad)
    with pytest.raises(ValueError, match="Please provide paths"):
        flow.xbitinfo_pipeline([])
    ds = xb.Dataset({"a": xb.DataArray(np.arange(24, dtype="float32").reshape(2, 3, 4), dims=("t", "y", "x"), name="a"),
                     "b": xb.DataArray(np.arange(6, dtype="int16").reshape(2, 3), dims=("t", "y"), name="b")})
    p = str(tmp_path / "x.npz")
    flow.save_npz(ds, p)
    back = flow.open_npz(p)
    assert list(back.data_vars) == ["a", "b"] and back["a"].dims == ("t", "y", "x") and back["b"].dims == ("t", "y")
    assert np.array_equal(back["a"].values, ds["a"].values) and back["b"].dtype == np.dtype("int16")


def test_bench_reference_arm_prints_the_contract_line():
    """``bench.py --impl reference`` (the CPU arm the driver runs beside ours) on a few seconds of budget:
    one JSON line with the contract's keys, no GPU needed."""
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "1",
                        "--warmup", "0", "--ref-budget-s", "4"], capture_output=True, text=True, cwd=root, timeout=600)
    assert r.returncode == 0, r.stderr
    line = json.loads(r.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "get_bitinformation_input_throughput"
    assert line["unit"] == "GB/s" and line["higher_is_better"] is True and line["value"] > 0
    assert line["cpu_baseline"]["kind"] == "port" and line["cpu_baseline"]["cores"] >= 1
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert line["config"]["workload"].startswith("configs[1]")


def test_chunk_grid_descriptions():
    """Host logic of the chunk-grid front end: dask-style chunk tuples -> what the batched kernels take."""
    from xbitinfo_b200.chunked import _chunk_spec, keepbits_of_information, keepbits_of_information_batched, normalize_chunks
    from xbitinfo_b200.device import chunk_grid_shape

    grid = normalize_chunks((120, 25, 53), {1: 5, 2: 10})
    assert _chunk_spec(grid) == (120, 5, 10)                      # regular: the last chunk may be shorter
    assert _chunk_spec(((120,), (5, 7, 13), (53,))) == (120, (5, 7, 13), 53)
    assert _chunk_spec(((4, 4, 5),)) == ((4, 4, 5),)              # a longer last chunk is not regular
    assert chunk_grid_shape((120, 25, 53), (120, 5, 10)) == (1, 5, 6)
    assert chunk_grid_shape((120, 25, 53), (120, (5, 7, 13), 53)) == (1, 3, 1)
    assert chunk_grid_shape((0, 4), (2, 2)) == (0, 2)
    assert chunk_grid_shape((7,), (100,)) == (1,)
    with pytest.raises(ValueError):
        chunk_grid_shape((7, 7), (3,))
    with pytest.raises(ValueError):
        chunk_grid_shape((7,), (0,))
    rng = np.random.default_rng(3)
    info = np.abs(rng.standard_normal((4, 5, 32))) ** 4
    info[1, 2] = np.nan                                            # a chunk without valid pairs
    batched = keepbits_of_information_batched(info, np.dtype("float32"), 0.99)
    for idx in np.ndindex(4, 5):
        assert batched[idx] == keepbits_of_information(info[idx], np.dtype("float32"), 0.99)
        if idx != (1, 2):
            assert batched[idx] == ko.keepbits_from_info(info[idx], "float32", 0.99)
    assert (keepbits_of_information_batched(info, np.dtype("float32"), 1.0) == 23).all()
    # what K3 is fed: negatives clamped to 0, but a chunk WITHOUT any valid pair (all-NaN information) is left
    # unrounded with a warning instead of being rounded to 0 mantissa bits
    from xbitinfo_b200.chunked import chunk_keepbits

    assert batched[1, 2] < 0
    with pytest.warns(RuntimeWarning, match="no valid pair"):
        fed = chunk_keepbits(info, np.dtype("float32"), 0.99)
    assert fed[1, 2] == 23 and fed.min() >= 0
    rest = np.ones((4, 5), bool)
    rest[1, 2] = False
    assert np.array_equal(fed[rest], np.clip(batched, 0, 23)[rest])
