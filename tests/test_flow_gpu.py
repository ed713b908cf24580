"""The multi-file flow (f3) after the reference's get_prefect_flow tests
(tests/test_bitinformation_pipeline.py:68-135): n files of identical structure, analysis on a
subset, every file rounded with the same keepbits and saved compressed."""
import os

import numpy as np
import pytest

import xbitinfo_b200 as xb
from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from oracle import keepbits_oracle as ko
from xbitinfo_b200 import flow

from conftest import smooth_field

pytestmark = pytest.mark.gpu
IMAX = 3


@pytest.fixture()
def flow_paths(tmp_path):
    rng = np.random.default_rng(2)
    x = smooth_field(rng, (36, 40, 56), "float32", axis=0, scale=0.7, base=5.0)
    x[:, rng.random((40, 56)) < 0.3] = np.nan
    stride = x.shape[0] // IMAX
    paths, parts = [], []
    for i in range(IMAX):
        part = x[stride * i: stride * (i + 1)]
        p = str(tmp_path / f"file_{i}.npz")
        flow.save_npz(xb.Dataset({"Tair": xb.DataArray(part, dims=("time", "y", "x"), name="Tair")}), p)
        paths.append(p)
        parts.append(part)
    return paths, parts


def _expected_keepbits(parts, axis):
    x = np.concatenate(parts, axis=0)
    _, _, mi = bo.bitinformation(x, axis, masked_value=float("nan"), set_zero_insignificant_flag=True, confidence=0.99)
    return max(0, ko.keepbits_from_info(mi, "float32", 0.99))


@pytest.mark.parametrize("analyse_paths,pick", [("first", [0]), ("first_last", [0, 2]), ("all", [0, 1, 2]), (2, [0, 2])])
def test_flow_rounds_every_file_with_the_keepbits_of_the_analysis_subset(flow_paths, analyse_paths, pick):
    paths, parts = flow_paths
    keepbits, done = flow.xbitinfo_pipeline(paths, analyse_paths=analyse_paths, axis=0, overwrite=True)
    k = _expected_keepbits([parts[i] for i in pick], 0)
    assert keepbits == {"Tair": k}
    assert len(done) == IMAX
    for new_path, part, path in zip(done, parts, paths):
        assert new_path == path.replace(".npz", "_bitrounded_compressed.xbz") and os.path.exists(new_path)
        got = flow.open_compressed(new_path)["Tair"]
        assert np.array_equal(got.values.view("u4"), bro.bitround(part, k).view("u4"))
        assert got.attrs["_QuantizeBitRoundNumberOfSignificantDigits"] == k
        assert os.path.getsize(new_path) < part.nbytes


def test_flow_dim_by_name_overwrite_and_errors(flow_paths):
    paths, parts = flow_paths
    kb_x, done = flow.xbitinfo_pipeline(paths, dim="x", inflevel=0.9999)
    _, _, mi = bo.bitinformation(parts[0], 2, masked_value=float("nan"), set_zero_insignificant_flag=True)
    assert kb_x == {"Tair": max(0, ko.keepbits_from_info(mi, "float32", 0.9999))}
    stamp = {p: os.path.getmtime(p) for p in done}
    _, again = flow.xbitinfo_pipeline(paths, dim="x", inflevel=0.9999)  # complete outputs exist: skipped
    assert again == [None] * IMAX and {p: os.path.getmtime(p) for p in done} == stamp
    with open(done[1], "wb") as f:  # a truncated output is deleted and recalculated
        f.write(b"garbage")
    _, redo = flow.xbitinfo_pipeline(paths, dim="x", inflevel=0.9999)
    assert redo == [None, done[1], None]
    assert np.array_equal(flow.open_compressed(done[1])["Tair"].values.view("u4"),
                          bro.bitround(parts[1], kb_x["Tair"]).view("u4"))
    with pytest.raises(ValueError, match="Please provide paths"):
        flow.xbitinfo_pipeline([])
    with pytest.raises(ValueError, match="analyse_paths"):
        flow.xbitinfo_pipeline(paths, analyse_paths="middle")


def test_negative_keepbits_are_clamped_unless_asked_otherwise(tmp_path):
    x = np.full((8, 16, 16), 2.0, dtype="float32")
    x[:, ::2] = -2.0  # only the sign carries information
    p = str(tmp_path / "const.npz")
    flow.save_npz(xb.Dataset({"c": xb.DataArray(x, dims=("t", "y", "x"), name="c")}), p)
    kb = flow.get_bitinformation_keepbits([p], dim="y", non_negative_keepbits=False)
    assert kb["c"] < 0
    assert flow.get_bitinformation_keepbits([p], dim="y")["c"] == 0
    _, done = flow.xbitinfo_pipeline([p], dim="y")
    assert np.array_equal(flow.open_compressed(done[0])["c"].values, x)  # 0 keepbits keep a power of two
