"""The batched per-chunk kernels (csrc/xbi_chunked.cu: xbi_count_pairs_chunked,
xbi_mutual_information_batched, xbi_bitround_chunked) against the CPU oracle applied to every
chunk on its own -- the per-block loop of the reference's docs/chunking.ipynb (cells 8-10).
Counts and rounded arrays bit-exact, information within 1e-12."""
import numpy as np
import pytest
import torch

from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from oracle import keepbits_oracle as ko
from xbitinfo_b200.chunked import bitinformation_by_chunks, bitround_by_chunks, normalize_chunks, slices_from_chunks
from xbitinfo_b200.device import bitround_chunked_, count_pairs_chunked, information_batched

from conftest import smooth_field
from test_counts_gpu import assert_mi_close

pytestmark = pytest.mark.gpu


def oracle_per_chunk(x, chunk_shape, axis, masked_value=None, set_zero=False):
    grid = normalize_chunks(x.shape, tuple(chunk_shape))
    shape = tuple(len(c) for c in grid)
    nbits = x.dtype.itemsize * 8
    counts = np.zeros(shape + (nbits, 2, 2), dtype=np.int64)
    mi = np.full(shape + (nbits,), np.nan)
    for idx, sl in zip(np.ndindex(*shape), slices_from_chunks(grid)):
        counts[idx], _, mi[idx] = bo.bitinformation(x[sl], axis, masked_value=masked_value,
                                                    set_zero_insignificant_flag=set_zero)
    return counts, mi


def check(x, chunk_shape, axis, masked_value=None, set_zero=False):
    t = torch.from_numpy(x).cuda()
    got = count_pairs_chunked(t, chunk_shape, axis, masked_value=masked_value)
    want, want_mi = oracle_per_chunk(x, chunk_shape, axis, masked_value, set_zero)
    assert tuple(got.shape) == want.shape
    assert np.array_equal(got.cpu().numpy(), want)
    # both chunk-grid kernels, whatever the planner would pick for this grid: the walkers (one load per element,
    # xbi_chunked_walk.cu) and the per-element odometers (xbi_chunked.cu)
    for flavour in ("walk", "walk2", "walk3", "odometer"):
        pinned = count_pairs_chunked(t, chunk_shape, axis, masked_value=masked_value, flavour=flavour)
        assert np.array_equal(pinned.cpu().numpy(), want), flavour
    mi = information_batched(got, set_zero, 0.99).cpu().numpy()
    for idx in np.ndindex(*want.shape[:-3]):
        assert_mi_close(mi[idx], want_mi[idx])


CASES = [
    # shape, chunk_shape, axis
    ((120, 25, 53), (120, 5, 10), 1),        # the chunking notebook's grid (ragged last chunk along lon)
    ((120, 25, 53), (120, 5, 10), 0),
    ((120, 25, 53), (120, 5, 10), 2),
    ((120, 25, 53), (7, 25, 53), 0),          # chunked along the analysed axis itself
    ((37, 41), (8, 16), 1),
    ((37, 41), (8, 16), 0),
    ((1000,), (64,), 0),
    ((5000,), (5000,), 0),                    # one chunk
    ((6, 4, 5, 8), (2, 4, 5, 8), 3),          # unchunked neighbours merge
    ((6, 4, 5, 8), (2, 4, 5, 8), 1),          # ... but not across the analysed axis
    ((6, 4, 5, 8), (6, 4, 5, 3), 0),
    ((3, 1, 7, 1, 9), (2, 1, 3, 1, 4), 2),    # dimensions of length one vanish
    ((3, 4, 5, 6, 7), (2, 3, 2, 4, 3), 2),    # five chunked dimensions
    ((2, 3, 4, 5, 6, 7), (1, 3, 4, 2, 6, 7), 3),  # six dimensions, three groups after merging
    ((9, 2), (4, 2), 1),                      # two elements along the axis: one pair per row
    # rows made of whole 16-byte vectors: the walkers take 16 bytes per lane along the outer axes
    ((40, 24, 64), (40, 8, 16), 0),
    ((40, 24, 64), (10, 24, 32), 1),
    ((33, 6, 72), (33, 4, 20), 0),             # ... ragged edge chunk (12 columns) and a single row group left over
    ((19, 5, 3, 16), (19, 2, 3, 8), 0),        # ... inner dimensions flattened into the vector columns
    ((40, 24, 64), (40, 8, 16), 2),            # the same grid along the innermost axis (lane-major ring)
    # irregular grids (dask-style tuples of chunk lengths), alone and mixed with regular dimensions
    ((120, 25, 53), (120, (5, 7, 13), (53,)), 1),
    ((120, 25, 53), ((100, 20), (5, 7, 13), (1, 30, 2, 20)), 2),
    ((120, 25, 53), ((1, 119), 25, (50, 3)), 0),
    ((6, 4, 5, 8), ((1, 2, 3), 4, 5, 8), 3),   # an irregular dimension takes the unchunked ones inside it
    ((6, 4, 5, 8), (6, (3, 1), 5, (2, 6)), 2),
    ((1000,), ((1, 2, 3, 994),), 0),
    ((9, 1), (4, 1), 1),                      # no pair at all
]


@pytest.mark.parametrize("shape,chunk_shape,axis", CASES)
def test_counts_and_information_per_chunk_float32(shape, chunk_shape, axis):
    rng = np.random.default_rng(len(shape) * 100 + axis)
    x = smooth_field(rng, shape, "float32", axis=axis, scale=0.3)
    check(x, chunk_shape, axis)


@pytest.mark.parametrize("dtype", ["float64", "float16", "int8", "uint16", "int32", "uint64"])
def test_counts_per_chunk_every_width(dtype):
    rng = np.random.default_rng(11)
    if np.dtype(dtype).kind == "f":
        x = smooth_field(rng, (40, 23, 37), dtype, axis=1, scale=0.3, base=3.0)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=(40, 23, 37), dtype=dtype, endpoint=True)
    check(x, (16, 8, 10), 1)
    check(x, (40, 23, 5), 2)
    # rows of whole 16-byte vectors for every width (8-byte types: two columns per lane, 4-byte types: four)
    y = np.ascontiguousarray(np.concatenate([x, x[:, :, :3]], axis=2))  # (40, 23, 40)
    check(y, (40, 8, 8), 0)
    check(y, (10, 23, 16), 1)


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16"])
@pytest.mark.parametrize("masked_value", [float("nan"), -999.0])
def test_masked_pairs_per_chunk(dtype, masked_value):
    rng = np.random.default_rng(5)
    x = smooth_field(rng, (30, 26, 44), dtype, axis=2, scale=0.3, base=3.0)
    x[:, 5:12, 7:30] = masked_value            # a "land" rectangle covering whole chunks and cutting others
    x[rng.random(x.shape) < 0.02] = masked_value
    check(x, (30, 6, 11), 2, masked_value=masked_value, set_zero=True)
    check(x, (10, 26, 44), 1, masked_value=masked_value)


def test_integer_mask_value_per_chunk():
    rng = np.random.default_rng(6)
    x = rng.integers(0, 50, size=(20, 33, 18), dtype="int16")
    check(x, (20, 8, 6), 1, masked_value=7)


def test_chunks_of_many_tiles_and_groups():
    """Chunks far larger than one 4096-element tile: several tiles per work item, several items per chunk,
    ragged edge chunks with fewer tiles than the full ones."""
    rng = np.random.default_rng(7)
    x = smooth_field(rng, (96, 200, 300), "float32", axis=1, scale=0.3)
    check(x, (96, 128, 256), 1)
    check(x, (96, 128, 256), 2)
    check(x, (96, 200, 300), 0)      # a single 5.8 M-element chunk split over many CTAs
    x = smooth_field(rng, (400, 64, 64), "float64", axis=0, scale=0.3)
    check(x, (400, 16, 16), 0)       # many chunks of 25 tiles each


def test_bitround_chunked_against_the_oracle():
    rng = np.random.default_rng(8)
    for dtype, mant in (("float32", 23), ("float64", 52), ("float16", 10)):
        x = smooth_field(rng, (50, 25, 53), dtype, axis=1, scale=0.3, base=3.0)
        chunk_shape = (50, 5, 10)
        grid = normalize_chunks(x.shape, chunk_shape)
        shape = tuple(len(c) for c in grid)
        keep = rng.integers(0, mant, size=shape, endpoint=True).astype("int32")
        keep.flat[0] = mant  # identity
        keep.flat[1] = 0
        t = torch.from_numpy(x).cuda()
        bitround_chunked_(t, chunk_shape, torch.from_numpy(keep.ravel()).cuda())
        want = np.empty_like(x)
        for idx, sl in zip(np.ndindex(*shape), slices_from_chunks(grid)):
            want[sl] = bro.bitround(np.ascontiguousarray(x[sl]), int(keep[idx]))
        u = f"u{x.dtype.itemsize}"
        assert np.array_equal(t.cpu().numpy().view(u), want.view(u)), dtype


def test_batched_route_equals_chunk_by_chunk_route_and_the_oracle():
    rng = np.random.default_rng(9)
    x = smooth_field(rng, (120, 25, 53), "float32", axis=1, scale=0.3, base=280.0)
    x[:, :10] *= 1e-3
    x[:, 20:, 40:] = np.nan
    chunks = {1: 5, 2: 10}
    r1, k1 = bitround_by_chunks(x, chunks, axis=1, batched=True)
    r2, k2 = bitround_by_chunks(x, chunks, axis=1, batched=False)
    assert np.array_equal(k1, k2) and np.array_equal(r1.view("u4"), r2.view("u4"))
    grid = normalize_chunks(x.shape, chunks)
    for idx, sl in zip(np.ndindex(*k1.shape), slices_from_chunks(grid)):
        _, n, mi = bo.bitinformation(x[sl], 1, masked_value=np.nan, set_zero_insignificant_flag=True)
        if n:
            assert k1[idx] == min(23, max(0, ko.keepbits_from_info(mi, "float32", 0.99))), idx
    assert len(np.unique(k1)) > 1
    info, counts = bitinformation_by_chunks(x, chunks, axis=1)
    assert tuple(info.shape) == k1.shape + (32,) and tuple(counts.shape) == k1.shape + (32, 2, 2)
    # irregular grids go through the same kernels (offset tables instead of a chunk length)
    for irregular in (((120,), (5, 7, 13), (53,)), ((7, 100, 13), (25,), (20, 33))):
        r3, k3 = bitround_by_chunks(x, irregular, axis=1)
        r4, k4 = bitround_by_chunks(x, irregular, axis=1, batched=False)
        assert k3.shape == tuple(len(c) for c in irregular)
        assert np.array_equal(k3, k4) and np.array_equal(r3.view("u4"), r4.view("u4"))


def test_labelled_front_end_of_the_chunking_notebook():
    import xbitinfo_b200 as xb
    from xbitinfo_b200.chunked import xr_bitround_by_chunks

    rng = np.random.default_rng(13)
    air = smooth_field(rng, (60, 25, 53), "float32", axis=1, scale=0.3, base=280.0)
    air[:, :10] *= 1e-3
    ds = xb.Dataset({
        "air": xb.DataArray(air, dims=("time", "lat", "lon"), name="air", attrs={"units": "K"}),
        "orog": xb.DataArray(smooth_field(rng, (25, 53), "float32", axis=0), dims=("lat", "lon"), name="orog"),
        "time_bnds": xb.DataArray(np.arange(60, dtype="float64"), dims=("time",), name="time_bnds"),
    })
    rounded, keep = xr_bitround_by_chunks(ds, {"lat": 5, "lon": (10, 10, 33)}, "lat", masked_value=None,
                                          set_zero_insignificant=False)
    assert keep["air"].dims == ("time", "lat", "lon") and keep["air"].shape == (1, 5, 3)
    assert keep["orog"].shape == (5, 3) and "time_bnds" not in keep.data_vars
    assert rounded["air"].attrs["units"] == "K"
    assert np.array_equal(rounded["time_bnds"].values, ds["time_bnds"].values)
    want, kwant = bitround_by_chunks(air, {1: 5, 2: (10, 10, 33)}, 1, masked_value=None, set_zero_insignificant=False)
    assert np.array_equal(rounded["air"].values.view("u4"), want.view("u4")) and np.array_equal(keep["air"].values, kwant)
    assert len(np.unique(keep["air"].values)) > 1
    # the input is untouched
    assert np.array_equal(ds["air"].values, air)


def test_launch_count_does_not_grow_with_the_number_of_chunks():
    from xbitinfo_b200 import _native as nat

    rng = np.random.default_rng(10)
    x = torch.from_numpy(smooth_field(rng, (64, 64, 64), "float32", axis=1)).cuda()
    n0 = nat.kernel_launches()
    bitround_by_chunks(x, {0: 64, 1: 4, 2: 4}, axis=0)  # 256 chunks
    few = nat.kernel_launches() - n0
    n0 = nat.kernel_launches()
    bitround_by_chunks(x, {0: 64, 1: 2, 2: 2}, axis=0)  # 1024 chunks
    assert nat.kernel_launches() - n0 == few <= 4


def test_bitround_chunked_irregular_grid():
    rng = np.random.default_rng(12)
    x = smooth_field(rng, (40, 25, 53), "float64", axis=1, scale=0.3, base=3.0)
    spec = ((13, 27), 5, (3, 30, 20))
    grid = normalize_chunks(x.shape, spec)
    shape = tuple(len(c) for c in grid)
    keep = rng.integers(0, 52, size=shape, endpoint=True).astype("int32")
    t = torch.from_numpy(x).cuda()
    bitround_chunked_(t, spec, torch.from_numpy(keep.ravel()).cuda())
    want = np.empty_like(x)
    for idx, sl in zip(np.ndindex(*shape), slices_from_chunks(grid)):
        want[sl] = bro.bitround(np.ascontiguousarray(x[sl]), int(keep[idx]))
    assert np.array_equal(t.cpu().numpy().view("u8"), want.view("u8"))


def test_single_element_and_all_ones_shapes():
    """Found by scripts/fuzz_gpu.py: an array whose every dimension has length 1 leaves no group after the
    planner drops them; the stand-in group must be fully initialised."""
    for shape in ((1,), (1, 1), (1, 1, 1, 1)):
        x = np.full(shape, 3.14159, dtype="float64")
        t = torch.from_numpy(x).cuda()
        bitround_chunked_(t, shape, torch.tensor([5], dtype=torch.int32, device="cuda"))
        assert np.array_equal(t.cpu().numpy().view("u8"), bro.bitround(x, 5).view("u8"))
        c = count_pairs_chunked(torch.from_numpy(x).cuda(), shape, 0)
        assert int(c.abs().sum()) == 0 and tuple(c.shape) == shape + (64, 2, 2)


def test_unsupported_and_bad_arguments():
    x = torch.zeros((2, 2, 2, 2, 2, 2), dtype=torch.float32, device="cuda")
    with pytest.raises(NotImplementedError):
        count_pairs_chunked(x, (1, 1, 1, 1, 1, 1), 0)   # six chunked dimensions
    with pytest.raises(ValueError):
        count_pairs_chunked(x, (1, 1, 1), 0)
    with pytest.raises(ValueError):
        count_pairs_chunked(x, (1, 1, 1, 1, 1, 0), 0)
    with pytest.raises(ValueError):
        bitround_chunked_(x, (2, 2, 2, 2, 2, 1), torch.zeros(3, dtype=torch.int32, device="cuda"))
    with pytest.raises(ValueError):
        count_pairs_chunked(x[0, 0, 0, 0], ((1, 2), 2), 0)   # lengths do not add up
    empty = torch.zeros((0, 4), dtype=torch.float32, device="cuda")
    assert count_pairs_chunked(empty, (2, 2), 1).shape == (0, 2, 32, 2, 2)


def test_arrays_of_more_than_2_to_32_elements_use_64_bit_offsets():
    """The kernels switch to 64-bit element offsets at 2^32 elements.  Checked against the single-array
    kernels (K1 / K3, themselves pinned to the oracle) applied to every chunk's contiguous copy."""
    from xbitinfo_b200.device import PairCounter, bitround_

    if torch.cuda.mem_get_info()[0] < 30 << 30:
        pytest.skip("needs 30 GB of free device memory")
    g = torch.Generator(device="cuda").manual_seed(3)
    shape, chunk = (66000, 256, 256), (66000, 64, 128)  # 4.33e9 elements, 8 chunks
    x = torch.empty(shape, dtype=torch.float16, device="cuda")
    for lo in range(0, shape[0], 6000):
        blk = torch.empty((min(6000, shape[0] - lo),) + shape[1:], dtype=torch.float32, device="cuda")
        blk.normal_(0.0, 0.05, generator=g)
        torch.cumsum(blk, dim=1, out=blk)
        x[lo:lo + blk.shape[0]] = (blk + 3.0).to(torch.float16)
        del blk
    assert x.numel() >= 1 << 32
    for axis in (0, 2):
        got = count_pairs_chunked(x, chunk, axis, masked_value=None)
        pc = PairCounter(x.dtype, x.device)
        for i in range(4):
            for j in range(2):
                blk = x[:, 64 * i:64 * (i + 1), 128 * j:128 * (j + 1)].contiguous()
                pc.reset()
                pc.add(blk, axis, masked_value=None)
                assert torch.equal(got[0, i, j], pc.counts), (axis, i, j)
                del blk
    keep = torch.tensor([0, 1, 2, 3, 4, 5, 7, 10], dtype=torch.int32, device="cuda")
    ref_blocks = []
    for i in range(4):
        for j in range(2):
            blk = x[:, 64 * i:64 * (i + 1), 128 * j:128 * (j + 1)].contiguous()
            bitround_(blk, int(keep[2 * i + j]))
            ref_blocks.append(blk[::977].clone())  # a sample of rows of the reference result
            del blk
    bitround_chunked_(x, chunk, keep)
    for i in range(4):
        for j in range(2):
            got_rows = x[::977, 64 * i:64 * (i + 1), 128 * j:128 * (j + 1)]
            assert torch.equal(got_rows.view(torch.int16), ref_blocks[2 * i + j].view(torch.int16)), (i, j)


def test_ragged_chunk_of_extent_one_along_the_axis_is_left_unrounded():
    """101 rows cut in chunks of 50 leave a last chunk of ONE row along the analysed axis: it has no pairs, its
    information is NaN, and it must come back unrounded (with a warning) -- not rounded to 0 mantissa bits."""
    rng = np.random.default_rng(11)
    x = (280.0 + np.cumsum(rng.standard_normal((101, 64)) * 0.3, axis=0)).astype("float32")
    for batched in (True, False):
        with pytest.warns(RuntimeWarning, match="no valid pair"):
            rounded, keep = bitround_by_chunks(x, {0: 50}, 0, inflevel=0.99, masked_value=None,
                                                       set_zero_insignificant=False, batched=batched)
        assert keep.shape == (3, 1) and keep[2, 0] == 23 and 0 <= keep[0, 0] < 23
        assert np.array_equal(rounded[100].view("u4"), x[100].view("u4"))
        assert np.array_equal(rounded[:50].view("u4"), bro.bitround(x[:50], int(keep[0, 0])).view("u4"))
