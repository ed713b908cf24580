"""K1/K2 through the C ABI (host-buffer entry point) vs the CPU oracle -- bit-exact counts."""
import ctypes as C

import numpy as np
import pytest

from oracle import bitinfo_oracle as bo
from xbitinfo_b200 import _native as nat

from conftest import smooth_field

pytestmark = pytest.mark.gpu


def run_host(x, axis, masked_value=None, set_zero=False, confidence=0.99, force_generic=False, transform=True):
    x = np.ascontiguousarray(x)
    axis = axis % x.ndim
    outer = int(np.prod(x.shape[:axis], dtype=np.int64))
    n = int(x.shape[axis])
    inner = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    nbits = x.dtype.itemsize * 8
    flags, bits = nat.mask_spec(x.dtype, masked_value)
    if x.dtype.kind == "f" and transform:
        flags |= nat.SIGNED_EXPONENT
    if force_generic:
        flags |= nat.FORCE_GENERIC
    mi = np.empty(nbits, dtype=np.float64)
    counts = np.empty(nbits * 4, dtype=np.uint64)
    rc = nat.lib().xbi_bitinformation_host(
        x.ctypes.data_as(C.c_void_p), nat.dtype_code(x.dtype), outer, n, inner, flags, bits,
        int(set_zero), float(confidence), mi.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), 0)
    nat.check(rc)
    return counts.reshape(nbits, 2, 2).astype(np.int64), mi


def assert_mi_close(mi, ref):
    """1e-12 relative where the information is significant; below that an absolute floor
    (mutual information of noise bits is a cancelling sum)."""
    if np.isnan(ref).all():
        assert np.isnan(mi).all()
        return
    scale = max(np.nanmax(np.abs(ref)), 1e-300)
    np.testing.assert_allclose(mi, ref, rtol=1e-12, atol=1e-12 * scale)


@pytest.mark.parametrize("force_generic", [True, False])
def test_golden_cases(golden, force_generic):
    for name in golden["case_names"]:
        x = golden[f"case_{name}_x"]
        axis = int(golden[f"case_{name}_axis"])
        counts, mi = run_host(x, axis, force_generic=force_generic)
        assert np.array_equal(counts, golden[f"case_{name}_counts"]), name
        assert_mi_close(mi, golden[f"case_{name}_mi"])


DTYPES = ["float32", "float64", "float16", "int8", "uint8", "int16", "uint16", "int32", "uint32", "int64", "uint64"]


@pytest.mark.parametrize("dtype", DTYPES)
@pytest.mark.parametrize("force_generic", [True, False])
def test_every_axis_every_dtype(dtype, force_generic):
    rng = np.random.default_rng(hash(dtype) % 2**32)
    shape = (5, 33, 7, 130)
    if np.dtype(dtype).kind == "f":
        x = smooth_field(rng, shape, dtype, axis=-1, scale=0.2, base=3.0)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
    for axis in range(4):
        counts, mi = run_host(x, axis, force_generic=force_generic)
        ref_counts, npairs, ref_mi = bo.bitinformation(x, axis)
        assert np.array_equal(counts, ref_counts), (dtype, axis)
        assert counts[0].sum() == npairs
        assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16"])
@pytest.mark.parametrize("masked_value", [float("nan"), -999.0])
@pytest.mark.parametrize("force_generic", [True, False])
def test_masked_pairs(dtype, masked_value, force_generic):
    rng = np.random.default_rng(7)
    x = smooth_field(rng, (6, 40, 96), dtype, axis=1, scale=0.3, base=5.0)
    land = rng.random((40, 96)) < 0.3
    x[:, land] = masked_value
    x[0, 0, :5] = np.nan  # NaNs that are not the masked value in the -999 variant
    for axis in range(3):
        counts, mi = run_host(x, axis, masked_value=masked_value, force_generic=force_generic)
        ref_counts, npairs, ref_mi = bo.bitinformation(x, axis, masked_value=masked_value)
        assert np.array_equal(counts, ref_counts), (dtype, axis)
        assert counts[0].sum() == npairs
        assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("confidence", [0.99, 0.5])
def test_set_zero_insignificant(confidence):
    rng = np.random.default_rng(11)
    x = smooth_field(rng, (50, 700), "float32", axis=1, scale=0.05, base=280.0)
    counts, mi = run_host(x, 1, set_zero=True, confidence=confidence)
    _, npairs, ref = bo.bitinformation(x, 1, set_zero_insignificant_flag=True, confidence=confidence)
    thr = bo.free_entropy_threshold(npairs, confidence)
    _, mi_raw = run_host(x, 1)
    # away from the threshold the decisions must agree exactly
    far = np.abs(mi_raw - thr) > 1e-6 * thr
    assert np.array_equal((mi == 0)[far], (ref == 0)[far])
    assert_mi_close(mi[far], ref[far])


def test_errors():
    x = np.zeros((4, 4), dtype="int32")
    with pytest.raises(ValueError):
        run_host(x, 1, masked_value=float("nan"), transform=False) if False else nat.check(
            nat.lib().xbi_bitinformation_host(x.ctypes.data_as(C.c_void_p), nat.U32, 4, 4, 1, nat.MASK_NAN, 0, 0, 0.99,
                                              np.empty(32).ctypes.data_as(C.c_void_p), None, 0))


STRIDED_DTYPES = ["float32", "float64", "float16", "uint8", "int16", "int32", "int64"]


@pytest.mark.parametrize("dtype", STRIDED_DTYPES)
@pytest.mark.parametrize("mask", [None, "nan", "value"])
def test_strided_fast_path(dtype, mask):
    """Adjacency along non-last axes with 16-byte aligned rows: the column-walker kernel
    (must actually be taken) vs the oracle."""
    isfloat = np.dtype(dtype).kind == "f"
    if mask == "nan" and not isfloat:
        pytest.skip("NaN mask needs floats")
    rng = np.random.default_rng(3)
    shape = (3, 301, 7, 64)  # rows of 64 elements: 16-byte aligned for every dtype
    if isfloat:
        x = smooth_field(rng, shape, dtype, axis=1, scale=0.2, base=3.0)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
    masked_value = None
    if mask is not None:
        masked_value = float("nan") if mask == "nan" else (-999.0 if isfloat else 7)
        land = rng.random(shape[2:]) < 0.3
        x[:, :, land] = masked_value
        x[1, 5:9] = masked_value  # whole rows masked
    for axis in (0, 1, 2):
        before = nat.fast_path_launches()
        counts, mi = run_host(x, axis, masked_value=masked_value)
        assert nat.fast_path_launches() > before, "fast path was not taken"
        ref_counts, npairs, ref_mi = bo.bitinformation(x, axis, masked_value=masked_value)
        assert np.array_equal(counts, ref_counts), (dtype, axis, mask)
        assert counts[0].sum() == npairs
        assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "uint8", "uint16"])
@pytest.mark.parametrize("mask", [None, "nan", "value"])
def test_last_axis_fast_path(dtype, mask):
    isfloat = np.dtype(dtype).kind == "f"
    if mask == "nan" and not isfloat:
        pytest.skip("NaN mask needs floats")
    rng = np.random.default_rng(5)
    shape = (37, 11, 1441)  # odd row length: rows are not vector aligned
    if isfloat:
        x = smooth_field(rng, shape, dtype, axis=-1, scale=0.2, base=3.0)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
    masked_value = None
    if mask is not None:
        masked_value = float("nan") if mask == "nan" else (-999.0 if isfloat else 7)
        x[rng.random(shape) < 0.2] = masked_value
    before = nat.fast_path_launches()
    counts, mi = run_host(x, 2, masked_value=masked_value)
    assert nat.fast_path_launches() > before, "fast path was not taken"
    ref_counts, npairs, ref_mi = bo.bitinformation(x, 2, masked_value=masked_value)
    assert np.array_equal(counts, ref_counts), (dtype, mask)
    assert counts[0].sum() == npairs
    assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("dtype,mask", [("float64", "nan"), ("float64", None), ("float32", "nan"), ("float16", None)])
def test_long_segments_and_ragged_strips_fast_equals_generic(dtype, mask):
    """Column walkers on a tall, narrow matrix: rows are not a multiple of the 512-byte strip
    (idle lanes) and segments are longer than a flush period.  At this size the checker is
    the generic kernel (itself checked against the oracle above)."""
    import torch

    from xbitinfo_b200.device import PairCounter

    rows, cols = 700_000, (40 if dtype == "float16" else 36)  # 16-byte aligned rows, not a multiple of 512 bytes
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.empty((rows, cols), dtype=torch.float32, device="cuda").normal_(0.0, 0.3, generator=g)
    x = (torch.cumsum(x, dim=0) + 3.0).to(getattr(torch, dtype))
    mv = None
    if mask == "nan":
        mv = float("nan")
        x[torch.rand((rows, cols), device="cuda", generator=g) < 0.1] = float("nan")
    pc = PairCounter(x.dtype, x.device)
    before = nat.fast_path_launches()
    pc.add(x, 0, masked_value=mv)
    assert nat.fast_path_launches() > before
    fast = pc.counts.clone()
    pc.reset()
    pc.add(x, 0, masked_value=mv, force_generic=True)
    assert torch.equal(fast, pc.counts)
    if mask is None:
        assert int(fast[0].sum()) == (rows - 1) * cols


def test_all_float16_patterns_through_the_packed_paths():
    """Every float16 bit pattern (incl. NaN payloads, infinities, subnormals) in random order,
    along the last axis (TMA row kernel, two halves per word) and along axis 0 (column
    walkers): the packed signed-exponent transform must match the reference bit for bit."""
    rng = np.random.default_rng(12)
    pats = np.arange(1 << 16, dtype="uint16")
    x = np.concatenate([rng.permutation(pats) for _ in range(6)]).view("float16").reshape(768, 512)
    for axis in (1, 0):
        before = nat.fast_path_launches()
        counts, _ = run_host(x, axis)
        assert nat.fast_path_launches() > before
        assert np.array_equal(counts, bo.bitinformation(x, axis)[0]), axis
    counts, _ = run_host(x, 1, masked_value=float("nan"))
    assert np.array_equal(counts, bo.bitinformation(x, 1, masked_value=float("nan"))[0])


@pytest.mark.parametrize("dtype", ["float32", "float16", "uint8", "float64"])
@pytest.mark.parametrize("n", [1920, 3840, 7680, 481, 250])
def test_row_ends_on_stage_boundaries(dtype, n):
    """Row lengths that put row ends exactly on TMA stage boundaries (the partner of the last
    element is the halo word), just above the in-kernel threshold, and below it."""
    rng = np.random.default_rng(n)
    rows = 70 if n >= 1920 else 300
    if np.dtype(dtype).kind == "f":
        x = smooth_field(rng, (rows, n), dtype, axis=1, scale=0.3, base=2.0)
    else:
        x = rng.integers(0, 255, size=(rows, n), dtype=dtype, endpoint=True)
    before = nat.fast_path_launches()
    counts, _ = run_host(x, 1)
    assert nat.fast_path_launches() > before
    ref, npairs, _ = bo.bitinformation(x, 1)
    assert np.array_equal(counts, ref)
    assert counts[0].sum() == npairs == rows * (n - 1)


# ---------------------------------------------------------------------------------------
# masked counting through the fast kernels: clean warps, fully masked warps, validity
# boundaries (the rare-event corrections of csrc/xbi_mask.cuh), probe false positives
# ---------------------------------------------------------------------------------------
def _mask_pattern(x, pattern, rng, mv):
    """Write the masked value into x (in place) following a named pattern."""
    flat = x.reshape(-1)
    if pattern == "none":
        pass
    elif pattern == "blocks":  # long contiguous runs: whole warps see only masked elements
        n = flat.size
        for lo in (0, n // 5, n // 2 + 13):
            flat[lo: lo + n // 8 + 7] = mv
    elif pattern == "alternating":  # every pair straddles a validity boundary
        flat[::2] = mv
    elif pattern == "sparse":  # isolated masked elements
        flat[rng.integers(0, flat.size, size=max(4, flat.size // 3000))] = mv
    elif pattern == "ends":  # first / last elements of the array and of rows
        flat[0] = mv
        flat[-1] = mv
        x[..., 0] = mv
        x[..., -1] = mv
    elif pattern == "all":
        flat[:] = mv
    elif pattern == "slabs":  # masked along every axis: slabs, rows and columns
        x[x.shape[0] // 3: x.shape[0] // 3 + max(1, x.shape[0] // 4)] = mv
        x[:, ..., 1::5] = mv
    else:
        raise AssertionError(pattern)


MASK_PATTERNS = ["none", "blocks", "alternating", "sparse", "ends", "all", "slabs"]


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16"])
@pytest.mark.parametrize("pattern", MASK_PATTERNS)
@pytest.mark.parametrize("mv", [float("nan"), -999.0])
def test_mask_patterns_rows_kernel(dtype, pattern, mv):
    rng = np.random.default_rng(21)
    x = smooth_field(rng, (48, 4099), dtype, axis=1, scale=0.3, base=2.0)
    x[3, 7] = np.inf  # probe false positive in the NaN variants; an ordinary value otherwise
    x[5, 100:140] = -np.inf
    _mask_pattern(x, pattern, rng, mv)
    before = nat.fast_path_launches()
    counts, mi = run_host(x, 1, masked_value=mv)
    assert nat.fast_path_launches() > before
    ref_counts, npairs, ref_mi = bo.bitinformation(x, 1, masked_value=mv)
    assert np.array_equal(counts, ref_counts), (dtype, pattern, mv)
    assert counts[0].sum() == npairs
    assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16"])
@pytest.mark.parametrize("pattern", MASK_PATTERNS)
@pytest.mark.parametrize("shape", [(1300, 4, 128), (900, 5, 72)])  # power-of-two pitch / other pitch (TMA ring)
def test_mask_patterns_column_walkers(dtype, pattern, shape):
    rng = np.random.default_rng(22)
    mv = float("nan") if shape[2] == 128 else -999.0
    x = smooth_field(rng, shape, dtype, axis=0, scale=0.3, base=2.0)
    x[7, 1, 3] = np.inf
    _mask_pattern(x, pattern, rng, mv)
    before = nat.fast_path_launches()
    counts, mi = run_host(x, 0, masked_value=mv)
    assert nat.fast_path_launches() > before
    ref_counts, npairs, ref_mi = bo.bitinformation(x, 0, masked_value=mv)
    assert np.array_equal(counts, ref_counts), (dtype, pattern, shape)
    assert counts[0].sum() == npairs
    assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("dtype,mv", [("uint8", 7), ("int16", -3), ("uint32", 0), ("int64", 2**40 + 5)])
@pytest.mark.parametrize("pattern", ["blocks", "alternating", "sparse", "ends"])
def test_mask_patterns_integers(dtype, mv, pattern):
    rng = np.random.default_rng(23)
    info = np.iinfo(dtype)
    for shape, axis in (((40, 8200), 1), ((1100, 3, 64), 0)):
        x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
        _mask_pattern(x, pattern, rng, mv)
        before = nat.fast_path_launches()
        counts, _ = run_host(x, axis, masked_value=mv)
        assert nat.fast_path_launches() > before
        ref_counts, npairs, _ = bo.bitinformation(x, axis, masked_value=mv)
        assert np.array_equal(counts, ref_counts), (dtype, pattern, shape)
        assert counts[0].sum() == npairs


def test_mask_probe_false_positives_float64():
    """Doubles >= 2^1017 and infinities look like NaN to the high-word probe; a masked value
    whose high word collides with ordinary data looks masked to the value probe.  Both must
    only cost time, never change the counts."""
    rng = np.random.default_rng(24)
    x = smooth_field(rng, (30, 4100), "float64", axis=1, scale=0.3, base=2.0)
    x[2, 50:90] = 1.5e306
    x[4, 10] = np.inf
    x[9, 1000:1100] = np.nan
    for mv in (float("nan"), 2.0):
        y = x.copy()
        if mv == 2.0:
            y[11, 5:50] = 2.0
            y[12, 5:50] = np.nextafter(2.0, 3.0)  # same high word as the masked value, not masked
        counts, _ = run_host(y, 1, masked_value=mv)
        ref_counts, npairs, _ = bo.bitinformation(y, 1, masked_value=mv)
        assert np.array_equal(counts, ref_counts)
        assert counts[0].sum() == npairs


@pytest.mark.parametrize("axis", [0, 1])
def test_mask_probe_false_positives_float32(axis):
    """The float32 NaN probe multiplies three words per FMA lane (a * b + c): products beyond 3.4e38, infinities of
    either sign and zeros next to them make NaNs out of NaN-free words.  Such stages only take the exact test; real
    NaNs -- alone, next to infinities, next to zeros, at the neighbour word -- are never missed."""
    rng = np.random.default_rng(26)
    x = smooth_field(rng, (40, 4096), "float32", axis=1, scale=0.3, base=2.0)
    x[1, 100:400] = 3e38            # every product overflows
    x[2, 100:400] = -3e38           # ... to the other infinity
    x[3, 10:300:2] = 0.0            # zero factors between ordinary values
    x[4, 10:300:3] = 0.0
    x[4, 11:300:3] = 3e38           # inf * 0 at the second level
    x[5, 50] = np.inf
    x[5, 51] = 0.0
    x[5, 52] = -np.inf
    x[6, 700:760] = np.nan
    x[7, 15] = np.nan               # a lone NaN among zeros and huge values
    x[7, 10:15] = 0.0
    x[7, 16:22] = 3e38
    x[8, 0:4096:16] = np.nan        # every thread's first word
    x[9, 15:4096:16] = np.nan       # ... last word / the neighbour word of the thread before
    x[10, 5::6] = np.nan            # one NaN per FMA group, every position in turn
    x[11, 4::6] = np.nan
    x[12, 3::6] = np.nan
    x[13, 2::6] = np.nan
    x[14, 1::6] = np.nan
    x[15, 0::6] = np.nan
    if axis == 0:
        x = np.ascontiguousarray(x.T)
    counts, _ = run_host(x, axis, masked_value=float("nan"))
    ref_counts, npairs, _ = bo.bitinformation(x, axis, masked_value=float("nan"))
    assert np.array_equal(counts, ref_counts)
    assert counts[0].sum() == npairs
    import torch

    from xbitinfo_b200.device import PairCounter

    pc = PairCounter(torch.float32, "cuda")
    for flavour in ("sparse", "dense"):
        pc.reset()
        pc.add(torch.from_numpy(x).cuda(), axis, masked_value=float("nan"), flavour=flavour)
        assert np.array_equal(pc.counts.cpu().numpy(), ref_counts), flavour


def test_masked_zero_matches_bit_pattern_only():
    """masked_value=0.0 masks +0.0 but not -0.0 (bit identity)."""
    rng = np.random.default_rng(25)
    x = smooth_field(rng, (20, 8000), "float32", axis=1, scale=0.3, base=0.0)
    x[3, 100:200] = 0.0
    x[4, 100:200] = -0.0
    counts, _ = run_host(x, 1, masked_value=0.0)
    ref_counts, npairs, _ = bo.bitinformation(x, 1, masked_value=0.0)
    assert np.array_equal(counts, ref_counts)
    assert counts[0].sum() == npairs == 20 * 7999 - 101


@pytest.mark.parametrize("dtype", ["float32", "float64", "int32", "uint64"])
@pytest.mark.parametrize("mask", [None, "nan", "value", "neutral"])
def test_strided_rows_that_are_only_element_aligned(dtype, mask):
    """Adjacency along non-last axes when a row is not a multiple of 16 bytes (and the base
    is not 16-byte aligned either): whole-word element types still take the column walkers,
    with element-sized copies; lane slots beyond the end of a row hold the neutral value.
    'neutral' masks exactly that padding value (1.0 / 0)."""
    import torch

    from xbitinfo_b200.device import PairCounter

    isfloat = np.dtype(dtype).kind == "f"
    if mask == "nan" and not isfloat:
        pytest.skip("NaN mask needs floats")
    rng = np.random.default_rng(31)
    shape = (3, 257, 5, 13)  # rows of 13 and of 65 elements
    n = int(np.prod(shape))
    if isfloat:
        flat = smooth_field(rng, (n + 3,), dtype, axis=0, scale=0.2, base=1.0)
    else:
        flat = rng.integers(0, 5, size=n + 3, dtype=dtype)
    masked_value = None
    if mask is not None:
        masked_value = {"nan": float("nan"), "value": (-999.0 if isfloat else 3), "neutral": (1.0 if isfloat else 0)}[mask]
        sel = rng.random(n + 3) < 0.25
        sel[1000:1400] = True
        flat[sel] = masked_value
    for off in (0, 1, 3):  # base address 0, 1 and 3 elements past a 16-byte boundary
        x = flat[off: off + n].reshape(shape)
        t = torch.from_numpy(flat).cuda()[off: off + n].view(shape)
        for axis in (0, 1, 2):
            pc = PairCounter(t.dtype, t.device)
            before = nat.fast_path_launches()
            pc.add(t, axis, masked_value=masked_value)
            assert nat.fast_path_launches() > before, "fast path was not taken"
            ref_counts, npairs, _ = bo.bitinformation(x, axis, masked_value=masked_value)
            assert np.array_equal(pc.counts.cpu().numpy(), ref_counts), (dtype, mask, off, axis)
            assert int(pc.counts[0].sum()) == npairs


@pytest.mark.parametrize("dtype", ["float32", "float64", "uint32"])
@pytest.mark.parametrize("mask", [None, "nan_or_value"])
def test_element_aligned_rows_strips_and_side_by_side_segments(dtype, mask):
    """The element-aligned column walkers give a warp a strip of 128 (64 for 8-byte types) columns and, when a whole
    row is at most half a strip, walk 2..32 row segments side by side in one warp.  Row lengths on both sides of
    every strip / power-of-two fraction of a strip x row counts with and without a tail group; the base address is
    one element past a 16-byte boundary, so that row lengths that ARE a multiple of 16 bytes take this kernel too."""
    import torch

    from xbitinfo_b200.device import PairCounter

    isfloat = np.dtype(dtype).kind == "f"
    rng = np.random.default_rng(77)
    mv = None
    if mask is not None:
        mv = float("nan") if isfloat else 3
    for inner in (2, 3, 31, 32, 33, 50, 64, 65, 96, 97, 128, 129, 161, 193, 257, 513):
        for rows in (2, 3, 4, 5, 6, 8, 9, 23, 70):
            n = rows * inner
            if isfloat:
                flat = smooth_field(rng, (n + 1,), dtype, axis=0, scale=0.2, base=1.0)
            else:
                flat = rng.integers(0, 5, size=n + 1, dtype=dtype)
            if mv is not None:
                flat[rng.random(n + 1) < 0.2] = mv
            x = flat[1:].reshape(rows, inner)
            t = torch.from_numpy(flat).cuda()[1:].view(rows, inner)
            pc = PairCounter(t.dtype, t.device)
            before = nat.fast_path_launches()
            pc.add(t, 0, masked_value=mv)
            assert nat.fast_path_launches() > before, "fast path was not taken"
            ref_counts, npairs, _ = bo.bitinformation(x, 0, masked_value=mv)
            assert np.array_equal(pc.counts.cpu().numpy(), ref_counts), (dtype, mask, rows, inner)
            assert int(pc.counts[0].sum()) == npairs


@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_element_aligned_rows_many_segments_fast_equals_generic(dtype):
    """Tall matrices with 513 / 161 / 1441 columns (many row segments per strip, segment lengths that differ by one,
    a last strip with few columns) and with 13 / 33 / 3 columns (several segments side by side in a warp).  The
    checker at this size is the generic kernel."""
    import torch

    from xbitinfo_b200.device import PairCounter

    g = torch.Generator(device="cuda").manual_seed(9)
    for rows, cols in ((40_001, 513), (90_003, 161), (20_002, 1441), (300_001, 13), (200_003, 33), (1_000_001, 3)):
        x = torch.empty((rows, cols), dtype=torch.float32, device="cuda").normal_(0.0, 0.3, generator=g)
        x = (torch.cumsum(x, dim=0) + 3.0).to(getattr(torch, dtype))
        x[torch.rand((rows, cols), device="cuda", generator=g) < 0.05] = float("nan")
        for mv in (None, float("nan")):
            pc = PairCounter(x.dtype, x.device)
            before = nat.fast_path_launches()
            pc.add(x, 0, masked_value=mv)
            assert nat.fast_path_launches() > before
            fast = pc.counts.clone()
            pc.reset()
            pc.add(x, 0, masked_value=mv, force_generic=True)
            assert torch.equal(fast, pc.counts), (dtype, rows, cols, mv)


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "int32", "uint8"])
@pytest.mark.parametrize("pattern", ["none", "blocks", "alternating", "sparse", "ends", "all", "slabs"])
def test_both_flavours_of_the_masked_kernels_agree_with_the_oracle(dtype, pattern):
    """The masked fast kernels exist in a boundary-correction ("sparse") and a three-stream
    ("dense") flavour, picked on the device from a sampled boundary density.  Forced either
    way, and left to choose, each must give the oracle's counts on every mask pattern."""
    import torch

    from xbitinfo_b200.device import PairCounter

    isfloat = np.dtype(dtype).kind == "f"
    rng = np.random.default_rng(41)
    for shape, axis in (((24, 8200), 1), ((900, 3, 64), 0)):
        if isfloat:
            x = smooth_field(rng, shape, dtype, axis=axis, scale=0.3, base=2.0)
        else:
            info = np.iinfo(dtype)
            x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
        mv = float("nan") if isfloat else 7
        _mask_pattern(x, pattern, rng, mv)
        ref_counts, npairs, _ = bo.bitinformation(x, axis, masked_value=mv)
        t = torch.from_numpy(x).cuda()
        for flavour in ("sparse", "dense", None):
            pc = PairCounter(t.dtype, t.device)
            before = nat.fast_path_launches()
            pc.add(t, axis, masked_value=mv, flavour=flavour)
            assert nat.fast_path_launches() > before
            assert np.array_equal(pc.counts.cpu().numpy(), ref_counts), (dtype, pattern, shape, flavour)
            assert int(pc.counts[0].sum()) == npairs
