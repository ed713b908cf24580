"""The fused device entry points (xbi_bitinformation_dev: K1 + fringe + combine + K2 in three launches;
xbi_bitinformation_dev_allreduce: the same with the all-reduce over peer memory inside the finish kernel)
against the CPU oracle -- counts bit-exact, information within 1e-12."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import bitinfo_oracle as bo
from xbitinfo_b200 import _native as nat
from xbitinfo_b200.device import PairCounter

from conftest import smooth_field
from test_counts_gpu import assert_mi_close

pytestmark = pytest.mark.gpu


def _torch_view(x):
    """CUDA tensor with the bits of a numpy array (unsigned 16/32/64 travel as the signed type of that width)."""
    name = {"uint16": "int16", "uint32": "int32", "uint64": "int64"}.get(x.dtype.name, x.dtype.name)
    return torch.from_numpy(x.view(name)).cuda()


def fused(x, axis, **kw):
    t = x if isinstance(x, torch.Tensor) else _torch_view(np.ascontiguousarray(x))
    pc = PairCounter(t.dtype, t.device)
    pc.counts.fill_(-12345)  # must be overwritten, not accumulated into
    mi = pc.bitinformation(t, axis, **kw)
    return pc.counts.cpu().numpy(), mi.cpu().numpy()


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16", "uint8", "int16", "int32", "int64"])
def test_every_axis(dtype):
    rng = np.random.default_rng(5)
    shape = (3, 37, 6, 1100)  # last axis: one TMA stage and a ragged tail; other axes: column walkers
    if np.dtype(dtype).kind == "f":
        x = smooth_field(rng, shape, dtype, axis=-1, scale=0.2, base=3.0)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
    for axis in range(4):
        counts, mi = fused(x, axis)
        ref_counts, npairs, ref_mi = bo.bitinformation(x, axis)
        assert np.array_equal(counts, ref_counts), (dtype, axis)
        assert counts[0].sum() == npairs
        assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("n", [2, 3, 17, 239, 240, 241, 1440, 7680, 7681, 20000])
@pytest.mark.parametrize("dtype", ["float32", "float64", "float16"])
def test_row_lengths_around_the_stage_size(n, dtype):
    """Last-axis analysis: row ends inside / outside the TMA range, rows shorter and longer than a stage, and the
    leftovers (head, tail, their row ends) that the finish kernel counts."""
    rng = np.random.default_rng(n)
    outer = max(2, 60000 // n)
    x = smooth_field(rng, (outer, n), dtype, axis=-1, scale=0.3, base=1.5)
    counts, mi = fused(x, 1)
    ref_counts, _, ref_mi = bo.bitinformation(x, 1)
    assert np.array_equal(counts, ref_counts)
    assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("offset", [1, 2, 3, 5])
def test_unaligned_base_pointer(offset):
    """A slice that starts off a 16-byte boundary: the head pairs go through the finish kernel."""
    rng = np.random.default_rng(offset)
    flat = smooth_field(rng, (offset + 40 * 1500,), "float32", axis=0, scale=0.3)
    t = torch.from_numpy(flat).cuda()[offset:].view(40, 1500)
    assert t.data_ptr() % 16 != 0 and t.is_contiguous()
    counts, mi = fused(t, 1)
    ref_counts, _, ref_mi = bo.bitinformation(flat[offset:].reshape(40, 1500), 1)
    assert np.array_equal(counts, ref_counts)
    assert_mi_close(mi, ref_mi)


@pytest.mark.parametrize("dtype", ["float32", "float64", "float16"])
@pytest.mark.parametrize("masked_value", [float("nan"), -999.0])
def test_masked_and_significance(dtype, masked_value):
    rng = np.random.default_rng(7)
    x = smooth_field(rng, (6, 40, 1000), dtype, axis=2, scale=0.3, base=5.0)
    land = rng.random((40, 1000)) < 0.3
    x[:, land] = masked_value
    x[5, 39, -3:] = masked_value  # masked elements in the ragged tail
    for axis in range(3):
        counts, mi = fused(x, axis, masked_value=masked_value)
        ref_counts, npairs, ref_mi = bo.bitinformation(x, axis, masked_value=masked_value)
        assert np.array_equal(counts, ref_counts), (dtype, axis)
        assert_mi_close(mi, ref_mi)
    # the significance filter inside the finish kernel == the stand-alone K2 on the same counts
    t = _torch_view(x)
    pc = PairCounter(t.dtype, t.device)
    a = pc.bitinformation(t, 2, masked_value=masked_value, set_zero_insignificant=True, confidence=0.99).cpu().numpy()
    b = pc.information(True, 0.99).cpu().numpy()
    assert np.array_equal(a, b, equal_nan=True)
    assert (a == 0).any()


def test_generic_route_and_large_leftovers():
    """FORCE_GENERIC leaves everything to the generic kernels (more leftovers than the finish kernel takes)."""
    rng = np.random.default_rng(3)
    x = smooth_field(rng, (300, 500), "float32", axis=1, scale=0.2)
    counts, mi = fused(x, 1, flavour="generic")
    ref_counts, _, ref_mi = bo.bitinformation(x, 1)
    assert np.array_equal(counts, ref_counts)
    assert_mi_close(mi, ref_mi)
    x = smooth_field(rng, (9000, 5), "float32", axis=1, scale=0.2)  # 8999 row ends: beyond the finish kernel's share
    for flavour in (None, "generic"):
        counts, _ = fused(x, 1, flavour=flavour)
        assert np.array_equal(counts, bo.bitinformation(x, 1)[0])


def test_no_pairs_and_repeat():
    x = np.ones((5, 1), dtype="float32")
    counts, mi = fused(x, 1)
    assert (counts == 0).all() and np.isnan(mi).all()
    counts, mi = fused(np.ones((0, 4), dtype="float32"), 1)
    assert (counts == 0).all() and np.isnan(mi).all()
    # overwrite semantics: a second call on the same counter gives the same table
    rng = np.random.default_rng(1)
    x = smooth_field(rng, (50, 3000), "float32", axis=1, scale=0.2)
    t = torch.from_numpy(x).cuda()
    pc = PairCounter(t.dtype, t.device)
    pc.bitinformation(t, 1)
    first = pc.counts.clone()
    pc.bitinformation(t, 1)
    assert torch.equal(first, pc.counts)
    # counts only
    assert pc.bitinformation(t, 1, with_information=False) is None
    assert torch.equal(first, pc.counts)
    # the accumulate-only entry point still accumulates
    pc.add(t, 1)
    assert torch.equal(2 * first, pc.counts)


def test_bad_arguments():
    lib = nat.lib()
    ws = torch.empty(lib.xbi_workspace_bytes(), dtype=torch.uint8, device="cuda")
    counts = torch.zeros(128, dtype=torch.int64, device="cuda")
    x = torch.zeros(64, dtype=torch.float32, device="cuda")
    args = [x.data_ptr(), nat.F32, 1, 64, 1, nat.SIGNED_EXPONENT, 0, 1, 1.5, counts.data_ptr(), counts.data_ptr(),
            ws.data_ptr(), ws.numel(), None]
    with pytest.raises(ValueError, match="confidence"):
        nat.check(lib.xbi_bitinformation_dev(*args))
    args[7] = 0
    args[12] = 8
    with pytest.raises(ValueError, match="workspace"):
        nat.check(lib.xbi_bitinformation_dev(*args))
    table = (C.c_void_p * 2)(ws.data_ptr(), ws.data_ptr())
    args[12] = ws.numel()
    for world, rank, epoch in ((9, 0, 1), (2, 2, 1), (2, 0, 0)):
        with pytest.raises(ValueError):
            nat.check(lib.xbi_bitinformation_dev_allreduce(*args, world, rank, epoch, table, 1.0))


def test_peer_allreduce_protocol_on_one_gpu():
    """The exchange protocol of xbi_bitinformation_dev_allreduce with both 'ranks' on this GPU: two exchange
    buffers, two streams, three epochs (both slot parities).  Every rank must end up with the sum of the two
    tables and the information of the whole.  (Small slabs: the two finish kernels wait for each other and must
    be resident at the same time.  Across real GPUs: tests/test_distributed_gpu.py.)"""
    lib = nat.lib()
    rng = np.random.default_rng(21)
    x = smooth_field(rng, (24, 2000), "float32", axis=1, scale=0.25)
    slabs = [x[:10], x[10:]]
    ref_counts, _, ref_mi = bo.bitinformation(x, 1)
    bufs, handle = [], C.create_string_buffer(int(lib.xbi_peer_handle_bytes()))
    for _ in range(2):
        b = C.c_void_p()
        nat.check(lib.xbi_peer_create(C.byref(b), handle))
        bufs.append(b)
    table = (C.c_void_p * 2)(bufs[0].value, bufs[1].value)
    try:
        streams = [torch.cuda.Stream() for _ in range(2)]
        ts = [torch.from_numpy(np.ascontiguousarray(s)).cuda() for s in slabs]
        pcs = [PairCounter(torch.float32, "cuda") for _ in range(2)]
        torch.cuda.synchronize()
        for epoch in (1, 2, 3):
            for r in range(2):
                pcs[r].counts.fill_(-1)
            torch.cuda.synchronize()
            for r in range(2):
                rc = lib.xbi_bitinformation_dev_allreduce(
                    ts[r].data_ptr(), nat.F32, ts[r].shape[0], ts[r].shape[1], 1, nat.SIGNED_EXPONENT, 0, 0, 0.99,
                    pcs[r].counts.data_ptr(), pcs[r].mi.data_ptr(), pcs[r].workspace.data_ptr(),
                    pcs[r].workspace.numel(), streams[r].cuda_stream, 2, r, epoch, table, 5.0)
                nat.check(rc)
            torch.cuda.synchronize()
            for r in range(2):
                assert np.array_equal(pcs[r].counts.cpu().numpy(), ref_counts), (epoch, r)
                assert_mi_close(pcs[r].mi.cpu().numpy(), ref_mi)
                failed = C.c_ulonglong(99)
                nat.check(lib.xbi_peer_error_epoch(bufs[r], C.byref(failed)))
                assert failed.value == 0
        # a peer that never shows up: the wait gives up after the timeout instead of hanging the GPU
        rc = lib.xbi_bitinformation_dev_allreduce(
            ts[0].data_ptr(), nat.F32, ts[0].shape[0], ts[0].shape[1], 1, nat.SIGNED_EXPONENT, 0, 0, 0.99,
            pcs[0].counts.data_ptr(), pcs[0].mi.data_ptr(), pcs[0].workspace.data_ptr(), pcs[0].workspace.numel(),
            streams[0].cuda_stream, 2, 0, 4, table, 0.2)
        nat.check(rc)
        torch.cuda.synchronize()
        failed = C.c_ulonglong(0)
        nat.check(lib.xbi_peer_error_epoch(bufs[0], C.byref(failed)))
        assert failed.value == 4
    finally:
        torch.cuda.synchronize()
        for b in bufs:
            lib.xbi_peer_destroy(b)


# ---------------------------------------------------------------------------------------
# several dimensions of one device-resident array in one call (the two innermost in ONE read)
# ---------------------------------------------------------------------------------------
from xbitinfo_b200 import engine  # noqa: E402
from xbitinfo_b200.device import MultiAxisCounter  # noqa: E402


def _multi(x, axes, **kw):
    t = x if isinstance(x, torch.Tensor) else _torch_view(np.ascontiguousarray(x))
    return engine.bitinformation_multi(t, axes, return_counts=True, **kw)


@pytest.mark.parametrize("shape", [(3, 70, 64), (5, 9, 33, 128), (2, 300, 8), (4, 97, 132), (1, 1, 130, 520), (70, 36)])
@pytest.mark.parametrize("dtype", ["float32", "int32", "uint32"])
def test_all_axes_in_one_call(shape, dtype):
    """xbi_bitinformation_dev_multi == one single-axis call per axis == the oracle; the shapes walk through ragged
    strips (rows that are not a multiple of 512 bytes), single-strip rows, tail groups and block boundaries."""
    rng = np.random.default_rng(len(shape) * 1000 + shape[-1])
    if dtype == "float32":
        x = smooth_field(rng, shape, dtype, axis=-1, scale=0.2, base=3.0)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
    axes = list(range(len(shape)))
    launches0 = nat.fast_path_launches()
    got = _multi(x, axes)
    for a, (mi, counts) in zip(axes, got):
        if shape[a] < 2:
            assert (counts == 0).all()
            continue
        ref_counts, npairs, ref_mi = bo.bitinformation(x, a)
        assert np.array_equal(counts, ref_counts), (shape, dtype, a)
        assert_mi_close(mi, ref_mi)
    assert nat.fast_path_launches() > launches0
    # a subset and a different order of the axes
    got = _multi(x, axes[::-1][:2])
    for a, (mi, counts) in zip(axes[::-1][:2], got):
        assert np.array_equal(counts, bo.bitinformation(x, a)[0] if shape[a] >= 2 else np.zeros_like(counts))


@pytest.mark.parametrize("masked_value", [float("nan"), -999.0])
def test_all_axes_masked_clean_and_dirty(masked_value):
    """With a mask the two-axis kernel counts optimistically and looks for masked elements: clean data takes the
    one-read route, data that holds masked elements is recounted by the masked kernels -- same answers."""
    rng = np.random.default_rng(17)
    x = smooth_field(rng, (6, 80, 256), "float32", axis=2, scale=0.3, base=5.0)
    for dirty in (False, True):
        if dirty:
            x[:, rng.random((80, 256)) < 0.3] = masked_value
            x[3, 40, 100] = np.nan  # a NaN that is not the masked value in the -999 variant
        got = _multi(x, [0, 1, 2], masked_value=masked_value, set_zero_insignificant=True, confidence=0.99)
        for a, (mi, counts) in enumerate(got):
            ref_counts, _, ref_mi = bo.bitinformation(x, a, masked_value=masked_value)
            assert np.array_equal(counts, ref_counts), (masked_value, dirty, a)
            single = engine.bitinformation(_torch_view(x), a, masked_value=masked_value, set_zero_insignificant=True)
            assert np.array_equal(mi, single, equal_nan=True)


def test_multi_axis_counter_reuse_and_errors():
    rng = np.random.default_rng(2)
    x = smooth_field(rng, (40, 50, 64), "float32", axis=1, scale=0.2)
    t = torch.from_numpy(x).cuda()
    mc = MultiAxisCounter(t.dtype, 3, t.device)
    a = mc.bitinformation(t, [0, 1, 2]).clone()
    c = mc.counts.clone()
    b = mc.bitinformation(t, [0, 1, 2])
    assert torch.equal(c, mc.counts) and torch.equal(a, b)
    with pytest.raises(ValueError):
        mc.bitinformation(t, [0, 1, 1])
    with pytest.raises(ValueError):
        mc.bitinformation(t, [0, 1])
    # other widths fall back to one pass per axis inside the same call
    for dtype in ("float64", "float16", "uint8"):
        y = smooth_field(rng, (9, 30, 64), dtype, axis=1, scale=0.2, base=40.0) if dtype != "uint8" else \
            rng.integers(0, 255, size=(9, 30, 64), dtype="uint8")
        for a, (mi, counts) in enumerate(_multi(y, [0, 1, 2])):
            assert np.array_equal(counts, bo.bitinformation(y, a)[0]), (dtype, a)
