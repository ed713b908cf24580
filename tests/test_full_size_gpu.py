"""BASELINE.json's configurations at (or near) their full sizes, checked through properties
that do not need the CPU oracle to finish (it runs at ~0.01 GB/s):

* every 2x2 table sums to the number of valid pairs, computed independently with torch;
* the marginals: #(a_k = 1) - #(b_k = 1) = bits of the first elements - bits of the last
  elements along the analysed axis (every interior element is both an a and a b), with the
  signed-exponent transform of those two small slices taken from the oracle on the CPU;
* additivity: the counts of the whole array equal the sum over slabs of a non-analysed axis
  (accumulate-only counts), and over slabs of the analysed axis with a one-row halo;
* reversal: flipping the analysed axis swaps a and b, i.e. transposes every table;
* the fast kernels equal the generic kernel on a slab (the generic kernel is checked
  against the oracle bit for bit in test_counts_gpu.py);
* bitround: idempotent, trailing mantissa bits zero, equal to the oracle on a sample and to an
  independent frexp / round-half-even / ldexp formulation on large slabs.

Sizes are cut down (never the code path) when the device has less free memory than the
configuration needs.
"""
import numpy as np
import pytest
import torch

from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from oracle import keepbits_oracle as ko
from xbitinfo_b200 import _native as nat
from xbitinfo_b200.device import PairCounter, bitround_

pytestmark = pytest.mark.gpu


@pytest.fixture(autouse=True)
def _release_cached_device_memory():
    import gc

    gc.collect()
    torch.cuda.empty_cache()
    yield
    gc.collect()
    torch.cuda.empty_cache()


def free_bytes():
    free, _ = torch.cuda.mem_get_info()
    return free


def synth(shape, dtype, axis, seed):
    """base + cumulative sum of noise along `axis`, generated in bounded pieces on the device."""
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty(shape, dtype=dtype, device="cuda")
    axis %= len(shape)
    step = max(1, (512 << 20) // max(1, x[0].numel() * 4)) if axis != 0 else shape[0]
    if axis == 0:
        # accumulate along the leading axis plane by plane
        acc = torch.full(shape[1:], 280.0, dtype=torch.float32, device="cuda")
        for i in range(shape[0]):
            acc += torch.empty_like(acc).normal_(0.0, 0.3, generator=g)
            x[i] = acc.to(dtype)
        return x
    for lo in range(0, shape[0], step):
        blk = torch.empty((min(step, shape[0] - lo),) + tuple(shape[1:]), dtype=torch.float32, device="cuda")
        blk.normal_(0.0, 0.3, generator=g)
        torch.cumsum(blk, dim=axis, out=blk)
        blk += 280.0
        x[lo: lo + step] = blk.to(dtype)
    return x


def counts_of(x, axis, masked_value=None, force_generic=False, dims=None):
    pc = PairCounter(x.dtype, x.device)
    pc.add(x, axis, masked_value=masked_value, force_generic=force_generic, dims=dims)
    return pc.counts.clone()


def bits_of_slice(t, nbits):
    """per-bit population (most significant bit first) of the analysis words of a small tensor"""
    w = np.ascontiguousarray(bo.to_analysis_words(t.cpu().numpy())).reshape(-1)
    one = w.dtype.type(1)
    return np.array([np.count_nonzero(w & (one << w.dtype.type(nbits - 1 - k))) for k in range(nbits)], dtype=np.int64)


def check_tables(x, axis, counts, masked_value=None):
    """sum and marginal properties of the (nbits,2,2) tables of `x` along `axis`"""
    nbits = counts.shape[0]
    a = x.narrow(axis, 0, x.shape[axis] - 1)
    b = x.narrow(axis, 1, x.shape[axis] - 1)
    if masked_value is None:
        npairs = a.numel()
    elif masked_value != masked_value:
        npairs = int((~(torch.isnan(a) | torch.isnan(b))).sum())
    else:
        npairs = int((~((a == masked_value) | (b == masked_value))).sum())
    c = counts.cpu().numpy()
    assert (c.reshape(nbits, 4).sum(axis=1) == npairs).all()
    assert (c >= 0).all()
    if masked_value is None:
        first = bits_of_slice(x.narrow(axis, 0, 1), nbits)
        last = bits_of_slice(x.narrow(axis, x.shape[axis] - 1, 1), nbits)
        A = c[:, 1, 0] + c[:, 1, 1]
        B = c[:, 0, 1] + c[:, 1, 1]
        assert np.array_equal(A - B, first - last)
    return npairs


def test_config1_era5_like_float32_longitude():
    """configs[1]: float32 (8760, 721, 1440), dim='longitude' -- the bench workload."""
    T = 8760
    while T > 64 and T * 721 * 1440 * 4 * 2.2 > free_bytes():  # array + a flipped copy
        T //= 2
    x = synth((T, 721, 1440), torch.float32, -1, 1)
    before = nat.fast_path_launches()
    whole = counts_of(x, -1)
    assert nat.fast_path_launches() > before
    assert check_tables(x, -1, whole) == T * 721 * 1439
    # additivity over time slabs (the multi-GPU sharding rule), ragged split
    pc = PairCounter(x.dtype, x.device)
    for lo, hi in ((0, T // 3), (T // 3, T // 3 + 1), (T // 3 + 1, T)):
        pc.add(x[lo:hi], -1)
    assert torch.equal(pc.counts, whole)
    # the default API call masks NaN: on NaN-free data the result must not change
    assert torch.equal(counts_of(x, -1, masked_value=float("nan")), whole)
    # fast kernels == generic kernel on a slab
    s = x[: min(T, 96)]
    assert torch.equal(counts_of(s, -1), counts_of(s, -1, force_generic=True))
    # reversal along the analysed axis transposes every table
    y = torch.flip(x, dims=(-1,)).contiguous()
    assert torch.equal(counts_of(y, -1), whole.transpose(1, 2))
    del y
    # bit-exact at full size: the whole array back on the host, counted by the C restatement of the
    # reference's algorithm (oracle/fast_count.c, pinned to the numpy oracle and the golden vectors)
    import psutil

    from oracle import fast_count as fc

    if psutil.virtual_memory().available > 1.3 * x.numel() * 4:
        xh = x.cpu().numpy()
        del x
        assert np.array_equal(fc.count_pairs(xh, 2), whole.cpu().numpy()), "device counts differ from the CPU oracle"


@pytest.mark.parametrize("masked_value", [float("nan"), -999.0])
def test_config2_ocean_float64_latitude_with_land_mask(masked_value):
    """configs[2]: float64 (365, 75, 1080, 1440), land mask, dim='latitude' (strided
    adjacency).  340 GB does not fit one GPU: the path streams time slabs through the
    accumulate-only counts, which is what is exercised here on as many days as fit."""
    days = 24
    while days > 2 and days * 75 * 1080 * 1440 * 8 * 1.3 > free_bytes():
        days //= 2
    shape = (days, 75, 1080, 1440)
    x = synth(shape, torch.float64, 2, 2)
    g = torch.Generator(device="cuda").manual_seed(3)
    # land: a few rectangles and scattered cells, constant over time and level (~30 %)
    land = torch.rand((1080, 1440), device="cuda", generator=g) < 0.02
    land[200:700, 100:600] = True
    land[50:120, 900:1400] = True
    land[800:1000, 700:1100] = True
    x[:, :, land] = masked_value
    before = nat.fast_path_launches()
    whole = counts_of(x, 2, masked_value=masked_value)
    assert nat.fast_path_launches() > before
    npairs = check_tables(x, 2, whole, masked_value=masked_value)
    assert 0 < npairs < days * 75 * 1079 * 1440
    pc = PairCounter(x.dtype, x.device)
    for lo in range(0, days, 5):
        pc.add(x[lo: lo + 5], 2, masked_value=masked_value)
    assert torch.equal(pc.counts, whole)
    s = x[:1, :6]
    assert torch.equal(counts_of(s, 1, masked_value=masked_value),
                       counts_of(s, 1, masked_value=masked_value, force_generic=True))
    assert torch.equal(counts_of(s, 2, masked_value=masked_value),
                       counts_of(s, 2, masked_value=masked_value, force_generic=True))
    # splitting the analysed axis itself: slabs of latitude rows with a one-row halo
    v = x[0, 0]  # (1080, 1440): outer = 1, the slab rows are contiguous
    pc = PairCounter(x.dtype, x.device)
    for lo, hi in ((0, 400), (400, 401), (401, 1080)):
        rows = min(hi + 1, 1080) - lo  # + halo row of the next slab, none after the last
        pc.add(v[lo: lo + rows], 0, masked_value=masked_value)
    assert torch.equal(pc.counts, counts_of(v, 0, masked_value=masked_value))
    # bit-exact against the CPU oracle (oracle/fast_count.c) on one day of the field: all 75 levels at the
    # full 1080 x 1440 grid with its land mask, along latitude (the configured axis) and along longitude
    from oracle import fast_count as fc

    day = x[:1].contiguous()
    xh = day.cpu().numpy()
    for axis in (2, 3):
        assert np.array_equal(fc.count_pairs(xh, axis, masked_value), counts_of(day, axis, masked_value=masked_value).cpu().numpy()), axis


def test_config3_multi_variable_float32_every_dim():
    """configs[3]: float32 (720, 137, 256, 512) variables analysed along every dimension."""
    T = 720
    while T > 16 and T * 137 * 256 * 512 * 4 * 2.2 > free_bytes():
        T //= 2
    shape = (T, 137, 256, 512)
    x = synth(shape, torch.float32, 1, 4)
    for axis in range(4):
        before = nat.fast_path_launches()
        c = counts_of(x, axis)
        assert nat.fast_path_launches() > before, axis
        n = shape[axis]
        assert check_tables(x, axis, c) == x.numel() // n * (n - 1)
        other = 0 if axis != 0 else 1  # additivity over a non-analysed axis
        pc = PairCounter(x.dtype, x.device)
        half = shape[other] // 2
        pc.add(x.narrow(other, 0, half).contiguous(), axis)
        pc.add(x.narrow(other, half, shape[other] - half).contiguous(), axis)
        assert torch.equal(pc.counts, c), axis
    s = x[:2, :9].contiguous()
    for axis in range(4):
        assert torch.equal(counts_of(s, axis), counts_of(s, axis, force_generic=True)), axis
    # bit-exact against the CPU oracle (oracle/fast_count.c) along every dimension of a 2.3 GB time slab
    from oracle import fast_count as fc

    slab = x[:32].contiguous()
    xh = slab.cpu().numpy()
    for axis in range(4):
        assert np.array_equal(fc.count_pairs(xh, axis), counts_of(slab, axis).cpu().numpy()), axis


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32])
def test_config4_pipeline_share_of_one_gpu(dtype):
    """configs[4]: bitinfo -> keepbits -> bitround on one GPU's share (5e9 elements) of the
    4e10-element arrays."""
    n = 5_000_000_000
    itemsize = 2 if dtype == torch.float16 else 4
    while n > 50_000_000 and n * itemsize * 1.6 > free_bytes():
        n //= 2
    rows = n // 1000
    shape = (rows, 1000)
    scale, base = (0.02, 8.0)
    g = torch.Generator(device="cuda").manual_seed(5)
    x = torch.empty(shape, dtype=dtype, device="cuda")
    step = max(1, (512 << 20) // (shape[1] * 4))
    for lo in range(0, rows, step):
        blk = torch.empty((min(step, rows - lo), shape[1]), dtype=torch.float32, device="cuda")
        blk.normal_(0.0, scale, generator=g)
        torch.cumsum(blk, dim=1, out=blk)
        blk += base
        x[lo: lo + step] = blk.to(dtype)
    pc = PairCounter(x.dtype, x.device)
    pc.add(x, 1, masked_value=float("nan"))
    assert check_tables(x, 1, pc.counts, masked_value=float("nan")) == rows * (shape[1] - 1)
    info = pc.information(True, 0.99).cpu().numpy()
    np_dtype = "float16" if dtype == torch.float16 else "float32"
    keep = ko.keepbits_from_info(info, np_dtype, 0.99)
    mant = 10 if dtype == torch.float16 else 23
    assert keep <= mant  # (negative when not even the exponent bits are all needed: reference semantics)
    keep = max(1, min(keep, mant - 1))  # make the rounding non-trivial whatever the spectrum
    sample = x[:3].cpu().numpy()
    orig = {lo: x[lo: lo + 65536].clone() for lo in {0, max(0, min(rows // 2, rows - 1)), max(0, rows - 65536)}}
    bitround_(x, keep)
    # exact against the oracle on a sample, then size-independent properties on everything
    u = "u2" if dtype == torch.float16 else "u4"
    assert np.array_equal(x[:3].cpu().numpy().view(u), bro.bitround(sample, keep).view(u))
    ints = x.view(torch.int16 if dtype == torch.float16 else torch.int32)
    drop = mant - keep
    assert int((ints & ((1 << drop) - 1)).count_nonzero()) == 0  # trailing bits are zero
    # an independent formulation on the device, on slabs from the start, middle and end of the
    # array: x = m * 2^ex with m in [0.5, 1); round-half-even of m * 2^(keep+1), scaled back
    for lo in (0, rows // 2, rows - 65536):
        lo = max(0, min(lo, rows - 1))
        before_rounding = orig[lo] if lo in orig else None
        if before_rounding is None:
            continue
        m, ex = torch.frexp(before_rounding.double())
        want = torch.ldexp(torch.round(torch.ldexp(m, torch.tensor(keep + 1, device="cuda"))), ex - (keep + 1)).to(dtype)
        assert torch.equal(want, x[lo: lo + want.shape[0]]), lo
    y = x[: max(1, rows // 8)].clone()
    bitround_(y, keep)
    assert torch.equal(y, x[: max(1, rows // 8)])  # idempotent


def test_bench_line_carries_every_key_of_the_contract():
    """A short ``bench.py`` run (64 time steps of the workload's grid) prints one JSON line with the keys the
    driver and the judge read."""
    import json
    import os
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--time-steps", "64", "--steps", "3", "--warmup", "3",
                        "--e2e-steps", "1", "--no-cpu-baseline"], capture_output=True, text=True, cwd=root, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    line = json.loads(r.stdout.strip().splitlines()[-1])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "roofline", "cpu_baseline", "e2e", "gpu_launches", "clocks"):
        assert key in line, key
    assert line["n_gpus"] == 1 and line["steps"] == 3 and line["value"] > 0 and line["gpu_launches"] > 0
    for key in ("bound", "achieved", "peak", "unit", "frac", "traffic"):
        assert key in line["roofline"], key
    assert line["roofline"]["bound"] == "hbm" and 0 < line["roofline"]["frac"] < 1.5
    for key in ("value", "unit", "h2d_bytes_per_step", "d2h_bytes_per_step"):
        assert key in line["e2e"], key
    assert line["e2e"]["h2d_bytes_per_step"] >= 64 * 721 * 1440 * 4 and line["e2e"]["value"] > 0
    assert "workload" in line["config"] and "sm_mhz" in line["clocks"]
