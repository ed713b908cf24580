"""Randomised differential check of the core kernels against the CPU oracle (run on a GPU box): K1 on
device tensors (random shapes, axes, dtypes, masks, misaligned base pointers, both masked flavours, the
generic kernel), K2, K3 in place (odd sizes, misaligned), and the host-buffer entry points with a small
staging buffer (XBI_STAGING_MB=1 exercises every slab plan).

    XBI_STAGING_MB=1 python scripts/fuzz_k1.py [--seconds 60] [--seed 0] [--max-elems 400000]
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
sys.path.insert(0, "tests")
from oracle import bitinfo_oracle as bo
from oracle import bitround_oracle as bro
from xbitinfo_b200 import engine
from xbitinfo_b200.device import MultiAxisCounter, PairCounter, bitround, bitround_

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--max-elems", type=int, default=400_000)
ap.add_argument("--verbose", action="store_true")
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
DTYPES = ["float32", "float64", "float16", "int8", "uint8", "int16", "int32", "int64"]
SIZES = [1, 2, 3, 4, 7, 16, 25, 33, 64, 100, 127, 128, 129, 256, 480, 513, 1024, 1441, 4096]
TORCH = {"uint8": torch.uint8, "int8": torch.int8, "int16": torch.int16, "int32": torch.int32, "int64": torch.int64,
         "float16": torch.float16, "float32": torch.float32, "float64": torch.float64}


def random_array(shape, dtype):
    if np.dtype(dtype).kind == "f":
        x = (3.0 + np.cumsum(rng.standard_normal(shape) * 0.2, axis=int(rng.integers(len(shape))))).astype(dtype)
        r = rng.random()
        if r < 0.3:
            x[rng.random(shape) < rng.choice([0.001, 0.05, 0.5])] = np.nan
        elif r < 0.5 and len(shape) >= 2:  # a masked slab ("land")
            sl = [slice(None)] * len(shape)
            ax = int(rng.integers(len(shape)))
            lo = int(rng.integers(shape[ax]))
            sl[ax] = slice(lo, lo + int(rng.integers(1, shape[ax] + 1)))
            x[tuple(sl)] = np.nan
        if rng.random() < 0.1:
            x.flat[rng.integers(x.size)] = np.inf
        return x
    info = np.iinfo(dtype)
    return rng.integers(max(info.min, -40), min(info.max, 40), size=shape, dtype=dtype, endpoint=True)


t_end = time.time() + args.seconds
n_cases = 0
while time.time() < t_end:
    nd = int(rng.integers(1, 5))
    shape = tuple(int(v) for v in rng.choice(SIZES, size=nd))
    if np.prod(shape) > args.max_elems:
        continue
    dtype = str(rng.choice(DTYPES))
    x = random_array(shape, dtype)
    axis = int(rng.integers(nd))
    if np.dtype(dtype).kind == "f":
        mv = [None, float("nan"), float(x.flat[0]), -999.0][int(rng.integers(4))]
    else:
        mv = [None, int(x.flat[0])][int(rng.integers(2))]
    offset = int(rng.integers(0, 9))  # elements: misaligns the base pointer
    if args.verbose:
        print("case", n_cases, shape, dtype, "axis", axis, "mv", mv, "offset", offset, flush=True)
    want_c, npairs, want_mi = bo.bitinformation(x, axis, masked_value=mv)
    backing = torch.empty(x.size + 16, dtype=TORCH[dtype], device="cuda")
    t = backing[offset:offset + x.size].view(shape)
    t.copy_(torch.from_numpy(x))
    pc = PairCounter(t.dtype, t.device)
    for kw in ({}, {"force_generic": True}) + (({"flavour": "sparse"}, {"flavour": "dense"}) if mv is not None else ()):
        pc.reset()
        pc.add(t, axis, masked_value=mv, **kw)
        assert np.array_equal(pc.counts.cpu().numpy(), want_c), ("K1", shape, dtype, axis, mv, offset, kw)
    # the fused entry point (xbi_bitinformation_dev: counts overwritten, K2 in the finish kernel) ...
    pc.counts.fill_(-7)
    mi_f = pc.bitinformation(t, axis, masked_value=mv).cpu().numpy()
    assert np.array_equal(pc.counts.cpu().numpy(), want_c), ("fused", shape, dtype, axis, mv, offset)
    # ... and every axis of the array in one call (the two innermost share one read where the kernel applies)
    if nd >= 2 and rng.random() < 0.5:
        axes = [int(a) for a in rng.permutation(nd)[: int(rng.integers(2, nd + 1))]]
        mc = MultiAxisCounter(t.dtype, len(axes), t.device)
        mc.bitinformation(t, axes, masked_value=mv)
        got = mc.counts.cpu().numpy()
        for i, a in enumerate(axes):
            assert np.array_equal(got[i], bo.bitinformation(x, a, masked_value=mv)[0]), ("multi", shape, dtype, axes, a, mv, offset)
    pc.reset()
    pc.add(t, axis, masked_value=mv)
    mi = pc.information(False).cpu().numpy()
    assert np.array_equal(mi, mi_f, equal_nan=True), ("fused K2", shape, dtype, axis, mv)
    if np.isnan(want_mi).all():
        assert np.isnan(mi).all()
    else:
        scale = max(np.nanmax(np.abs(want_mi)), 1e-300)
        np.testing.assert_allclose(mi, want_mi, rtol=1e-12, atol=1e-12 * scale)
    # host-buffer entry point (slab plans depend on XBI_STAGING_MB)
    mi_h, c_h = engine.bitinformation(x, axis, masked_value=mv, return_counts=True)
    assert np.array_equal(c_h, want_c), ("host", shape, dtype, axis, mv)
    if np.dtype(dtype).kind == "f":
        mant = {"float16": 10, "float32": 23, "float64": 52}[dtype]
        kb = int(rng.integers(0, mant + 1))
        u = f"u{x.dtype.itemsize}"
        want_r = bro.bitround(x, kb).view(u)
        r = t.clone() if rng.random() < 0.5 else None
        if r is None:  # in place on the misaligned view
            bitround_(t, kb)
            r = t
        else:
            bitround_(r, kb)
        assert np.array_equal(r.cpu().numpy().view(u), want_r), ("K3", shape, dtype, kb, offset)
        t.copy_(torch.from_numpy(x))
        o = bitround(t, kb)  # out of place: the input stays as it is
        assert np.array_equal(o.cpu().numpy().view(u), want_r) and np.array_equal(t.cpu().numpy().view(u), x.view(u)), \
            ("K3 out of place", shape, dtype, kb, offset)
        assert np.array_equal(engine.bitround(x, kb).view(u), want_r), ("K3 host", shape, dtype, kb)
    n_cases += 1
print(f"fuzz ok: {n_cases} random cases in {args.seconds:.0f} s (seed {args.seed})")
