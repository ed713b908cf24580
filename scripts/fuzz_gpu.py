"""Randomised differential check of the late-round kernels against the CPU oracle (run on a GPU box):
chunk-grid counting / rounding (regular and irregular grids), the multi-axis host entry point, K4 inverses.

    python scripts/fuzz_gpu.py [--seconds 60] [--seed 0]
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
from oracle import shuffle_oracle as so
from xbitinfo_b200 import engine
from xbitinfo_b200 import shuffle as xs
from xbitinfo_b200.chunked import normalize_chunks, slices_from_chunks
from xbitinfo_b200.device import bitround_chunked_, count_pairs_chunked

ap = argparse.ArgumentParser()
ap.add_argument("--seconds", type=float, default=60.0)
ap.add_argument("--seed", type=int, default=0)
ap.add_argument("--verbose", action="store_true", help="print every case before it runs")
args = ap.parse_args()
rng = np.random.default_rng(args.seed)
DTYPES = ["float32", "float64", "float16", "int8", "uint16", "int32", "int64"]


def random_array(shape, dtype):
    if np.dtype(dtype).kind == "f":
        x = (3.0 + np.cumsum(rng.standard_normal(shape) * 0.2, axis=int(rng.integers(len(shape))))).astype(dtype)
        if rng.random() < 0.5:
            x[rng.random(shape) < rng.choice([0.01, 0.2])] = np.nan
        return x
    info = np.iinfo(dtype)
    return rng.integers(max(info.min, -50), min(info.max, 50), size=shape, dtype=dtype, endpoint=True)


def random_spec(shape):
    spec = []
    for n in shape:
        kind = rng.integers(3)
        if kind == 0 or n == 1:
            spec.append(int(n))
        elif kind == 1:
            spec.append(int(rng.integers(1, n + 1)))
        else:
            cuts = np.sort(rng.choice(np.arange(1, n), size=int(rng.integers(0, min(n - 1, 4) + 1)), replace=False))
            spec.append(tuple(int(v) for v in np.diff(np.concatenate([[0], cuts, [n]]))))
    return tuple(spec)


t_end = time.time() + args.seconds
n_cases = 0
while time.time() < t_end:
    nd = int(rng.integers(1, 6))
    shape = tuple(int(v) for v in rng.choice([1, 2, 3, 5, 8, 13, 31, 64, 100], size=nd))
    if np.prod(shape) > 300_000:
        continue
    dtype = str(rng.choice(DTYPES))
    x = random_array(shape, dtype)
    axis = int(rng.integers(nd))
    spec = random_spec(shape)
    mv = None
    if np.dtype(dtype).kind == "f":
        mv = [None, float("nan"), float(x.flat[0])][int(rng.integers(3))]
    elif rng.random() < 0.5:
        mv = int(x.flat[0])
    if args.verbose:
        print("case", n_cases, shape, dtype, "axis", axis, "spec", spec, "mv", mv, flush=True)
    t = torch.from_numpy(x).cuda()
    got = count_pairs_chunked(t, spec, axis, masked_value=mv).cpu().numpy()
    # both chunk-grid kernels whatever the planner picks: the column walkers and the per-element odometers
    got_walk = count_pairs_chunked(t, spec, axis, masked_value=mv, flavour="walk").cpu().numpy()
    got_walk2 = count_pairs_chunked(t, spec, axis, masked_value=mv, flavour="walk2").cpu().numpy()
    got_walk3 = count_pairs_chunked(t, spec, axis, masked_value=mv, flavour="walk3").cpu().numpy()
    got_odo = count_pairs_chunked(t, spec, axis, masked_value=mv, flavour="odometer").cpu().numpy()
    grid = normalize_chunks(shape, spec)
    gshape = tuple(len(c) for c in grid)
    for idx, sl in zip(np.ndindex(*gshape), slices_from_chunks(grid)):
        want = bo.bitinformation(x[sl], axis, masked_value=mv)[0]
        assert np.array_equal(got[idx], want), ("counts", shape, dtype, spec, axis, mv, idx)
        assert np.array_equal(got_walk[idx], want), ("counts (walkers)", shape, dtype, spec, axis, mv, idx)
        assert np.array_equal(got_walk2[idx], want), ("counts (walkers, two streams)", shape, dtype, spec, axis, mv, idx)
        assert np.array_equal(got_walk3[idx], want), ("counts (walkers, three streams)", shape, dtype, spec, axis, mv, idx)
        assert np.array_equal(got_odo[idx], want), ("counts (odometers)", shape, dtype, spec, axis, mv, idx)
    if np.dtype(dtype).kind == "f":
        mant = {"float16": 10, "float32": 23, "float64": 52}[dtype]
        keep = rng.integers(0, mant, size=gshape, endpoint=True).astype("int32")
        r = t.clone()
        bitround_chunked_(r, spec, torch.from_numpy(keep.ravel()).cuda())
        torch.cuda.synchronize()
        r = r.cpu().numpy()
        u = f"u{x.dtype.itemsize}"
        for idx, sl in zip(np.ndindex(*gshape), slices_from_chunks(grid)):
            want = bro.bitround(np.ascontiguousarray(x[sl]), int(keep[idx]))
            assert np.array_equal(np.ascontiguousarray(r[sl]).view(u), want.view(u)), ("round", shape, dtype, spec, idx)
    # every axis in one pass over the host array
    axes = [int(a) for a in rng.permutation(nd)]
    for a, (mi, c) in zip(axes, engine.bitinformation_multi(x, axes, masked_value=mv, return_counts=True)):
        assert np.array_equal(c, bo.bitinformation(x, a, masked_value=mv)[0]), ("multi", shape, dtype, a, mv)
    # K4 round trips and layouts on a flat view (float types only; bit shuffle needs a multiple of 8)
    if np.dtype(dtype).kind == "f":
        flat = np.ascontiguousarray(x).ravel()
        flat = flat[: flat.size - flat.size % 8]
        if flat.size:
            tt = torch.from_numpy(flat).cuda()
            for fwd, inv, ref in ((xs.byte_shuffle, xs.byte_unshuffle, so.byte_shuffle), (xs.bit_shuffle, xs.bit_unshuffle, so.bit_shuffle)):
                sh = fwd(tt)
                assert np.array_equal(sh.cpu().numpy(), ref(flat)), ("shuffle", fwd.__name__, dtype, flat.size)
                back = inv(sh, tt.dtype)
                assert np.array_equal(back.cpu().numpy().view(f"u{flat.dtype.itemsize}"), flat.view(f"u{flat.dtype.itemsize}")), ("unshuffle", dtype, flat.size)
    n_cases += 1
print(f"fuzz ok: {n_cases} random cases in {args.seconds:.0f} s (seed {args.seed})")
