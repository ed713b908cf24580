"""Per-chunk workflow (docs/chunking.ipynb) on the device: batched kernels vs the chunk-by-chunk route.

    python scripts/perf_chunked.py [--shape 744,721,1440] [--dtype float32] [--loop]
"""
import argparse
import sys
import time

import torch

sys.path.insert(0, ".")
from xbitinfo_b200 import _native as nat
from xbitinfo_b200.chunked import bitround_by_chunks
from xbitinfo_b200.device import bitround_chunked_, count_pairs_chunked, information_batched

ap = argparse.ArgumentParser()
ap.add_argument("--shape", default="744,721,1440")
ap.add_argument("--dtype", default="float32")
ap.add_argument("--loop", action="store_true", help="also time the chunk-by-chunk route (slow for many chunks)")
ap.add_argument("--compare", action="store_true", help="only K1: default kernel vs the three-stream walkers (walk3)")
ap.add_argument("--nan", default=None, help="land: 30 %% NaN rectangle; a number: that fraction of NaNs at random")
ap.add_argument("--chunk", default=None, help="only this chunk shape (comma separated), e.g. for an ncu capture")
args = ap.parse_args()
shape = tuple(int(v) for v in args.shape.split(","))
dt = getattr(torch, args.dtype)
g = torch.Generator(device="cuda").manual_seed(1)
x = torch.empty(shape, dtype=torch.float32, device="cuda").normal_(0, 0.35, generator=g)
torch.cumsum(x, dim=-1, out=x)
x += 287.0
x = x.to(dt)
nbytes = x.numel() * x.element_size()
if args.nan == "land":
    x[:, : int(shape[1] * 0.55), : int(shape[2] * 0.55)] = float("nan")
elif args.nan:
    x[torch.rand(shape, device="cuda", generator=g) < float(args.nan)] = float("nan")


def timed(fn, reps=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts)


grids = [
    (shape[0], shape[1], shape[2]),
    (shape[0], -(-shape[1] // 7), -(-shape[2] // 8)),
    (shape[0], 32, 32),
    (shape[0], 10, 10),
    (shape[0], 4, 4),
    (max(1, shape[0] // 12), 64, 64),
]
if args.chunk:
    grids = [tuple(int(v) for v in args.chunk.split(","))]
print(f"{args.dtype} {shape}: {nbytes/1e9:.2f} GB")
for chunk in grids:
    for axis in (0, len(shape) - 1):
        if args.compare:
            ref = count_pairs_chunked(x, chunk, axis, masked_value=float("nan"), flavour="walk3")
            new = count_pairs_chunked(x, chunk, axis, masked_value=float("nan"))
            same = bool(torch.equal(ref, new))
            t_new = timed(lambda: count_pairs_chunked(x, chunk, axis, masked_value=float("nan")))
            t_two = timed(lambda: count_pairs_chunked(x, chunk, axis, masked_value=float("nan"), flavour="walk2"))
            t_old = timed(lambda: count_pairs_chunked(x, chunk, axis, masked_value=float("nan"), flavour="walk3"))
            t_plain = timed(lambda: count_pairs_chunked(x, chunk, axis, masked_value=None))
            print(f"chunk {str(chunk):22s} axis {axis}: default {nbytes/t_new/1e6:6.0f} GB/s | two-stream walkers {nbytes/t_two/1e6:6.0f} GB/s | three-stream walkers "
                  f"{nbytes/t_old/1e6:6.0f} GB/s | no mask {nbytes/t_plain/1e6:6.0f} GB/s | equal {same}", flush=True)
            continue
        counts = count_pairs_chunked(x, chunk, axis, masked_value=float("nan"))
        nchunks = counts.numel() // (counts.shape[-3] * 4)
        t_count = timed(lambda: count_pairs_chunked(x, chunk, axis, masked_value=float("nan")))
        t_mi = timed(lambda: information_batched(counts, True, 0.99))
        keep = torch.full((nchunks,), 7, dtype=torch.int32, device="cuda")
        y = x.clone()
        t_round = timed(lambda: bitround_chunked_(y, chunk, keep))
        del y
        t0 = time.perf_counter()
        n0 = nat.kernel_launches()
        bitround_by_chunks(x, chunk, axis)
        torch.cuda.synchronize()
        t_all = (time.perf_counter() - t0) * 1e3
        launches = nat.kernel_launches() - n0
        line = (f"chunk {str(chunk):22s} axis {axis} {nchunks:7d} chunks: count {t_count:8.3f} ms = {nbytes/t_count/1e6:6.0f} GB/s"
                f" | MI {t_mi:6.3f} ms | round {t_round:8.3f} ms = {2*nbytes/t_round/1e6:6.0f} GB/s r+w"
                f" | whole workflow {t_all:9.2f} ms wall, {launches} launches")
        if args.loop and nchunks <= 20000:
            t0 = time.perf_counter()
            n0 = nat.kernel_launches()
            bitround_by_chunks(x, chunk, axis, batched=False)
            torch.cuda.synchronize()
            line += f" | chunk-by-chunk {(time.perf_counter()-t0)*1e3:9.1f} ms wall, {nat.kernel_launches()-n0} launches"
        print(line, flush=True)
