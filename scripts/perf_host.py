"""Host-buffer entry points (what the numpy-facing API calls): pageable vs pinned source arrays.

    python scripts/perf_host.py [--gb 4]
"""
import argparse
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from xbitinfo_b200 import engine

ap = argparse.ArgumentParser()
ap.add_argument("--gb", type=float, default=4.0)
args = ap.parse_args()
rows = max(1, int(args.gb * 1e9 / (721 * 1440 * 4)))
rng = np.random.default_rng(0)
x = (287.0 + np.cumsum(rng.standard_normal((rows, 721, 1440), dtype=np.float32) * np.float32(0.35), axis=-1)).astype(np.float32)
pinned_t = torch.empty(x.shape, dtype=torch.float32, pin_memory=True)
pinned = pinned_t.numpy()
pinned[...] = x


def timed(fn, reps=3):
    fn()
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return min(ts)


gb = x.nbytes / 1e9
for name, arr in (("pageable", x), ("pinned", pinned)):
    t = timed(lambda: engine.bitinformation(arr, -1))
    print(f"bitinformation {name:9s} {gb:.2f} GB: {t*1e3:8.1f} ms = {gb/t:6.1f} GB/s")
    t = timed(lambda: engine.bitround(arr, 7))
    print(f"bitround       {name:9s} {gb:.2f} GB: {t*1e3:8.1f} ms = {gb/t:6.1f} GB/s (in) + same out")
