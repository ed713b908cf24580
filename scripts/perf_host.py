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
# the pageable route (pinned ring + copy threads) must give what the direct route gives, also beyond 2^32 bytes
mi_a, c_a = engine.bitinformation(x, -1, return_counts=True)
mi_b, c_b = engine.bitinformation(pinned, -1, return_counts=True)
assert np.array_equal(c_a, c_b) and np.array_equal(mi_a, mi_b), "pageable and pinned routes disagree"
r_a, r_b = engine.bitround(x, 7), engine.bitround(pinned, 7)
assert np.array_equal(r_a.view("u4"), r_b.view("u4")), "bitround: pageable and pinned routes disagree"
probe = slice(0, x.shape[0], max(1, x.shape[0] // 7))
import importlib
bro = importlib.import_module("oracle.bitround_oracle")
assert np.array_equal(r_a[probe].view("u4"), bro.bitround(x[probe], 7).view("u4")), "bitround differs from the oracle"
del r_a, r_b
print(f"routes agree on {gb:.2f} GB ({x.nbytes} bytes)")
for name, arr in (("pageable", x), ("pinned", pinned)):
    t = timed(lambda: engine.bitinformation(arr, -1))
    print(f"bitinformation {name:9s} {gb:.2f} GB: {t*1e3:8.1f} ms = {gb/t:6.1f} GB/s")
    t = timed(lambda: engine.bitround(arr, 7))
    print(f"bitround       {name:9s} {gb:.2f} GB: {t*1e3:8.1f} ms = {gb/t:6.1f} GB/s (in) + same out")

# every dimension of a configs[3]-like variable: one PCIe pass for all axes vs one per axis
y = (287.0 + np.cumsum(rng.standard_normal((24, 137, 256, 512), dtype=np.float32) * np.float32(0.35), axis=-1)).astype(np.float32)
gb = y.nbytes / 1e9
t1 = timed(lambda: [engine.bitinformation(y, a, masked_value=float("nan"), set_zero_insignificant=True) for a in range(4)])
t2 = timed(lambda: engine.bitinformation_multi(y, [0, 1, 2, 3], masked_value=float("nan"), set_zero_insignificant=True))
print(f"all 4 dims of {y.shape} float32 ({gb:.2f} GB, pageable): axis by axis {t1*1e3:8.1f} ms, one pass {t2*1e3:8.1f} ms "
      f"= {4*gb/t2:6.1f} GB/s of (array x dims)")
