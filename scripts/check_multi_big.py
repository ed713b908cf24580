import sys, time
import numpy as np
sys.path.insert(0, ".")
from xbitinfo_b200 import engine
rng = np.random.default_rng(5)
y = (287.0 + np.cumsum(rng.standard_normal((66, 137, 256, 512), dtype=np.float32) * np.float32(0.35), axis=-1)).astype(np.float32)
y[:, :, 100:140, 200:300] = np.nan
print(y.nbytes, y.nbytes > 2**32)
t0 = time.perf_counter()
multi = engine.bitinformation_multi(y, [0, 1, 2, 3], masked_value=float("nan"), set_zero_insignificant=True, return_counts=True)
t1 = time.perf_counter()
for a in range(4):
    mi, c = engine.bitinformation(y, a, masked_value=float("nan"), set_zero_insignificant=True, return_counts=True)
    assert np.array_equal(c, multi[a][1]) and np.array_equal(mi, multi[a][0], equal_nan=True), a
    assert int(c[0].sum()) == int(c[5].sum())
print("multi == per axis on", y.nbytes / 1e9, "GB;", (t1 - t0) * 1e3, "ms for the one-pass call")
