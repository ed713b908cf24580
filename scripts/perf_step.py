"""Where does the fixed per-step cost go?  Times the fused step (xbi_bitinformation_dev) on arrays chosen to
isolate its parts (CUDA events, 200 calls each, queue kept full so that launch gaps -- not host latency -- show):

    nothing   n = 1: workspace clear + finish kernel with empty ranges      -> the floor of one call
    fringe    7,000 elements in one row: no TMA stage, all pairs in the finish kernel
    edges     the same with rows of 100 elements (69 row ends for the finish kernel)
    stage+31  one TMA stage and 31 leftover pairs;  148 stages: one stage per SM, 3 leftover pairs
    full      (T,721,1440): the bench step; also reports step - kernel

    python scripts/perf_step.py [T]
"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitinfo_b200 import _native as nat  # noqa: E402
from xbitinfo_b200.device import PairCounter  # noqa: E402


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3  # us


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 1095
    dev = torch.device("cuda", 0)
    pc = PairCounter(torch.float32, dev)
    cases = {
        "nothing": torch.ones((4, 1), device=dev),
        "fringe": torch.randn((1, 7000), device=dev),
        "edges": torch.randn((70, 100), device=dev),
        "stage+31": torch.randn((1, 7680 + 32), device=dev),
        "148 stages": torch.randn((1, 148 * 7680 + 4), device=dev),
    }
    for name, x in cases.items():
        us = timeit(lambda: pc.bitinformation(x, -1, masked_value=None))
        us_cnt = timeit(lambda: pc.bitinformation(x, -1, masked_value=None, with_information=False))
        print(f"{name:12s} shape {tuple(x.shape)}: {us:7.2f} us per call   (counts only: {us_cnt:7.2f})")
    ws = pc.workspace
    us = timeit(lambda: ws.zero_())
    print(f"torch zero_ of the {ws.numel()}-byte workspace alone: {us:.2f} us")
    x = torch.randn((T, 721, 1440), device=dev)
    for mv in (None, float("nan")):
        nat.profile_enable(True)
        reps = 30
        us = timeit(lambda: pc.bitinformation(x, -1, masked_value=mv), reps=reps)
        kms, kn, kb = nat.profile_collect()
        nat.profile_enable(False)
        kus = kms / kn * 1e3
        print(f"full ({T},721,1440) masked_value={mv}: step {us:.1f} us, kernel {kus:.1f} us, step - kernel {us - kus:.1f} us "
              f"({x.numel() * 4 / us / 1e3:.0f} GB/s)")
        us2 = timeit(lambda: pc.bitinformation(x, -1, masked_value=mv), reps=reps)
        print(f"   the same without the event pairs: step {us2:.1f} us")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        pc.bitinformation(x, -1, masked_value=None)
    us = timeit(g.replay, reps=30)
    print(f"full, CUDA graph replay: {us:.1f} us")


if __name__ == "__main__":
    main()
