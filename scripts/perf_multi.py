"""Times xbi_bitinformation_dev_multi on a configs[3]-like float32 (T,137,256,512) variable: all four axes, the two
innermost in one read, and each axis on its own.   python scripts/perf_multi.py [T] [masked: nan|none]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xbitinfo_b200.device import MultiAxisCounter, PairCounter  # noqa: E402


def timeit(fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    T = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mv = None if (len(sys.argv) > 2 and sys.argv[2] == "none") else float("nan")
    dev = torch.device("cuda", 0)
    x = torch.empty((T, 137, 256, 512), device=dev)
    for t in range(T):
        x[t].normal_(0.0, 0.3)
    torch.cumsum(x, dim=3, out=x)
    x += 12.0
    gb = x.numel() * 4 / 1e9
    kw = dict(masked_value=mv, set_zero_insignificant=True, confidence=0.99)
    m4, m2, pc = MultiAxisCounter(x.dtype, 4, dev), MultiAxisCounter(x.dtype, 2, dev), PairCounter(x.dtype, dev)
    ms = timeit(lambda: m4.bitinformation(x, [0, 1, 2, 3], **kw))
    print(f"all four axes, one call: {ms:.3f} ms ({gb / ms * 1e3:.0f} GB/s of array bytes)")
    ms = timeit(lambda: m2.bitinformation(x, [2, 3], **kw))
    print(f"two innermost axes, one read: {ms:.3f} ms ({gb / ms * 1e3:.0f} GB/s of array bytes)")
    for a in range(4):
        ms = timeit(lambda a=a: pc.bitinformation(x, a, **kw))
        print(f"axis {a} alone: {ms:.3f} ms ({gb / ms * 1e3:.0f} GB/s)")


if __name__ == "__main__":
    main()
