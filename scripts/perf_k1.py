"""Quick K1/K3 timing on one GPU (development aid; bench.py is the judged harness)."""
import argparse
import os
import sys
import time

import torch

sys.path.insert(0, ".")
from xbitinfo_b200 import _native as nat
from xbitinfo_b200.device import PairCounter, bitround_


def synth(shape, dtype, axis=-1, seed=1234):
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = torch.empty(shape, dtype=torch.float32, device="cuda")
    x.normal_(0.0, 0.3, generator=g)
    x = torch.cumsum(x, dim=axis, out=x)
    x += 280.0
    return x.to(dtype)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    if os.environ.get('XBI_PERF_VERBOSE'):
        print('   iters ms:', ' '.join(f'{t*1e3:.3f}' for t in ts))
    return min(ts), sorted(ts)[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shape", default="1024,721,1440")
    ap.add_argument("--dtype", default="float32")
    ap.add_argument("--axis", type=int, default=-1)
    ap.add_argument("--mask", default="none")
    ap.add_argument("--land", type=float, default=0.0, help="fraction of the last two dims set to the masked value (a rectangle)")
    ap.add_argument("--randmask", type=float, default=0.0, help="fraction of elements masked at random positions")
    ap.add_argument("--flavour", default=None, choices=[None, "sparse", "dense"])
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--bitround", action="store_true")
    a = ap.parse_args()
    shape = tuple(int(s) for s in a.shape.split(","))
    dtype = getattr(torch, a.dtype)
    if dtype.is_floating_point:
        x = synth(shape, dtype, a.axis)
    else:
        x = torch.randint(-2**30, 2**30, shape, device="cuda", dtype=torch.int64).to(dtype)
    mv = None if a.mask == "none" else float(a.mask)
    if a.land > 0 and mv is not None:
        h, w = x.shape[-2], x.shape[-1]
        x[..., int(0.2 * h): int(0.7 * h), : int(min(1.0, a.land / 0.5) * w)] = mv
        print(f"masked fraction: {float((x != x).float().mean() if mv != mv else (x == mv).float().mean()):.3f}")
    if a.randmask > 0 and mv is not None:
        g = torch.Generator(device="cuda").manual_seed(7)
        sel = torch.rand(x.shape, device="cuda", generator=g) < a.randmask
        x[sel] = mv
        del sel
    nbytes = x.numel() * x.element_size()
    pc = PairCounter(x.dtype, x.device)
    l0 = nat.kernel_launches()
    pc.add(x, a.axis, masked_value=mv, flavour=a.flavour)
    torch.cuda.synchronize()
    print("launches per call:", nat.kernel_launches() - l0)
    best, med = timeit(lambda: pc.add(x, a.axis, masked_value=mv, flavour=a.flavour))
    nat.profile_enable(True)
    for _ in range(10):
        pc.add(x, a.axis, masked_value=mv, flavour=a.flavour)
    kms, kn, kb = nat.profile_collect()
    nat.profile_enable(False)
    if kn:
        print(f"   fast-path kernel alone: {kms/kn:.3f} ms avg over {kn} launches = {kb/kn/(kms/kn)/1e6:.0f} GB/s")
    print(f"K1 {a.dtype} {shape} axis={a.axis} mask={a.mask}: best {best*1e3:.3f} ms  median {med*1e3:.3f} ms  "
          f"{nbytes/best/1e9:.1f} GB/s best, {nbytes/med/1e9:.1f} GB/s median ({nbytes/1e9:.2f} GB)")
    if a.check:
        pc.reset(); pc.add(x, a.axis, masked_value=mv, flavour=a.flavour); fast = pc.counts.clone()
        pc.reset(); pc.add(x, a.axis, masked_value=mv, force_generic=True); gen = pc.counts.clone()
        bg, _ = timeit(lambda: pc.add(x, a.axis, masked_value=mv, force_generic=True), iters=2, warm=1)
        print("fast == generic:", bool(torch.equal(fast, gen)), f"(generic {nbytes/bg/1e9:.1f} GB/s)")
        print("pairs:", int(fast[0].sum()), "expected", x.numel() // shape[a.axis] * (shape[a.axis] - 1))
    if a.bitround and dtype.is_floating_point:
        y = x.clone()
        best, med = timeit(lambda: bitround_(y, 7))
        print(f"K3 bitround: best {best*1e3:.3f} ms  {2*nbytes/best/1e9:.1f} GB/s (read+write)")


if __name__ == "__main__":
    main()
