import sys, torch
sys.path.insert(0, ".")
from xbitinfo_b200.device import bitround_
for dt, n in ((torch.float32, 2**31), (torch.float64, 2**30), (torch.float16, 2**32)):
    y = torch.randn(n // 8, device="cuda", dtype=torch.float32).to(dt).repeat(8)
    for _ in range(3): bitround_(y, 7)
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); bitround_(y, 7); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    b = y.numel() * y.element_size()
    print(f"{str(dt):16s} {b/1e9:.1f} GB  best {min(ts):.3f} ms = {2*b/min(ts)/1e6:.0f} GB/s r+w, median {2*b/sorted(ts)[5]/1e6:.0f}")
    del y
