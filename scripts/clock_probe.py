import os, sys, time, threading
import torch
sys.path.insert(0, ".")
from xbitinfo_b200.device import PairCounter
import pynvml
pynvml.nvmlInit(); h = pynvml.nvmlDeviceGetHandleByIndex(0)
shape = tuple(int(s) for s in sys.argv[1].split(",")); dtype = getattr(torch, sys.argv[2]); axis = int(sys.argv[3])
x = (torch.cumsum(torch.randn(shape, device="cuda"), dim=axis) * 0.3 + 280).to(dtype)
pc = PairCounter(x.dtype, x.device)
samples = []; stop = False
def poll():
    while not stop:
        samples.append((pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM), pynvml.nvmlDeviceGetPowerUsage(h) / 1000.0,
                        pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)))
        time.sleep(0.05)
for _ in range(3): pc.add(x, axis)
torch.cuda.synchronize()
t = threading.Thread(target=poll); t.start()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(sys.argv[4]); e0.record()
for _ in range(n): pc.add(x, axis)
e1.record(); torch.cuda.synchronize(); stop = True; t.join()
ms = e0.elapsed_time(e1) / n
clk = sorted(s[0] for s in samples); pw = sorted(s[1] for s in samples)
print(f"{os.environ.get('XBI_COLS_LDG','TMA')=} {ms:.3f} ms/iter {x.numel()*x.element_size()/ms/1e6:.0f} GB/s  clk med {clk[len(clk)//2]} min {clk[0]}  power med {pw[len(pw)//2]:.0f} max {pw[-1]:.0f} W reasons {set(hex(s[2]) for s in samples)} n={len(samples)}")
