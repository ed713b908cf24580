"""K4 timing (byte / bit shuffle, forward and inverse) on one GPU."""
import sys

import torch

sys.path.insert(0, ".")
from xbitinfo_b200 import _native as nat
from xbitinfo_b200.device import _stream_ptr

n_bytes = int(float(sys.argv[1])) if len(sys.argv) > 1 else 4 << 30
for dtype in (torch.float32, torch.float64, torch.float16):
    es = torch.empty((), dtype=dtype).element_size()
    n = n_bytes // es
    x = torch.empty(n, dtype=dtype, device="cuda").normal_(280.0, 5.0)
    out = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    back = torch.empty_like(x)
    for kind, name in ((nat.SHUFFLE_BYTES, "byte"), (nat.SHUFFLE_BITS, "bit")):
        for inverse in (0, 1):
            src, dst = (x, out) if not inverse else (out, back)
            def run():
                nat.check(nat.lib().xbi_shuffle(src.data_ptr(), dst.data_ptr(), es, n, kind, inverse, _stream_ptr(x.device)))
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(10):
                run()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 10
            print(f"K4 {name:4s} {'inverse' if inverse else 'forward'} {str(dtype):14s} {n_bytes/1e9:.2f} GB: {ms:.3f} ms  "
                  f"{2*n_bytes/ms/1e6:.0f} GB/s read+write")
        if kind == nat.SHUFFLE_BITS:
            assert torch.equal(back, x)
