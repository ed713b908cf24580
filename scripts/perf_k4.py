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

# fused K3+K4 against the two-pass sequence
from xbitinfo_b200.device import bitround_
from xbitinfo_b200 import shuffle as xs
for dtype in (torch.float32, torch.float64, torch.float16):
    es = torch.empty((), dtype=dtype).element_size()
    n = n_bytes // es
    x = torch.empty(n, dtype=dtype, device="cuda").normal_(280.0, 5.0)
    y = x.clone()
    out = torch.empty(n_bytes, dtype=torch.uint8, device="cuda")
    def two_pass():
        bitround_(y, 7)
        nat.check(nat.lib().xbi_shuffle(y.data_ptr(), out.data_ptr(), es, n, nat.SHUFFLE_BYTES, 0, _stream_ptr(x.device)))
    def fused(kind):
        nat.check(nat.lib().xbi_bitround_shuffle(x.data_ptr(), out.data_ptr(), {2: nat.F16, 4: nat.F32, 8: nat.F64}[es], n, 7, kind, _stream_ptr(x.device)))
    for name, fn in (("bitround then byte shuffle (2 passes)", two_pass), ("fused bitround+byte shuffle", lambda: fused(nat.SHUFFLE_BYTES)),
                     ("fused bitround+bit shuffle", lambda: fused(nat.SHUFFLE_BITS))):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print(f"{name:40s} {str(dtype):14s} {n_bytes/1e9:.2f} GB: {ms:.3f} ms = {n_bytes/ms/1e6:.0f} GB/s of input")
