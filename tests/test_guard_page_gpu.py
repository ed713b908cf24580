"""Out-of-bounds reads do not change results, so parity tests cannot see them -- and
compute-sanitizer is not available on the GPU pool.  Here the input arrays are placed with the
CUDA virtual-memory API so that they END exactly at the end of the mapped range, followed by
reserved but UNMAPPED address space (and, in the second variant, START exactly at the start
of the mapping, preceded by unmapped space).  Any kernel that reads a byte past either end
faults with an illegal-address error instead of silently reading a neighbour's memory.

Covered: the TMA rows kernel (stage + halo copies, block-0 corrections), the cp.async column
walkers (aligned and element-aligned rows, tail groups, shadow lanes), the generic kernel
(heads, tails, edges), K3 and K4 -- at shapes chosen to sit on and around their tile sizes.
The counts are compared with the same call on an ordinary torch allocation.
"""
import ctypes as C

import numpy as np
import pytest
import torch

from xbitinfo_b200 import _native as nat
from xbitinfo_b200.device import PairCounter
from xbitinfo_b200.streaming import _TORCH_OF_NP

from conftest import smooth_field

pytestmark = pytest.mark.gpu

drv = pytest.importorskip("cuda.bindings.driver")


def _ok(res):
    err = res[0]
    if int(err) != 0:
        raise RuntimeError(f"CUDA driver error {err}")
    return res[1] if len(res) == 2 else res[1:]


class GuardedBuffer:
    """`nbytes` of device memory that end (where='end') or start (where='start') on the edge of
    the mapped range; the other side of that edge is reserved, unmapped address space."""

    def __init__(self, nbytes, where="end"):
        torch.cuda.init()
        torch.zeros(1, device="cuda")  # make sure the primary context is current
        dev = torch.cuda.current_device()
        prop = drv.CUmemAllocationProp()
        prop.type = drv.CUmemAllocationType.CU_MEM_ALLOCATION_TYPE_PINNED
        prop.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        prop.location.id = dev
        gran = int(_ok(drv.cuMemGetAllocationGranularity(prop, drv.CUmemAllocationGranularity_flags.CU_MEM_ALLOC_GRANULARITY_MINIMUM)))
        self.mapped = max(gran, (nbytes + gran - 1) // gran * gran)
        self.reserved = self.mapped + 2 * gran  # one unmapped granule on each side
        self.base = int(_ok(drv.cuMemAddressReserve(self.reserved, gran, 0, 0)))
        self.handle = _ok(drv.cuMemCreate(self.mapped, prop, 0))
        self.map_at = self.base + gran
        _ok(drv.cuMemMap(self.map_at, self.mapped, 0, self.handle, 0) + (None,))
        acc = drv.CUmemAccessDesc()
        acc.location.type = drv.CUmemLocationType.CU_MEM_LOCATION_TYPE_DEVICE
        acc.location.id = dev
        acc.flags = drv.CUmemAccess_flags.CU_MEM_ACCESS_FLAGS_PROT_READWRITE
        _ok(drv.cuMemSetAccess(self.map_at, self.mapped, [acc], 1) + (None,))
        self.nbytes = nbytes
        self.ptr = self.map_at + (self.mapped - nbytes if where == "end" else 0)

    def upload(self, arr):
        arr = np.ascontiguousarray(arr)
        assert arr.nbytes == self.nbytes
        _ok(drv.cuMemcpyHtoD(self.ptr, arr.ctypes.data, arr.nbytes) + (None,))

    def download(self, dtype, shape):
        out = np.empty(shape, dtype=dtype)
        assert out.nbytes == self.nbytes
        _ok(drv.cuMemcpyDtoH(out.ctypes.data, self.ptr, out.nbytes) + (None,))
        return out

    def close(self):
        torch.cuda.synchronize()
        _ok(drv.cuMemUnmap(self.map_at, self.mapped) + (None,))
        _ok(drv.cuMemRelease(self.handle) + (None,))
        _ok(drv.cuMemAddressFree(self.base, self.reserved) + (None,))


def _count_at(ptr, x, axis, masked_value):
    """K1 on raw device memory at `ptr` holding the bytes of x."""
    pc = PairCounter(_TORCH_OF_NP[x.dtype.name], "cuda")
    outer = int(np.prod(x.shape[:axis], dtype=np.int64))
    n = x.shape[axis]
    inner = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    flags, bits = nat.mask_spec(x.dtype, masked_value)
    if x.dtype.kind == "f":
        flags |= nat.SIGNED_EXPONENT
    rc = nat.lib().xbi_count_pairs(ptr, nat.dtype_code(x.dtype), outer, n, inner, flags, bits, pc.counts.data_ptr(),
                                   pc.workspace.data_ptr(), pc.workspace.numel(), None)
    nat.check(rc)
    torch.cuda.synchronize()  # an illegal address surfaces here
    return pc.counts.cpu().numpy()


def _count_plain(x, axis, masked_value):
    t = torch.from_numpy(x).cuda()
    pc = PairCounter(t.dtype, t.device)
    pc.add(t, axis, masked_value=masked_value)
    return pc.counts.cpu().numpy()


CASES = [
    # dtype, shape, axis: rows kernel on and around its 7680-element (float32) stages
    ("float32", (4, 7680), 1), ("float32", (3, 7681), 1), ("float32", (5, 7679), 1), ("float32", (1, 30720 + 4), 1),
    ("float32", (9, 1923), 1), ("float64", (3, 3840 * 2 + 1), 1), ("float16", (2, 15360 + 8), 1), ("uint8", (2, 30720 + 16), 1),
    ("float32", (1, 100), 1), ("float64", (7, 33), 1),
    # column walkers: 4-row groups + tails, one strip / several strips / ragged strips, segments
    ("float32", (2, 70, 128), 1), ("float32", (1, 261, 132), 1), ("float32", (3, 66, 8), 1), ("float64", (2, 130, 66), 1),
    ("float16", (1, 200, 72), 1), ("float32", (700, 4, 36), 0), ("int64", (5, 9, 34), 1),
    # element-aligned rows (element-sized copies, a lane owns single columns)
    ("float32", (2, 130, 13), 1), ("float32", (3, 67, 65), 1), ("float64", (2, 70, 33), 1), ("int32", (1, 300, 7), 1),
    # ... a last strip with few columns, rows of a few columns (row segments side by side in a warp)
    ("float32", (2, 71, 161), 1), ("float64", (1, 67, 65), 1), ("float32", (1, 3, 513), 1), ("float32", (1, 135, 513), 1),
    # packed types on unaligned rows: generic kernel
    ("float16", (2, 50, 13), 1), ("uint8", (3, 40, 7), 1),
]


@pytest.mark.parametrize("dtype,shape,axis", CASES)
@pytest.mark.parametrize("where", ["end", "start"])
@pytest.mark.parametrize("mask", [None, "nan_or_value"])
def test_k1_never_reads_outside_the_array(dtype, shape, axis, where, mask):
    rng = np.random.default_rng(1000 * len(dtype) + sum(shape))
    if np.dtype(dtype).kind == "f":
        x = smooth_field(rng, shape, dtype, axis=axis, scale=0.3, base=2.0)
    else:
        info = np.iinfo(dtype)
        x = rng.integers(info.min, info.max, size=shape, dtype=dtype, endpoint=True)
    mv = None
    if mask is not None:
        mv = float("nan") if np.dtype(dtype).kind == "f" else 7
        sel = rng.random(shape) < 0.2
        x[sel] = mv
        x.reshape(-1)[-3:] = mv  # masked elements at the very end
        x.reshape(-1)[0] = mv
    buf = GuardedBuffer(x.nbytes, where)
    try:
        buf.upload(x)
        got = _count_at(buf.ptr, x, axis, mv)
    finally:
        buf.close()
    assert np.array_equal(got, _count_plain(x, axis, mv))


@pytest.mark.parametrize("dtype", ["float16", "float32", "float64"])
@pytest.mark.parametrize("n", [8, 1000, 2048, 4096 + 64, 65536 + 24])
@pytest.mark.parametrize("where", ["end", "start"])
def test_k3_k4_never_touch_memory_outside_their_buffers(dtype, n, where):
    rng = np.random.default_rng(n)
    x = smooth_field(rng, (n,), dtype, axis=0, scale=0.3, base=3.0)
    src, dst = GuardedBuffer(x.nbytes, where), GuardedBuffer(x.nbytes, where)
    try:
        src.upload(x)
        es = x.dtype.itemsize
        for kind in (nat.SHUFFLE_BYTES, nat.SHUFFLE_BITS):
            nat.check(nat.lib().xbi_shuffle(src.ptr, dst.ptr, es, n, kind, 0, None))
            torch.cuda.synchronize()
            shuffled = dst.download("u1", (x.nbytes,))
            nat.check(nat.lib().xbi_shuffle(dst.ptr, src.ptr, es, n, kind, 1, None))
            torch.cuda.synchronize()
            assert np.array_equal(src.download(dtype, (n,)).view(f"u{es}"), x.view(f"u{es}"))
            assert shuffled.size == x.nbytes
        nat.check(nat.lib().xbi_bitround_inplace(src.ptr, nat.dtype_code(x.dtype), n, 3, None))
        torch.cuda.synchronize()
        from oracle import bitround_oracle as bro

        assert np.array_equal(src.download(dtype, (n,)).view(f"u{es}"), bro.bitround(x, 3).view(f"u{es}"))
    finally:
        src.close()
        dst.close()


@pytest.mark.skipif(not __import__("os").environ.get("XBI_RUN_GUARD_CANARY"),
                    reason="provokes a GPU page fault on purpose; opt in with XBI_RUN_GUARD_CANARY=1 "
                           "(verified on a B200 box: the overrun faults, see profiles/README.md)")
def test_the_guard_bites():
    """The harness itself: a call that is told the array is longer than it is must fault (run in
    a child process, an illegal address poisons the CUDA context)."""
    import os
    import subprocess
    import sys

    code = (
        "import sys, numpy as np, torch\n"
        "sys.path.insert(0, %r); sys.path.insert(0, %r)\n"
        "from test_guard_page_gpu import GuardedBuffer\n"
        "from xbitinfo_b200 import _native as nat\n"
        "b = GuardedBuffer(4096 * 4, 'end'); b.upload(np.ones(4096, dtype='float32'))\n"
        "rc = nat.lib().xbi_bitround_inplace(b.ptr, nat.F32, 4096 + (8 << 20), 3, None)\n"
        "torch.cuda.synchronize()\n"
        "print('NO FAULT')\n"
    ) % (os.path.dirname(os.path.dirname(os.path.abspath(__file__))), os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "NO FAULT" not in r.stdout
    assert r.returncode != 0 and ("illegal memory access" in r.stderr or "CUDA error" in r.stderr or "AcceleratorError" in r.stderr)
