"""Multi-GPU: shard planning and the one collective of the path.

Pairs only couple neighbours along the analysed axis and the pair histogram is a plain sum
over pairs (``xbitinfo/_py_bitinfo.py:124, 155-160``), so the pair set partitions along any
other axis and integer counts add.  One process per GPU (``torch.distributed``, NCCL over
NVLink/NVSwitch): every rank counts its slab locally with K1, then a single
``all_reduce(SUM)`` of the ``int64[nbits,2,2]`` table (<= 2 KiB, latency-bound) gives every
rank the global counts; K2 runs redundantly on each rank.  No data-path collective.

When the array can only be split along the analysed axis itself, each shard but the last
additionally reads one halo row (the first row of the next shard) so that the pair across
the cut is counted exactly once.
"""
from __future__ import annotations

from dataclasses import dataclass


def shard_bounds(n_units: int, world_size: int, rank: int):
    """Contiguous split of ``n_units`` into ``world_size`` parts; the first ``n_units %
    world_size`` parts are one unit longer.  Returns (lo, hi)."""
    base, extra = divmod(n_units, world_size)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


@dataclass(frozen=True)
class ShardPlan:
    split_axis: int   # axis of the global array that is cut
    lo: int           # first index of this rank's slab along split_axis
    hi: int           # one past the last index it owns
    halo: int         # extra trailing rows it must also hold (0 or 1; only when split_axis == axis)

    @property
    def read_slice(self):
        return slice(self.lo, self.hi + self.halo)


def plan_shard(shape, axis: int, world_size: int, rank: int) -> ShardPlan:
    """Which slab of a C-contiguous array of ``shape`` rank ``rank`` analyses for adjacency
    along ``axis``.  Prefers the outermost non-analysed axis that is long enough to give
    every rank work; falls back to cutting the analysed axis with a one-row halo."""
    ndim = len(shape)
    axis %= ndim
    for ax in range(ndim):
        if ax != axis and shape[ax] >= world_size:
            lo, hi = shard_bounds(shape[ax], world_size, rank)
            return ShardPlan(ax, lo, hi, 0)
    # cut the analysed axis: pairs (j, j+1) with lo <= j < hi belong to this rank, so it
    # reads rows lo .. hi (the halo) unless it owns the end of the axis
    n = shape[axis]
    lo, hi = shard_bounds(n, world_size, rank)
    halo = 1 if (hi < n and hi > lo) else 0
    return ShardPlan(axis, lo, hi, halo)


def is_initialized() -> bool:
    try:
        import torch.distributed as dist

        return dist.is_available() and dist.is_initialized()
    except Exception:
        return False


def world():
    """(rank, world_size) of the default process group, (0, 1) when not initialised."""
    if not is_initialized():
        return 0, 1
    import torch.distributed as dist

    return dist.get_rank(), dist.get_world_size()


def bind_to_gpu_numa_node(device_index: int):
    """Restrict this process to the CPUs of the NUMA node its GPU hangs off, so that host staging memory
    allocated afterwards (pinned buffers are placed by first touch) and the copy threads are local to the
    GPU's PCIe root.  With one process per GPU on a two-socket box, unbound ranks end up reading half of
    their host buffers across the socket interconnect.  Returns the node, or None when the topology is not
    exposed (containers often hide it) or the CPUs are not available to this process; never raises."""
    import os

    try:
        import torch

        p = torch.cuda.get_device_properties(device_index)
        sys_id = f"{p.pci_domain_id:04x}:{p.pci_bus_id:02x}:{p.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{sys_id}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        return node
    except Exception:
        return None


def allreduce_counts(counts, group=None):
    """In-place SUM all-reduce of a counts tensor (int64).  NCCL for CUDA tensors, gloo for
    CPU tensors (tests).  A no-op without an initialised process group."""
    if not is_initialized():
        return counts
    import torch.distributed as dist

    dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=group)
    return counts


class PeerExchange:
    """Exchange buffers of the fused all-reduce (``xbi_bitinformation_dev_allreduce``): one small device buffer
    per rank, allocated by the library, shared with the other ranks of the node through CUDA IPC handles that
    travel over ``torch.distributed.all_gather_object`` (plumbing), and mapped into this process.  The finish
    kernel of every counting call then pushes the rank's ``int64[nbits,2,2]`` table into the peers' buffers over
    NVLink and sums what the peers pushed -- no NCCL call, no extra launch.

    Raises (``XbiError`` / ``RuntimeError``) where the ranks are not GPUs of one node or IPC is not permitted;
    callers fall back to :func:`allreduce_counts` (NCCL) and say so."""

    def __init__(self, device, group=None, timeout_s: float = 20.0):
        import ctypes as C

        import torch
        import torch.distributed as dist

        from . import _native as nat

        if not is_initialized():
            raise RuntimeError("PeerExchange needs an initialised torch.distributed process group")
        self.rank = dist.get_rank(group)
        self.world = dist.get_world_size(group)
        if self.world > 8:
            raise RuntimeError("the peer all-reduce spans the GPUs of one node (world <= 8)")
        self.timeout_s = float(timeout_s)
        self.device = torch.device(device)
        self._epoch = 0
        self._lib = nat.lib()
        self._own = C.c_void_p()
        self._mapped = []
        handle = C.create_string_buffer(int(self._lib.xbi_peer_handle_bytes()))
        with torch.cuda.device(self.device):
            nat.check(self._lib.xbi_peer_create(C.byref(self._own), handle))
            gathered = [None] * self.world
            dist.all_gather_object(gathered, (self.rank, bytes(handle.raw)), group=group)
            ptrs = [None] * self.world
            failure = None
            for r, raw in gathered:
                if r == self.rank:
                    ptrs[r] = self._own.value
                    continue
                mapped = C.c_void_p()
                rc = self._lib.xbi_peer_open(C.create_string_buffer(raw, len(raw)), C.byref(mapped))
                if rc != 0:
                    failure = self._lib.xbi_last_error().decode("utf-8", "replace")
                    break
                self._mapped.append(mapped)
                ptrs[r] = mapped.value
            # every rank must know whether every rank succeeded, or some would wait for flags that never come
            ok = [None] * self.world
            dist.all_gather_object(ok, failure, group=group)
            if any(f is not None for f in ok):
                self.close()
                raise RuntimeError("CUDA IPC mapping of the peer exchange buffers failed: "
                                   + "; ".join(f"rank {i}: {f}" for i, f in enumerate(ok) if f is not None))
        self.table = (C.c_void_p * self.world)(*ptrs)

    def next_epoch(self) -> int:
        self._epoch += 1
        return self._epoch

    def failed_epoch(self) -> int:
        """Epoch of a wait that timed out (0: none): results from that call on are invalid."""
        import ctypes as C

        out = C.c_ulonglong(0)
        from . import _native as nat

        nat.check(self._lib.xbi_peer_error_epoch(self._own, C.byref(out)))
        return int(out.value)

    def close(self) -> None:
        import torch

        with torch.cuda.device(self.device):
            torch.cuda.synchronize()
            if is_initialized():  # a peer may still be pushing into this rank's buffer
                import torch.distributed as dist

                try:
                    dist.barrier()
                except Exception:
                    pass
            for m in self._mapped:
                self._lib.xbi_peer_close(m)
            self._mapped = []
            if self._own:
                self._lib.xbi_peer_destroy(self._own)
                self._own = None


def sharded_bitinformation(local_slab, axis: int, masked_value=None, set_zero_insignificant: bool = False,
                           confidence: float = 0.99, return_counts: bool = False, device=None):
    """Information per bit of the GLOBAL array whose rank-local slab (as cut by
    :func:`plan_shard`, halo row included) is ``local_slab``.  A CUDA tensor is counted in
    place; a numpy array is streamed through the host-buffer entry point of the C ABI.
    The per-rank ``int64[nbits,2,2]`` counts are summed with one all-reduce and K2 runs on
    every rank, so every rank returns the same numpy result."""
    import numpy as np
    import torch

    from . import engine
    from .device import PairCounter

    if type(local_slab).__module__.startswith("torch"):
        return engine.bitinformation(local_slab, axis, masked_value=masked_value,
                                     set_zero_insignificant=set_zero_insignificant, confidence=confidence,
                                     return_counts=return_counts, sharded=True)
    x = np.asarray(local_slab)
    _, counts = engine.bitinformation(x, axis, masked_value=masked_value, return_counts=True, device=device)
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    tdtype = getattr(torch, {"int8": "uint8"}.get(x.dtype.name, x.dtype.name), None)
    if tdtype is None:  # unsigned types torch lacks: same width signed
        tdtype = getattr(torch, x.dtype.name.replace("uint", "int"))
    pc = PairCounter(tdtype, dev)
    pc.counts.copy_(torch.from_numpy(counts))
    allreduce_counts(pc.counts)
    mi = pc.information(set_zero_insignificant, confidence).cpu().numpy()
    if return_counts:
        return mi, pc.counts.cpu().numpy()
    return mi
