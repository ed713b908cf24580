"""Out-of-core front end: bit information of arrays that fit neither in HBM nor in host RAM.

SURVEY.md row f1.  The reference's python path builds one graph over the whole variable
(``xbitinfo/xbitinfo.py:349``) and its julia path calls ``.values`` (``310``): neither
streams.  Here any array-like that supports numpy-style slicing along its first axis
(``np.memmap``, ``h5py`` / ``netCDF4`` / ``zarr`` variables, a dask array) is cut into slabs
along that axis; each slab goes pinned host buffer -> device (double buffered on two
streams) -> K1, and the accumulate-only ``counts`` make the slabs compose.  When the
analysed axis is the sliced one, every slab carries the first row of the next slab as a
halo so that the pair across the cut is counted exactly once.
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .device import PairCounter

_TORCH_OF_NP = {
    "float16": torch.float16, "float32": torch.float32, "float64": torch.float64,
    "uint8": torch.uint8, "int8": torch.int8, "int16": torch.int16, "int32": torch.int32, "int64": torch.int64,
    # torch has no arithmetic on wide unsigned types; only the bits matter here
    "uint16": torch.int16, "uint32": torch.int32, "uint64": torch.int64,
}


def is_lazy_array(obj) -> bool:
    """A sliceable array-like that is neither a numpy array nor a torch tensor: np.memmap is
    a numpy array and stays on the in-memory path unless the caller streams it explicitly;
    dask / zarr / h5py / netCDF4 variables (``shape``, ``dtype``, ``__getitem__``) are lazy."""
    if isinstance(obj, np.ndarray) or type(obj).__module__.startswith("torch"):
        return False
    return all(hasattr(obj, a) for a in ("shape", "dtype", "__getitem__")) and not np.isscalar(obj)


def _chunk_edges_axis0(source):
    """Boundaries of the source's own chunks along axis 0 (dask ``.chunks`` tuples, zarr / h5py
    ``.chunks`` chunk shape), or None."""
    ch = getattr(source, "chunks", None)
    if not ch:
        return None
    first = ch[0]
    n0 = int(source.shape[0])
    if isinstance(first, (tuple, list)):
        edges = np.concatenate([[0], np.cumsum(first)])
    else:
        step = int(first)
        if step <= 0:
            return None
        edges = np.arange(0, n0 + step, step)
        edges[-1] = n0
    return [int(e) for e in edges if e <= n0]


def slab_plan(shape, axis: int, itemsize: int, slab_bytes: int, chunk_edges=None):
    """[(lo, hi, halo)] rows of axis 0 per slab; ``halo`` (0/1) extra trailing rows to read.
    With ``chunk_edges`` (the source's own chunk boundaries along axis 0) slabs end on chunk
    boundaries whenever a whole chunk fits, so that no stored chunk is decoded twice."""
    n0 = int(shape[0])
    row_bytes = int(np.prod(shape[1:], dtype=np.int64)) * itemsize
    rows = max(1 if axis != 0 else 2, int(slab_bytes // max(row_bytes, 1)))
    plan = []
    if axis != 0 and chunk_edges and len(chunk_edges) > 1:
        lo = 0
        while lo < n0:
            limit = min(lo + rows, n0)
            fit = [e for e in chunk_edges if lo < e <= limit]
            hi = fit[-1] if fit else limit  # a chunk larger than a slab is cut
            plan.append((lo, hi, 0))
            lo = hi
        return plan
    if axis != 0:
        for lo in range(0, n0, rows):
            plan.append((lo, min(lo + rows, n0), 0))
    else:
        step = rows - 1  # pairs (j, j+1) with lo <= j < hi
        lo = 0
        while lo < n0 - 1:
            hi = min(lo + step, n0 - 1)
            plan.append((lo, hi, 1))
            lo = hi
    return plan


def bitinformation_streamed(source, axis: int, slab_bytes: int = 1 << 30, masked_value=None,
                            set_zero_insignificant: bool = False, confidence: float = 0.99, device=None,
                            return_counts: bool = False):
    """float64[nbits] information per bit of ``source`` along ``axis``, reading it slab by slab."""
    shape = tuple(int(s) for s in source.shape)
    if len(shape) == 0:
        raise ValueError("cannot analyse a 0-d array")
    axis %= len(shape)
    plan = slab_plan(shape, axis, np.dtype(source.dtype).itemsize, slab_bytes, _chunk_edges_axis0(source))
    if shape[axis] < 2:
        plan = []  # no pairs along the axis: nothing to read
    return _stream(source, [axis], plan, masked_value, set_zero_insignificant, confidence, device, return_counts)[0]


def bitinformation_streamed_multi(source, axes, slab_bytes: int = 1 << 30, masked_value=None,
                                  set_zero_insignificant: bool = False, confidence: float = 0.99, device=None,
                                  return_counts: bool = False):
    """Information per bit along every axis in ``axes`` from ONE pass over ``source``: every slab is read
    (decoded, for a compressed store) once and counted along each axis while it is on the device.  When the
    sliced first axis is among them, every slab but the last carries the first row of the next one as halo;
    the other axes count the slab's own rows only.  Returns a list in the order of ``axes``."""
    shape = tuple(int(s) for s in source.shape)
    if len(shape) == 0:
        raise ValueError("cannot analyse a 0-d array")
    axes = [int(a) % len(shape) for a in axes]
    if len(axes) == 1:
        return [bitinformation_streamed(source, axes[0], slab_bytes, masked_value, set_zero_insignificant,
                                        confidence, device, return_counts)]
    n0 = shape[0]
    itemsize = np.dtype(source.dtype).itemsize
    # slabs partition the first axis (on the source's chunk boundaries where possible); one more row is read
    # as halo when the first axis itself is analysed
    base = slab_plan(shape, 1 if len(shape) > 1 else 0, itemsize, slab_bytes, _chunk_edges_axis0(source)) \
        if len(shape) > 1 else slab_plan(shape, 0, itemsize, slab_bytes)
    if len(shape) == 1:  # only axis 0 exists
        return [bitinformation_streamed(source, 0, slab_bytes, masked_value, set_zero_insignificant, confidence,
                                        device, return_counts) for _ in axes]
    halo_wanted = 0 in axes
    plan = [(lo, hi, 1 if (halo_wanted and hi < n0) else 0) for lo, hi, _ in base]
    return _stream(source, axes, plan, masked_value, set_zero_insignificant, confidence, device, return_counts)


def _stream(source, axes, plan, masked_value, set_zero_insignificant, confidence, device, return_counts):
    """Slabs ``[(lo, hi, halo)]`` of the first axis: pinned host buffer -> device (double buffered on two
    streams) -> K1 per analysed axis.  With one axis the plan of ``slab_plan`` applies as is (rows
    ``lo..hi+halo`` counted along it).  With several, the slab's own rows ``lo..hi`` are counted along the
    axes other than the first and rows ``lo..hi+halo`` along the first."""
    shape = tuple(int(s) for s in source.shape)
    dtype = np.dtype(source.dtype)
    nat.dtype_code(dtype)  # ValueError for unsupported dtypes
    tdtype = _TORCH_OF_NP[dtype.name]
    dev = torch.device("cuda", torch.cuda.current_device() if device is None else device)
    single = len(axes) == 1
    inner_elems = int(np.prod(shape[1:], dtype=np.int64))
    max_rows = max((hi - lo + halo for lo, hi, halo in plan), default=0)
    counters = [[PairCounter(tdtype, dev) for _ in axes] for _ in range(2)]

    def results():
        out = []
        for k in range(len(axes)):
            counters[0][k].counts += counters[1][k].counts
            mi = counters[0][k].information(set_zero_insignificant, confidence).cpu().numpy()
            out.append((mi, counters[0][k].counts.cpu().numpy()) if return_counts else mi)
        return out

    if max_rows == 0:
        return results()
    pinned = [torch.empty(max_rows * inner_elems, dtype=tdtype).pin_memory() for _ in range(2)]
    staged = [torch.empty(max_rows * inner_elems, dtype=tdtype, device=dev) for _ in range(2)]
    streams = [torch.cuda.Stream(device=dev) for _ in range(2)]
    # the counters were zero-filled and the staging buffers allocated on the current stream: the side streams
    # are not ordered after it unless told to be
    for s_ in streams:
        s_.wait_stream(torch.cuda.current_stream(dev))
    for buf in staged:
        for s_ in streams:
            buf.record_stream(s_)
    done = [None, None]
    for i, (lo, hi, halo) in enumerate(plan):
        b = i & 1
        if done[b] is not None:
            done[b].synchronize()  # the pinned buffer is free again
        rows = hi - lo + halo
        slab = np.asarray(source[lo: hi + halo])
        n = rows * inner_elems
        host_view = pinned[b].numpy()[:n].view(dtype)
        np.copyto(host_view, slab.reshape(-1), casting="equiv")
        with torch.cuda.stream(streams[b]):
            staged[b][:n].copy_(pinned[b][:n], non_blocking=True)
            x = staged[b][:n].view((rows,) + shape[1:])
            for k, axis in enumerate(axes):
                if single or axis == 0:
                    if shape[axis] >= 2 and rows >= (2 if axis == 0 else 1):
                        counters[b][k].add(x, axis, masked_value=masked_value)
                elif shape[axis] >= 2:
                    counters[b][k].add(x[: hi - lo], axis, masked_value=masked_value)
            ev = torch.cuda.Event()
            ev.record(streams[b])
        done[b] = ev
    for s in streams:
        s.synchronize()
    return results()


def bitround_streamed(source, keepbits: int, out, slab_bytes: int = 1 << 30, device=None):
    """Out-of-core ``xr_bitround``: ``out[lo:hi] = bitround(source[lo:hi], keepbits)`` slab by slab along the
    first axis (on the source's chunk boundaries), so that neither the input nor the result has to fit in host
    memory.  ``source``: anything sliceable (``np.memmap``, zarr / h5py / netCDF4 variables, dask arrays);
    ``out``: anything that accepts slice assignment of numpy arrays (a ``np.memmap``, a zarr array, an h5py
    dataset ...) with the shape and dtype of ``source``.  Every slab goes through ``xbi_bitround_host`` (pinned
    ring + copy threads + K3).  Returns ``out``."""
    from . import engine

    shape = tuple(int(s) for s in source.shape)
    if len(shape) == 0:
        raise ValueError("cannot round a 0-d array")
    dtype = np.dtype(source.dtype)
    if dtype.kind != "f" or dtype.itemsize > 8:
        raise TypeError("Only float arrays (16-64bit) can be bit-rounded")
    if tuple(int(s) for s in out.shape) != shape or np.dtype(out.dtype) != dtype:
        raise ValueError("`out` must have the shape and dtype of `source`")
    plan = slab_plan(shape, 1, dtype.itemsize, slab_bytes, _chunk_edges_axis0(source)) if len(shape) > 1 else \
        [(lo, min(lo + max(1, slab_bytes // dtype.itemsize), shape[0]), 0)
         for lo in range(0, shape[0], max(1, slab_bytes // dtype.itemsize))]
    for lo, hi, _ in plan:
        out[lo:hi] = engine.bitround(np.asarray(source[lo:hi]), keepbits, device=device)
    return out
