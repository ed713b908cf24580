"""One variable, one axis: raw array in, information per bit out (K1 -> K2).

This is the function behind the ``implementation="cuda"`` branch; it is the B200
counterpart of ``pb.bitinformation(X, axis).compute()`` at ``xbitinfo/xbitinfo.py:366``
together with the transform/cast that precedes it (``349-355``).

* numpy (host) arrays go through the C ABI's host-buffer entry point
  (``xbi_bitinformation_host``): slabs are staged host->device, counted, and only the
  ``nbits`` doubles come back.
* torch CUDA tensors are counted in place in HBM (``xbi_count_pairs`` +
  ``xbi_mutual_information``); with an initialised ``torch.distributed`` NCCL group and
  ``sharded=True`` the per-rank counts are summed with one all-reduce first.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _native as nat


def _is_torch_tensor(x) -> bool:
    return type(x).__module__.startswith("torch") and hasattr(x, "data_ptr")


def _native_contiguous(x: np.ndarray) -> np.ndarray:
    if not x.dtype.isnative:
        x = x.astype(x.dtype.newbyteorder("="))
    return np.ascontiguousarray(x)


def bitinformation(data, axis: int, masked_value=None, set_zero_insignificant: bool = False,
                   confidence: float = 0.99, device: int | None = None, return_counts: bool = False,
                   sharded: bool = False):
    """float64[nbits] information per bit (most significant bit first) of ``data`` along
    ``axis``; with ``return_counts`` also the (nbits,2,2) int64 pair counts."""
    if _is_torch_tensor(data):
        return _bitinformation_device(data, axis, masked_value, set_zero_insignificant, confidence,
                                      return_counts, sharded)
    if sharded:
        raise ValueError("sharded=True needs per-rank CUDA tensors")
    from .streaming import bitinformation_streamed, is_lazy_array

    if is_lazy_array(data):
        # dask / zarr / h5py ...: never materialised on the host; slabs along the first axis
        return bitinformation_streamed(data, axis, masked_value=masked_value,
                                       set_zero_insignificant=set_zero_insignificant, confidence=confidence,
                                       device=device, return_counts=return_counts)
    x = np.asarray(data)
    code = nat.dtype_code(x.dtype)  # raises ValueError for unsupported dtypes
    x = _native_contiguous(x)
    if x.ndim == 0:
        raise ValueError("cannot analyse a 0-d array")
    axis = axis % x.ndim
    outer = int(np.prod(x.shape[:axis], dtype=np.int64))
    n = int(x.shape[axis])
    inner = int(np.prod(x.shape[axis + 1:], dtype=np.int64))
    nbits = 8 * x.dtype.itemsize
    flags, bits = nat.mask_spec(x.dtype, masked_value)
    if x.dtype.kind == "f":
        flags |= nat.SIGNED_EXPONENT
    mi = np.empty(nbits, dtype=np.float64)
    counts = np.empty(nbits * 4, dtype=np.uint64)
    if device is None:
        device = _current_device()
    rc = nat.lib().xbi_bitinformation_host(
        x.ctypes.data_as(C.c_void_p), code, outer, n, inner, flags, bits, int(bool(set_zero_insignificant)),
        float(confidence), mi.ctypes.data_as(C.c_void_p), counts.ctypes.data_as(C.c_void_p), int(device))
    nat.check(rc)
    if return_counts:
        return mi, counts.reshape(nbits, 2, 2).astype(np.int64)
    return mi


def bitinformation_multi(data, axes, masked_value=None, set_zero_insignificant: bool = False,
                         confidence: float = 0.99, device: int | None = None, return_counts: bool = False):
    """Information per bit along several axes of one array: list of float64[nbits] in the order of
    ``axes`` (with ``return_counts`` a list of ``(mi, counts)``).  A host numpy array crosses PCIe once for
    all the axes (``xbi_bitinformation_host_multi``), a lazy array (dask / zarr / h5py) is read slab by slab
    from its source once for all the axes; a device tensor goes through ``xbi_bitinformation_dev_multi`` (the two
    innermost axes in one read of the array, the others one pass each)."""
    from .streaming import is_lazy_array

    axes = [int(a) for a in axes]
    if is_lazy_array(data) and len(axes) >= 2:
        # dask / zarr / h5py ...: every slab is read from the source once for all the axes
        from .streaming import bitinformation_streamed_multi

        return bitinformation_streamed_multi(data, axes, masked_value=masked_value,
                                             set_zero_insignificant=set_zero_insignificant, confidence=confidence,
                                             device=device, return_counts=return_counts)
    if _is_torch_tensor(data) and len(axes) >= 2 and 0 < data.dim() <= 8 and len(set(a % data.dim() for a in axes)) == len(axes):
        # device-resident: one call of the C ABI; the two innermost axes share one read of the array
        from .device import bitinformation_multi_device

        return bitinformation_multi_device(data, axes, masked_value=masked_value,
                                           set_zero_insignificant=set_zero_insignificant, confidence=confidence,
                                           return_counts=return_counts)
    if _is_torch_tensor(data) or is_lazy_array(data) or len(axes) < 2:
        return [bitinformation(data, a, masked_value=masked_value, set_zero_insignificant=set_zero_insignificant,
                               confidence=confidence, device=device, return_counts=return_counts) for a in axes]
    x = np.asarray(data)
    code = nat.dtype_code(x.dtype)
    x = _native_contiguous(x)
    if x.ndim == 0:
        raise ValueError("cannot analyse a 0-d array")
    if x.ndim > 8:
        return [bitinformation(x, a, masked_value=masked_value, set_zero_insignificant=set_zero_insignificant,
                               confidence=confidence, device=device, return_counts=return_counts) for a in axes]
    axes = [a % x.ndim for a in axes]
    nbits = 8 * x.dtype.itemsize
    flags, bits = nat.mask_spec(x.dtype, masked_value)
    if x.dtype.kind == "f":
        flags |= nat.SIGNED_EXPONENT
    mi = np.empty((len(axes), nbits), dtype=np.float64)
    counts = np.empty((len(axes), nbits * 4), dtype=np.uint64)
    if device is None:
        device = _current_device()
    rc = nat.lib().xbi_bitinformation_host_multi(
        x.ctypes.data_as(C.c_void_p), code, x.ndim, nat.int64_array(x.shape), len(axes), (C.c_int * len(axes))(*axes),
        flags, bits, int(bool(set_zero_insignificant)), float(confidence), mi.ctypes.data_as(C.c_void_p),
        counts.ctypes.data_as(C.c_void_p), int(device))
    nat.check(rc)
    if return_counts:
        return [(mi[i], counts[i].reshape(nbits, 2, 2).astype(np.int64)) for i in range(len(axes))]
    return [mi[i] for i in range(len(axes))]


def _current_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return torch.cuda.current_device()
    except Exception:  # torch is plumbing, not a requirement of the host path
        pass
    return 0


def _bitinformation_device(x, axis, masked_value, set_zero, confidence, return_counts, sharded):
    from .device import PairCounter

    if x.dim() == 0:
        raise ValueError("cannot analyse a 0-d array")
    if not x.is_contiguous():
        x = x.contiguous()
    pc = PairCounter(x.dtype, x.device)
    if sharded:
        from .distributed import allreduce_counts

        # per-rank table (counts only), one all-reduce, K2 on every rank
        pc.bitinformation(x, axis, masked_value=masked_value, with_information=False)
        allreduce_counts(pc.counts)
        mi = pc.information(set_zero, confidence).cpu().numpy()
    else:
        # K1 + K2 in one call of the C ABI (workspace clear, counting kernel, finish kernel)
        mi = pc.bitinformation(x, axis, masked_value=masked_value, set_zero_insignificant=set_zero,
                               confidence=confidence).cpu().numpy()
    if return_counts:
        return mi, pc.counts.cpu().numpy()
    return mi


def bitround(data, keepbits: int, device: int | None = None):
    """New array rounded to ``keepbits`` mantissa bits (K3); the input is not modified
    (``xbitinfo/bitround.py:10``).  numpy in -> numpy out, CUDA tensor in -> CUDA tensor out."""
    if _is_torch_tensor(data):
        from .device import bitround as bitround_to

        return bitround_to(data.contiguous(), keepbits)  # one read, one write; the input is left as it is
    x = np.asarray(data)
    if x.dtype.kind != "f" or x.dtype.itemsize > 8:
        raise TypeError("Only float arrays (16-64bit) can be bit-rounded")
    code = nat.dtype_code(x.dtype)
    src = _native_contiguous(x)
    out = _new_output_like(src)
    if device is None:
        device = _current_device()
    rc = nat.lib().xbi_bitround_host(src.ctypes.data_as(C.c_void_p), out.ctypes.data_as(C.c_void_p), code,
                                     src.size, int(keepbits), int(device))
    nat.check(rc)
    return out


def _pinned_output_cap_bytes() -> int:
    import os

    try:
        return int(float(os.environ.get("XBI_PINNED_OUTPUT_MB", "8192")) * (1 << 20))
    except ValueError:
        return 8192 << 20


def _new_output_like(src: np.ndarray) -> np.ndarray:
    """The array ``bitround`` returns.  A fresh ``np.empty_like`` is first touched while the result is copied out: its
    page faults cap the call at ~20 GB/s however many threads copy (profiles/r01_host_path.md).  Outputs from 1 MB up
    to ``XBI_PINNED_OUTPUT_MB`` (default 8192; 0 disables) are therefore taken from torch's caching page-locked host
    allocator -- the DMA engine writes the rounded slabs straight into them, no staging copy, no faults -- and given
    back to that cache when the caller drops the array (the numpy array owns a reference to the storage).  It is an
    ordinary numpy array in every other respect."""
    cap = _pinned_output_cap_bytes()
    if (1 << 20) <= src.nbytes <= cap:
        try:
            import torch

            if torch.cuda.is_available():
                buf = torch.empty(src.nbytes, dtype=torch.uint8, pin_memory=True)
                return buf.numpy().view(src.dtype).reshape(src.shape)
        except Exception:  # torch is plumbing: any trouble here and the plain allocation serves
            pass
    return np.empty_like(src)
