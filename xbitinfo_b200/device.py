"""Device-level plumbing: torch tensors in HBM <-> the C ABI.

torch is used only for what it is good at here -- owning device memory, streams and the
NCCL process group.  Every kernel is ours and is reached through ``_native.lib()``.
"""
from __future__ import annotations

import math

import numpy as np
import torch

from . import _native as nat

_TORCH_CODES = {
    torch.uint8: nat.U8, torch.int8: nat.U8,
    torch.int16: nat.U16, torch.int32: nat.U32, torch.int64: nat.U64,
    torch.float16: nat.F16, torch.float32: nat.F32, torch.float64: nat.F64,
}
for _name, _code in (("uint16", nat.U16), ("uint32", nat.U32), ("uint64", nat.U64)):
    if hasattr(torch, _name):
        _TORCH_CODES[getattr(torch, _name)] = _code

_NP_OF_TORCH = {
    torch.uint8: "uint8", torch.int8: "int8", torch.int16: "int16", torch.int32: "int32",
    torch.int64: "int64", torch.float16: "float16", torch.float32: "float32", torch.float64: "float64",
}
for _name in ("uint16", "uint32", "uint64"):
    if hasattr(torch, _name):
        _NP_OF_TORCH[getattr(torch, _name)] = _name


def numpy_dtype_of(t: torch.Tensor) -> np.dtype:
    return np.dtype(_NP_OF_TORCH[t.dtype])


def torch_dtype_code(dtype: torch.dtype) -> int:
    try:
        return _TORCH_CODES[dtype]
    except KeyError:
        raise ValueError(f"dtype {dtype} neither known nor implemented.") from None


def require_cuda_tensor(x: torch.Tensor) -> None:
    if not isinstance(x, torch.Tensor) or not x.is_cuda:
        raise TypeError("expected a CUDA torch.Tensor (xbitinfo_b200 has no CPU path)")
    if not x.is_contiguous():
        raise ValueError("the device array must be C-contiguous")


def split_axis(shape, axis: int):
    """(outer, n, inner) view of a C-contiguous array for adjacency along ``axis``."""
    ndim = len(shape)
    if not -ndim <= axis < ndim:
        raise ValueError(f"axis {axis} out of range for {ndim}-d array")
    axis %= ndim
    return math.prod(shape[:axis]), int(shape[axis]), math.prod(shape[axis + 1:])


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class PairCounter:
    """Reusable device state for one (dtype) counting job: the accumulate-only
    ``counts`` buffer (nbits x 2 x 2, int64) and the scratch workspace of K1."""

    def __init__(self, dtype: torch.dtype, device):
        self.code = torch_dtype_code(dtype)
        self.nbits = 8 * torch.empty((), dtype=dtype).element_size()
        self.device = torch.device(device)
        self.counts = torch.zeros((self.nbits, 2, 2), dtype=torch.int64, device=self.device)
        self.workspace = torch.empty(nat.lib().xbi_workspace_bytes(), dtype=torch.uint8, device=self.device)
        self.mi = torch.empty(self.nbits, dtype=torch.float64, device=self.device)

    def reset(self) -> None:
        self.counts.zero_()

    def add(self, x: torch.Tensor, axis: int = -1, *, transform: bool = True, masked_value=None,
            force_generic: bool = False, dims=None, flavour: str | None = None) -> None:
        """counts += pair histogram of ``x`` along ``axis`` (K1).  ``dims`` overrides the
        (outer, n, inner) split, e.g. to describe a slab with a halo row."""
        require_cuda_tensor(x)
        if torch_dtype_code(x.dtype) != self.code:
            raise ValueError("dtype differs from the counter's dtype")
        outer, n, inner = dims if dims is not None else split_axis(tuple(x.shape), axis)
        np_dtype = numpy_dtype_of(x)
        flags, bits = nat.mask_spec(np_dtype, masked_value)
        if np_dtype.kind == "f" and transform:
            flags |= nat.SIGNED_EXPONENT
        if force_generic:
            flags |= nat.FORCE_GENERIC
        if flavour is not None:  # testing: "sparse" / "dense" flavour of the masked fast kernels
            flags |= {"sparse": nat.FORCE_SPARSE, "dense": nat.FORCE_DENSE}[flavour]
        with torch.cuda.device(x.device):
            rc = nat.lib().xbi_count_pairs(
                x.data_ptr(), self.code, outer, n, inner, flags, bits, self.counts.data_ptr(),
                self.workspace.data_ptr(), self.workspace.numel(), _stream_ptr(x.device))
        nat.check(rc)

    def bitinformation(self, x: torch.Tensor, axis: int = -1, *, transform: bool = True, masked_value=None,
                       set_zero_insignificant: bool = False, confidence: float = 0.99, dims=None,
                       with_information: bool = True, peers=None, flavour: str | None = None):
        """K1 + K2 in ONE call of the C ABI (``xbi_bitinformation_dev``): ``counts`` is OVERWRITTEN with the pair
        histogram of ``x`` along ``axis`` and the information per bit lands in ``mi`` (returned; ``None`` with
        ``with_information=False``).  Three launches: workspace clear, counting kernel, finish kernel.

        ``peers`` (a :class:`xbitinfo_b200.distributed.PeerExchange`): ``x`` is this rank's slab of a sharded
        variable and the finish kernel all-reduces the table over NVLink peer memory before K2
        (``xbi_bitinformation_dev_allreduce``); ``counts`` and ``mi`` then hold the global result on every rank."""
        require_cuda_tensor(x)
        if torch_dtype_code(x.dtype) != self.code:
            raise ValueError("dtype differs from the counter's dtype")
        outer, n, inner = dims if dims is not None else split_axis(tuple(x.shape), axis)
        np_dtype = numpy_dtype_of(x)
        flags, bits = nat.mask_spec(np_dtype, masked_value)
        if np_dtype.kind == "f" and transform:
            flags |= nat.SIGNED_EXPONENT
        if flavour is not None:
            flags |= {"sparse": nat.FORCE_SPARSE, "dense": nat.FORCE_DENSE, "generic": nat.FORCE_GENERIC}[flavour]
        mi_ptr = self.mi.data_ptr() if with_information else None
        with torch.cuda.device(x.device):
            common = (x.data_ptr(), self.code, outer, n, inner, flags, bits, int(bool(set_zero_insignificant)),
                      float(confidence), self.counts.data_ptr(), mi_ptr, self.workspace.data_ptr(),
                      self.workspace.numel(), _stream_ptr(x.device))
            if peers is None or peers.world == 1:
                rc = nat.lib().xbi_bitinformation_dev(*common)
            else:
                rc = nat.lib().xbi_bitinformation_dev_allreduce(*common, peers.world, peers.rank, peers.next_epoch(),
                                                                peers.table, float(peers.timeout_s))
        nat.check(rc)
        return self.mi if with_information else None

    def information(self, set_zero_insignificant: bool = False, confidence: float = 0.99) -> torch.Tensor:
        """K2 on the current counts -> float64[nbits] on the device."""
        with torch.cuda.device(self.device):
            rc = nat.lib().xbi_mutual_information(
                self.counts.data_ptr(), self.nbits, int(bool(set_zero_insignificant)), float(confidence),
                self.mi.data_ptr(), _stream_ptr(self.device))
        nat.check(rc)
        return self.mi


class MultiAxisCounter:
    """Reusable device state for the analysis of several dimensions of one array in one call
    (``xbi_bitinformation_dev_multi``): ``counts`` ``(naxes, nbits, 2, 2)`` int64 and ``mi`` ``(naxes, nbits)``
    float64, both overwritten by every call, plus the scratch workspace."""

    def __init__(self, dtype: torch.dtype, naxes: int, device):
        self.code = torch_dtype_code(dtype)
        self.nbits = 8 * torch.empty((), dtype=dtype).element_size()
        self.naxes = int(naxes)
        self.device = torch.device(device)
        self.counts = torch.zeros((self.naxes, self.nbits, 2, 2), dtype=torch.int64, device=self.device)
        self.mi = torch.empty((self.naxes, self.nbits), dtype=torch.float64, device=self.device)
        self.workspace = torch.empty(nat.lib().xbi_workspace_bytes_multi(self.naxes), dtype=torch.uint8, device=self.device)

    def bitinformation(self, x: torch.Tensor, axes, *, transform: bool = True, masked_value=None,
                       set_zero_insignificant: bool = False, confidence: float = 0.99, flavour: str | None = None):
        """Information per bit along each of ``axes`` (distinct) -> ``mi`` (device tensor ``(naxes, nbits)``)."""
        require_cuda_tensor(x)
        if torch_dtype_code(x.dtype) != self.code:
            raise ValueError("dtype differs from the counter's dtype")
        nd = x.dim()
        axes = [int(a) % nd for a in axes]
        if len(axes) != self.naxes or len(set(axes)) != len(axes):
            raise ValueError("`axes` must name naxes distinct dimensions")
        np_dtype = numpy_dtype_of(x)
        flags, bits = nat.mask_spec(np_dtype, masked_value)
        if np_dtype.kind == "f" and transform:
            flags |= nat.SIGNED_EXPONENT
        if flavour is not None:
            flags |= {"sparse": nat.FORCE_SPARSE, "dense": nat.FORCE_DENSE, "generic": nat.FORCE_GENERIC}[flavour]
        with torch.cuda.device(x.device):
            rc = nat.lib().xbi_bitinformation_dev_multi(
                x.data_ptr(), self.code, nd, nat.int64_array(x.shape), self.naxes, (nat.C.c_int * self.naxes)(*axes),
                flags, bits, int(bool(set_zero_insignificant)), float(confidence), self.counts.data_ptr(),
                self.mi.data_ptr(), self.workspace.data_ptr(), self.workspace.numel(), _stream_ptr(x.device))
        nat.check(rc)
        return self.mi


def bitinformation_multi_device(x: torch.Tensor, axes, masked_value=None, set_zero_insignificant: bool = False,
                                confidence: float = 0.99, return_counts: bool = False):
    """list of float64[nbits] (with ``return_counts``: of ``(mi, counts)``) in the order of ``axes``."""
    if not x.is_contiguous():
        x = x.contiguous()
    mc = MultiAxisCounter(x.dtype, len(axes), x.device)
    mi = mc.bitinformation(x, axes, masked_value=masked_value, set_zero_insignificant=set_zero_insignificant,
                           confidence=confidence).cpu().numpy()
    if return_counts:
        counts = mc.counts.cpu().numpy()
        return [(mi[i], counts[i]) for i in range(len(axes))]
    return [mi[i] for i in range(len(axes))]


def bitround_(x: torch.Tensor, keepbits: int) -> torch.Tensor:
    """K3 in place on a CUDA tensor."""
    require_cuda_tensor(x)
    code = torch_dtype_code(x.dtype)
    with torch.cuda.device(x.device):
        rc = nat.lib().xbi_bitround_inplace(x.data_ptr(), code, x.numel(), int(keepbits), _stream_ptr(x.device))
    nat.check(rc)
    return x


def bitround(x: torch.Tensor, keepbits: int, out: torch.Tensor | None = None) -> torch.Tensor:
    """K3 into a second tensor (``xbi_bitround``): ``x`` is read once and left untouched, the rounded values are
    written once -- the "new array, input not modified" contract of ``_np_bitround`` without a clone."""
    require_cuda_tensor(x)
    code = torch_dtype_code(x.dtype)
    if out is None:
        out = torch.empty_like(x)
    else:
        require_cuda_tensor(out)
        if out.dtype != x.dtype or out.numel() != x.numel() or out.device != x.device:
            raise ValueError("`out` must match the input in dtype, size and device")
    with torch.cuda.device(x.device):
        rc = nat.lib().xbi_bitround(x.data_ptr(), out.data_ptr(), code, x.numel(), int(keepbits), _stream_ptr(x.device))
    nat.check(rc)
    return out


# --------------------------------------------------------------------------------------
# batched over a chunk grid, regular or irregular (docs/chunking.ipynb workflow, csrc/xbi_chunked.cu)
# --------------------------------------------------------------------------------------
def _grid_tables(shape, chunk_shape, device):
    """None for a regular grid (every entry of ``chunk_shape`` an int), else ``(nchunks per dimension,
    offsets on the host, the same offsets as an int64 CUDA tensor)`` for per-dimension tuples of chunk
    lengths (dask's ``.chunks``); ints and tuples may be mixed."""
    if len(shape) != len(chunk_shape):
        raise ValueError("chunk_shape needs one entry per dimension")
    if not any(isinstance(c, (tuple, list)) for c in chunk_shape):
        return None
    nchunks, offsets = [], []
    for n, c in zip(shape, chunk_shape):
        n = int(n)
        if isinstance(c, (tuple, list)):
            lens = [int(v) for v in c]
        else:
            c = min(int(c), max(n, 1))
            if c < 1:
                raise ValueError("chunk extents must be positive")
            lens = [c] * (n // c) + ([n % c] if n % c else [])
        if sum(lens) != n or any(v < 1 for v in lens):
            raise ValueError(f"chunk lengths {tuple(lens)} do not tile an extent of {n}")
        nchunks.append(len(lens))
        offsets.extend(np.concatenate([[0], np.cumsum(lens)]).astype(np.int64).tolist())
    host = np.asarray(offsets, dtype=np.int64)
    return nchunks, host, torch.from_numpy(host).to(device)


def chunk_grid_shape(shape, chunk_shape):
    """Chunks per dimension of a grid given by chunk lengths (ints: regular, the last chunk of a dimension
    may be shorter; tuples: explicit lengths)."""
    if len(shape) != len(chunk_shape):
        raise ValueError("chunk_shape needs one entry per dimension")
    out = []
    for n, c in zip(shape, chunk_shape):
        if isinstance(c, (tuple, list)):
            out.append(len(c))
            continue
        c = int(c)
        if c < 1:
            raise ValueError("chunk extents must be positive")
        out.append(-(-int(n) // min(c, max(int(n), 1))) if n else 0)
    return tuple(out)


def count_pairs_chunked(x: torch.Tensor, chunk_shape, axis: int, *, transform: bool = True,
                        masked_value=None, flavour: str | None = None) -> torch.Tensor:
    """K1 on every chunk of a grid on its own, one launch: int64 tensor of shape
    ``chunks-per-dimension + (nbits, 2, 2)`` (chunks in C order of the grid).  ``chunk_shape``: one int
    (regular) or one tuple of chunk lengths (dask's ``.chunks``) per dimension."""
    require_cuda_tensor(x)
    code = torch_dtype_code(x.dtype)
    nd = x.dim()
    if not -nd <= axis < nd:
        raise ValueError(f"axis {axis} out of range for {nd}-d array")
    axis %= nd
    grid = chunk_grid_shape(tuple(x.shape), chunk_shape)
    tables = _grid_tables(tuple(x.shape), chunk_shape, x.device)
    nbits = 8 * x.element_size()
    counts = torch.empty(grid + (nbits, 2, 2), dtype=torch.int64, device=x.device)
    np_dtype = numpy_dtype_of(x)
    flags, bits = nat.mask_spec(np_dtype, masked_value)
    if np_dtype.kind == "f" and transform:
        flags |= nat.SIGNED_EXPONENT
    if flavour is not None:  # testing: pin one of the two chunk-grid kernels
        # (4- / 8-byte types: "walk" lets the sampled mask density choose between the walkers on two streams +
        # boundary corrections and those on three explicit streams; "walk2" / "walk3" pin one of them)
        flags |= {"walk": nat.FORCE_WALK, "walk2": nat.FORCE_WALK | nat.FORCE_SPARSE,
                  "walk3": nat.FORCE_WALK | nat.FORCE_DENSE, "odometer": nat.FORCE_GENERIC}[flavour]
    with torch.cuda.device(x.device):
        if tables is None:
            rc = nat.lib().xbi_count_pairs_chunked(
                x.data_ptr(), code, nd, nat.int64_array(x.shape), nat.int64_array(chunk_shape), axis, flags, bits,
                counts.data_ptr(), _stream_ptr(x.device))
        else:
            nchunks, host, dev = tables
            rc = nat.lib().xbi_count_pairs_chunked_v(
                x.data_ptr(), code, nd, nat.int64_array(x.shape), nat.int64_array(nchunks),
                host.ctypes.data_as(nat.C.POINTER(nat.C.c_int64)), dev.data_ptr(), axis, flags, bits,
                counts.data_ptr(), _stream_ptr(x.device))
            dev.record_stream(torch.cuda.current_stream(x.device))
    nat.check(rc)
    return counts


def information_batched(counts: torch.Tensor, set_zero_insignificant: bool = False,
                        confidence: float = 0.99) -> torch.Tensor:
    """K2 on a stack of (nbits, 2, 2) tables -> float64 tensor ``counts.shape[:-3] + (nbits,)``."""
    require_cuda_tensor(counts)
    if counts.dtype != torch.int64 or counts.dim() < 3 or tuple(counts.shape[-2:]) != (2, 2):
        raise ValueError("counts must be an int64 tensor of shape (..., nbits, 2, 2)")
    nbits = int(counts.shape[-3])
    nbatch = counts.numel() // (nbits * 4) if nbits else 0
    mi = torch.empty(tuple(counts.shape[:-2]), dtype=torch.float64, device=counts.device)
    with torch.cuda.device(counts.device):
        rc = nat.lib().xbi_mutual_information_batched(
            counts.data_ptr(), nbits, nbatch, int(bool(set_zero_insignificant)), float(confidence),
            mi.data_ptr(), _stream_ptr(counts.device))
    nat.check(rc)
    return mi


def bitround_chunked_(x: torch.Tensor, chunk_shape, keepbits: torch.Tensor) -> torch.Tensor:
    """K3 in place, chunk c of the grid rounded to ``keepbits[c]`` mantissa bits (int32 CUDA tensor
    with one entry per chunk, C order of the grid; ``chunk_shape`` as for :func:`count_pairs_chunked`)."""
    require_cuda_tensor(x)
    require_cuda_tensor(keepbits)
    code = torch_dtype_code(x.dtype)
    grid = chunk_grid_shape(tuple(x.shape), chunk_shape)
    if keepbits.dtype != torch.int32 or keepbits.numel() != math.prod(grid):
        raise ValueError("keepbits must be an int32 tensor with one entry per chunk")
    tables = _grid_tables(tuple(x.shape), chunk_shape, x.device)
    with torch.cuda.device(x.device):
        if tables is None:
            rc = nat.lib().xbi_bitround_chunked(
                x.data_ptr(), code, x.dim(), nat.int64_array(x.shape), nat.int64_array(chunk_shape),
                keepbits.data_ptr(), _stream_ptr(x.device))
        else:
            nchunks, host, dev = tables
            rc = nat.lib().xbi_bitround_chunked_v(
                x.data_ptr(), code, x.dim(), nat.int64_array(x.shape), nat.int64_array(nchunks),
                host.ctypes.data_as(nat.C.POINTER(nat.C.c_int64)), dev.data_ptr(), keepbits.data_ptr(),
                _stream_ptr(x.device))
            dev.record_stream(torch.cuda.current_stream(x.device))
    nat.check(rc)
    return x
