"""Per-chunk bit information, keepbits and bitround: the workflow of the reference's
``docs/chunking.ipynb`` (cells 8-10: for every dask block run ``get_bitinformation`` ->
``get_keepbits`` -> ``xr_bitround`` on that block alone and write it to its region), with the
blocks cut and processed on the device.

Every block is an independent K1 -> K2 -> (host: keepbits) -> K3 job; the array stays in HBM,
only ``nbits`` doubles per block come back to the host for the keepbits decision.  Regular chunk
grids go through the batched kernels of ``csrc/xbi_chunked.cu`` (all blocks in one launch per stage).
"""
from __future__ import annotations

from itertools import product

import numpy as np
import torch

from .device import PairCounter, bitround_
from .keepbits import _cdf_from_info_per_bit
from .bitinfo import bit_partitioning

_MANT = {torch.float16: 10, torch.float32: 23, torch.float64: 52}


def slices_from_chunks(chunks):
    """Chunks tuple -> slices in product order (``docs/chunking.ipynb`` cell 8).

    >>> slices_from_chunks(((2, 2), (3, 3, 3)))[1]
    (slice(0, 2, None), slice(3, 6, None))
    """
    cumdims = []
    for bds in chunks:
        out = np.empty(len(bds) + 1, dtype=int)
        out[0] = 0
        np.cumsum(bds, out=out[1:])
        cumdims.append(out)
    slices = [[slice(int(s), int(s + d)) for s, d in zip(starts, shapes)] for starts, shapes in zip(cumdims, chunks)]
    return list(product(*slices))


def normalize_chunks(shape, chunks):
    """``{axis: size}`` / per-axis sizes / explicit tuples -> tuple of per-axis chunk-length tuples."""
    out = []
    for ax, n in enumerate(shape):
        c = chunks.get(ax, n) if isinstance(chunks, dict) else chunks[ax]
        if isinstance(c, (tuple, list)):
            if sum(c) != n:
                raise ValueError(f"chunks along axis {ax} do not add up to {n}")
            out.append(tuple(int(v) for v in c))
        else:
            c = int(n if c in (None, -1) else c)
            out.append(tuple([c] * (n // c) + ([n % c] if n % c else [])))
    return tuple(out)


def keepbits_of_information(info, np_dtype, inflevel=0.99) -> int:
    """``get_keepbits`` on one information vector (``xbitinfo/xbitinfo.py:686-717``)."""
    n_bits, _, _, n_mant = bit_partitioning(np_dtype)
    if inflevel == 1.0:
        return n_mant
    cdf = _cdf_from_info_per_bit(np.asarray(info, dtype="float64"))
    with np.errstate(invalid="ignore"):
        return int((cdf > inflevel).argmax() + 1 - (n_bits - n_mant))


def _chunk_spec(grid):
    """Per-dimension chunk description for the batched kernels: an int where the dimension is cut regularly
    (equal chunks, the last possibly shorter), else the tuple of chunk lengths."""
    out = []
    for c in grid:
        if len(c) == 0:
            out.append(1)
        elif all(v == c[0] for v in c[:-1]) and c[-1] <= c[0]:
            out.append(int(c[0]))
        else:
            out.append(tuple(int(v) for v in c))
    return tuple(out)


def keepbits_of_information_batched(info, np_dtype, inflevel=0.99) -> np.ndarray:
    """``get_keepbits`` on a stack of information vectors ``(..., nbits)`` -> int64 ``(...)``."""
    n_bits, _, _, n_mant = bit_partitioning(np_dtype)
    info = np.asarray(info, dtype="float64")
    if inflevel == 1.0:
        return np.full(info.shape[:-1], n_mant, dtype="int64")
    cdf = _cdf_from_info_per_bit(info)
    with np.errstate(invalid="ignore"):
        return ((cdf > inflevel).argmax(axis=-1) + 1 - (n_bits - n_mant)).astype("int64")


def chunk_keepbits(info, np_dtype, inflevel=0.99) -> np.ndarray:
    """Keepbits per chunk from a stack of information vectors ``(..., nbits)``, ready for K3: negative values
    are clamped to 0 as the reference's multi-file flow does (``xbitinfo/xbitinfo.py:867-868``).  A chunk whose
    information is all NaN had NO valid pair along the analysed axis (extent 1 along it -- e.g. the ragged last
    chunk of 101 cut in 50s -- or everything masked): there is nothing to base a decision on, so it keeps every
    mantissa bit (left unrounded) and a warning says so; clamping its meaningless negative keepbits to 0 would
    silently wipe the mantissa."""
    import warnings

    _, _, _, n_mant = bit_partitioning(np_dtype)
    info = np.asarray(info, dtype="float64")
    keep = np.clip(keepbits_of_information_batched(info, np_dtype, inflevel), 0, n_mant)
    undecided = np.isnan(info).all(axis=-1)
    if undecided.any():
        warnings.warn(f"{int(undecided.sum())} chunk(s) hold no valid pair along the analysed axis; they are left "
                      f"unrounded (keepbits={n_mant})", RuntimeWarning, stacklevel=3)
        keep = np.where(undecided, n_mant, keep)
    return keep.astype("int64")


def bitinformation_by_chunks(x, chunks, axis: int, masked_value=float("nan"), set_zero_insignificant: bool = True,
                             confidence: float = 0.99):
    """Information per bit of every chunk of the grid on its own: ``(info, counts)`` as CUDA tensors of
    shape ``chunks-per-dimension + (nbits,)`` and ``+ (nbits, 2, 2)``.  One counting launch for all chunks
    (``xbi_count_pairs_chunked``) and one for the mutual information."""
    from .device import count_pairs_chunked, information_batched

    t = torch.from_numpy(np.ascontiguousarray(x)).cuda() if isinstance(x, np.ndarray) else x.contiguous()
    grid = normalize_chunks(tuple(t.shape), chunks)
    counts = count_pairs_chunked(t, _chunk_spec(grid), axis, masked_value=masked_value)
    return information_batched(counts, set_zero_insignificant, confidence), counts


def bitround_by_chunks(x, chunks, axis: int, inflevel: float = 0.99, masked_value=float("nan"),
                       set_zero_insignificant: bool = True, confidence: float = 0.99, batched: bool | None = None):
    """Round every chunk of ``x`` to its own keepbits.

    ``x``: numpy array or CUDA tensor (floats); ``chunks``: ``{axis: size}``, per-axis sizes or
    explicit per-axis tuples (dask's ``.chunks``); ``axis``: the dimension analysed inside each
    chunk.  Returns ``(rounded, keepbits)`` with ``rounded`` of the type of ``x`` (a new array)
    and ``keepbits`` an int64 array with one entry per chunk (shape = chunks per axis).
    Negative keepbits (not even all exponent bits carry information) are clamped to 0, as the
    reference's multi-file flow does (``xbitinfo/xbitinfo.py:867-868``); a chunk without any valid pair
    along ``axis`` is left unrounded, with a warning (:func:`chunk_keepbits`).

    The grid (regular, or dask-style tuples of uneven chunk lengths) is processed in a constant number
    of launches whatever the number of chunks: ``xbi_count_pairs_chunked[_v]`` ->
    ``xbi_mutual_information_batched`` -> (one D2H of the information, keepbits on the host, one H2D of
    ``nchunks`` ints) -> ``xbi_bitround_chunked[_v]``.  ``batched=False`` walks the grid chunk by chunk with
    K1/K2/K3 instead (the reference notebook's loop; kept for comparison).
    """
    is_np = isinstance(x, np.ndarray)
    t = torch.from_numpy(np.ascontiguousarray(x)).cuda() if is_np else x.contiguous().clone()
    if t.dtype not in _MANT:
        raise TypeError("Only float arrays (16-64bit) can be bit-rounded")
    np_dtype = np.dtype(str(t.dtype).replace("torch.", ""))
    grid = normalize_chunks(tuple(t.shape), chunks)
    keep = np.zeros([len(c) for c in grid], dtype="int64")
    axis %= t.dim()
    chunk_shape = _chunk_spec(grid)
    if batched is None:
        batched = True
    if batched:
        from .device import bitround_chunked_, count_pairs_chunked, information_batched

        counts = count_pairs_chunked(t, chunk_shape, axis, masked_value=masked_value)
        info = information_batched(counts, set_zero_insignificant, confidence).cpu().numpy()
        keep = chunk_keepbits(info, np_dtype, inflevel).reshape(keep.shape)
        if keep.size:
            bitround_chunked_(t, chunk_shape, torch.from_numpy(keep.astype("int32").ravel()).to(t.device))
        return (t.cpu().numpy() if is_np else t), keep
    pc = PairCounter(t.dtype, t.device)
    for idx, sl in zip(np.ndindex(*keep.shape), slices_from_chunks(grid)):
        blk = t[sl].contiguous()
        pc.reset()
        pc.add(blk, axis, masked_value=masked_value)
        info = pc.information(set_zero_insignificant, confidence).cpu().numpy()
        k = int(chunk_keepbits(info[None, :], np_dtype, inflevel)[0])
        keep[idx] = k
        bitround_(blk, k)
        t[sl] = blk
    return (t.cpu().numpy() if is_np else t), keep


def xr_bitround_by_chunks(obj, chunks, dim, inflevel: float = 0.99, **kwargs):
    """The chunking notebook on a labelled object: every block of the grid ``chunks`` (``{dim name: chunk
    length or tuple of lengths}``, dimensions left out are not cut -- what ``ds.chunk(chunks)`` would give)
    is analysed along ``dim`` and rounded to its own keepbits at ``inflevel``.

    Returns ``(rounded, keepbits)``: ``rounded`` like ``obj`` (Dataset or DataArray, attrs kept), ``keepbits``
    an int64 DataArray per variable with the dimensions of the variable and one entry per block.  Variables
    without ``dim`` are returned unchanged (as ``get_bitinformation`` skips them, ``xbitinfo.py:358-363``).
    ``kwargs``: ``masked_value``, ``set_zero_insignificant``, ``confidence`` as for ``get_bitinformation``.
    """
    from . import _xr
    from .bitinfo import _resolve_kwargs

    ns = _xr.namespace_for(obj)
    if _xr.is_dataset(obj):
        rounded, keep = obj.copy(), ns.Dataset()
        for v in obj.data_vars:
            if dim not in obj[v].dims:
                continue
            rounded[v], keep[v] = xr_bitround_by_chunks(obj[v], chunks, dim, inflevel, **kwargs)
        return rounded, keep
    da = obj
    if dim not in da.dims:
        raise ValueError(f"{dim} is not a dimension of {da.name}")
    chunks = {d: c for d, c in chunks.items() if d in da.dims}  # a Dataset-wide grid may name other dimensions
    masked_value, set_zero, confidence = _resolve_kwargs(da.dtype, kwargs)
    axis = list(da.dims).index(dim)
    by_axis = {list(da.dims).index(d): c for d, c in chunks.items()}
    data = _xr.raw_data(da)
    new, keep = bitround_by_chunks(data, by_axis, axis, inflevel=inflevel, masked_value=masked_value,
                                   set_zero_insignificant=set_zero, confidence=confidence)
    if _xr.is_xarray(da):  # pragma: no cover - xarray is absent from the build image
        out = da.copy(data=new)
    else:
        out = da._replace(new)
    return out, ns.DataArray(keep, dims=list(da.dims), name=da.name)
