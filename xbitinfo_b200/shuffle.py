"""Byte / bit shuffle on the device (K4) -- the lossless-compression pre-filters that follow
bitround on the reference's save path.

The reference does not shuffle itself: it asks the storage library to
(``xbitinfo/save_compressed.py:44-80``: netCDF ``shuffle=True`` -> HDF5 shuffle filter;
``:155-209``: zarr ``Blosc(cname="zstd", shuffle=BITSHUFFLE)``).  Here the permutation runs on
the GPU right after K3, so the rounded data leaves HBM already plane-ordered and the host
only has to run the entropy coder (``Blosc(shuffle=NOSHUFFLE)`` / plain zstd / zlib).

* device tensors: :func:`byte_shuffle`, :func:`bit_shuffle` and their inverses;
* host arrays: :class:`Shuffle` and :class:`BitShuffle` follow the numcodecs codec protocol
  (``codec_id``, ``encode``, ``decode``, ``get_config``, ``from_config``) so they can be used
  as zarr ``filters`` once numcodecs is present (``numcodecs.register_codec(Shuffle)``).
"""
from __future__ import annotations

import numpy as np
import torch

from . import _native as nat
from .device import _stream_ptr, require_cuda_tensor


def _run(x: torch.Tensor, itemsize: int, kind: int, inverse: bool, out_dtype) -> torch.Tensor:
    require_cuda_tensor(x)
    nbytes = x.numel() * x.element_size()
    if nbytes % itemsize:
        raise ValueError("buffer size is not a multiple of the element size")
    out = torch.empty(nbytes // torch.empty((), dtype=out_dtype).element_size(), dtype=out_dtype, device=x.device)
    with torch.cuda.device(x.device):
        rc = nat.lib().xbi_shuffle(x.data_ptr(), out.data_ptr(), itemsize, nbytes // itemsize, kind, int(inverse),
                                   _stream_ptr(x.device))
    nat.check(rc)
    return out


def byte_shuffle(x: torch.Tensor) -> torch.Tensor:
    """uint8 tensor of ``x.nbytes`` bytes: plane j holds byte j of every element of ``x``."""
    return _run(x.contiguous(), x.element_size(), nat.SHUFFLE_BYTES, False, torch.uint8)


def byte_unshuffle(buf: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return _run(buf.contiguous(), torch.empty((), dtype=dtype).element_size(), nat.SHUFFLE_BYTES, True, dtype)


def bit_shuffle(x: torch.Tensor) -> torch.Tensor:
    """uint8 tensor of ``x.nbytes`` bytes: plane p (= 8*byte + bit from the least significant
    bit) holds bit p of every element, eight elements per byte, first element in bit 0.
    ``x.numel()`` must be a multiple of 8."""
    return _run(x.contiguous(), x.element_size(), nat.SHUFFLE_BITS, False, torch.uint8)


def bit_unshuffle(buf: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
    return _run(buf.contiguous(), torch.empty((), dtype=dtype).element_size(), nat.SHUFFLE_BITS, True, dtype)


def bitround_shuffle(x: torch.Tensor, keepbits: int, bits: bool = False) -> torch.Tensor:
    """``shuffle(bitround(x, keepbits))`` in one pass over HBM (K3 fused into K4): the uint8
    planes of the rounded array, ``x`` itself untouched.  ``bits`` selects the bit shuffle."""
    from .device import torch_dtype_code

    x = x.contiguous()
    require_cuda_tensor(x)
    out = torch.empty(x.numel() * x.element_size(), dtype=torch.uint8, device=x.device)
    with torch.cuda.device(x.device):
        rc = nat.lib().xbi_bitround_shuffle(x.data_ptr(), out.data_ptr(), torch_dtype_code(x.dtype), x.numel(),
                                            int(keepbits), nat.SHUFFLE_BITS if bits else nat.SHUFFLE_BYTES,
                                            _stream_ptr(x.device))
    nat.check(rc)
    return out


class _DeviceFilter:
    """numcodecs-style codec whose permutation runs on the GPU (host buffer in, host buffer out)."""

    codec_id = None
    _kind = None

    def __init__(self, elementsize: int = 4, device: int = 0):
        if elementsize not in (2, 4, 8):
            raise ValueError("elementsize must be 2, 4 or 8")
        self.elementsize = int(elementsize)
        self.device = int(device)

    def _apply(self, buf, inverse: bool) -> np.ndarray:
        arr = np.ascontiguousarray(np.frombuffer(memoryview(buf).cast("B"), dtype="u1") if not isinstance(buf, np.ndarray)
                                   else buf).reshape(-1).view("u1")
        if arr.size % self.elementsize:
            raise ValueError("buffer size is not a multiple of elementsize")
        t = torch.from_numpy(arr.copy()).to(f"cuda:{self.device}")
        out = _run(t, self.elementsize, self._kind, inverse, torch.uint8)
        return out.cpu().numpy()

    def encode(self, buf) -> np.ndarray:
        return self._apply(buf, False)

    def decode(self, buf, out=None):
        dec = self._apply(buf, True)
        if out is not None:
            np.frombuffer(memoryview(out).cast("B"), dtype="u1")[:] = dec
            return out
        return dec

    def get_config(self) -> dict:
        return {"id": self.codec_id, "elementsize": self.elementsize}

    @classmethod
    def from_config(cls, config: dict):
        return cls(elementsize=config.get("elementsize", 4))

    def __repr__(self) -> str:
        return f"{type(self).__name__}(elementsize={self.elementsize})"


class Shuffle(_DeviceFilter):
    """Byte shuffle (HDF5 shuffle filter / ``numcodecs.Shuffle`` layout)."""

    codec_id = "shuffle"
    _kind = nat.SHUFFLE_BYTES


class BitShuffle(_DeviceFilter):
    """Bit shuffle over the whole buffer (the bit transpose Blosc's BITSHUFFLE applies per block)."""

    codec_id = "xbitinfo_b200_bitshuffle"
    _kind = nat.SHUFFLE_BITS
