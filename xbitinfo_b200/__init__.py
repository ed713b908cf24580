"""xbitinfo_b200 -- the bit-information hot path of xbitinfo on NVIDIA B200 (sm_100a).

Public names follow ``xbitinfo/__init__.py:10-13`` for the path this package covers:
``get_bitinformation`` (with ``implementation="cuda"``), ``get_keepbits`` and
``xr_bitround``.  Everything numerical runs in hand-written CUDA kernels behind the C ABI
declared in ``include/xbitinfo_b200.h``; importing the package does not need a GPU, calling
it does (no CPU fallback).
"""
from ._version import __version__
from ._labelled import DataArray, Dataset
from .bitinfo import get_bitinformation, get_bit_coords, bit_partitioning, load_bitinformation
from .bitround import xr_bitround
from .keepbits import get_keepbits

__all__ = [
    "get_bitinformation",
    "get_keepbits",
    "xr_bitround",
    "get_bit_coords",
    "bit_partitioning",
    "load_bitinformation",
    "DataArray",
    "Dataset",
    "__version__",
]
