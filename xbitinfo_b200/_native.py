"""ctypes binding of libxbitinfo_b200.so (C ABI: include/xbitinfo_b200.h).

There is deliberately no fallback: if the shared library is missing or no CUDA device is
present, calls raise.  Build the library with ``python -m xbitinfo_b200.csrc.build`` (or
``__graft_entry__.build()``).
"""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "csrc", "libxbitinfo_b200.so")

# enum xbi_dtype
U8, U16, U32, U64, F16, F32, F64 = range(7)
# flags
SIGNED_EXPONENT, MASK_NAN, MASK_VALUE, FORCE_GENERIC = 1, 2, 4, 8
FORCE_WALK = 64  # testing: the walker kernel of the chunk-grid entry points (FORCE_GENERIC: the odometer kernel)
FORCE_SPARSE, FORCE_DENSE = 16, 32  # testing: flavour of the masked fast kernels (default: sampled on the device)
# xbi_shuffle kinds
SHUFFLE_BYTES, SHUFFLE_BITS = 0, 1
# error codes
ERR_ARG, ERR_UNSUPPORTED, ERR_NO_DEVICE, ERR_CUDA_BASE = 1, 2, 3, 1000

_DTYPE_CODES = {
    "uint8": U8, "int8": U8,
    "uint16": U16, "int16": U16,
    "uint32": U32, "int32": U32,
    "uint64": U64, "int64": U64,
    "float16": F16, "float32": F32, "float64": F64,
}


class XbiError(RuntimeError):
    """CUDA-side failure reported by libxbitinfo_b200."""


def dtype_code(dtype) -> int:
    name = np.dtype(dtype).name
    try:
        return _DTYPE_CODES[name]
    except KeyError:
        raise ValueError(f"dtype {name} neither known nor implemented.") from None


_lib = None


def lib() -> C.CDLL:
    """The loaded library (loads on first use; raises if it has not been built)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: the CUDA extension has not been built "
            "(run `python -m xbitinfo_b200.csrc.build`). xbitinfo_b200 has no CPU fallback."
        )
    L = C.CDLL(LIB_PATH)
    i64, u64, vp, i32, u32, dbl = C.c_int64, C.c_uint64, C.c_void_p, C.c_int, C.c_uint, C.c_double
    L.xbi_abi_version.restype = i32
    L.xbi_abi_version.argtypes = []
    L.xbi_last_error.restype = C.c_char_p
    L.xbi_last_error.argtypes = []
    L.xbi_kernel_launches.restype = C.c_ulonglong
    L.xbi_kernel_launches.argtypes = []
    L.xbi_fast_path_launches.restype = C.c_ulonglong
    L.xbi_fast_path_launches.argtypes = []
    L.xbi_profile_enable.restype = None
    L.xbi_profile_enable.argtypes = [i32]
    L.xbi_profile_collect.restype = i32
    L.xbi_profile_collect.argtypes = [vp, vp, vp]
    L.xbi_workspace_bytes.restype = C.c_size_t
    L.xbi_workspace_bytes.argtypes = []
    L.xbi_count_pairs.restype = i32
    L.xbi_count_pairs.argtypes = [vp, i32, i64, i64, i64, u32, u64, vp, vp, C.c_size_t, vp]
    L.xbi_mutual_information.restype = i32
    L.xbi_mutual_information.argtypes = [vp, i32, i32, dbl, vp, vp]
    L.xbi_bitinformation_dev.restype = i32
    L.xbi_bitinformation_dev.argtypes = [vp, i32, i64, i64, i64, u32, u64, i32, dbl, vp, vp, vp, C.c_size_t, vp]
    L.xbi_workspace_bytes_multi.restype = C.c_size_t
    L.xbi_workspace_bytes_multi.argtypes = [i32]
    L.xbi_bitinformation_dev_multi.restype = i32
    L.xbi_bitinformation_dev_multi.argtypes = [vp, i32, i32, C.POINTER(i64), i32, C.POINTER(i32), u32, u64, i32, dbl, vp, vp,
                                               vp, C.c_size_t, vp]
    L.xbi_bitinformation_dev_allreduce.restype = i32
    L.xbi_bitinformation_dev_allreduce.argtypes = [vp, i32, i64, i64, i64, u32, u64, i32, dbl, vp, vp, vp, C.c_size_t, vp,
                                                   i32, i32, u64, C.POINTER(vp), dbl]
    L.xbi_peer_bytes.restype = C.c_size_t
    L.xbi_peer_bytes.argtypes = []
    L.xbi_peer_handle_bytes.restype = C.c_size_t
    L.xbi_peer_handle_bytes.argtypes = []
    L.xbi_peer_create.restype = i32
    L.xbi_peer_create.argtypes = [C.POINTER(vp), vp]
    L.xbi_peer_open.restype = i32
    L.xbi_peer_open.argtypes = [vp, C.POINTER(vp)]
    L.xbi_peer_close.restype = i32
    L.xbi_peer_close.argtypes = [vp]
    L.xbi_peer_destroy.restype = i32
    L.xbi_peer_destroy.argtypes = [vp]
    L.xbi_peer_error_epoch.restype = i32
    L.xbi_peer_error_epoch.argtypes = [vp, C.POINTER(C.c_ulonglong)]
    L.xbi_mutual_information_batched.restype = i32
    L.xbi_mutual_information_batched.argtypes = [vp, i32, i64, i32, dbl, vp, vp]
    L.xbi_count_pairs_chunked.restype = i32
    L.xbi_count_pairs_chunked.argtypes = [vp, i32, i32, C.POINTER(i64), C.POINTER(i64), i32, u32, u64, vp, vp]
    L.xbi_bitround_chunked.restype = i32
    L.xbi_bitround_chunked.argtypes = [vp, i32, i32, C.POINTER(i64), C.POINTER(i64), vp, vp]
    L.xbi_count_pairs_chunked_v.restype = i32
    L.xbi_count_pairs_chunked_v.argtypes = [vp, i32, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), vp, i32, u32, u64,
                                            vp, vp]
    L.xbi_bitround_chunked_v.restype = i32
    L.xbi_bitround_chunked_v.argtypes = [vp, i32, i32, C.POINTER(i64), C.POINTER(i64), C.POINTER(i64), vp, vp, vp]
    L.xbi_bitround_inplace.restype = i32
    L.xbi_bitround_inplace.argtypes = [vp, i32, i64, i32, vp]
    L.xbi_bitround.restype = i32
    L.xbi_bitround.argtypes = [vp, vp, i32, i64, i32, vp]
    L.xbi_shuffle.restype = i32
    L.xbi_shuffle.argtypes = [vp, vp, i32, i64, i32, i32, vp]
    L.xbi_bitround_shuffle.restype = i32
    L.xbi_bitround_shuffle.argtypes = [vp, vp, i32, i64, i32, i32, vp]
    L.xbi_bitinformation_host.restype = i32
    L.xbi_bitinformation_host.argtypes = [vp, i32, i64, i64, i64, u32, u64, i32, dbl, vp, vp, i32]
    L.xbi_bitinformation_host_multi.restype = i32
    L.xbi_bitinformation_host_multi.argtypes = [vp, i32, i32, C.POINTER(i64), i32, C.POINTER(i32), u32, u64, i32, dbl,
                                                vp, vp, i32]
    L.xbi_bitround_host.restype = i32
    L.xbi_bitround_host.argtypes = [vp, vp, i32, i64, i32, i32]
    L.xbi_host_release.restype = None
    L.xbi_host_release.argtypes = []
    if L.xbi_abi_version() != 1:
        raise ImportError("libxbitinfo_b200.so ABI version mismatch; rebuild it")
    _lib = L
    # the host-buffer entry points keep a staging context per host thread; worker threads free theirs when they
    # exit (the library's thread_local destructor), the main thread's goes here, while the runtime is still up
    import atexit

    atexit.register(L.xbi_host_release)
    return L


EXPORTED_SYMBOLS = (
    "xbi_abi_version", "xbi_last_error", "xbi_kernel_launches", "xbi_fast_path_launches",
    "xbi_workspace_bytes", "xbi_profile_enable", "xbi_profile_collect",
    "xbi_count_pairs", "xbi_mutual_information", "xbi_bitround_inplace", "xbi_shuffle", "xbi_bitround_shuffle",
    "xbi_bitinformation_host", "xbi_bitround_host", "xbi_host_release",
    "xbi_mutual_information_batched", "xbi_count_pairs_chunked", "xbi_bitround_chunked",
    "xbi_bitinformation_host_multi", "xbi_count_pairs_chunked_v", "xbi_bitround_chunked_v",
    "xbi_bitinformation_dev", "xbi_bitinformation_dev_allreduce", "xbi_peer_bytes", "xbi_peer_handle_bytes",
    "xbi_bitround", "xbi_workspace_bytes_multi", "xbi_bitinformation_dev_multi", "xbi_peer_create", "xbi_peer_open", "xbi_peer_close", "xbi_peer_destroy", "xbi_peer_error_epoch",
)


def check(rc: int) -> None:
    """Translate a return code into the exception the reference interface would raise."""
    if rc == 0:
        return
    msg = lib().xbi_last_error().decode("utf-8", "replace")
    if rc == ERR_ARG:
        raise ValueError(msg)
    if rc == ERR_UNSUPPORTED:
        raise NotImplementedError(msg)
    raise XbiError(f"[xbi rc={rc}] {msg}")


def int64_array(values):
    """ctypes int64[] of a sequence of ints (shape / chunk_shape arguments)."""
    values = [int(v) for v in values]
    return (C.c_int64 * len(values))(*values)


def kernel_launches() -> int:
    return int(lib().xbi_kernel_launches())


def fast_path_launches() -> int:
    return int(lib().xbi_fast_path_launches())


def profile_enable(on: bool) -> None:
    lib().xbi_profile_enable(int(bool(on)))


def profile_collect():
    """(total kernel ms, timed launches, input bytes covered) since profile_enable(True)."""
    ms, n, b = C.c_double(0.0), C.c_ulonglong(0), C.c_ulonglong(0)
    check(lib().xbi_profile_collect(C.byref(ms), C.byref(n), C.byref(b)))
    return ms.value, int(n.value), int(b.value)


def mask_spec(dtype, masked_value):
    """(flag, mask_bits) for a ``masked_value`` keyword (None / "nothing" -> no mask,
    NaN -> any NaN, other -> bit identity with the value cast to ``dtype``)."""
    dtype = np.dtype(dtype)
    if masked_value is None or (isinstance(masked_value, str) and masked_value == "nothing"):
        return 0, 0
    if dtype.kind == "f":
        try:
            is_nan = bool(np.isnan(masked_value))
        except TypeError:
            is_nan = False
        if is_nan:
            return MASK_NAN, 0
    mv = np.array(masked_value).astype(dtype)
    bits = int(mv.view(f"u{dtype.itemsize}"))
    return MASK_VALUE, bits
