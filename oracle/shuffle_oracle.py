"""CPU restatement of the shuffle pre-filters that follow bitround on the reference's save
path (TEST INFRASTRUCTURE: only tests/, smoke() and bench.py's CPU legs may import this).

* byte shuffle: the HDF5 shuffle filter that netCDF4's ``shuffle=True`` switches on
  (xbitinfo/save_compressed.py:44-80 passes the flag; the arithmetic lives in libhdf5's
  H5Zshuffle.c, not vendored) and Blosc's SHUFFLE: byte j of every element is gathered into
  plane j.
* bit shuffle: Blosc's BITSHUFFLE / the bitshuffle library (xbitinfo/save_compressed.py:155-209
  selects it through ``BloscShuffle.bitshuffle``; c-blosc is not vendored): the element x bit
  matrix is transposed, bit planes ordered from the least significant bit of byte 0, eight
  elements per output byte with the first element in the least significant bit.  Blosc applies
  it per internal block and only to whole groups of eight elements; here the block is the
  buffer handed in.

PARITY UNPINNED: neither libhdf5 nor c-blosc (nor numcodecs/h5py/bitshuffle) exists in this
image, so these restate the published layouts; the tests pin the CUDA kernels to this
restatement and to round trips.
"""
import numpy as np


def byte_shuffle(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    s = x.dtype.itemsize
    return np.ascontiguousarray(x.reshape(-1).view("u1").reshape(-1, s).T).reshape(-1)


def byte_unshuffle(buf: np.ndarray, dtype) -> np.ndarray:
    dtype = np.dtype(dtype)
    s = dtype.itemsize
    b = np.ascontiguousarray(buf).view("u1").reshape(s, -1)
    return np.ascontiguousarray(b.T).reshape(-1).view(dtype)


def bit_shuffle(x: np.ndarray) -> np.ndarray:
    x = np.ascontiguousarray(x)
    s = x.dtype.itemsize
    n = x.size
    if n % 8:
        raise ValueError("bit shuffle needs a multiple of 8 elements")
    bits = np.unpackbits(x.reshape(-1).view("u1").reshape(n, s), axis=1, bitorder="little")  # [element, 8*byte+bit]
    return np.packbits(np.ascontiguousarray(bits.T), axis=1, bitorder="little").reshape(-1)   # [plane, n/8]


def bit_unshuffle(buf: np.ndarray, dtype) -> np.ndarray:
    dtype = np.dtype(dtype)
    s = dtype.itemsize
    planes = np.ascontiguousarray(buf).view("u1").reshape(8 * s, -1)
    bits = np.unpackbits(planes, axis=1, bitorder="little")  # [plane, element]
    return np.packbits(np.ascontiguousarray(bits.T), axis=1, bitorder="little").reshape(-1).view(dtype)
