/*
 * xbitinfo_b200 -- C ABI of the B200-native bit-information hot path.
 *
 * This is the drop-in boundary.  The reference (observingClouds/xbitinfo) has no FFI of
 * its own: its "operator interface" for this path is a pair of Python functions,
 *
 *   _py_get_bitinformation(ds, var, axis, dim, kwargs) -> {"bitinfo": f64[nbits], ...}
 *       xbitinfo/xbitinfo.py:337-370, selected by the `implementation=` string at
 *       xbitinfo/xbitinfo.py:424-439; arithmetic in xbitinfo/_py_bitinfo.py:42-160
 *   _np_bitround(data, keepbits) -> ndarray
 *       xbitinfo/bitround.py:7-12 (arithmetic: numcodecs.bitround.BitRound)
 *
 * The entry points below are what a ctypes binding of those two functions calls
 * (INTEGRATION.md shows the stub).  Conventions:
 *   - plain pointers and sizes only; every function returns 0 on success, otherwise a
 *     non-zero code (XBI_ERR_* or a cudaError_t value offset by XBI_ERR_CUDA_BASE) and
 *     xbi_last_error() returns a thread-local message;
 *   - *_dev pointers are device memory owned by the caller; kernels are ordered on the
 *     given stream (a cudaStream_t passed as void*, NULL = default stream); nothing is
 *     allocated or cached by the device-level calls, so they are re-entrant from
 *     several host threads (dask workers);
 *   - an array is described as a C-contiguous (outer, n, inner) view: the analysed
 *     axis has length n, `outer` is the product of the dimensions before it, `inner`
 *     the product of those after it;
 *   - counts are ACCUMULATED (+=) into counts_dev so slabs, shards and halo rows
 *     compose; the caller zeroes it once.  When a caller splits the analysed axis
 *     itself it passes an n that includes one halo row of the next slab.
 */
#ifndef XBITINFO_B200_H
#define XBITINFO_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XBI_ABI_VERSION 1

/* element types.  Signed integers are passed as the unsigned type of the same width
 * (the reference reinterprets them: xbitinfo/xbitinfo.py:355). */
enum xbi_dtype {
    XBI_U8 = 0,
    XBI_U16 = 1,
    XBI_U32 = 2,
    XBI_U64 = 3,
    XBI_F16 = 4,
    XBI_F32 = 5,
    XBI_F64 = 6
};

/* flags of xbi_count_pairs */
#define XBI_SIGNED_EXPONENT 1u /* floats: biased -> sign-magnitude exponent (_py_bitinfo.py:42-94) */
#define XBI_MASK_NAN 2u        /* floats: a pair with a NaN element (any payload) is not counted   */
#define XBI_MASK_VALUE 4u      /* a pair with an element bit-identical to mask_bits is not counted */
#define XBI_FORCE_GENERIC 8u   /* testing: skip the TMA fast paths                                  */
#define XBI_FORCE_SPARSE 16u   /* testing: masked fast kernels in their boundary-correction flavour */
#define XBI_FORCE_DENSE 32u    /* testing: ... in their three-stream flavour (default: sampled)     */
#define XBI_FORCE_WALK 64u     /* testing (chunk grids): the walker kernel whatever the grid;
                                  XBI_FORCE_GENERIC there: the per-element odometer kernel;
                                  XBI_FORCE_SPARSE / XBI_FORCE_DENSE there: the walkers of 4- / 8-byte types on
                                  two streams + boundary corrections / on three explicit streams      */

/* error codes */
#define XBI_OK 0
#define XBI_ERR_ARG 1
#define XBI_ERR_UNSUPPORTED 2 /* also: the TMA fast path cannot run on this driver (never a silent slow path) */
#define XBI_ERR_NO_DEVICE 3
#define XBI_ERR_CUDA_BASE 1000 /* + cudaError_t */

int xbi_abi_version(void);
const char* xbi_last_error(void);
/* number of kernels this library has launched so far in this process (all threads) */
unsigned long long xbi_kernel_launches(void);
/* how many of those were fast-path counting kernels (TMA rows / column walkers) */
unsigned long long xbi_fast_path_launches(void);

/*
 * Measurement aid (per host thread): while enabled, xbi_count_pairs brackets its fast-path
 * counting kernel with CUDA events on the caller's stream.  xbi_profile_collect waits for
 * them and returns the summed kernel time, the number of timed launches and the input
 * bytes those launches covered.  Enabling resets the tallies.
 */
void xbi_profile_enable(int on);
int xbi_profile_collect(double* total_ms, unsigned long long* launches, unsigned long long* bytes);

/* bytes of device scratch xbi_count_pairs needs (contents are overwritten) */
size_t xbi_workspace_bytes(void);

/*
 * K1: per-bit adjacent-pair histogram.   Replaces pb.bitinformation's counting
 * (xbitinfo/_py_bitinfo.py:97-140, 155-160) fused with pb.signed_exponent (42-94).
 *
 * counts_dev: nbits*4 unsigned 64-bit counters, nbits = 8*sizeof(element);
 *   counts_dev[4*k + 2*a + b] += #pairs whose bit k (k = 0 is the most significant bit)
 *   is `a` in the first and `b` in the second element -- the (nbits,2,2) layout of
 *   bitpaircount (_py_bitinfo.py:128-140).
 * mask_bits: element bit pattern for XBI_MASK_VALUE (low bits used for narrow types).
 */
int xbi_count_pairs(const void* x_dev, int dtype, int64_t outer, int64_t n, int64_t inner,
                    unsigned flags, uint64_t mask_bits, unsigned long long* counts_dev,
                    void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * K2: mutual information per bit from the counts, in bits.  Replaces
 * pb.mutual_information (xbitinfo/_py_bitinfo.py:143-152): p = c/N, zero cells skipped,
 * sum p*log(p/(pr*ps)) in the reference's order, / log(2), IEEE double, no FMA
 * contraction.  N = sum of the four cells (0 pairs -> NaN).  With
 * set_zero_insignificant != 0 every value <= the binomial free entropy at `confidence`
 * is set to 0 (BitInformation.jl semantics).
 */
int xbi_mutual_information(const unsigned long long* counts_dev, int nbits,
                           int set_zero_insignificant, double confidence, double* mi_dev,
                           void* stream);

/*
 * K1 + K2 of one device-resident variable along one axis in ONE call: what the backend function
 * _py_get_bitinformation (xbitinfo/xbitinfo.py:337-370: transform, cast, pb.bitinformation(X, axis).compute())
 * does for an array that already lives in HBM.  counts_dev (nbits*4) is OVERWRITTEN with the pair histogram
 * (layout as for xbi_count_pairs) and mi_dev (nbits doubles; NULL: counts only) receives the information per
 * bit (semantics of xbi_mutual_information).  Two or three launches in all: workspace clear, the counting
 * kernel, and one finish kernel that counts the few pairs the fast kernel leaves over, combines the raw sums
 * and runs K2.  No pairs along the axis: counts 0, information NaN.
 */
int xbi_bitinformation_dev(const void* x_dev, int dtype, int64_t outer, int64_t n, int64_t inner,
                           unsigned flags, uint64_t mask_bits, int set_zero_insignificant, double confidence,
                           unsigned long long* counts_dev, double* mi_dev, void* workspace_dev,
                           size_t workspace_bytes, void* stream);

/*
 * Several analysed dimensions of ONE device-resident array: get_bitinformation(ds) with dim=None, which the
 * reference serves by looping the dimensions and running the backend once per dimension
 * (xbitinfo/xbitinfo.py:373-406).  The C-contiguous array has `ndim` extents `shape`; counts_dev receives
 * naxes*nbits*4 counters and mi_dev (may be NULL) naxes*nbits doubles in the order of `axes`; results are
 * identical to naxes calls of xbi_bitinformation_dev.  When the two innermost dimensions are both asked for
 * (4-byte element types, rows a multiple of 16 bytes) they are counted in ONE read of the array by a
 * two-axis kernel; the remaining dimensions take one pass each.  With a mask flag the two-axis kernel counts
 * as if nothing were masked and looks for masked elements on the way: the call then waits for the stream
 * once, and if there are any recounts the two dimensions with the masked single-axis kernels.
 * workspace: xbi_workspace_bytes_multi(naxes) bytes of device scratch.
 */
size_t xbi_workspace_bytes_multi(int naxes);
int xbi_bitinformation_dev_multi(const void* x_dev, int dtype, int ndim, const int64_t* shape, int naxes,
                                 const int* axes, unsigned flags, uint64_t mask_bits, int set_zero_insignificant,
                                 double confidence, unsigned long long* counts_dev, double* mi_dev,
                                 void* workspace_dev, size_t workspace_bytes, void* stream);

/*
 * The same across the GPUs of one node, one process per GPU (every rank counts its slab, the
 * int64[nbits,2,2] tables are summed over the ranks, K2 runs on every rank): the all-reduce is part of the
 * finish kernel.  Each rank owns an exchange buffer (xbi_peer_create: xbi_peer_bytes() of device memory plus
 * a CUDA IPC handle of xbi_peer_handle_bytes() bytes to hand to the other ranks by whatever means the host
 * program has -- torch.distributed.all_gather_object in this package); a rank maps the buffers of the others
 * with xbi_peer_open.  peer_buffers[p] is rank p's buffer as addressable by THIS process (own buffer at
 * [rank]).  The finish kernel stores the rank's table into its slot of every peer's buffer over NVLink, raises
 * a flag there, waits for the flags of all peers in its own buffer, sums the slots: counts_dev and mi_dev then
 * hold the GLOBAL result on every rank.
 *   epoch: 1 for the first call, +1 for every later call, the same sequence on every rank.
 *   timeout_s: a rank that does not see its peers within this time stops waiting (no hung GPU); the call's
 *   results are then invalid and xbi_peer_error_epoch() returns the epoch of the failed wait (0: none so far).
 *   <= 0: 20 s.
 * All ranks must call it the same number of times; world <= 8.
 */
size_t xbi_peer_bytes(void);
size_t xbi_peer_handle_bytes(void);
int xbi_peer_create(void** buffer_dev, void* handle_out);
int xbi_peer_open(const void* handle, void** mapped_dev);
int xbi_peer_close(void* mapped_dev);
int xbi_peer_destroy(void* buffer_dev);
int xbi_peer_error_epoch(const void* own_buffer_dev, unsigned long long* epoch_out);
int xbi_bitinformation_dev_allreduce(const void* x_dev, int dtype, int64_t outer, int64_t n, int64_t inner,
                                     unsigned flags, uint64_t mask_bits, int set_zero_insignificant,
                                     double confidence, unsigned long long* counts_dev, double* mi_dev,
                                     void* workspace_dev, size_t workspace_bytes, void* stream, int world, int rank,
                                     uint64_t epoch, void* const* peer_buffers, double timeout_s);

/* K2 on `nbatch` consecutive (nbits,2,2) tables: mi_dev receives nbatch*nbits doubles. */
int xbi_mutual_information_batched(const unsigned long long* counts_dev, int nbits, int64_t nbatch,
                                   int set_zero_insignificant, double confidence, double* mi_dev, void* stream);

/*
 * K1 / K3 over a regular chunk grid, in a constant number of launches: the per-block workflow of
 * the reference's docs/chunking.ipynb (cells 8-10: get_bitinformation -> get_keepbits ->
 * xr_bitround on every dask block on its own; backend xbitinfo/xbitinfo.py:337-370 and
 * xbitinfo/bitround.py:7-12 called once per block there).
 *   The array is C-contiguous with `ndim` (<= XBI_MAX_DIMS) extents `shape`; chunk_shape[d] is the
 *   chunk length along dimension d (the last chunk of a dimension may be shorter; values above
 *   shape[d] mean "not chunked").  Chunks are numbered in C order of the chunk grid.
 *   xbi_count_pairs_chunked: counts_dev[chunk][nbits][4] is OVERWRITTEN with the pair histogram of
 *   every chunk taken on its own along `axis` (pairs never cross a chunk boundary); flags and
 *   mask_bits as for xbi_count_pairs (the XBI_FORCE_* flags pin one of the chunk-grid kernels for the
 *   tests, see their definitions).
 *   xbi_bitround_chunked: chunk c is rounded in place to keepbits_dev[c] mantissa bits (values are
 *   clamped to 0..mantissa bits; mantissa bits = unchanged).
 * XBI_ERR_UNSUPPORTED: more than 5 dimensions remain after merging unchunked neighbours, a chunk
 * extent or the number of chunks reaches 2^31.
 */
#define XBI_MAX_DIMS 8
int xbi_count_pairs_chunked(const void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* chunk_shape,
                            int axis, unsigned flags, uint64_t mask_bits, unsigned long long* counts_dev, void* stream);
int xbi_bitround_chunked(void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* chunk_shape,
                         const int32_t* keepbits_dev, void* stream);
/*
 * The same over an IRREGULAR grid (dask's `.chunks` after a concatenation or an uneven rechunk):
 * nchunks_dim[d] chunks along dimension d, whose offsets -- nchunks_dim[d]+1 ascending values from 0 to
 * shape[d], dimension after dimension in one table -- are passed twice: bounds_host for planning,
 * bounds_dev (the same values in device memory, alive until the kernels have run) for the kernels.
 */
int xbi_count_pairs_chunked_v(const void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* nchunks_dim,
                              const int64_t* bounds_host, const int64_t* bounds_dev, int axis, unsigned flags,
                              uint64_t mask_bits, unsigned long long* counts_dev, void* stream);
int xbi_bitround_chunked_v(void* x_dev, int dtype, int ndim, const int64_t* shape, const int64_t* nchunks_dim,
                           const int64_t* bounds_host, const int64_t* bounds_dev, const int32_t* keepbits_dev,
                           void* stream);

/*
 * K3: round to nearest (ties to even) keeping `keepbits` explicit mantissa bits, in
 * place.  Replaces the arithmetic of _np_bitround (xbitinfo/bitround.py:7-12).
 * dtype must be XBI_F16/F32/F64; keepbits == mantissa bits is the identity;
 * keepbits < 0 or > mantissa bits -> XBI_ERR_ARG (numcodecs raises ValueError).
 */
int xbi_bitround_inplace(void* x_dev, int dtype, int64_t numel, int keepbits, void* stream);
/*
 * The same into a second array: dst = bitround(src), src untouched -- _np_bitround's contract (it copies its
 * input first, xbitinfo/bitround.py:10) in one read and one write instead of a copy plus an in-place pass.
 * src_dev == dst_dev is the in-place call; otherwise the arrays must not overlap.
 */
int xbi_bitround(const void* src_dev, void* dst_dev, int dtype, int64_t numel, int keepbits, void* stream);

/*
 * K4: the lossless-compression pre-filters that follow bitround on the reference's save path
 * (xbitinfo/save_compressed.py:44-80: netCDF `shuffle=True`, i.e. the HDF5 shuffle filter;
 * :155-209: zarr Blosc with SHUFFLE / BITSHUFFLE), out of place on the device.
 *   kind XBI_SHUFFLE_BYTES: dst[j*numel + i] = src[i*itemsize + j]
 *   kind XBI_SHUFFLE_BITS:  bit (i%8) of dst[p*numel/8 + i/8] = bit p of element i, with
 *                           p = 8*byte + bit counted from the least significant bit of the
 *                           little-endian element; numel must be a multiple of 8
 *   inverse != 0 undoes the transform.  itemsize in {2, 4, 8}; src and dst are numel*itemsize
 *   bytes each and must not overlap.
 */
#define XBI_SHUFFLE_BYTES 0
#define XBI_SHUFFLE_BITS 1
int xbi_shuffle(const void* src_dev, void* dst_dev, int itemsize, int64_t numel, int kind, int inverse,
                void* stream);
/*
 * K3 + K4 fused: dst = shuffle(bitround(src, keepbits)) in ONE pass over HBM (read once, write
 * once) -- xr_bitround (xbitinfo/bitround.py:69-101) followed by the save path's shuffle
 * (xbitinfo/save_compressed.py:44-80, 155-209).  src is not modified.  dtype XBI_F16/F32/F64,
 * keepbits as for xbi_bitround_inplace, kind as for xbi_shuffle.
 */
int xbi_bitround_shuffle(const void* src_dev, void* dst_dev, int dtype, int64_t numel, int keepbits, int kind,
                         void* stream);

/*
 * Host-buffer entry points: what the numpy-facing plugin functions call.  Data is
 * streamed host -> device in slabs through a double-buffered staging area on `device`
 * (allocated on first use per host thread, released by xbi_host_release), counted /
 * rounded there, and only nbits doubles (or the rounded array) come back.
 *   xbi_bitinformation_host: counts_host (nbits*4, may be NULL) is overwritten, not
 *   accumulated; mi_host receives nbits doubles.
 *   xbi_bitround_host: src_host and dst_host may be the same buffer.
 */
int xbi_bitinformation_host(const void* x_host, int dtype, int64_t outer, int64_t n,
                            int64_t inner, unsigned flags, uint64_t mask_bits,
                            int set_zero_insignificant, double confidence, double* mi_host,
                            unsigned long long* counts_host, int device);
int xbi_bitround_host(const void* src_host, void* dst_host, int dtype, int64_t numel,
                      int keepbits, int device);
/*
 * Several analysed dimensions of ONE host array in a single pass over PCIe: what
 * get_bitinformation(ds, dim=None) asks for (xbitinfo/xbitinfo.py:373-406 loops the dimensions and
 * runs the backend once per dimension).  The C-contiguous array (`ndim` extents `shape`) is staged
 * in slabs of whole rows of its first dimension; every slab is counted along each of the `naxes`
 * axes in `axes` while it is on the device (the first dimension with a one-row halo).  mi_host
 * receives naxes*nbits doubles, counts_host (may be NULL) naxes*nbits*4 counters, in the order of
 * `axes`.  Results are identical to naxes calls of xbi_bitinformation_host.
 */
int xbi_bitinformation_host_multi(const void* x_host, int dtype, int ndim, const int64_t* shape, int naxes,
                                  const int* axes, unsigned flags, uint64_t mask_bits,
                                  int set_zero_insignificant, double confidence, double* mi_host,
                                  unsigned long long* counts_host, int device);
void xbi_host_release(void);

#ifdef __cplusplus
}
#endif
#endif /* XBITINFO_B200_H */
