// Host-side declarations shared by the translation units of libxbitinfo_b200.so.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cuda_runtime.h>

#include "../../include/xbitinfo_b200.h"

namespace xbi {

struct Workspace;
struct RawSums;

// element kinds as template parameters (== enum xbi_dtype)
constexpr int kU8 = XBI_U8, kU16 = XBI_U16, kU32 = XBI_U32, kU64 = XBI_U64;
constexpr int kF16 = XBI_F16, kF32 = XBI_F32, kF64 = XBI_F64;

// how masked elements are recognised
constexpr int kMaskNone = 0, kMaskNaN = 1, kMaskValue = 2;
// The fast kernels exist in two masked flavours.  "Sparse" (kMaskNaN / kMaskValue): two streams
// over zeroed elements + corrections at validity boundaries -- near the unmasked speed when
// masked elements come in runs (land masks) or not at all.  "Dense" (the values below): three
// explicit streams with keep-masks, flat 1.8x the unmasked cost whatever the data -- better when
// masked elements are scattered.  Both are launched; each reads the sampled boundary density from
// the workspace (k_mask_density) and returns at once unless it is the one for this data.
constexpr int kMaskNaNDense = 3, kMaskValueDense = 4;
template <int M> struct MaskMode {
    static constexpr bool kDense = (M >= kMaskNaNDense);
    static constexpr int kTest = kDense ? M - 2 : M;  // which test: kMaskNone / kMaskNaN / kMaskValue
};

inline int dtype_bytes(int dtype) {
    switch (dtype) {
        case XBI_U8: return 1;
        case XBI_U16: case XBI_F16: return 2;
        case XBI_U32: case XBI_F32: return 4;
        case XBI_U64: case XBI_F64: return 8;
        default: return 0;
    }
}
inline bool dtype_is_float(int dtype) { return dtype >= XBI_F16 && dtype <= XBI_F64; }

// Description of one counting job on a C-contiguous (outer, n, inner) array.
struct CountJob {
    const void* x;       // device pointer to element 0
    int dtype;           // enum xbi_dtype
    bool transform;      // signed-exponent transform (floats)
    int mask_mode;       // kMaskNone / kMaskNaN / kMaskValue
    uint64_t mask_bits;  // for kMaskValue
    int64_t outer, n, inner;
};

void note_launch();       // bumps the process-wide launch counter
void note_fast_launch();  // ... and the fast-path counter

// Flat pairs (f, f+inner) for f in [lo, hi) accumulated into `dst` -- any alignment.
cudaError_t launch_pairs_generic_flat(const CountJob& job, int64_t lo, int64_t hi, RawSums* dst,
                                      cudaStream_t stream);
// The flat pairs that straddle two consecutive blocks of the analysed axis:
// f = (o*n + n-1)*inner + i for o in [0, outer-1), i in [0, inner).
// restricted to edge indices t = o*inner + i in [t_begin, t_end)
cudaError_t launch_pairs_generic_edges(const CountJob& job, int64_t t_begin, int64_t t_end, RawSums* dst,
                                       cudaStream_t stream);

// Bit planes of the single elements f in [lo, lo+count) added to dst->sums[S_A] (transform
// applied, masked elements skipped): the a-stream correction of the column walkers.
// With `dense_probe` the pass returns at once when the sampled mask density (dense_probe->pad) says
// the dense flavour of the column walkers ran (which needs no a-stream correction).
cudaError_t launch_elem_planes(const CountJob& job, int64_t lo, int64_t count, RawSums* dst, cudaStream_t stream,
                               const RawSums* dense_probe = nullptr, unsigned long long dense_one_in = 0);

// TMA fast paths.  Each consumes as much of [0, numel-inner) as its alignment rules
// allow, reports the flat range it covered in [*done_lo, *done_hi) (empty if none), and
// leaves the rest to the generic kernel.
// *edges_done: the kernel also removed the row-straddling pairs inside [*done_lo, *done_hi).
// *leftovers: what else the kernel counted itself (its producer warps, see RowsFringe in xbi_count_tma.cu):
//   flat  -- the flat pairs outside [*done_lo, *done_hi) as well, i.e. every flat pair of the job;
//   edges -- every row-straddling pair that *edges_done does not cover, i.e. all of them.
struct RowsLeftovers {
    bool flat = false, edges = false;
};
cudaError_t launch_pairs_tma(const CountJob& job, int64_t npairs_flat, Workspace* ws, cudaStream_t stream,
                             int64_t* done_lo, int64_t* done_hi, bool* edges_done, RowsLeftovers* leftovers);

// Column walkers for inner > 1 (16-byte aligned rows); covers flat pairs [*done_lo, *done_hi).
// *edges_done: the block-boundary pairs have been subtracted as well (all of them).
cudaError_t launch_pairs_cols(const CountJob& job, Workspace* ws, cudaStream_t stream, int64_t* done_lo,
                              int64_t* done_hi, bool* edges_done);

// The follow-up passes of the walkers alone (row planes, block-boundary pairs) for an unmasked job of a 4-byte type
// with 16-byte aligned rows.
cudaError_t launch_cols_followups(const CountJob& job, Workspace* ws, cudaStream_t stream);

// Two innermost axes in one read (xbi_count_xy.cu).  jy = the job of the second-to-last axis.
bool pairs_xy_eligible(const CountJob& jy);
cudaError_t launch_pairs_xy(const CountJob& jy, Workspace* ws_y, Workspace* ws_x, unsigned int* dirty_flag,
                            cudaStream_t stream);

// Samples ~2^18 flat pairs and leaves (#pairs whose two elements differ in validity, #pairs sampled)
// in ws->plus.pad[0], pad[1] for the masked fast kernels to choose their flavour.
cudaError_t launch_mask_density(const CountJob& job, int64_t npairs_flat, Workspace* ws, cudaStream_t stream);
// ... the two numbers added to any two (zeroed) device words
cudaError_t launch_mask_density_to(const CountJob& job, int64_t npairs_flat, unsigned long long* out_bnd,
                                   unsigned long long* out_pairs, cudaStream_t stream);


// ---------------------------------------------------------------------------------
// finish kernel (xbi_finish.cu): fringe pairs + combine (+ all-reduce over peer memory) (+ K2), one launch
// ---------------------------------------------------------------------------------
// Exchange buffer of the peer all-reduce, one per rank, in 64-bit words:
//   slots [2 parities][kPeerMaxWorld source ranks][kPeerCells]   the (nbits,2,2) table of every rank
//   flags [2 parities][kPeerMaxWorld source ranks]                epoch of the table in that slot
//   error word                                                    epoch of a wait that timed out (0: none)
// Epochs start at 1 and grow by one per call on every rank; consecutive epochs use different parities, and a
// rank can only reach epoch e+2 after every peer has finished reading epoch e (it needs their flag of e+1,
// raised by a kernel that is stream-ordered after the one that read e), so the slots are never overwritten
// while they are being read.
constexpr int kPeerMaxWorld = 8;
constexpr int kPeerCells = 256;
constexpr size_t kPeerErrorWord = (size_t)2 * kPeerMaxWorld * kPeerCells + 2 * kPeerMaxWorld;
constexpr size_t kPeerWords = kPeerErrorWord + 8;
struct PeerExchange {
    int world;  // <= 1: no exchange
    int rank;
    unsigned long long epoch;
    unsigned long long timeout_ns;
    unsigned long long* peer[kPeerMaxWorld];  // exchange buffer of every rank as mapped into this process (own included)
};
struct FinishJob {
    const void* x;
    long long n, inner;
    // fringe: [0], [1] ranges of flat pairs (f, f+inner), f in [lo, lo+cnt) -> added;
    //         [2], [3] ranges of edge indices t = o*inner + i (pair f = (o*n + n-1)*inner + i) -> subtracted
    long long range_lo[4], range_cnt[4];
    uint32_t mask_lo, mask_hi;
    const Workspace* ws;
    unsigned long long* counts;  // nbits*4
    int accumulate;              // counts += table (atomics) instead of counts = table
    double* mi;                  // nbits doubles, or nullptr: no K2
    int set_zero;
    double z;
    PeerExchange px;
};
// up to this many fringe pairs go through the finish kernel; larger leftovers keep their own generic launches
constexpr long long kFinishMaxFlat = 16384, kFinishMaxEdges = 4096;
cudaError_t launch_finish(const CountJob& job, const FinishJob& J, cudaStream_t stream);
// K2 on `nbatch` consecutive (nbits,2,2) tables -> mi[nbatch][nbits]
cudaError_t launch_mutual_information(const unsigned long long* counts, int nbits, long long nbatch, int set_zero,
                                      double z_quantile, double* mi, cudaStream_t stream);
cudaError_t launch_bitround(void* x, int dtype, int64_t numel, int keepbits, cudaStream_t stream);
// dst = bitround(src); src == dst: in place, otherwise the arrays must not overlap
cudaError_t launch_bitround_to(const void* src, void* dst, int dtype, int64_t numel, int keepbits, cudaStream_t stream);
// K4: byte (bits == false) or bit (bits == true) shuffle of numel elements of itemsize bytes,
// forward or inverse; src and dst must not overlap.  drop_bits > 0 (forward only): the elements are
// bitrounded (K3's arithmetic, that many trailing mantissa bits dropped) on their way through.
cudaError_t launch_shuffle(const void* src, void* dst, int itemsize, int64_t numel, bool bits, bool forward,
                           int drop_bits, cudaStream_t stream);

// ---------------------------------------------------------------------------------
// batched per-chunk kernels (xbi_chunked.cu)
// ---------------------------------------------------------------------------------
constexpr int kMaxChunkDims = 5;  // after merging unchunked neighbours (xbi_api.cu: plan_chunk_grid)
// A regular chunk grid over a C-contiguous array.  Unused LEADING dimensions hold extent 1.
struct ChunkGrid {
    long long shape[kMaxChunkDims];   // array extents
    long long stride[kMaxChunkDims];  // element strides
    unsigned chunk[kMaxChunkDims];    // chunk extents (< 2^31; the last chunk of a dimension may be shorter)
    unsigned nchunk[kMaxChunkDims];   // chunks per dimension
    // irregular dimensions: device table of nchunk+1 ascending offsets (nullptr: regular, `chunk` applies);
    // chunk q covers [bounds[q], bounds[q+1]) * scale (scale: the unchunked inner dimensions merged into it);
    // `chunk` then holds the largest extent
    const long long* bounds[kMaxChunkDims];
    long long scale[kMaxChunkDims];
    int axis;                         // analysed dimension, -1: none (bitround)
    int lo;                           // first dimension in use (the ones before it are padding)
    unsigned groups;                  // work items per chunk
    unsigned tiles_per_group;         // 4096-element tiles per work item
    unsigned items;                   // nchunks * groups
    int vec16;                        // (walkers) every row of every chunk is made of whole, aligned 16-byte vectors
};
// counts[nchunks][nbits][4] must be zero on entry and holds the 2x2 tables on completion
cudaError_t launch_pairs_chunked(const void* x, int dtype, const ChunkGrid& g, unsigned nchunks, bool transform,
                                 int mask_mode, uint64_t mask_bits, unsigned long long* counts, cudaStream_t stream,
                                 unsigned flags = 0);
// the fast formulation (xbi_chunked_walk.cu); *taken = false: the grid does not suit it, nothing was launched.
// flavour (4- and 8-byte types): 0 -- masked modes sample the density of validity boundaries and run the walkers on
// two streams + boundary corrections or, for scattered masks, on three explicit streams under keep-masks;
// 1 / 2 -- pinned to the two-stream / three-stream walkers (tests)
cudaError_t launch_pairs_chunked_walk(const void* x, int dtype, const ChunkGrid& g, unsigned nchunks, bool transform,
                                      int mask_mode, uint64_t mask_bits, unsigned long long* counts, cudaStream_t stream,
                                      bool force, int flavour, bool* taken);
cudaError_t launch_bitround_chunked(void* x, int dtype, const ChunkGrid& g, const int* keepbits, cudaStream_t stream);

// xbi_hostpool.cu: rows of row_bytes bytes copied host -> host by a small pool of threads
// (XBI_HOST_THREADS, default min(8, cores)); blocks until done
void host_copy_parallel(void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t row_bytes, size_t rows);
int host_copy_threads();

double normal_quantile(double p);  // AS241, host

}  // namespace xbi
