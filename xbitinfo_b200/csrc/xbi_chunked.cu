// K1 and K3 batched over a chunk grid (regular, or irregular through per-dimension offset tables): the per-chunk workflow of the reference's
// docs/chunking.ipynb (cells 8-10: get_bitinformation -> get_keepbits -> xr_bitround on every
// dask block on its own) in a constant number of launches, whatever the number of chunks.
//
// The array stays as the caller has it (C-contiguous, N-d).  A work item is (chunk, group of
// 4096-element tiles of that chunk); a CTA walks its tiles in the chunk's own row-major order,
// 256 consecutive elements of that order per step, so a warp reads runs of the chunk's rows
// (coalesced up to the chunk's width).  Every thread keeps its position as a mixed-radix
// "odometer" over the chunk's extents and advances it by 256 elements per step with one
// compare/subtract per dimension that moves -- no divisions after the per-item set-up.
//
// Counting: a pair (x[c], x[c + e_axis]) exists when the element is not the last of ITS CHUNK
// along the analysed axis (xbitinfo/_py_bitinfo.py:155-160 applied to the block alone); the
// three word streams a, b, a&b of the valid pairs go into vertical counters (xbi_common.cuh) as
// in the generic kernel, and each item ends with one warp butterfly, a block reduction and one
// atomic per counter into the chunk's table.
#include <cstdlib>
#include <type_traits>

#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

constexpr int kThreads = 256;
constexpr int kPerThread = 16;
constexpr unsigned kTile = kThreads * kPerThread;
constexpr int kD = kMaxChunkDims;

// Position of one thread inside one chunk.  A thread's 16 elements of a tile are 4 "chains" x 4
// rounds: chain j sits at local index l0 + 256 j and every chain advances by 1024 per round, so the
// four odometers are independent instruction streams (one chain stepped 16 times in a row was a
// 16-deep dependent sequence per tile).  WIDE: 64-bit element offsets (arrays of 2^32 elements or
// more); otherwise offsets are 32-bit and wrap-around arithmetic is exact.
constexpr int kChains = 4;
template <bool WIDE>
struct Walker {
    using Off = typename std::conditional<WIDE, long long, unsigned>::type;
    unsigned c[kChains][kD];  // local coordinates (the outermost one may run past its extent: past the end)
    Off off[kChains];         // element offset of the chain's current element in the array
    unsigned e[kD];           // extents of this chunk (ragged at the upper array edges)
    unsigned inc[kD];         // kChains * kThreads as a mixed-radix number over e
    Off wrap[kD];             // offset correction when dimension d wraps and carries into d-1
    Off incoff;               // offset change of one step without wraps
    long long left;           // elements from chain 0's current one to the end of the chunk
    long long numel;

    static __device__ __forceinline__ void decompose(unsigned v, const unsigned (&e)[kD], unsigned (&out)[kD]) {
#pragma unroll
        for (int d = kD - 1; d >= 1; --d) { out[d] = v % e[d]; v /= e[d]; }
        out[0] = v;
    }
    // (c, off) += step, branch-free per dimension; `lo` = first dimension in use (uniform).  Returns the
    // new coordinate along `axis` (uniform; -1: nobody asks), picked up where that dimension is stepped.
    __device__ __forceinline__ unsigned advance(unsigned (&cc)[kD], Off& o, const unsigned (&step)[kD], Off stepoff, int lo,
                                                int axis) const {
        o += stepoff;
        unsigned carry = 0, ca = 0;
#pragma unroll
        for (int d = kD - 1; d >= 1; --d) {
            if (d >= lo) {
                const unsigned v = cc[d] + step[d] + carry;  // < 2 e[d]
                const bool wr = v >= e[d];
                cc[d] = wr ? v - e[d] : v;
                o += wr ? wrap[d] : (Off)0;
                carry = wr ? 1u : 0u;
                if (d == axis) ca = cc[d];
            }
        }
        cc[0] += step[0] + carry;  // (a carry out of dimension `lo` only happens past the end of the chunk)
        if (axis == 0) ca = cc[0];
        return ca;
    }

    __device__ __forceinline__ void init(const ChunkGrid& g, unsigned cid, unsigned long long l0) {
        long long base = 0;
        numel = 1;
        unsigned r = cid;
#pragma unroll
        for (int d = kD - 1; d >= 0; --d) {
            unsigned q = r;
            if (d > 0) { q = r % g.nchunk[d]; r /= g.nchunk[d]; }
            long long start;
            if (g.bounds[d] != nullptr) {  // irregular dimension (uniform branch)
                const long long b0 = g.bounds[d][q], b1 = g.bounds[d][q + 1];
                start = b0 * g.scale[d];
                e[d] = (unsigned)((b1 - b0) * g.scale[d]);
            } else {
                start = (long long)q * g.chunk[d];
                const long long ext = g.shape[d] - start;
                e[d] = ext < (long long)g.chunk[d] ? (unsigned)ext : g.chunk[d];
            }
            base += start * g.stride[d];
            numel *= e[d];
        }
        if (numel < (1ll << 31)) {
            decompose((unsigned)l0, e, c[0]);
        } else {
            unsigned long long l = l0;
#pragma unroll
            for (int d = kD - 1; d >= 1; --d) { c[0][d] = (unsigned)(l % e[d]); l /= e[d]; }
            c[0][0] = (unsigned)l;
        }
        long long o0 = base;
#pragma unroll
        for (int d = 0; d < kD; ++d) {
            o0 += (long long)c[0][d] * g.stride[d];
            wrap[d] = d > 0 ? (Off)(g.stride[d - 1] - (long long)e[d] * g.stride[d]) : (Off)0;
        }
        off[0] = (Off)o0;
        // chains 1.. : one step of kThreads each
        unsigned s1[kD];
        decompose(kThreads, e, s1);
        long long s1off = 0;
#pragma unroll
        for (int d = 0; d < kD; ++d) s1off += (long long)s1[d] * g.stride[d];
#pragma unroll
        for (int j = 1; j < kChains; ++j) {
#pragma unroll
            for (int d = 0; d < kD; ++d) c[j][d] = c[j - 1][d];
            off[j] = off[j - 1];
            advance(c[j], off[j], s1, (Off)s1off, g.lo, -1);
        }
        decompose(kChains * kThreads, e, inc);
        long long io = 0;
#pragma unroll
        for (int d = 0; d < kD; ++d) io += (long long)inc[d] * g.stride[d];
        incoff = (Off)io;
        left = numel - (long long)l0;
        // The per-chunk constants are uniform, and left alone ptxas re-derives them from the kernel
        // parameters in every step (7 uniform-pipe instructions per dimension and step); make them opaque
        // so that they are computed once per item and stay in registers.
#pragma unroll
        for (int d = 0; d < kD; ++d) {
            asm volatile("" : "+r"(e[d]), "+r"(inc[d]));
            pin(wrap[d]);
        }
        pin(incoff);
    }
    static __device__ __forceinline__ void pin(unsigned& v) { asm volatile("" : "+r"(v)); }
    static __device__ __forceinline__ void pin(long long& v) { asm volatile("" : "+l"(v)); }

    // chain j to its next round
    __device__ __forceinline__ unsigned step(int j, int lo, int axis) { return advance(c[j], off[j], inc, incoff, lo, axis); }
    // coordinate along a dimension known only at run time (uniform), without indexing the register arrays dynamically
    __device__ __forceinline__ unsigned coord(int j, int axis) const {
        unsigned v = 0;
#pragma unroll
        for (int d = 0; d < kD; ++d) if (d == axis) v = c[j][d];
        return v;
    }
    __device__ __forceinline__ bool inside(int j) const { return left > (long long)j * kThreads; }
    __device__ __forceinline__ void next_round() { left -= kChains * kThreads; }

    // extent along a dimension known only at run time (uniform), without indexing the register array dynamically
    __device__ __forceinline__ unsigned extent(int axis) const {
        unsigned v = 1;
#pragma unroll
        for (int d = 0; d < kD; ++d) if (d == axis) v = e[d];
        return v;
    }
};

// counts: [nchunks][nbits][4] accumulated RAW here (slot 0 of bit 0: pairs; slots 1, 2, 3 of bit k:
// B, A, AB plane sums), turned into the 2x2 tables by k_chunk_finalize.
template <int KIND, int MASK, bool XFORM, bool WIDE>
__global__ void __launch_bounds__(kThreads, Elem<KIND>::kWords == 1 ? 2 : 1)
k_pairs_chunked(const void* __restrict__ x, const __grid_constant__ ChunkGrid g, uint32_t mask_lo, uint32_t mask_hi,
                unsigned long long* __restrict__ counts) {
    using E = Elem<KIND>;
    constexpr int NW = E::kWords;
    constexpr int NBITS = E::kBits;
    using L = Ladder<3, 6>;
    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_n;
    const uint32_t lane = threadIdx.x & 31u;
    using Off = typename Walker<WIDE>::Off;
    Off pstride = 0;
#pragma unroll
    for (int d = 0; d < kD; ++d) if (d == g.axis) pstride = (Off)g.stride[d];

    for (unsigned item = blockIdx.x; item < g.items; item += gridDim.x) {
        const unsigned cid = item / g.groups, grp = item - cid * g.groups;
        const unsigned long long group_base = (unsigned long long)grp * g.tiles_per_group * kTile;
        Walker<WIDE> w;
        w.init(g, cid, group_base + threadIdx.x);
        long long tiles_left = (w.numel - (long long)group_base + (long long)kTile - 1) / (long long)kTile;  // uniform
        if (tiles_left <= 0) continue;  // an edge chunk with fewer tiles than the full ones
        if (tiles_left > (long long)g.tiles_per_group) tiles_left = g.tiles_per_group;
        const unsigned ea = w.extent(g.axis);
        unsigned ca[kChains];  // coordinate of every chain along the analysed axis
#pragma unroll
        for (int j = 0; j < kChains; ++j) ca[j] = w.coord(j, g.axis);

        L A[NW], B[NW], AB[NW];
        unsigned long long accA[NW], accB[NW], accAB[NW];
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            A[k].clear(); B[k].clear(); AB[k].clear();
            accA[k] = 0; accB[k] = 0; accAB[k] = 0;
        }
        uint32_t nvalid = 0, it = 0;
        unsigned long long nvalid_total = 0;
        auto flush = [&]() {
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                accA[k] += A[k].warp_totals_bounded(lane, it);
                accB[k] += B[k].warp_totals_bounded(lane, it);
                accAB[k] += AB[k].warp_totals_bounded(lane, it);
                A[k].clear(); B[k].clear(); AB[k].clear();
            }
            nvalid_total += nvalid;
            nvalid = 0;
            it = 0;
        };

        for (long long t = 0; t < tiles_left; ++t) {
            uint32_t wa[NW][kPerThread], wb[NW][kPerThread];
            // phase 1: where this thread's 16 pairs are (pure index arithmetic); pairs that do not exist
            // (past the end of the chunk, last element along the axis) read element 0 and its partner instead
            Off offs[kPerThread];
            uint32_t live = 0;
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) {
                const int j = u % kChains;  // element u = chain j in round u / kChains
                const bool pair = w.inside(j) && ca[j] + 1u < ea;
                offs[u] = pair ? w.off[j] : (Off)0;
                live |= (pair ? 1u : 0u) << u;
                ca[j] = w.step(j, g.lo, g.axis);
                if (j == kChains - 1) w.next_round();
            }
            // phase 2: all 32 loads in flight
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) {
                uint32_t a[NW], b[NW];
                E::load(x, (long long)offs[u], a);
                E::load(x, (long long)(Off)(offs[u] + pstride), b);
#pragma unroll
                for (int k = 0; k < NW; ++k) { wa[k][u] = a[k]; wb[k][u] = b[k]; }
            }
            // phase 3: mask test, transform, dead pairs zeroed
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) {
                uint32_t a[NW], b[NW];
#pragma unroll
                for (int k = 0; k < NW; ++k) { a[k] = wa[k][u]; b[k] = wb[k][u]; }
                bool ok = (live >> u) & 1u;
                if (MASK == kMaskNaN) ok = ok && !(E::is_nan(a) || E::is_nan(b));
                if (MASK == kMaskValue) ok = ok && !(E::equals(a, mask_lo, mask_hi) || E::equals(b, mask_lo, mask_hi));
                if (XFORM) { E::transform(a); E::transform(b); }
#pragma unroll
                for (int k = 0; k < NW; ++k) { wa[k][u] = ok ? a[k] : 0u; wb[k][u] = ok ? b[k] : 0u; }
                nvalid += ok ? 1u : 0u;
            }
#pragma unroll
            for (int k = 0; k < NW; ++k) {
                uint32_t wab[kPerThread];
#pragma unroll
                for (int u = 0; u < kPerThread; ++u) wab[u] = wa[k][u] & wb[k][u];
                A[k].add16(wa[k], it);
                B[k].add16(wb[k], it);
                AB[k].add16(wab, it);
            }
            ++it;
            if (it == L::kFlushChunks) flush();
        }
        flush();

        // block reduction, then one atomic per non-zero counter into the chunk's table
        for (int i = threadIdx.x; i < 3 * 64; i += kThreads) (&s_sum[0][0])[i] = 0;
        if (threadIdx.x == 0) s_n = 0;
        __syncthreads();
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int j = k * 32 + lane;
            if (accA[k]) atomicAdd(&s_sum[S_A][j], accA[k]);
            if (accB[k]) atomicAdd(&s_sum[S_B][j], accB[k]);
            if (accAB[k]) atomicAdd(&s_sum[S_AB][j], accAB[k]);
        }
        if (nvalid_total) atomicAdd(&s_n, nvalid_total);
        __syncthreads();
        unsigned long long* dst = counts + (size_t)cid * NBITS * 4;
        for (int i = threadIdx.x; i < 3 * NBITS; i += kThreads) {
            const int s = i / NBITS, j = i - s * NBITS;  // stream, bit from the least significant one
            const unsigned long long v = s_sum[s][j];
            const int slot = (s == S_A) ? 2 : (s == S_B) ? 1 : 3;
            if (v) atomicAdd(dst + 4 * (NBITS - 1 - j) + slot, v);
        }
        if (threadIdx.x == 0 && s_n) atomicAdd(dst, s_n);
        __syncthreads();  // s_sum is zeroed again by the next item
    }
}

// raw sums -> [[N-A-B+AB, B-AB], [A-AB, AB]] in place (bitpaircount layout, _py_bitinfo.py:128-140)
__global__ void __launch_bounds__(kThreads)
k_chunk_finalize(unsigned long long* __restrict__ counts, int nbits, unsigned nchunks) {
    const unsigned per = kThreads / nbits;
    const unsigned ci = blockIdx.x * per + threadIdx.x / nbits;
    const unsigned k = threadIdx.x % nbits;
    const bool ok = ci < nchunks;
    unsigned long long N = 0, A = 0, B = 0, AB = 0;
    unsigned long long* c = counts + (size_t)ci * nbits * 4;
    if (ok) {
        N = c[0];
        B = c[4 * k + 1];
        A = c[4 * k + 2];
        AB = c[4 * k + 3];
    }
    __syncthreads();  // every N of the block's chunks has been read before slot 0 of bit 0 is rewritten
    if (ok) {
        c[4 * k + 0] = N - A - B + AB;
        c[4 * k + 1] = B - AB;
        c[4 * k + 2] = A - AB;
    }
}

// K3 per chunk: keepbits[chunk] mantissa bits kept (clamped to 0..mant; mant = identity)
template <int BYTES, bool WIDE>
__global__ void __launch_bounds__(kThreads, 2)
k_bitround_chunked(void* __restrict__ x, const __grid_constant__ ChunkGrid g, const int* __restrict__ keepbits, int mant) {
    using T = typename std::conditional<BYTES == 2, uint16_t, typename std::conditional<BYTES == 4, uint32_t, unsigned long long>::type>::type;
    T* __restrict__ p = static_cast<T*>(x);
    for (unsigned item = blockIdx.x; item < g.items; item += gridDim.x) {
        const unsigned cid = item / g.groups, grp = item - cid * g.groups;
        int kb = keepbits[cid];
        kb = kb < 0 ? 0 : kb > mant ? mant : kb;
        const uint32_t drop = (uint32_t)(mant - kb);
        if (drop == 0) continue;
        const unsigned long long full = BYTES == 8 ? ~0ull : ((1ull << (8 * BYTES)) - 1ull);
        const unsigned long long keep = (full >> drop) << drop;
        const unsigned long long half_m1 = (1ull << (drop - 1)) - 1ull;
        const unsigned long long group_base = (unsigned long long)grp * g.tiles_per_group * kTile;
        Walker<WIDE> w;
        w.init(g, cid, group_base + threadIdx.x);
        long long tiles_left = (w.numel - (long long)group_base + (long long)kTile - 1) / (long long)kTile;
        if (tiles_left > (long long)g.tiles_per_group) tiles_left = g.tiles_per_group;
        for (long long t = 0; t < tiles_left; ++t) {
            typename Walker<WIDE>::Off offs[kPerThread];
            uint32_t live = 0;
            T v[kPerThread];
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) {
                const int j = u % kChains;
                offs[u] = w.off[j];
                live |= (w.inside(j) ? 1u : 0u) << u;
                w.step(j, g.lo, -1);
                if (j == kChains - 1) w.next_round();
            }
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) if ((live >> u) & 1u) v[u] = p[offs[u]];
#pragma unroll
            for (int u = 0; u < kPerThread; ++u) {
                if ((live >> u) & 1u) {
                    const unsigned long long b = v[u];
                    p[offs[u]] = (T)((b + ((b >> drop) & 1ull) + half_m1) & keep);
                }
            }
        }
    }
}

// 32-bit element offsets are exact when the whole array has fewer than 2^32 elements
bool needs_wide_offsets(const ChunkGrid& g) {
    double numel = 1.0;
    for (int d = 0; d < kD; ++d) numel *= (double)g.shape[d];
    return numel >= 4294967296.0;
}

unsigned grid_for(const ChunkGrid& g, int per_sm) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned cap = (unsigned)(sms * per_sm);
    return g.items < cap ? g.items : cap;
}

template <int KIND>
cudaError_t launch_count_kind(const void* x, const ChunkGrid& g, bool transform, int mask_mode, uint64_t mask_bits,
                              unsigned long long* counts, cudaStream_t stream) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    const unsigned grid = grid_for(g, 4);
    const uint32_t mlo = (uint32_t)(mask_bits & 0xFFFFFFFFu), mhi = (uint32_t)(mask_bits >> 32);
    const bool wide = needs_wide_offsets(g);
#define XBI_LAUNCH_CHUNKED(MASK, XF)                                                                   \
    do {                                                                                               \
        if (wide) k_pairs_chunked<KIND, MASK, XF, true><<<grid, kThreads, 0, stream>>>(x, g, mlo, mhi, counts);  \
        else k_pairs_chunked<KIND, MASK, XF, false><<<grid, kThreads, 0, stream>>>(x, g, mlo, mhi, counts);      \
    } while (0)
    if (transform) {
        if constexpr (isf) {
            switch (mask_mode) {
                case kMaskNone: XBI_LAUNCH_CHUNKED(kMaskNone, true); break;
                case kMaskNaN: XBI_LAUNCH_CHUNKED(kMaskNaN, true); break;
                default: XBI_LAUNCH_CHUNKED(kMaskValue, true); break;
            }
        } else {
            return cudaErrorInvalidValue;
        }
    } else {
        switch (mask_mode) {
            case kMaskNone: XBI_LAUNCH_CHUNKED(kMaskNone, false); break;
            case kMaskNaN:
                if constexpr (isf) { XBI_LAUNCH_CHUNKED(kMaskNaN, false); break; }
                return cudaErrorInvalidValue;
            default: XBI_LAUNCH_CHUNKED(kMaskValue, false); break;
        }
    }
#undef XBI_LAUNCH_CHUNKED
    note_launch();
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_pairs_chunked(const void* x, int dtype, const ChunkGrid& g, unsigned nchunks, bool transform,
                                 int mask_mode, uint64_t mask_bits, unsigned long long* counts, cudaStream_t stream,
                                 unsigned flags) {
    cudaError_t e = cudaSuccess;
    // the walkers of xbi_chunked_walk.cu (one load per element, warp-level indexing) where the grid gives them
    // enough columns / runs; the per-element odometers below otherwise (tiny chunks, very short rows)
    bool walked = false;
    static const bool allow_walk = [] {
        const char* env = std::getenv("XBI_CHUNKED_WALK");
        return !(env && env[0] == '0');
    }();
    if ((allow_walk || (flags & XBI_FORCE_WALK)) && !(flags & XBI_FORCE_GENERIC)) {
        e = launch_pairs_chunked_walk(x, dtype, g, nchunks, transform, mask_mode, mask_bits, counts, stream,
                                      (flags & XBI_FORCE_WALK) != 0,
                                      (flags & XBI_FORCE_DENSE) ? 2 : (flags & XBI_FORCE_SPARSE) ? 1 : 0, &walked);
        if (e != cudaSuccess) return e;
    }
    if (!walked) switch (dtype) {
        case XBI_U8: e = launch_count_kind<kU8>(x, g, transform, mask_mode, mask_bits, counts, stream); break;
        case XBI_U16: e = launch_count_kind<kU16>(x, g, transform, mask_mode, mask_bits, counts, stream); break;
        case XBI_U32: e = launch_count_kind<kU32>(x, g, transform, mask_mode, mask_bits, counts, stream); break;
        case XBI_U64: e = launch_count_kind<kU64>(x, g, transform, mask_mode, mask_bits, counts, stream); break;
        case XBI_F16: e = launch_count_kind<kF16>(x, g, transform, mask_mode, mask_bits, counts, stream); break;
        case XBI_F32: e = launch_count_kind<kF32>(x, g, transform, mask_mode, mask_bits, counts, stream); break;
        case XBI_F64: e = launch_count_kind<kF64>(x, g, transform, mask_mode, mask_bits, counts, stream); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    const int nbits = 8 * dtype_bytes(dtype);
    const unsigned per = kThreads / nbits;
    k_chunk_finalize<<<(nchunks + per - 1) / per, kThreads, 0, stream>>>(counts, nbits, nchunks);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_bitround_chunked(void* x, int dtype, const ChunkGrid& g, const int* keepbits, cudaStream_t stream) {
    const unsigned grid = grid_for(g, 8);
    const bool wide = needs_wide_offsets(g);
    switch (dtype) {
#define XBI_LAUNCH_ROUND(B, M)                                                                        \
    do {                                                                                              \
        if (wide) k_bitround_chunked<B, true><<<grid, kThreads, 0, stream>>>(x, g, keepbits, M);      \
        else k_bitround_chunked<B, false><<<grid, kThreads, 0, stream>>>(x, g, keepbits, M);          \
    } while (0)
        case XBI_F16: XBI_LAUNCH_ROUND(2, 10); break;
        case XBI_F32: XBI_LAUNCH_ROUND(4, 23); break;
        case XBI_F64: XBI_LAUNCH_ROUND(8, 52); break;
#undef XBI_LAUNCH_ROUND
        default: return cudaErrorInvalidValue;
    }
    note_launch();
    return cudaGetLastError();
}

}  // namespace xbi
