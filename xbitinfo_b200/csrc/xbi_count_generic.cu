// K1, generic variant: bit-pair counting for any shape and any alignment.
//
// Every thread loads the two elements of a flat pair (f, f+inner) with scalar loads
// (coalesced across the warp), applies the signed-exponent transform
// (xbitinfo/_py_bitinfo.py:42-94) and the mask test, and folds the three word streams
// a, b, a&b into vertical counters (xbi_common.cuh).  It serves
//   * the ranges the TMA fast paths cannot take (unaligned bases, ragged tails),
//   * the "edge" pairs between consecutive blocks of the analysed axis, which the
//     flat formulation counts although the reference (_py_bitinfo.py:155-160) does not.
#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

constexpr int kThreads = 256;
constexpr int kPairsPerTile = 16 * kThreads;

// MODE 0: flat pairs (lo+t, lo+t+inner); 1: edge pairs; 2: single elements lo+t (a-stream only)
template <int KIND, int MASK, bool XFORM, int MODE>
__global__ void __launch_bounds__(kThreads)
k_pairs_generic(const void* __restrict__ x, long long lo, long long count, long long n, long long inner,
                uint32_t mask_lo, uint32_t mask_hi, RawSums* __restrict__ dst, const RawSums* dense_probe,
                unsigned long long dense_one_in) {
    using E = Elem<KIND>;
    if (MODE == 2 && dense_probe != nullptr && sampled_mask_is_dense(dense_probe, dense_one_in)) return;
    constexpr int NW = E::kWords;
    using L = Ladder<3, 6>;
    L A[NW], B[NW], AB[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) { A[w].clear(); B[w].clear(); AB[w].clear(); }
    unsigned long long accA[NW], accB[NW], accAB[NW];
#pragma unroll
    for (int w = 0; w < NW; ++w) { accA[w] = 0; accB[w] = 0; accAB[w] = 0; }
    uint32_t nvalid = 0;
    unsigned long long nvalid_total = 0;
    const uint32_t lane = threadIdx.x & 31u;
    uint32_t it = 0;

    auto flush = [&]() {
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            accA[w] += A[w].warp_totals(lane, it);
            accB[w] += B[w].warp_totals(lane, it);
            accAB[w] += AB[w].warp_totals(lane, it);
            A[w].clear(); B[w].clear(); AB[w].clear();
        }
        nvalid_total += nvalid;
        nvalid = 0;
        it = 0;
    };

    const long long tiles = (count + kPairsPerTile - 1) / kPairsPerTile;
    for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
        uint32_t wa[NW][16], wb[NW][16];
        // edge mode: (block o, column i) of this thread's first edge, then advanced without divisions
        long long eo = 0, ei = 0;
        if (MODE == 1 && inner != 1) {
            const long long te0 = lo + tile * kPairsPerTile + threadIdx.x;
            eo = te0 / inner;
            ei = te0 - eo * inner;
        }
#pragma unroll
        for (int u = 0; u < 16; ++u) {
            const long long t = tile * kPairsPerTile + (long long)u * kThreads + threadIdx.x;
            uint32_t a[NW], b[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) { a[w] = 0; b[w] = 0; }
            if (t < count) {
                long long f;
                if (MODE == 1) {
                    if (inner == 1) {
                        f = (lo + t) * n + (n - 1);
                    } else {
                        f = (eo * n + (n - 1)) * inner + ei;
                        ei += kThreads;
                        while (ei >= inner) { ei -= inner; ++eo; }
                    }
                } else {
                    f = lo + t;
                }
                E::load(x, f, a);
                if (MODE != 2) E::load(x, f + inner, b);
                bool ok = true;
                if (MASK == kMaskNaN) ok = !(E::is_nan(a) || (MODE != 2 && E::is_nan(b)));
                if (MASK == kMaskValue)
                    ok = !(E::equals(a, mask_lo, mask_hi) || (MODE != 2 && E::equals(b, mask_lo, mask_hi)));
                if (XFORM) { E::transform(a); if (MODE != 2) E::transform(b); }
                if (!ok) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) { a[w] = 0; b[w] = 0; }
                }
                nvalid += (ok && MODE != 2) ? 1u : 0u;
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) { wa[w][u] = a[w]; wb[w][u] = b[w]; }
        }
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            uint32_t wab[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) wab[u] = wa[w][u] & wb[w][u];
            A[w].add16(wa[w], it);
            B[w].add16(wb[w], it);
            AB[w].add16(wab, it);
        }
        ++it;
        if (it == L::kFlushChunks) flush();
    }
    flush();

    // block reduction, then one atomic per non-zero counter and block
    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_n;
    for (int i = threadIdx.x; i < 3 * 64; i += kThreads) (&s_sum[0][0])[i] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
#pragma unroll
    for (int w = 0; w < NW; ++w) {
        const int j = w * 32 + lane;
        if (accA[w]) atomicAdd(&s_sum[S_A][j], accA[w]);
        if (accB[w]) atomicAdd(&s_sum[S_B][j], accB[w]);
        if (accAB[w]) atomicAdd(&s_sum[S_AB][j], accAB[w]);
    }
    if (nvalid_total) atomicAdd(&s_n, nvalid_total);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 64; i += kThreads) {
        const unsigned long long v = (&s_sum[0][0])[i];
        if (v) atomicAdd(&dst->sums[0][0] + i, v);
    }
    if (threadIdx.x == 0 && s_n) atomicAdd(&dst->npairs, s_n);
}

// validity boundaries among `nsamples` flat pairs (f, f + inner), f = i * stride
template <int KIND, int MASK>
__global__ void __launch_bounds__(kThreads)
k_mask_density(const void* __restrict__ x, long long nsamples, long long stride, long long inner, uint32_t mask_lo,
               uint32_t mask_hi, unsigned long long* __restrict__ out_bnd, unsigned long long* __restrict__ out_pairs) {
    using E = Elem<KIND>;
    unsigned bnd = 0, pairs = 0;
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < nsamples; i += (long long)gridDim.x * kThreads) {
        const long long f = i * stride;
        uint32_t a[E::kWords], b[E::kWords];
        E::load(x, f, a);
        E::load(x, f + inner, b);
        const bool ma = (MASK == kMaskNaN) ? E::is_nan(a) : E::equals(a, mask_lo, mask_hi);
        const bool mb = (MASK == kMaskNaN) ? E::is_nan(b) : E::equals(b, mask_lo, mask_hi);
        bnd += (ma != mb) ? 1u : 0u;
        pairs += 1u;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        bnd += __shfl_xor_sync(0xFFFFFFFFu, bnd, o);
        pairs += __shfl_xor_sync(0xFFFFFFFFu, pairs, o);
    }
    if ((threadIdx.x & 31u) == 0u) {
        if (bnd) atomicAdd(out_bnd, (unsigned long long)bnd);
        if (pairs) atomicAdd(out_pairs, (unsigned long long)pairs);
    }
}

template <int KIND>
cudaError_t launch_density_kind(const CountJob& job, long long nsamples, long long stride, unsigned long long* out_bnd,
                                unsigned long long* out_pairs, cudaStream_t stream) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    const unsigned grid = (unsigned)((nsamples + kThreads - 1) / kThreads < 592 ? (nsamples + kThreads - 1) / kThreads : 592);
    const uint32_t mlo = (uint32_t)(job.mask_bits & 0xFFFFFFFFu), mhi = (uint32_t)(job.mask_bits >> 32);
    if (job.mask_mode == kMaskNaN) {
        if constexpr (isf) k_mask_density<KIND, kMaskNaN><<<grid, kThreads, 0, stream>>>(job.x, nsamples, stride, job.inner, mlo, mhi, out_bnd, out_pairs);
        else return cudaErrorInvalidValue;
    } else {
        k_mask_density<KIND, kMaskValue><<<grid, kThreads, 0, stream>>>(job.x, nsamples, stride, job.inner, mlo, mhi, out_bnd, out_pairs);
    }
    note_launch();
    return cudaGetLastError();
}

template <int KIND, int MASK, bool XFORM, int MODE>
cudaError_t launch_one(const CountJob& job, long long lo, long long count, RawSums* dst, cudaStream_t stream,
                       const RawSums* dense_probe, unsigned long long dense_one_in) {
    if (count <= 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const long long tiles = (count + kPairsPerTile - 1) / kPairsPerTile;
    long long grid = (long long)sms * 4;
    if (grid > tiles) grid = tiles;
    k_pairs_generic<KIND, MASK, XFORM, MODE><<<(unsigned)grid, kThreads, 0, stream>>>(
        job.x, lo, count, job.n, job.inner, (uint32_t)(job.mask_bits & 0xFFFFFFFFu),
        (uint32_t)(job.mask_bits >> 32), dst, dense_probe, dense_one_in);
    note_launch();
    return cudaGetLastError();
}

template <int KIND, int MODE>
cudaError_t dispatch_mask(const CountJob& job, long long lo, long long count, RawSums* dst, cudaStream_t stream,
                          const RawSums* dense_probe, unsigned long long dense_one_in) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    if (job.transform) {
        if constexpr (isf) {
            switch (job.mask_mode) {
                case kMaskNone: return launch_one<KIND, kMaskNone, true, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
                case kMaskNaN: return launch_one<KIND, kMaskNaN, true, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
                default: return launch_one<KIND, kMaskValue, true, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
            }
        } else {
            return cudaErrorInvalidValue;
        }
    }
    switch (job.mask_mode) {
        case kMaskNone: return launch_one<KIND, kMaskNone, false, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        case kMaskNaN:
            if constexpr (isf) return launch_one<KIND, kMaskNaN, false, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
            return cudaErrorInvalidValue;
        default: return launch_one<KIND, kMaskValue, false, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
    }
}

template <int MODE>
cudaError_t dispatch_kind(const CountJob& job, long long lo, long long count, RawSums* dst, cudaStream_t stream,
                          const RawSums* dense_probe = nullptr, unsigned long long dense_one_in = 0) {
    switch (job.dtype) {
        case XBI_U8: return dispatch_mask<kU8, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        case XBI_U16: return dispatch_mask<kU16, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        case XBI_U32: return dispatch_mask<kU32, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        case XBI_U64: return dispatch_mask<kU64, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        case XBI_F16: return dispatch_mask<kF16, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        case XBI_F32: return dispatch_mask<kF32, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        case XBI_F64: return dispatch_mask<kF64, MODE>(job, lo, count, dst, stream, dense_probe, dense_one_in);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace

cudaError_t launch_pairs_generic_flat(const CountJob& job, int64_t lo, int64_t hi, RawSums* dst,
                                      cudaStream_t stream) {
    return dispatch_kind<0>(job, lo, hi - lo, dst, stream);
}

cudaError_t launch_pairs_generic_edges(const CountJob& job, int64_t t_begin, int64_t t_end, RawSums* dst,
                                       cudaStream_t stream) {
    const int64_t total = (job.outer - 1) * job.inner;
    if (t_begin < 0) t_begin = 0;
    if (t_end > total) t_end = total;
    if (t_end <= t_begin) return cudaSuccess;
    return dispatch_kind<1>(job, t_begin, t_end - t_begin, dst, stream);
}

cudaError_t launch_mask_density(const CountJob& job, int64_t npairs_flat, Workspace* ws, cudaStream_t stream) {
    return launch_mask_density_to(job, npairs_flat, &ws->plus.pad[0], &ws->plus.pad[1], stream);
}

cudaError_t launch_mask_density_to(const CountJob& job, int64_t npairs_flat, unsigned long long* out_bnd,
                                   unsigned long long* out_pairs, cudaStream_t stream) {
    if (job.mask_mode == kMaskNone || npairs_flat <= 0) return cudaSuccess;
    long long nsamples = npairs_flat < (1ll << 18) ? npairs_flat : (1ll << 18);
    long long stride = npairs_flat / nsamples;
    if (stride > 1) stride |= 1;  // odd: does not lock onto even row lengths
    while (stride > 1 && (nsamples - 1) * stride >= npairs_flat) --nsamples;
    switch (job.dtype) {
        case XBI_U8: return launch_density_kind<kU8>(job, nsamples, stride, out_bnd, out_pairs, stream);
        case XBI_U16: return launch_density_kind<kU16>(job, nsamples, stride, out_bnd, out_pairs, stream);
        case XBI_U32: return launch_density_kind<kU32>(job, nsamples, stride, out_bnd, out_pairs, stream);
        case XBI_U64: return launch_density_kind<kU64>(job, nsamples, stride, out_bnd, out_pairs, stream);
        case XBI_F16: return launch_density_kind<kF16>(job, nsamples, stride, out_bnd, out_pairs, stream);
        case XBI_F32: return launch_density_kind<kF32>(job, nsamples, stride, out_bnd, out_pairs, stream);
        case XBI_F64: return launch_density_kind<kF64>(job, nsamples, stride, out_bnd, out_pairs, stream);
        default: return cudaErrorInvalidValue;
    }
}

cudaError_t launch_elem_planes(const CountJob& job, int64_t lo, int64_t count, RawSums* dst, cudaStream_t stream,
                               const RawSums* dense_probe, unsigned long long dense_one_in) {
    return dispatch_kind<2>(job, lo, count, dst, stream, dense_probe, dense_one_in);
}

}  // namespace xbi
