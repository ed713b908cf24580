// The finish kernel: everything that follows the bulk counting kernel of one (variable, axis) job, in ONE
// single-CTA launch instead of five or six --
//   1. the fringe: the few flat pairs the fast kernel could not take (unaligned head, ragged tail shorter than
//      one stage) and the few row-straddling pairs among them, counted with the three explicit streams of the
//      generic kernel (xbi_count_generic.cu);
//   2. combine: raw plane sums of the workspace (+ fringe) -> the (nbits,2,2) table of bitpaircount
//      (xbitinfo/_py_bitinfo.py:128-140), added to or written over the caller's counts;
//   3. optionally the all-reduce over the GPUs of one node (the path's one collective) as a
//      push over NVLink peer memory: every rank stores its table into a slot of every peer's exchange buffer,
//      raises a flag there, waits for the flags of all peers in its own buffer and sums the slots;
//   4. optionally K2, the mutual information per bit (xbitinfo/_py_bitinfo.py:143-152).
// On configs[1] this replaced a (1,1,1)-grid tail kernel, an edge kernel, the combine kernel, the K2 kernel and a
// fill kernel (66 us of fixed cost per step; with the NCCL all-reduce 200 us at 8 GPUs).
#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_epilogue.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

constexpr int kFinishThreads = 512;
constexpr int kFinishTile = 16 * kFinishThreads;

__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_volatile(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

template <int KIND, int MASK, bool XFORM>
__global__ void __launch_bounds__(kFinishThreads, 1) k_finish(const FinishJob J) {
    using E = Elem<KIND>;
    constexpr int NW = E::kWords;
    constexpr int NBITS = E::kBits;
    using L = Ladder<2, 3>;
    const uint32_t lane = threadIdx.x & 31u;

    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_n;
    __shared__ int s_err;
    for (int i = threadIdx.x; i < 3 * 64; i += kFinishThreads) (&s_sum[0][0])[i] = 0;
    if (threadIdx.x == 0) { s_n = 0; s_err = 0; }
    __syncthreads();

    // ---- 1. fringe: kind 0 = flat pairs (added), kind 1 = row-straddling pairs (subtracted) ----
#pragma unroll 1
    for (int kind = 0; kind < 2; ++kind) {
        if (J.range_cnt[2 * kind] + J.range_cnt[2 * kind + 1] <= 0) continue;  // uniform
        L A[NW], B[NW], AB[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) { A[w].clear(); B[w].clear(); AB[w].clear(); }
        unsigned long long accA[NW], accB[NW], accAB[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) { accA[w] = 0; accB[w] = 0; accAB[w] = 0; }
        uint32_t nvalid = 0, it = 0;
        unsigned long long nvalid_total = 0;
        auto flush = [&]() {
            if (it == 0) return;
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                accA[w] += A[w].warp_totals_bounded(lane, it);
                accB[w] += B[w].warp_totals_bounded(lane, it);
                accAB[w] += AB[w].warp_totals_bounded(lane, it);
                A[w].clear(); B[w].clear(); AB[w].clear();
            }
            nvalid_total += nvalid;
            nvalid = 0;
            it = 0;
        };
#pragma unroll 1
        for (int r = 2 * kind; r < 2 * kind + 2; ++r) {
            const long long lo = J.range_lo[r], count = J.range_cnt[r];
#pragma unroll 1
            for (long long t0 = 0; t0 < count; t0 += kFinishTile) {
                uint32_t wa[NW][16], wb[NW][16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const long long t = t0 + (long long)u * kFinishThreads + threadIdx.x;
                    uint32_t a[NW], b[NW];
#pragma unroll
                    for (int w = 0; w < NW; ++w) { a[w] = 0; b[w] = 0; }
                    if (t < count) {
                        long long f = lo + t;
                        if (kind == 1) {  // edge index -> flat index of the last element of its block
                            const long long o = f / J.inner, i = f - o * J.inner;
                            f = (o * J.n + (J.n - 1)) * J.inner + i;
                        }
                        E::load(J.x, f, a);
                        E::load(J.x, f + J.inner, b);
                        bool ok = true;
                        if (MASK == kMaskNaN) ok = !(E::is_nan(a) || E::is_nan(b));
                        if (MASK == kMaskValue) ok = !(E::equals(a, J.mask_lo, J.mask_hi) || E::equals(b, J.mask_lo, J.mask_hi));
                        if (XFORM) { E::transform(a); E::transform(b); }
                        if (!ok) {
#pragma unroll
                            for (int w = 0; w < NW; ++w) { a[w] = 0; b[w] = 0; }
                        }
                        nvalid += ok ? 1u : 0u;
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
        }
        flush();
        // unsigned wrap-around: subtracting is adding the two's complement; the combined totals are exact
        const bool neg = kind == 1;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int j = w * 32 + (int)lane;
            if (accA[w]) atomicAdd(&s_sum[S_A][j], neg ? 0ull - accA[w] : accA[w]);
            if (accB[w]) atomicAdd(&s_sum[S_B][j], neg ? 0ull - accB[w] : accB[w]);
            if (accAB[w]) atomicAdd(&s_sum[S_AB][j], neg ? 0ull - accAB[w] : accAB[w]);
        }
        if (nvalid_total) atomicAdd(&s_n, neg ? 0ull - nvalid_total : nvalid_total);
    }
    __syncthreads();

    // ---- 2. combine: thread j owns bit j (counted from the least significant bit) ----
    const int j = threadIdx.x;
    const int k = NBITS - 1 - j;  // MSB-first position in the counts table
    unsigned long long c[4] = {0, 0, 0, 0};
    if (j < NBITS) {
        const Workspace* ws = J.ws;
        const unsigned long long A = ws->plus.sums[S_A][j] - ws->minus.sums[S_A][j] + s_sum[S_A][j];
        const unsigned long long B = ws->plus.sums[S_B][j] - ws->minus.sums[S_B][j] + s_sum[S_B][j];
        const unsigned long long AB = ws->plus.sums[S_AB][j] - ws->minus.sums[S_AB][j] + s_sum[S_AB][j];
        const unsigned long long N = ws->plus.npairs - ws->minus.npairs + s_n;
        c[0] = N - A - B + AB;
        c[1] = B - AB;
        c[2] = A - AB;
        c[3] = AB;
    }

    // ---- 3. all-reduce over peer memory ----
    if (J.px.world > 1) {
        const PeerExchange& px = J.px;
        const unsigned par = (unsigned)(px.epoch & 1ull);
        if (j < NBITS) {
            const size_t slot = ((size_t)par * kPeerMaxWorld + (size_t)px.rank) * kPeerCells + 4 * (size_t)k;
            for (int p = 0; p < px.world; ++p) {
                unsigned long long* dst = px.peer[p] + slot;
                dst[0] = c[0]; dst[1] = c[1]; dst[2] = c[2]; dst[3] = c[3];
            }
            __threadfence_system();
        }
        __syncthreads();
        if (j < px.world) {
            const size_t flag_base = (size_t)2 * kPeerMaxWorld * kPeerCells + (size_t)par * kPeerMaxWorld;
            st_release_sys(px.peer[j] + flag_base + px.rank, px.epoch);       // "my table for this epoch is in your buffer"
            const unsigned long long* mine = px.peer[px.rank] + flag_base + j;  // ... and rank j's in mine?
            const unsigned long long t0 = global_timer_ns();
            while (ld_acquire_sys(mine) < px.epoch) {
                if (global_timer_ns() - t0 > px.timeout_ns) {  // a peer never arrived: report, do not hang the GPU
                    s_err = 1;
                    break;
                }
                __nanosleep(100);
            }
        }
        __syncthreads();
        if (s_err) {
            if (threadIdx.x == 0) px.peer[px.rank][kPeerErrorWord] = px.epoch;
        }
        if (j < NBITS) {
            c[0] = c[1] = c[2] = c[3] = 0;
            for (int q = 0; q < px.world; ++q) {
                const unsigned long long* src =
                    px.peer[px.rank] + ((size_t)par * kPeerMaxWorld + (size_t)q) * kPeerCells + 4 * (size_t)k;
#pragma unroll
                for (int i = 0; i < 4; ++i) c[i] += ld_volatile(src + i);
            }
        }
    }

    // ---- 4. counts out, K2 ----
    if (j < NBITS) {
        unsigned long long* out = J.counts + 4 * k;
        if (J.accumulate) {
            // atomics: several streams may add slabs of the same variable into one counts buffer
#pragma unroll
            for (int i = 0; i < 4; ++i) c[i] += atomicAdd(out + i, c[i]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i) out[i] = c[i];
        }
        if (J.mi) J.mi[k] = mutual_information_bit(c[0], c[1], c[2], c[3], J.set_zero, J.z);
    }
}

template <int KIND, int MASK, bool XFORM>
cudaError_t launch_one(const FinishJob& J, cudaStream_t stream) {
    k_finish<KIND, MASK, XFORM><<<1, kFinishThreads, 0, stream>>>(J);
    note_launch();
    return cudaGetLastError();
}

template <int KIND>
cudaError_t dispatch_mask(const CountJob& job, const FinishJob& J, cudaStream_t stream) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    if (job.transform) {
        if constexpr (isf) {
            switch (job.mask_mode) {
                case kMaskNone: return launch_one<KIND, kMaskNone, true>(J, stream);
                case kMaskNaN: return launch_one<KIND, kMaskNaN, true>(J, stream);
                default: return launch_one<KIND, kMaskValue, true>(J, stream);
            }
        } else {
            return cudaErrorInvalidValue;
        }
    }
    switch (job.mask_mode) {
        case kMaskNone: return launch_one<KIND, kMaskNone, false>(J, stream);
        case kMaskNaN:
            if constexpr (isf) return launch_one<KIND, kMaskNaN, false>(J, stream);
            return cudaErrorInvalidValue;
        default: return launch_one<KIND, kMaskValue, false>(J, stream);
    }
}

}  // namespace

cudaError_t launch_finish(const CountJob& job, const FinishJob& J, cudaStream_t stream) {
    switch (job.dtype) {
        case XBI_U8: return dispatch_mask<kU8>(job, J, stream);
        case XBI_U16: return dispatch_mask<kU16>(job, J, stream);
        case XBI_U32: return dispatch_mask<kU32>(job, J, stream);
        case XBI_U64: return dispatch_mask<kU64>(job, J, stream);
        case XBI_F16: return dispatch_mask<kF16>(job, J, stream);
        case XBI_F32: return dispatch_mask<kF32>(job, J, stream);
        case XBI_F64: return dispatch_mask<kF64>(job, J, stream);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace xbi
