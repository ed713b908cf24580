// K1 over a chunk grid, fast formulation (docs/chunking.ipynb cells 8-10: get_bitinformation on every dask block on
// its own).  The first chunk-grid kernel (k_pairs_chunked, xbi_chunked.cu) is the generic formulation batched: every
// thread steps a mixed-radix odometer per ELEMENT and loads both elements of every pair -- 81 instructions per
// element, 0.9-1.25 TB/s.  Here a WARP owns the indexing and a lane owns a column:
//
//   analysed axis is not the innermost dimension (kWalk)
//       a lane takes one column of the chunk -- fixed coordinates in every dimension but the analysed one, 32
//       neighbouring lanes = 32 neighbouring elements of the innermost dimension -- and walks down the analysed
//       axis 16 rows per step, the previous row in registers: ONE load per element, one address add per row.
//       The column's offset is worked out once per column (divisions by the chunk's extents), not per element.
//   analysed axis is the innermost dimension (kRun)
//       a warp takes one run of the chunk (its contiguous innermost extent) at a time, 32 consecutive elements per
//       step; the position (run, step) is warp-uniform and advanced by a uniform odometer once per step; a lane
//       reads its element and the next one (the second read hits the same line in L1).
//
// Three explicit word streams a, b, a&b of the valid pairs go into the bit-sliced counters of xbi_common.cuh.  A
// warp works through a CONTIGUOUS range of (chunk, group) items, so it changes chunk a few times in its whole life
// and only then (and at the end) pays the butterfly flush and the atomics into that chunk's table.
#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

constexpr int kWThreads = 256;
constexpr int kWWarps = kWThreads / 32;
constexpr int kD = kMaxChunkDims;

// extents and base offset of chunk `cid` (uniform)
struct ChunkBox {
    unsigned e[kD];
    long long base;
    __device__ __forceinline__ void init(const ChunkGrid& g, unsigned cid) {
        base = 0;
        unsigned r = cid;
#pragma unroll
        for (int d = kD - 1; d >= 0; --d) {
            unsigned q = r;
            if (d > 0) { q = r % g.nchunk[d]; r /= g.nchunk[d]; }
            long long start;
            if (g.bounds[d] != nullptr) {
                const long long b0 = g.bounds[d][q], b1 = g.bounds[d][q + 1];
                start = b0 * g.scale[d];
                e[d] = (unsigned)((b1 - b0) * g.scale[d]);
            } else {
                start = (long long)q * g.chunk[d];
                const long long ext = g.shape[d] - start;
                e[d] = ext < (long long)g.chunk[d] ? (unsigned)ext : g.chunk[d];
            }
            base += start * g.stride[d];
        }
    }
};

template <int KIND, int MASK, bool XFORM, bool RUN>
__global__ void __launch_bounds__(kWThreads, Elem<KIND>::kWords == 1 ? 2 : 1)
k_pairs_chunked_walk(const void* __restrict__ x, const __grid_constant__ ChunkGrid g, uint32_t mask_lo, uint32_t mask_hi,
                     unsigned long long* __restrict__ counts, unsigned long long items, unsigned groups) {
    using E = Elem<KIND>;
    constexpr int NW = E::kWords;
    constexpr int NBITS = E::kBits;
    using L = Ladder<3, 8>;
    const uint32_t lane = threadIdx.x & 31u;
    const unsigned long long gw = (unsigned long long)blockIdx.x * kWWarps + (threadIdx.x >> 5);
    const unsigned long long nw = (unsigned long long)gridDim.x * kWWarps;
    const unsigned long long item_lo = gw * items / nw, item_hi = (gw + 1) * items / nw;  // this warp's items
    const int axis = g.axis;
    long long astride = 0;
#pragma unroll
    for (int d = 0; d < kD; ++d) if (d == axis) astride = g.stride[d];

    L A[NW], B[NW], AB[NW];
#pragma unroll
    for (int k = 0; k < NW; ++k) { A[k].clear(); B[k].clear(); AB[k].clear(); }
    uint32_t it = 0, nvalid = 0;
    unsigned long long accA[NW], accB[NW], accAB[NW], nvalid_total = 0;
#pragma unroll
    for (int k = 0; k < NW; ++k) { accA[k] = 0; accB[k] = 0; accAB[k] = 0; }
    auto drain = [&]() {  // ladders -> per-lane totals (lane L: bit L)
        if (it == 0) return;
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
    auto flush_chunk = [&](unsigned cid) {  // per-lane totals -> the chunk's raw table (see k_chunk_finalize)
        drain();
        unsigned long long* dst = counts + (size_t)cid * NBITS * 4;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int j = k * 32 + (int)lane;
            if (j < NBITS) {
                if (accB[k]) atomicAdd(dst + 4 * (NBITS - 1 - j) + 1, accB[k]);
                if (accA[k]) atomicAdd(dst + 4 * (NBITS - 1 - j) + 2, accA[k]);
                if (accAB[k]) atomicAdd(dst + 4 * (NBITS - 1 - j) + 3, accAB[k]);
            }
            accA[k] = 0; accB[k] = 0; accAB[k] = 0;
        }
        nvalid_total += __shfl_xor_sync(0xFFFFFFFFu, nvalid_total, 16);
        nvalid_total += __shfl_xor_sync(0xFFFFFFFFu, nvalid_total, 8);
        nvalid_total += __shfl_xor_sync(0xFFFFFFFFu, nvalid_total, 4);
        nvalid_total += __shfl_xor_sync(0xFFFFFFFFu, nvalid_total, 2);
        nvalid_total += __shfl_xor_sync(0xFFFFFFFFu, nvalid_total, 1);
        if (lane == 0 && nvalid_total) atomicAdd(dst, nvalid_total);
        nvalid_total = 0;
    };
    // fold 16 pairs per lane: words of the a and b elements (already transformed), `ok` bit u = pair u counts
    auto fold = [&](uint32_t (&wa)[NW][16], uint32_t (&wb)[NW][16], uint32_t ok) {
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            uint32_t wab[16];
#pragma unroll
            for (int u = 0; u < 16; ++u) {
                const bool on = (ok >> u) & 1u;
                wa[k][u] = on ? wa[k][u] : 0u;
                wb[k][u] = on ? wb[k][u] : 0u;
                wab[u] = wa[k][u] & wb[k][u];
            }
            A[k].add16(wa[k], it);
            B[k].add16(wb[k], it);
            AB[k].add16(wab, it);
        }
        nvalid += __popc(ok);
        ++it;
        if (it == L::kFlushChunks) drain();
    };
    auto is_masked = [&](const uint32_t (&w)[NW]) {
        if (MASK == kMaskNaN) return E::is_nan(w);
        if (MASK == kMaskValue) return E::equals(w, mask_lo, mask_hi);
        return false;
    };

    unsigned cur_chunk = 0xFFFFFFFFu;
    ChunkBox box;
    for (unsigned long long item = item_lo; item < item_hi; ++item) {
        const unsigned cid = (unsigned)(item / groups), grp = (unsigned)(item - (unsigned long long)cid * groups);
        if (cid != cur_chunk) {
            if (cur_chunk != 0xFFFFFFFFu) flush_chunk(cur_chunk);
            cur_chunk = cid;
            box.init(g, cid);
        }
        unsigned ea = 1;
#pragma unroll
        for (int d = 0; d < kD; ++d) if (d == axis) ea = box.e[d];
        if (ea < 2) continue;  // no pairs in this chunk

        if constexpr (!RUN) {
            // ---------------- column walkers ----------------
            unsigned long long O = 1, I = 1;
#pragma unroll
            for (int d = 0; d < kD; ++d) {
                if (d < axis) O *= box.e[d];
                if (d > axis) I *= box.e[d];
            }
            const unsigned long long tiles_per_o = (I + 31) / 32, tiles = O * tiles_per_o;
            const unsigned long long t_lo = tiles * grp / groups, t_hi = tiles * (grp + 1) / groups;
            for (unsigned long long tile = t_lo; tile < t_hi; ++tile) {
                const unsigned long long o = tile / tiles_per_o, ti = tile - o * tiles_per_o;
                const unsigned long long i = ti * 32 + lane;
                const bool live = i < I;
                // offset of this lane's column at row 0 of the chunk
                long long off = box.base;
                {
                    unsigned long long ro = o, ri = live ? i : 0;
#pragma unroll
                    for (int d = kD - 1; d >= 0; --d) {
                        if (d > axis) { off += (long long)(ri % box.e[d]) * g.stride[d]; ri /= box.e[d]; }
                        if (d < axis) { off += (long long)(ro % box.e[d]) * g.stride[d]; ro /= box.e[d]; }
                    }
                }
                uint32_t prev[NW];
                bool pv = false;
#pragma unroll
                for (int k = 0; k < NW; ++k) prev[k] = 0;
                if (live) {
                    E::load(x, off, prev);
                    pv = !is_masked(prev);
                    if (XFORM) E::transform(prev);
                }
                const unsigned nsteps = (ea - 1 + 15) / 16;
                long long roff = off + astride;  // row 1
                for (unsigned s = 0; s < nsteps; ++s) {
                    const unsigned r0 = 1 + 16 * s;
                    uint32_t raw[NW][16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        uint32_t w[NW];
#pragma unroll
                        for (int k = 0; k < NW; ++k) w[k] = 0;
                        if (live && r0 + u < ea) E::load(x, roff + (long long)u * astride, w);
#pragma unroll
                        for (int k = 0; k < NW; ++k) raw[k][u] = w[k];
                    }
                    roff += 16 * astride;
                    uint32_t wa[NW][16], wb[NW][16], ok = 0;
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const bool present = live && r0 + u < ea;
                        uint32_t w[NW];
#pragma unroll
                        for (int k = 0; k < NW; ++k) w[k] = raw[k][u];
                        const bool v = present && !is_masked(w);
                        if (XFORM) E::transform(w);
                        ok |= ((present && v && pv) ? 1u : 0u) << u;
#pragma unroll
                        for (int k = 0; k < NW; ++k) { wa[k][u] = prev[k]; wb[k][u] = w[k]; }
                        if (present) {
#pragma unroll
                            for (int k = 0; k < NW; ++k) prev[k] = w[k];
                            pv = v;
                        }
                    }
                    fold(wa, wb, ok);
                }
            }
        } else {
            // ---------------- runs along the innermost dimension ----------------
            unsigned long long R = 1;
#pragma unroll
            for (int d = 0; d < kD - 1; ++d) R *= box.e[d];
            const unsigned long long r_lo = R * grp / groups, r_hi = R * (grp + 1) / groups;
            if (r_hi <= r_lo) continue;
            const unsigned e4 = box.e[kD - 1];
            // uniform odometer over dimensions 0 .. kD-2, started at run r_lo
            unsigned c[kD - 1];
            long long roff = box.base;
            {
                unsigned long long rr = r_lo;
#pragma unroll
                for (int d = kD - 2; d >= 0; --d) {
                    c[d] = (unsigned)(rr % box.e[d]);
                    rr /= box.e[d];
                    roff += (long long)c[d] * g.stride[d];
                }
            }
            unsigned long long runs_left = r_hi - r_lo;
            unsigned j0 = 0;
            while (runs_left > 0) {
                uint32_t wa[NW][16], wb[NW][16], ok = 0;
                long long offs[16];
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    const unsigned j = j0 + lane;
                    const bool pair = runs_left > 0 && j + 1 < e4;
                    offs[u] = pair ? roff + j : -1;
                    ok |= (pair ? 1u : 0u) << u;
                    // next (run, step): uniform
                    j0 += 32;
                    if (j0 + 1 >= e4 && runs_left > 0) {
                        j0 = 0;
                        --runs_left;
#pragma unroll
                        for (int d = kD - 2; d >= 0; --d) {
                            if (d >= g.lo) {
                                roff += g.stride[d];
                                if (++c[d] < box.e[d]) break;
                                roff -= (long long)box.e[d] * g.stride[d];
                                c[d] = 0;
                            }
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    uint32_t a[NW], b[NW];
#pragma unroll
                    for (int k = 0; k < NW; ++k) { a[k] = 0; b[k] = 0; }
                    if (offs[u] >= 0) {
                        E::load(x, offs[u], a);
                        E::load(x, offs[u] + 1, b);
                    }
#pragma unroll
                    for (int k = 0; k < NW; ++k) { wa[k][u] = a[k]; wb[k][u] = b[k]; }
                }
#pragma unroll
                for (int u = 0; u < 16; ++u) {
                    uint32_t a[NW], b[NW];
#pragma unroll
                    for (int k = 0; k < NW; ++k) { a[k] = wa[k][u]; b[k] = wb[k][u]; }
                    if (MASK != kMaskNone) {
                        if (is_masked(a) || is_masked(b)) ok &= ~(1u << u);
                    }
                    if (XFORM) { E::transform(a); E::transform(b); }
#pragma unroll
                    for (int k = 0; k < NW; ++k) { wa[k][u] = a[k]; wb[k][u] = b[k]; }
                }
                fold(wa, wb, ok);
            }
        }
    }
    if (cur_chunk != 0xFFFFFFFFu) flush_chunk(cur_chunk);
}

template <int KIND, int MASK, bool XFORM>
cudaError_t launch_walk(const void* x, const ChunkGrid& g, uint32_t mlo, uint32_t mhi, unsigned long long* counts,
                        unsigned long long items, unsigned groups, unsigned grid, cudaStream_t stream) {
    if (g.axis == kD - 1)
        k_pairs_chunked_walk<KIND, MASK, XFORM, true><<<grid, kWThreads, 0, stream>>>(x, g, mlo, mhi, counts, items, groups);
    else
        k_pairs_chunked_walk<KIND, MASK, XFORM, false><<<grid, kWThreads, 0, stream>>>(x, g, mlo, mhi, counts, items, groups);
    return cudaGetLastError();
}

template <int KIND>
cudaError_t dispatch_walk(const void* x, const ChunkGrid& g, bool transform, int mask_mode, uint32_t mlo, uint32_t mhi,
                          unsigned long long* counts, unsigned long long items, unsigned groups, unsigned grid,
                          cudaStream_t stream) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    if (transform) {
        if constexpr (isf) {
            switch (mask_mode) {
                case kMaskNone: return launch_walk<KIND, kMaskNone, true>(x, g, mlo, mhi, counts, items, groups, grid, stream);
                case kMaskNaN: return launch_walk<KIND, kMaskNaN, true>(x, g, mlo, mhi, counts, items, groups, grid, stream);
                default: return launch_walk<KIND, kMaskValue, true>(x, g, mlo, mhi, counts, items, groups, grid, stream);
            }
        }
        return cudaErrorInvalidValue;
    }
    switch (mask_mode) {
        case kMaskNone: return launch_walk<KIND, kMaskNone, false>(x, g, mlo, mhi, counts, items, groups, grid, stream);
        case kMaskNaN:
            if constexpr (isf) return launch_walk<KIND, kMaskNaN, false>(x, g, mlo, mhi, counts, items, groups, grid, stream);
            return cudaErrorInvalidValue;
        default: return launch_walk<KIND, kMaskValue, false>(x, g, mlo, mhi, counts, items, groups, grid, stream);
    }
}

}  // namespace

// Counts into the raw per-chunk tables (the caller runs k_chunk_finalize) when the grid suits the walkers:
// *taken = false leaves everything to the odometer kernel of xbi_chunked.cu.
cudaError_t launch_pairs_chunked_walk(const void* x, int dtype, const ChunkGrid& g, unsigned nchunks, bool transform,
                                      int mask_mode, uint64_t mask_bits, unsigned long long* counts, cudaStream_t stream,
                                      bool force, bool* taken) {
    *taken = false;
    if (g.axis < 0 || nchunks == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = dtype_bytes(dtype) == 8 ? 1 : 2;
    const unsigned long long warps = (unsigned long long)sms * per_sm * kWWarps;
    // the work inside a (full) chunk that the warps share: column tiles, or runs
    unsigned long long units = 1;
    if (g.axis == kD - 1) {
        if (g.chunk[kD - 1] < 16 && !force) return cudaSuccess;  // runs too short to give a warp's lanes work
        for (int d = 0; d < kD - 1; ++d) units *= g.chunk[d];
    } else {
        unsigned long long O = 1, I = 1;
        for (int d = 0; d < kD; ++d) {
            if (d < g.axis) O *= g.chunk[d];
            if (d > g.axis) I *= g.chunk[d];
        }
        units = O * ((I + 31) / 32);
    }
    if (units * nchunks < 2 * warps && !force) return cudaSuccess;  // too few columns / runs to occupy the GPU this way
    // groups per chunk: about 8 items per warp overall, at least one unit per item
    unsigned long long groups = (8 * warps + nchunks - 1) / nchunks;
    if (groups > units) groups = units;
    if (groups < 1) groups = 1;
    if (groups > 0xFFFFFFFFull) return cudaSuccess;
    const unsigned long long items = groups * nchunks;
    const unsigned grid = (unsigned)(sms * per_sm);
    const uint32_t mlo = (uint32_t)(mask_bits & 0xFFFFFFFFu), mhi = (uint32_t)(mask_bits >> 32);
    cudaError_t e;
    switch (dtype) {
        case XBI_U8: e = dispatch_walk<kU8>(x, g, transform, mask_mode, mlo, mhi, counts, items, (unsigned)groups, grid, stream); break;
        case XBI_U16: e = dispatch_walk<kU16>(x, g, transform, mask_mode, mlo, mhi, counts, items, (unsigned)groups, grid, stream); break;
        case XBI_U32: e = dispatch_walk<kU32>(x, g, transform, mask_mode, mlo, mhi, counts, items, (unsigned)groups, grid, stream); break;
        case XBI_U64: e = dispatch_walk<kU64>(x, g, transform, mask_mode, mlo, mhi, counts, items, (unsigned)groups, grid, stream); break;
        case XBI_F16: e = dispatch_walk<kF16>(x, g, transform, mask_mode, mlo, mhi, counts, items, (unsigned)groups, grid, stream); break;
        case XBI_F32: e = dispatch_walk<kF32>(x, g, transform, mask_mode, mlo, mhi, counts, items, (unsigned)groups, grid, stream); break;
        case XBI_F64: e = dispatch_walk<kF64>(x, g, transform, mask_mode, mlo, mhi, counts, items, (unsigned)groups, grid, stream); break;
        default: return cudaErrorInvalidValue;
    }
    if (e != cudaSuccess) return e;
    note_fast_launch();
    *taken = true;
    return cudaSuccess;
}

}  // namespace xbi
