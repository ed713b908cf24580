// K1 fast path for adjacency along a NON-last axis (inner > 1): column walkers.
//
// The C-contiguous array is a matrix of R = outer*n rows of `inner` elements; the partner
// of an element is the one directly below it.  A thread owns one 16-byte column chunk and
// walks down a segment of rows with 128-bit coalesced, L1-bypassing loads (a warp reads
// 512 contiguous bytes per row), keeping the previous row in registers -- every element
// is read from HBM exactly once and needs no shared-memory staging or halo.  Four rows
// (16 words) are folded per step into the bit-sliced counters of xbi_common.cuh, with the
// next four rows already in flight.
//
// Unmasked, only two word streams are needed: b (every row below the first) and a&b.
// The a-stream total is the b-stream total + bits(first row of the matrix) - bits(its last
// row); two small row-plane passes (xbi_count_generic.cu) add those to the plus / minus sums.
// Pairs that straddle two blocks of the analysed axis are subtracted by the edge pass
// (xbi_count_generic.cu).
#include <cstdlib>

#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"
#include "xbi_tma.cuh"
#include "xbi_tmap.h"

namespace xbi {

namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ldg_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}

// lane L gets the number of lanes whose word has bit L set (rare use: once per thread)
__device__ __forceinline__ uint32_t warp_bit_totals(uint32_t w, uint32_t lane) {
    uint32_t mine = 0;
#pragma unroll 4
    for (uint32_t k = 0; k < 32; ++k) {
        const uint32_t b = __ballot_sync(0xFFFFFFFFu, (w >> k) & 1u);
        if (lane == k) mine = __popc(b);
    }
    return mine;
}

// transform the four words of a 16-byte chunk in place
template <int KIND, bool XFORM>
__device__ __forceinline__ void xform_chunk(uint32_t (&w)[4]) {
    if (!XFORM) return;
    if (KIND == kF32) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = sexp_f32(w[i]);
    } else if (KIND == kF16) {
#pragma unroll
        for (int i = 0; i < 4; ++i) w[i] = sexp_f16x2(w[i]);
    } else if (KIND == kF64) {
        w[1] = sexp_f64_hi(w[1]);
        w[3] = sexp_f64_hi(w[3]);
    }
}

// all-ones in the field of every element of the chunk that is NOT masked (raw words)
template <int KIND, int MASK>
__device__ __forceinline__ void keep_chunk(const uint32_t (&w)[4], uint32_t mlo, uint32_t mhi, uint32_t (&k)[4]) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    if (BYTES == 8) {
#pragma unroll
        for (int e = 0; e < 2; ++e) {
            const uint32_t lo = w[2 * e], hi = w[2 * e + 1];
            bool masked;
            if (MASK == kMaskNaN) {
                const uint32_t m = hi & 0x7FFFFFFFu;
                masked = m > 0x7FF00000u || (m == 0x7FF00000u && lo != 0u);
            } else {
                masked = (lo == mlo) && (hi == mhi);
            }
            k[2 * e] = k[2 * e + 1] = masked ? 0u : 0xFFFFFFFFu;
        }
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t x = w[i];
            if (BYTES == 4) {
                if (MASK == kMaskNaN) k[i] = ((x & 0x7FFFFFFFu) > 0x7F800000u) ? 0u : 0xFFFFFFFFu;
                else k[i] = (x == mlo) ? 0u : 0xFFFFFFFFu;
            } else if (BYTES == 2) {
                if (MASK == kMaskNaN) {
                    const uint32_t nan = (((x & 0x7FFF7FFFu) + 0x03FF03FFu) >> 15) & 0x00010001u;
                    k[i] = ~(nan * 0xFFFFu);
                } else {
                    const uint32_t d = x ^ (mlo * 0x00010001u);
                    k[i] = (((((d & 0x7FFF7FFFu) + 0x7FFF7FFFu) | d) >> 15) & 0x00010001u) * 0xFFFFu;
                }
            } else {
                const uint32_t d = x ^ (mlo * 0x01010101u);
                k[i] = (((((d & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | d) >> 7) & 0x01010101u) * 0xFFu;
            }
        }
    }
}
template <int BYTES>
__device__ __forceinline__ uint32_t count_chunk(const uint32_t (&v)[4]) {
    if (BYTES == 8) return (v[0] & 1u) + (v[2] & 1u);
    if (BYTES == 4) return (v[0] & 1u) + (v[1] & 1u) + (v[2] & 1u) + (v[3] & 1u);
    constexpr uint32_t m = (BYTES == 2) ? 0x00010001u : 0x01010101u;
    return __popc(v[0] & m) + __popc(v[1] & m) + __popc(v[2] & m) + __popc(v[3] & m);
}

// Work item = (segment, chunk) = divmod(item, chunks_per_row): one 16-byte column chunk
// walked down one segment of rows.  The pair rows [0, nseg*seg_min + seg_extra) are split
// into nseg segments; the first seg_extra segments are one row longer.  A warp takes 32
// consecutive items at a time (grid-stride over "warp items"), so all its lanes run the
// same number of full 4-row groups -- the flush shuffles rely on that -- plus one
// predicated tail group for the remaining 0..4 rows of each lane's segment.
template <int KIND, int MASK, bool XFORM>
__global__ void __launch_bounds__(kThreads, (MASK == kMaskNone) ? 2 : 1)
k_pairs_cols(const uint4* __restrict__ x, long long chunks_per_row, long long nseg, long long seg_min,
             long long seg_extra, uint32_t mask_lo, uint32_t mask_hi, Workspace* __restrict__ ws) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    constexpr bool K64 = (BYTES == 8);
    constexpr bool MASKED = (MASK != kMaskNone);
    constexpr int NL = K64 ? 2 : 1;             // ladders per stream (low / high words)
    constexpr int LOGC = K64 ? 3 : 4;           // 4 rows give 8 low + 8 high words, or 16 words
    constexpr int CH = 1 << LOGC;
    constexpr int kNumLadders = NL * (MASKED ? 3 : 2);
    using L = Ladder<3, (kNumLadders >= 6 ? 4 : kNumLadders >= 4 ? 5 : 8), LOGC>;

    const uint32_t lane = threadIdx.x & 31u;
    const long long total = chunks_per_row * nseg;
    const long long n_witems = (total + 31) >> 5;
    const long long warps_total = (long long)gridDim.x * (kThreads / 32);
    const long long groups = seg_min >> 2;  // full 4-row groups, the same for every item
    const long long stride = chunks_per_row * 16;  // bytes between rows

    L Bl[NL], ABl[NL], Al[MASKED ? NL : 1];
#pragma unroll
    for (int i = 0; i < NL; ++i) { Bl[i].clear(); ABl[i].clear(); }
    if (MASKED) {
#pragma unroll
        for (int i = 0; i < NL; ++i) Al[i].clear();
    }
    unsigned long long accA[NL], accB[NL], accAB[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { accA[i] = 0; accB[i] = 0; accAB[i] = 0; }
    uint32_t nvalid = 0;
    unsigned long long nvalid_total = 0;
    uint32_t it = 0;
    bool lane_live = true;  // false while this lane shadows an item beyond the end (last warp item)

    auto flush = [&]() {
        if (!lane_live) {
#pragma unroll
            for (int i = 0; i < NL; ++i) { Bl[i].clear(); ABl[i].clear(); if (MASKED) Al[i].clear(); }
            nvalid = 0;
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            accB[i] += Bl[i].warp_totals(lane, it);
            accAB[i] += ABl[i].warp_totals(lane, it);
            Bl[i].clear(); ABl[i].clear();
            if (MASKED) { accA[i] += Al[i].warp_totals(lane, it); Al[i].clear(); }
        }
        nvalid_total += nvalid;
        nvalid = 0;
        it = 0;
    };

    for (long long wi = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); wi < n_witems; wi += warps_total) {
        long long item = (wi << 5) + lane;
        const bool active = item < total;
        // only the very last warp item can have idle lanes; it is the last one its warp
        // handles.  Idle lanes shadow a valid item; what they add is discarded below.
        const bool ragged = __any_sync(0xFFFFFFFFu, !active);
        if (ragged) flush();
        lane_live = active;
        if (!active) item = total - 1;
        long long seg = 0, chunk = item;
        if (nseg > 1) { seg = item / chunks_per_row; chunk = item - seg * chunks_per_row; }
        const long long row_begin = seg * seg_min + (seg < seg_extra ? seg : seg_extra);
        const int tail_rows = (int)(seg_min + (seg < seg_extra ? 1 : 0) - (groups << 2));  // 0..4

        const char* q = reinterpret_cast<const char*>(x + row_begin * chunks_per_row + chunk);
        uint32_t prev[4], prev_keep[4];
        {
            const uint4 v = ldg_stream(reinterpret_cast<const uint4*>(q));
            prev[0] = v.x; prev[1] = v.y; prev[2] = v.z; prev[3] = v.w;
            if (MASKED) keep_chunk<KIND, MASK>(prev, mask_lo, mask_hi, prev_keep);
            xform_chunk<KIND, XFORM>(prev);
        }
        q += stride;  // first b-row

        auto load_group = [&](uint4 (&buf)[4]) {
            const char* q1 = q + stride;
            const char* q2 = q1 + stride;
            const char* q3 = q2 + stride;
            buf[0] = ldg_stream(reinterpret_cast<const uint4*>(q));
            buf[1] = ldg_stream(reinterpret_cast<const uint4*>(q1));
            buf[2] = ldg_stream(reinterpret_cast<const uint4*>(q2));
            buf[3] = ldg_stream(reinterpret_cast<const uint4*>(q3));
            q = q3 + stride;
        };

        // fold 4 rows; rows >= nrows are absent (tail group only)
        auto process = [&](const uint4 (&buf)[4], int nrows) {
            uint32_t wb[4][4], wab[4][4], wa[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                uint32_t cur[4] = {buf[r].x, buf[r].y, buf[r].z, buf[r].w};
                uint32_t keep[4];
                if (MASKED) keep_chunk<KIND, MASK>(cur, mask_lo, mask_hi, keep);
                xform_chunk<KIND, XFORM>(cur);
                const bool present = r < nrows;
                if (!MASKED) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        wb[r][i] = present ? cur[i] : 0u;
                        wab[r][i] = wb[r][i] & prev[i];
                    }
                } else {
                    uint32_t v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i] = present ? (keep[i] & prev_keep[i]) : 0u;
                        wa[r][i] = prev[i] & v[i];
                        wb[r][i] = cur[i] & v[i];
                        wab[r][i] = wa[r][i] & wb[r][i];
                        if (present) prev_keep[i] = keep[i];
                    }
                    nvalid += count_chunk<BYTES>(v);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (present) prev[i] = cur[i];
            }
            if constexpr (!K64) {
                uint32_t f[16];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[4 * r + i] = wb[r][i];
                Bl[0].add_chunk(f, it);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[4 * r + i] = wab[r][i];
                ABl[0].add_chunk(f, it);
                if constexpr (MASKED) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int i = 0; i < 4; ++i) f[4 * r + i] = wa[r][i];
                    Al[0].add_chunk(f, it);
                }
            } else {
                // words 0,2 of a chunk are low halves, words 1,3 high halves
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t f[CH];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { f[2 * r] = wb[r][h]; f[2 * r + 1] = wb[r][h + 2]; }
                    Bl[h % NL].add_chunk(f, it);
#pragma unroll
                    for (int r = 0; r < 4; ++r) { f[2 * r] = wab[r][h]; f[2 * r + 1] = wab[r][h + 2]; }
                    ABl[h % NL].add_chunk(f, it);
                    if constexpr (MASKED) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) { f[2 * r] = wa[r][h]; f[2 * r + 1] = wa[r][h + 2]; }
                        Al[h % NL].add_chunk(f, it);
                    }
                }
            }
            ++it;
            if (it == L::kFlushChunks) flush();
        };

        // ping-pong: the next group's four loads are in flight while the current one is folded
        uint4 bufA[4], bufB[4];
        long long g = 0;
        if (groups > 0) load_group(bufA);
        while (g + 2 <= groups) {
            load_group(bufB);
            process(bufA, 4);
            if (g + 2 < groups) load_group(bufA);
            process(bufB, 4);
            g += 2;
        }
        if (g < groups) process(bufA, 4);
        // tail: up to 4 remaining rows; absent rows re-read an in-bounds row and are zeroed
        if (__any_sync(0xFFFFFFFFu, tail_rows > 0)) {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                const long long rr = (r < tail_rows) ? r : (long long)tail_rows - 1;  // -1: last full-group row
                bufA[r] = ldg_stream(reinterpret_cast<const uint4*>(q + rr * stride));
            }
            process(bufA, tail_rows);
        }
        if (ragged) {
            flush();
            lane_live = true;
        }
    }
    flush();

    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_n;
    for (int i = threadIdx.x; i < 3 * 64; i += kThreads) (&s_sum[0][0])[i] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int j = K64 ? (i * 32 + (int)lane) : ((int)lane % (8 * BYTES));
        // unmasked: the a-stream is the b-stream shifted up one row; the launcher adds the
        // first row of the matrix and subtracts the last one (row-plane passes)
        const unsigned long long a = MASKED ? accA[i] : accB[i];
        if (a) atomicAdd(&s_sum[S_A][j], a);
        if (accB[i]) atomicAdd(&s_sum[S_B][j], accB[i]);
        if (accAB[i]) atomicAdd(&s_sum[S_AB][j], accAB[i]);
    }
    if (MASKED && nvalid_total) atomicAdd(&s_n, nvalid_total);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 64; i += kThreads) {
        const unsigned long long v = (&s_sum[0][0])[i];
        if (v) atomicAdd(&ws->plus.sums[0][0] + i, v);
    }
    if (MASKED && threadIdx.x == 0 && s_n) atomicAdd(&ws->plus.npairs, s_n);
    if (!MASKED && blockIdx.x == 0 && threadIdx.x == 0) {
        constexpr long long kElemsPerChunk = 16 / BYTES;
        atomicAdd(&ws->plus.npairs, (unsigned long long)((nseg * seg_min + seg_extra) * chunks_per_row * kElemsPerChunk));
    }
}

// ---------------------------------------------------------------------------------
// TMA variant of the column walkers.  Same work split (a warp item = one row segment x one
// 512-byte column strip, lane = 16-byte column chunk, previous row carried in registers),
// but the rows arrive through a per-warp ring of 6 shared-memory tiles of 4 rows x 512 B
// filled by 2-D TMA loads that the warp's own lane 0 issues five tiles ahead.  No producer
// warp and no block-wide barrier: each warp is its own software pipeline, and the loads in
// flight (16 warps x 5 tiles x 2 KB = 160 KB per SM) no longer live in registers.
// ---------------------------------------------------------------------------------
constexpr int kTmaThreads = 512;
constexpr int kTmaWarps = kTmaThreads / 32;
constexpr int kRing = 6;
constexpr int kTileRows = 4;
constexpr int kStripBytes = 512;
constexpr int kTileBytes = kTileRows * kStripBytes;
constexpr size_t kColsSmem = 128 + (size_t)kTmaWarps * kRing * kTileBytes + kTmaWarps * kRing * 8;

template <int KIND, int MASK, bool XFORM>
__global__ void __launch_bounds__(kTmaThreads, 1)
k_pairs_cols_tma(const __grid_constant__ CUtensorMap tmap, const uint4* __restrict__ x, long long chunks_per_row,
                 long long strips, long long nseg, long long seg_min, long long seg_extra, uint32_t mask_lo,
                 uint32_t mask_hi, Workspace* __restrict__ ws) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    constexpr bool K64 = (BYTES == 8);
    constexpr bool MASKED = (MASK != kMaskNone);
    constexpr int NL = K64 ? 2 : 1;
    constexpr int LOGC = K64 ? 3 : 4;
    constexpr int CH = 1 << LOGC;
    constexpr int kNumLadders = NL * (MASKED ? 3 : 2);
    using L = Ladder<3, (kNumLadders >= 6 ? 4 : kNumLadders >= 4 ? 5 : 8), LOGC>;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_addr(smem_raw) + 127u) & ~127u;
    // broadcast from lane 0: tells the compiler the value is warp-uniform (uniform branches, no
    // reconvergence bookkeeping around the ladder pushes)
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t ring0 = smem0 + warp * (kRing * kTileBytes);
    const uint32_t bar0 = smem0 + kTmaWarps * kRing * kTileBytes + warp * (kRing * 8);
    if (lane == 0) {
        for (int i = 0; i < kRing; ++i) mbar_init(bar0 + 8 * i, 1);
        mbar_fence_init();
    }
    if (threadIdx.x == 0) tma_prefetch_descriptor(&tmap);
    __syncthreads();

    const long long n_witems = strips * nseg;
    const long long warps_total = (long long)gridDim.x * kTmaWarps;

    L Bl[NL], ABl[NL], Al[MASKED ? NL : 1];
#pragma unroll
    for (int i = 0; i < NL; ++i) { Bl[i].clear(); ABl[i].clear(); }
    if (MASKED) {
#pragma unroll
        for (int i = 0; i < NL; ++i) Al[i].clear();
    }
    unsigned long long accA[NL], accB[NL], accAB[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { accA[i] = 0; accB[i] = 0; accAB[i] = 0; }
    uint32_t nvalid = 0;
    unsigned long long nvalid_total = 0;
    uint32_t it = 0;
    uint32_t t_issued = 0, t_consumed = 0;  // tiles of this warp since kernel start (ring position / phase)
    bool lane_live = true;  // false while this lane shadows a column chunk beyond the row (ragged strip)

    auto flush = [&]() {
        if (!lane_live) {  // whatever an idle lane gathered since the last flush is discarded
#pragma unroll
            for (int i = 0; i < NL; ++i) { Bl[i].clear(); ABl[i].clear(); if (MASKED) Al[i].clear(); }
            nvalid = 0;
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            accB[i] += Bl[i].warp_totals(lane, it);
            accAB[i] += ABl[i].warp_totals(lane, it);
            Bl[i].clear(); ABl[i].clear();
            if (MASKED) { accA[i] += Al[i].warp_totals(lane, it); Al[i].clear(); }
        }
        nvalid_total += nvalid;
        nvalid = 0;
        it = 0;
    };

    for (long long wi = (long long)blockIdx.x * kTmaWarps + warp; wi < n_witems; wi += warps_total) {
        const long long seg = wi / strips;
        const long long strip = wi - seg * strips;
        const long long chunk = strip * 32 + lane;
        const bool active = chunk < chunks_per_row;
        const bool ragged = __any_sync(0xFFFFFFFFu, !active);  // last strip of a row only
        if (ragged) flush();  // bank what the previous items gathered before lanes go idle
        lane_live = active;
        const long long row_begin = seg * seg_min + (seg < seg_extra ? seg : seg_extra);
        const long long seg_len = seg_min + (seg < seg_extra ? 1 : 0);
        const long long ntiles = (seg_len + kTileRows - 1) / kTileRows;

        uint32_t prev[4] = {0u, 0u, 0u, 0u}, prev_keep[4] = {0u, 0u, 0u, 0u};
        if (active) {
            const uint4 v = ldg_stream(x + row_begin * chunks_per_row + chunk);
            prev[0] = v.x; prev[1] = v.y; prev[2] = v.z; prev[3] = v.w;
            if (MASKED) keep_chunk<KIND, MASK>(prev, mask_lo, mask_hi, prev_keep);
            xform_chunk<KIND, XFORM>(prev);
        }

        auto issue_tile = [&](long long k) {
            if (lane == 0) {
                const uint32_t slot = t_issued % kRing;
                // the slot was last read (LDS) by this warp before the __syncwarp that ended that
                // iteration, and those values have been consumed: safe to overwrite
                mbar_arrive_expect_tx(bar0 + 8 * slot, kTileBytes);
                tma_load_2d(ring0 + slot * kTileBytes, &tmap, (int32_t)(strip * (kStripBytes / 4)),
                            (int32_t)(row_begin + 1 + k * kTileRows), bar0 + 8 * slot);
            }
            ++t_issued;
        };

        auto process = [&](const uint4 (&buf)[4], int nrows) {
            uint32_t wb[4][4], wab[4][4], wa[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                uint32_t cur[4] = {buf[r].x, buf[r].y, buf[r].z, buf[r].w};
                uint32_t keep[4];
                if (MASKED) keep_chunk<KIND, MASK>(cur, mask_lo, mask_hi, keep);
                xform_chunk<KIND, XFORM>(cur);
                const bool present = r < nrows;
                if (!MASKED) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        wb[r][i] = present ? cur[i] : 0u;
                        wab[r][i] = wb[r][i] & prev[i];
                    }
                } else {
                    uint32_t v[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        v[i] = present ? (keep[i] & prev_keep[i]) : 0u;
                        wa[r][i] = prev[i] & v[i];
                        wb[r][i] = cur[i] & v[i];
                        wab[r][i] = wa[r][i] & wb[r][i];
                        if (present) prev_keep[i] = keep[i];
                    }
                    nvalid += count_chunk<BYTES>(v);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i)
                    if (present) prev[i] = cur[i];
            }
            if constexpr (!K64) {
                uint32_t f[16];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[4 * r + i] = wb[r][i];
                Bl[0].add_chunk(f, it);
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[4 * r + i] = wab[r][i];
                ABl[0].add_chunk(f, it);
                if constexpr (MASKED) {
#pragma unroll
                    for (int r = 0; r < 4; ++r)
#pragma unroll
                        for (int i = 0; i < 4; ++i) f[4 * r + i] = wa[r][i];
                    Al[0].add_chunk(f, it);
                }
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t f[CH];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { f[2 * r] = wb[r][h]; f[2 * r + 1] = wb[r][h + 2]; }
                    Bl[h % NL].add_chunk(f, it);
#pragma unroll
                    for (int r = 0; r < 4; ++r) { f[2 * r] = wab[r][h]; f[2 * r + 1] = wab[r][h + 2]; }
                    ABl[h % NL].add_chunk(f, it);
                    if constexpr (MASKED) {
#pragma unroll
                        for (int r = 0; r < 4; ++r) { f[2 * r] = wa[r][h]; f[2 * r + 1] = wa[r][h + 2]; }
                        Al[h % NL].add_chunk(f, it);
                    }
                }
            }
            ++it;
            if (it == L::kFlushChunks) flush();
        };

        const long long pre = ntiles < (kRing - 1) ? ntiles : (kRing - 1);
        for (long long k = 0; k < pre; ++k) issue_tile(k);
        for (long long k = 0; k < ntiles; ++k) {
            if (k + (kRing - 1) < ntiles) issue_tile(k + (kRing - 1));
            const uint32_t slot = t_consumed % kRing;
            mbar_wait(bar0 + 8 * slot, (t_consumed / kRing) & 1u);
            const uint32_t base = ring0 + slot * kTileBytes + lane * 16;
            uint4 buf[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) buf[r] = lds128(base + r * kStripBytes);
            if (k + 1 < ntiles) {
                process(buf, 4);
            } else {
                process(buf, (int)(seg_len - k * kTileRows));
            }
            ++t_consumed;
            __syncwarp();
        }
        if (ragged) {
            flush();
            lane_live = true;
        }
    }
    flush();

    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_n;
    for (int i = threadIdx.x; i < 3 * 64; i += kTmaThreads) (&s_sum[0][0])[i] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int j = K64 ? (i * 32 + (int)lane) : ((int)lane % (8 * BYTES));
        const unsigned long long a = MASKED ? accA[i] : accB[i];  // unmasked: corrected by the row-plane passes
        if (a) atomicAdd(&s_sum[S_A][j], a);
        if (accB[i]) atomicAdd(&s_sum[S_B][j], accB[i]);
        if (accAB[i]) atomicAdd(&s_sum[S_AB][j], accAB[i]);
    }
    if (MASKED && nvalid_total) atomicAdd(&s_n, nvalid_total);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 64; i += kTmaThreads) {
        const unsigned long long v = (&s_sum[0][0])[i];
        if (v) atomicAdd(&ws->plus.sums[0][0] + i, v);
    }
    if (MASKED && threadIdx.x == 0 && s_n) atomicAdd(&ws->plus.npairs, s_n);
    if (!MASKED && blockIdx.x == 0 && threadIdx.x == 0) {
        constexpr long long kElemsPerChunk = 16 / BYTES;
        atomicAdd(&ws->plus.npairs, (unsigned long long)((nseg * seg_min + seg_extra) * chunks_per_row * kElemsPerChunk));
    }
}

// returns false when the TMA variant cannot be used (no encoder, extents beyond 32-bit coordinates)
template <int KIND, int MASK, bool XFORM>
bool try_launch_cols_tma(const CountJob& job, Workspace* ws, cudaStream_t stream, int64_t chunks, int64_t rows,
                         int64_t pair_rows, int sms, cudaError_t* err) {
    *err = cudaSuccess;
    EncodeFn encode = get_encode_fn();
    const int64_t row_words = chunks * 4;
    if (!encode || rows >= (1ll << 31) || row_words >= (1ll << 31)) return false;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {(cuuint64_t)row_words, (cuuint64_t)rows};
    const cuuint64_t gstride[1] = {(cuuint64_t)row_words * 4};
    const cuuint32_t box[2] = {kStripBytes / 4, kTileRows};
    const cuuint32_t estride[2] = {1, 1};
    if (encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<void*>(job.x), gdim, gstride, box, estride,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    auto kern = k_pairs_cols_tma<KIND, MASK, XFORM>;
    int dev = 0;
    cudaGetDevice(&dev);
    static bool attr_set[64] = {};
    if (dev < 0 || dev >= 64 || !attr_set[dev]) {
        *err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kColsSmem);
        if (*err != cudaSuccess) return true;
        if (dev >= 0 && dev < 64) attr_set[dev] = true;
    }
    const int64_t strips = (chunks + 31) / 32;
    // a few items per warp for balance, segments of at least 64 rows
    const int64_t want_items = (int64_t)sms * kTmaWarps * 4;
    int64_t nseg = (want_items + strips - 1) / strips;
    const int64_t max_seg = pair_rows / 64 > 0 ? pair_rows / 64 : 1;
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    const int64_t seg_min = pair_rows / nseg;
    const int64_t seg_extra = pair_rows - seg_min * nseg;
    int64_t blocks = (strips * nseg + kTmaWarps - 1) / kTmaWarps;
    if (blocks > sms) blocks = sms;
    kern<<<(unsigned)blocks, kTmaThreads, kColsSmem, stream>>>(tmap, static_cast<const uint4*>(job.x), (long long)chunks,
                                                                (long long)strips, (long long)nseg, (long long)seg_min,
                                                                (long long)seg_extra, (uint32_t)job.mask_bits,
                                                                (uint32_t)(job.mask_bits >> 32), ws);
    *err = cudaGetLastError();
    return true;
}

// ---------------------------------------------------------------------------------
// Block-boundary pairs for inner > 1: (last row of block o, first row of block o+1) for
// o < outer-1.  The flat formulation counts them, the reference does not; they are two
// adjacent rows in memory, read here with the same 128-bit column chunks and subtracted
// through the `minus` sums.  A work item is (group of 4 boundaries, column chunk); all three
// streams are formed explicitly, so absent boundaries / idle lanes simply contribute zeros.
// ---------------------------------------------------------------------------------
template <int KIND, int MASK, bool XFORM>
__global__ void __launch_bounds__(kThreads)
k_block_edges(const uint4* __restrict__ x, long long chunks_per_row, long long n, long long nbound, uint32_t mask_lo,
              uint32_t mask_hi, RawSums* __restrict__ dst) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    constexpr bool K64 = (BYTES == 8);
    constexpr bool MASKED = (MASK != kMaskNone);
    constexpr int NL = K64 ? 2 : 1;
    constexpr int LOGC = K64 ? 3 : 4;
    constexpr int CH = 1 << LOGC;
    using L = Ladder<3, 4, LOGC>;
    const uint32_t lane = threadIdx.x & 31u;
    const long long groups = (nbound + 3) >> 2;
    const long long total = groups * chunks_per_row;
    const long long n_witems = (total + 31) >> 5;
    const long long warps_total = (long long)gridDim.x * (kThreads / 32);

    L Al[NL], Bl[NL], ABl[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { Al[i].clear(); Bl[i].clear(); ABl[i].clear(); }
    unsigned long long accA[NL], accB[NL], accAB[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { accA[i] = 0; accB[i] = 0; accAB[i] = 0; }
    uint32_t nvalid = 0, it = 0;
    unsigned long long nvalid_total = 0;
    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            accA[i] += Al[i].warp_totals(lane, it);
            accB[i] += Bl[i].warp_totals(lane, it);
            accAB[i] += ABl[i].warp_totals(lane, it);
            Al[i].clear(); Bl[i].clear(); ABl[i].clear();
        }
        nvalid_total += nvalid;
        nvalid = 0;
        it = 0;
    };

    for (long long wi = (long long)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5); wi < n_witems; wi += warps_total) {
        const long long item = (wi << 5) + lane;
        const bool live = item < total;
        const long long g = live ? item / chunks_per_row : 0;
        const long long chunk = live ? item - g * chunks_per_row : 0;
        uint4 va[4], vb[4];
        bool present[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const long long o = 4 * g + q;
            present[q] = live && o < nbound;
            const long long row = ((present[q] ? o : 0) + 1) * n - 1;  // last row of block o
            const uint4* pa = x + row * chunks_per_row + chunk;
            va[q] = ldg_stream(pa);
            vb[q] = ldg_stream(pa + chunks_per_row);
        }
        uint32_t wa[4][4], wb[4][4], wab[4][4];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            uint32_t a[4] = {va[q].x, va[q].y, va[q].z, va[q].w};
            uint32_t b[4] = {vb[q].x, vb[q].y, vb[q].z, vb[q].w};
            uint32_t v[4];
            if (MASKED) {
                uint32_t ka[4], kb[4];
                keep_chunk<KIND, MASK>(a, mask_lo, mask_hi, ka);
                keep_chunk<KIND, MASK>(b, mask_lo, mask_hi, kb);
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = present[q] ? (ka[i] & kb[i]) : 0u;
                nvalid += count_chunk<BYTES>(v);
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i) v[i] = present[q] ? 0xFFFFFFFFu : 0u;
            }
            xform_chunk<KIND, XFORM>(a);
            xform_chunk<KIND, XFORM>(b);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                wa[q][i] = a[i] & v[i];
                wb[q][i] = b[i] & v[i];
                wab[q][i] = wa[q][i] & wb[q][i];
            }
        }
        if constexpr (!K64) {
            uint32_t f[16];
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) f[4 * q + i] = wa[q][i];
            Al[0].add_chunk(f, it);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) f[4 * q + i] = wb[q][i];
            Bl[0].add_chunk(f, it);
#pragma unroll
            for (int q = 0; q < 4; ++q)
#pragma unroll
                for (int i = 0; i < 4; ++i) f[4 * q + i] = wab[q][i];
            ABl[0].add_chunk(f, it);
        } else {
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t f[CH];
#pragma unroll
                for (int q = 0; q < 4; ++q) { f[2 * q] = wa[q][h]; f[2 * q + 1] = wa[q][h + 2]; }
                Al[h % NL].add_chunk(f, it);
#pragma unroll
                for (int q = 0; q < 4; ++q) { f[2 * q] = wb[q][h]; f[2 * q + 1] = wb[q][h + 2]; }
                Bl[h % NL].add_chunk(f, it);
#pragma unroll
                for (int q = 0; q < 4; ++q) { f[2 * q] = wab[q][h]; f[2 * q + 1] = wab[q][h + 2]; }
                ABl[h % NL].add_chunk(f, it);
            }
        }
        ++it;
        if (it == L::kFlushChunks) flush();
    }
    flush();

    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_n;
    for (int i = threadIdx.x; i < 3 * 64; i += kThreads) (&s_sum[0][0])[i] = 0;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int j = K64 ? (i * 32 + (int)lane) : ((int)lane % (8 * BYTES));
        if (accA[i]) atomicAdd(&s_sum[S_A][j], accA[i]);
        if (accB[i]) atomicAdd(&s_sum[S_B][j], accB[i]);
        if (accAB[i]) atomicAdd(&s_sum[S_AB][j], accAB[i]);
    }
    if (MASKED && nvalid_total) atomicAdd(&s_n, nvalid_total);
    __syncthreads();
    for (int i = threadIdx.x; i < 3 * 64; i += kThreads) {
        const unsigned long long v = (&s_sum[0][0])[i];
        if (v) atomicAdd(&dst->sums[0][0] + i, v);
    }
    if (MASKED && threadIdx.x == 0 && s_n) atomicAdd(&dst->npairs, s_n);
    if (!MASKED && blockIdx.x == 0 && threadIdx.x == 0) {
        constexpr long long kElemsPerChunk = 16 / BYTES;
        atomicAdd(&dst->npairs, (unsigned long long)(nbound * chunks_per_row * kElemsPerChunk));
    }
}

template <int KIND, int MASK, bool XFORM>
cudaError_t launch_cols(const CountJob& job, Workspace* ws, cudaStream_t stream, int64_t* done_lo, int64_t* done_hi,
                        bool* edges_done) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    const int64_t row_bytes = job.inner * BYTES;
    if (row_bytes % 16 != 0 || reinterpret_cast<uintptr_t>(job.x) % 16 != 0) return cudaSuccess;
    const int64_t rows = job.outer * job.n;
    const int64_t pair_rows = rows - 1;
    if (pair_rows < 1) return cudaSuccess;
    const int64_t chunks = row_bytes / 16;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaError_t terr = cudaSuccess;
    // Measured on B200 (same box, kernel alone): with a power-of-two row pitch the TMA ring
    // runs at 3.7-3.9 TB/s where the register-staged variant reaches 5.4-5.5 TB/s; with other
    // pitches the TMA ring wins (5.5-5.8 vs 4.9-5.1 TB/s).  XBI_COLS_LDG / XBI_COLS_TMA force one.
    const bool pow2_pitch = (row_bytes & (row_bytes - 1)) == 0;
    const bool use_ldg = std::getenv("XBI_COLS_LDG") != nullptr || (pow2_pitch && std::getenv("XBI_COLS_TMA") == nullptr);
    if (!use_ldg && try_launch_cols_tma<KIND, MASK, XFORM>(job, ws, stream, chunks, rows, pair_rows, sms, &terr)) {
        if (terr != cudaSuccess) return terr;
    } else {
    // enough threads to fill the machine about twice, but segments of at least 64 rows
    const int64_t want_threads = (int64_t)sms * 2048 * 2;
    int64_t nseg = (want_threads + chunks - 1) / chunks;
    const int64_t max_seg = pair_rows / 64 > 0 ? pair_rows / 64 : 1;
    if (nseg > max_seg) nseg = max_seg;
    if (nseg < 1) nseg = 1;
    const int64_t seg_min = pair_rows / nseg;
    const int64_t seg_extra = pair_rows - seg_min * nseg;  // this many segments get one more row
    const int64_t total = chunks * nseg;
    int64_t blocks = (total + kThreads - 1) / kThreads;
    const int64_t cap = (int64_t)sms * 16;  // a few waves; warps loop over the remaining items
    if (blocks > cap) blocks = cap;
    k_pairs_cols<KIND, MASK, XFORM><<<(unsigned)blocks, kThreads, 0, stream>>>(
        static_cast<const uint4*>(job.x), (long long)chunks, (long long)nseg, (long long)seg_min, (long long)seg_extra,
        (uint32_t)job.mask_bits, (uint32_t)(job.mask_bits >> 32), ws);
    }
    if (MASK == kMaskNone) {
        // a-stream correction: + planes(first row) - planes(row after the last pair row)
        cudaError_t e = launch_elem_planes(job, 0, job.inner, &ws->plus, stream);
        if (e != cudaSuccess) return e;
        e = launch_elem_planes(job, pair_rows * job.inner, job.inner, &ws->minus, stream);
        if (e != cudaSuccess) return e;
    }
    if (job.outer > 1) {
        const int64_t nbound = job.outer - 1;
        const int64_t items = ((nbound + 3) / 4) * chunks;
        // the end-of-kernel flush costs a warp ~10 items' worth of instructions: give every
        // warp at least 16 items before adding more blocks
        int64_t eblocks = (items + (int64_t)kThreads * 16 - 1) / ((int64_t)kThreads * 16);
        if (eblocks > (int64_t)sms * 8) eblocks = (int64_t)sms * 8;
        if (eblocks < 1) eblocks = 1;
        k_block_edges<KIND, MASK, XFORM><<<(unsigned)eblocks, kThreads, 0, stream>>>(
            static_cast<const uint4*>(job.x), (long long)chunks, (long long)job.n, (long long)nbound,
            (uint32_t)job.mask_bits, (uint32_t)(job.mask_bits >> 32), &ws->minus);
        note_launch();
    }
    *edges_done = true;
    note_fast_launch();
    *done_lo = 0;
    *done_hi = pair_rows * job.inner;  // every flat pair
    return cudaGetLastError();
}

template <int KIND>
cudaError_t dispatch_cols(const CountJob& job, Workspace* ws, cudaStream_t stream, int64_t* lo, int64_t* hi, bool* ed) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    if (job.transform) {
        if constexpr (isf) {
            switch (job.mask_mode) {
                case kMaskNone: return launch_cols<KIND, kMaskNone, true>(job, ws, stream, lo, hi, ed);
                case kMaskNaN: return launch_cols<KIND, kMaskNaN, true>(job, ws, stream, lo, hi, ed);
                default: return launch_cols<KIND, kMaskValue, true>(job, ws, stream, lo, hi, ed);
            }
        }
        return cudaSuccess;
    }
    if (job.mask_mode == kMaskNaN) return cudaSuccess;  // stays on the generic kernel
    constexpr int UK = (KIND == kF16) ? kU16 : (KIND == kF32) ? kU32 : (KIND == kF64) ? kU64 : KIND;
    if (job.mask_mode == kMaskNone) return launch_cols<UK, kMaskNone, false>(job, ws, stream, lo, hi, ed);
    return launch_cols<UK, kMaskValue, false>(job, ws, stream, lo, hi, ed);
}

}  // namespace

cudaError_t launch_pairs_cols(const CountJob& job, Workspace* ws, cudaStream_t stream, int64_t* done_lo,
                              int64_t* done_hi, bool* edges_done) {
    *done_lo = 0;
    *done_hi = 0;
    *edges_done = false;
    switch (job.dtype) {
        case XBI_U8: return dispatch_cols<kU8>(job, ws, stream, done_lo, done_hi, edges_done);
        case XBI_U16: return dispatch_cols<kU16>(job, ws, stream, done_lo, done_hi, edges_done);
        case XBI_U32: return dispatch_cols<kU32>(job, ws, stream, done_lo, done_hi, edges_done);
        case XBI_U64: return dispatch_cols<kU64>(job, ws, stream, done_lo, done_hi, edges_done);
        case XBI_F16: return dispatch_cols<kF16>(job, ws, stream, done_lo, done_hi, edges_done);
        case XBI_F32: return dispatch_cols<kF32>(job, ws, stream, done_lo, done_hi, edges_done);
        case XBI_F64: return dispatch_cols<kF64>(job, ws, stream, done_lo, done_hi, edges_done);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace xbi
