// K1 over a chunk grid, fast formulation (docs/chunking.ipynb cells 8-10: get_bitinformation on every dask block on
// its own).  The first chunk-grid kernel (k_pairs_chunked, xbi_chunked.cu) is the generic formulation batched: every
// thread steps a mixed-radix odometer per ELEMENT and loads both elements of every pair -- 0.9-1.25 TB/s.  Here a
// lane owns a COLUMN of a chunk: fixed coordinates in every dimension but the analysed one, worked out once per
// column (divisions by the chunk's extents), then one address add per row.  The columns of a chunk, flattened over
// all its other dimensions, are cut into tiles of 32 (one per warp); a lane walks down the analysed axis 8 rows per
// fold with the previous row in registers: ONE load per element.  When the analysed axis is not the innermost
// dimension, neighbouring lanes are neighbouring elements in memory (coalesced rows); when it is, they sit on
// neighbouring runs and each lane uses up its own 32-byte sectors as it goes.
//
// 4- and 8-byte element types do not load into registers (16 scalar loads per lane = 2 KB per warp in flight: the
// kernel was latency-bound at 1.3-1.5 TB/s) but through a per-warp shared-memory ring filled with element-sized
// cp.async copies seven 8-row steps ahead of the fold (RING): 7 KB per warp in flight, no registers tied up; a lane
// reads back only what it copied, so there is no barrier at all.
//
// Three explicit word streams a, b, a&b of the valid pairs go into the bit-sliced counters of xbi_common.cuh.  A
// CTA works through a CONTIGUOUS range of tiles dealt to its warps in turn, so a warp changes chunk a few times in
// its whole life and only then (and at the end) pays the butterfly flush and the atomics into that chunk's table.
#include <atomic>
#include <type_traits>

#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"
#include "xbi_mask.cuh"

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

// Masked modes, 4- and 8-byte types: a sampling kernel leaves (#pairs whose elements differ in validity, #pairs
// sampled) in two scratch words of chunk 0's table (slot 0 of bits 1 and 2: unused until k_chunk_finalize overwrites
// them); both walker kernels are launched and the one that does not suit the data returns at once.  More than one
// boundary in kWalkOneIn pairs: masked elements are scattered, nearly every step of the two-stream walkers would
// take the exact per-element path -- the three explicit streams are cheaper then.
constexpr unsigned long long kWalkOneIn = 128;
constexpr int kDensBnd = 4, kDensPairs = 8;  // word offsets into chunk 0's table
__device__ __forceinline__ bool walk_density_is_scattered(const unsigned long long* dens) {
    const volatile unsigned long long* p = dens;
    return p[kDensBnd] * kWalkOneIn > p[kDensPairs];
}

constexpr int kLogR = 3;
constexpr int kRows = 1 << kLogR;    // rows (column walkers) / 32-element steps (runs) folded at a time: 16 needed
                                     // ~30 registers more than the 128 a thread of a 2-CTA/SM kernel may have (spills)
constexpr int kSlots = 8;            // ring of kRows-row steps per warp
constexpr int kAheadW = kSlots - 1;  // steps in flight

template <int BYTES>
__device__ __forceinline__ void cpa_elem(uint32_t dst, const void* src) {
    if (BYTES == 8) asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
    else asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cpa_commit_w() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait_w() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
template <int NW>
__device__ __forceinline__ void lds_elem_w(uint32_t addr, uint32_t (&w)[NW]) {
    if constexpr (NW == 2) {
        asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(w[0]), "=r"(w[1]) : "r"(addr));
    } else {
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(w[0]) : "r"(addr));
    }
}
__device__ __forceinline__ uint32_t smem_addr_w(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

template <int KIND, int MASK, bool XFORM, bool RING, bool LASTAX>
__global__ void __launch_bounds__(kWThreads, Elem<KIND>::kWords == 1 ? 2 : 1)
k_pairs_chunked_walk(const void* __restrict__ x, const __grid_constant__ ChunkGrid g, uint32_t mask_lo, uint32_t mask_hi,
                     unsigned long long* __restrict__ counts, unsigned long long items, unsigned groups,
                     const unsigned long long* dens, unsigned cblock) {
    if (dens != nullptr && !walk_density_is_scattered(dens)) return;  // the two-stream walkers take this array
    constexpr int BYTES = Elem<KIND>::kBytes;
    static_assert(!RING || BYTES >= 4, "cp.async moves 4 bytes at least");
    // the warp's ring, kSlots steps of kRows rows x 32 lanes.  Row-major ([row][lane], element-sized copies, coalesced
    // rows) when the analysed axis is not the innermost dimension; LASTAX: lane-major ([lane][row]) so that a lane
    // whose run starts on a 16-byte boundary copies 16 bytes (kRows * BYTES / 16 copies per step instead of kRows)
    // and reads them back with 128-bit loads -- lane pitch 48 / 80 bytes keeps those conflict-free
    constexpr uint32_t kLanePitch = LASTAX ? (uint32_t)(kRows * BYTES + 16) : (uint32_t)BYTES;
    constexpr uint32_t kRowPitch = LASTAX ? (uint32_t)BYTES : 32u * (uint32_t)BYTES;
    constexpr uint32_t kSlotBytes = LASTAX ? 32u * kLanePitch : (uint32_t)kRows * 32u * (uint32_t)BYTES;
    extern __shared__ uint8_t smem_ring[];
    const uint32_t ring0 = RING ? ((smem_addr_w(smem_ring) + 15u) & ~15u) + (threadIdx.x >> 5) * (kSlots * kSlotBytes) : 0u;
    uint32_t slot_w = 0, slot_r = 0;
    using E = Elem<KIND>;
    constexpr int NW = E::kWords;
    constexpr int NBITS = E::kBits;
    using L = Ladder<3, 8, kLogR>;
    uint32_t lane = threadIdx.x & 31u;
    uint32_t warp = threadIdx.x >> 5;
    // (values the hot loops use are made opaque below: under register pressure ptxas otherwise re-derives them in
    // every step -- S2R for the lane, a select chain over the dimensions for the extent along the axis ...)
    asm volatile("" : "+r"(lane), "+r"(warp));
    // An item is one column tile / one block of runs of one chunk.  A CTA works through a CONTIGUOUS range of items
    // and deals them to its warps in turn: a warp changes chunk a few times in its life, and neighbouring tiles
    // are walked at the same time by neighbouring warps, so lines they share (tiles do not start on 128-byte
    // boundaries) come from DRAM once
    const unsigned long long item_lo = (unsigned long long)blockIdx.x * items / gridDim.x;
    const unsigned long long item_hi = (unsigned long long)(blockIdx.x + 1) * items / gridDim.x;
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
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid_total += __shfl_xor_sync(0xFFFFFFFFu, nvalid_total, o);
        if (lane == 0 && nvalid_total) atomicAdd(dst, nvalid_total);
        nvalid_total = 0;
    };
    // kRows pairs per lane, words already transformed and zeroed where the pair does not count
    auto fold = [&](const uint32_t (&wa)[NW][kRows], const uint32_t (&wb)[NW][kRows]) {
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            uint32_t wab[kRows];
#pragma unroll
            for (int u = 0; u < kRows; ++u) wab[u] = wa[k][u] & wb[k][u];
            A[k].add_chunk(wa[k], it);
            B[k].add_chunk(wb[k], it);
            AB[k].add_chunk(wab, it);
        }
        ++it;
        if (it == L::kFlushChunks) drain();
    };
    // all-ones when the element is not masked
    auto keep_of = [&](const uint32_t (&w)[NW]) -> uint32_t {
        if (MASK == kMaskNaN) return E::is_nan(w) ? 0u : 0xFFFFFFFFu;
        if (MASK == kMaskValue) return E::equals(w, mask_lo, mask_hi) ? 0u : 0xFFFFFFFFu;
        return 0xFFFFFFFFu;
    };

    unsigned cur_chunk = 0xFFFFFFFFu;
    ChunkBox box;
    for (unsigned long long item = item_lo + warp; item < item_hi; item += kWWarps) {
        // (cblock > 1: the items of `cblock` neighbouring chunks interleaved tile by tile, see k_pairs_chunked_walk2)
        unsigned cid, grp;
        if (cblock <= 1u) {
            cid = (unsigned)(item / groups);
            grp = (unsigned)(item - (unsigned long long)cid * groups);
        } else {
            const unsigned long long per = (unsigned long long)cblock * groups;
            const unsigned long long blk = item / per, r = item - blk * per;
            const unsigned long long first = blk * cblock, nchunks = items / groups;
            const unsigned bs = (unsigned)(nchunks - first < cblock ? nchunks - first : cblock);
            grp = (unsigned)(r / bs);
            cid = (unsigned)(first + (r - (unsigned long long)grp * bs));
        }
        if (cid != cur_chunk) {
            if (cur_chunk != 0xFFFFFFFFu) flush_chunk(cur_chunk);
            cur_chunk = cid;
            box.init(g, cid);
        }
        unsigned ea = 1;
#pragma unroll
        for (int d = 0; d < kD; ++d) if (d == axis) ea = box.e[d];
        asm volatile("" : "+r"(ea));
        if (ea < 2) continue;  // no pairs in this chunk

        // column walkers: the chunk's columns (every dimension but the analysed one, flattened) in tiles of 32 lanes.
        // A lane takes EPV neighbouring columns: 1, or -- outer axes, when the chunk's rows are whole 16-byte vectors
        // (`vec_allowed` from the planner, and this chunk's innermost extent a multiple of EPV) -- 16 bytes' worth:
        // 16-byte copies and 128-bit reads, a warp's row 512 contiguous bytes instead of 128.
        unsigned long long O = 1, I = 1;
#pragma unroll
        for (int d = 0; d < kD; ++d) {
            if (d < axis) O *= box.e[d];
            if (d > axis) I *= box.e[d];
        }
        long long step_bytes = astride * BYTES;
        asm volatile("" : "+l"(step_bytes));

        auto walk = [&](auto epv_tag) {
            constexpr int EPV = decltype(epv_tag)::value;   // columns per lane
            constexpr int RS = kRows / EPV;                  // rows per fold
            constexpr bool V16 = EPV * BYTES == 16 && !LASTAX;
            constexpr uint32_t kCell = LASTAX ? kLanePitch : (uint32_t)(EPV * BYTES);           // a lane's cell
            constexpr uint32_t kRowP = LASTAX ? (uint32_t)BYTES : 32u * (uint32_t)(EPV * BYTES);  // a row of the step
            const unsigned long long ncols = O * I / EPV, tiles = (ncols + 31) / 32;
            const unsigned long long t_lo = tiles * grp / groups, t_hi = tiles * (grp + 1) / groups;
            for (unsigned long long tile = t_lo; tile < t_hi; ++tile) {
                const unsigned long long col = tile * 32 + lane;
                unsigned live_u = col < ncols ? 1u : 0u;
                asm volatile("" : "+r"(live_u));  // (else re-derived from the extents in every step)
                const bool live = live_u != 0u;
                // offset of this lane's (first) column at row 0 of the chunk
                long long off = box.base;
                {
                    const unsigned long long cc = (live ? col : 0) * EPV;
                    unsigned long long ro = cc / I, ri = cc - ro * I;
#pragma unroll
                    for (int d = kD - 1; d >= 0; --d) {
                        if (d > axis) { off += (long long)(ri % box.e[d]) * g.stride[d]; ri /= box.e[d]; }
                        if (d < axis) { off += (long long)(ro % box.e[d]) * g.stride[d]; ro /= box.e[d]; }
                    }
                }
                const char* p = static_cast<const char*>(x) + off * BYTES;  // row 0 of the column(s)
                // the walk starts AT row 0 with "no row above" (pk = 0): the first row is a current row like the others
                uint32_t prev[EPV][NW], pk[EPV];  // previous row (transformed), its keep-masks (0: masked / no such row)
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
                    pk[e] = 0u;
#pragma unroll
                    for (int k = 0; k < NW; ++k) prev[e][k] = 0;
                }
                unsigned nsteps = (ea + RS - 1) / RS;
                asm volatile("" : "+r"(nsteps));
                // LASTAX: 16-byte copies when every lane's run starts on a 16-byte boundary (warp-uniform)
                const bool vec = LASTAX && RING && __all_sync(0xFFFFFFFFu, !live || (reinterpret_cast<uintptr_t>(p) & 15u) == 0u);
                const char* q = p;  // advanced by the copies (RING) or the loads
                unsigned issued = 0;
                // RING: copy step `issued` (RS rows of this lane's columns) into the ring; a lane reads back only its own cells
                auto issue = [&]() {
                    if (issued < nsteps && live) {
                        const uint32_t dst = ring0 + slot_w * kSlotBytes + lane * kCell;
                        const unsigned left = ea - RS * issued;  // rows from this step on
                        if (LASTAX && vec && left >= (unsigned)RS) {
#pragma unroll
                            for (int v = 0; v < kRows * BYTES / 16; ++v)
                                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + 16 * v), "l"(q + 16 * v) : "memory");
                            q += kRows * BYTES;
                        } else {
#pragma unroll
                            for (int u = 0; u < RS; ++u) {
                                if (left >= (unsigned)RS || (unsigned)u < left) {
                                    if constexpr (V16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * kRowP), "l"(q) : "memory");
                                    else cpa_elem<BYTES>(dst + u * kRowP, q);
                                    q += step_bytes;
                                }
                            }
                        }
                    }
                    cpa_commit_w();
                    slot_w = (slot_w + 1 == kSlots) ? 0 : slot_w + 1;
                    ++issued;
                };
                if constexpr (RING) {
#pragma unroll 1
                    for (int k = 0; k < kAheadW; ++k) issue();
                }
#pragma unroll 1
                for (unsigned s = 0; s < nsteps; ++s) {
                    const unsigned left = live ? ea - RS * s : 0u;  // rows this lane has from this step on
                    uint32_t raw[NW][kRows];  // [word][row * EPV + column]
                    if constexpr (RING) {
                        issue();
                        cpa_wait_w<kAheadW>();
                        const uint32_t src = ring0 + slot_r * kSlotBytes + lane * kCell;
                        slot_r = (slot_r + 1 == kSlots) ? 0 : slot_r + 1;
                        if constexpr (LASTAX || V16) {
                            // 128-bit reads: LASTAX -- the lane's kRows elements are contiguous; V16 -- one row's EPV
                            // elements are (absent rows: stale cells, zeroed below)
                            uint32_t flat[kRows * NW];
#pragma unroll
                            for (int v = 0; v < kRows * BYTES / 16; ++v) {
                                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                             : "=r"(flat[4 * v]), "=r"(flat[4 * v + 1]), "=r"(flat[4 * v + 2]), "=r"(flat[4 * v + 3])
                                             : "r"(src + (LASTAX ? 16u * v : (uint32_t)v * kRowP)));
                            }
#pragma unroll
                            for (int u = 0; u < kRows; ++u)
#pragma unroll
                                for (int k = 0; k < NW; ++k) raw[k][u] = flat[u * NW + k];
                        } else {
#pragma unroll
                            for (int u = 0; u < kRows; ++u) {
                                uint32_t w[NW];
                                lds_elem_w<NW>(src + u * kRowP, w);  // (absent rows: stale cells, zeroed below)
#pragma unroll
                                for (int k = 0; k < NW; ++k) raw[k][u] = w[k];
                            }
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kRows; ++u) {
                            uint32_t w[NW];
#pragma unroll
                            for (int k = 0; k < NW; ++k) w[k] = 0;
                            if ((unsigned)u < left) { E::load(q, 0, w); q += step_bytes; }
#pragma unroll
                            for (int k = 0; k < NW; ++k) raw[k][u] = w[k];
                        }
                    }
                    uint32_t wa[NW][kRows], wb[NW][kRows];
                    if (s > 0 && __all_sync(0xFFFFFFFFu, left >= (unsigned)RS)) {
                        // every lane has all RS rows and a row above the first of them (the common case): no
                        // "does the row exist" tests
                        uint32_t vbits = 0;
#pragma unroll
                        for (int u = 0; u < kRows; ++u) {
                            const int e = u % EPV;
                            uint32_t w[NW];
#pragma unroll
                            for (int k = 0; k < NW; ++k) w[k] = raw[k][u];
                            if (MASK != kMaskNone) {
                                const uint32_t kk = keep_of(w);
                                if (XFORM) E::transform(w);
                                const uint32_t m = kk & pk[e];  // the pair (row above, this row) counts
                                vbits |= m & (1u << u);
#pragma unroll
                                for (int k = 0; k < NW; ++k) { wa[k][u] = prev[e][k] & m; wb[k][u] = w[k] & m; }
                                pk[e] = kk;
                            } else {
                                if (XFORM) E::transform(w);
#pragma unroll
                                for (int k = 0; k < NW; ++k) { wa[k][u] = prev[e][k]; wb[k][u] = w[k]; }
                            }
#pragma unroll
                            for (int k = 0; k < NW; ++k) prev[e][k] = w[k];
                        }
                        nvalid += (MASK != kMaskNone) ? (uint32_t)__popc(vbits) : (uint32_t)kRows;
                    } else {
#pragma unroll
                        for (int u = 0; u < kRows; ++u) {
                            const int e = u % EPV;
                            uint32_t w[NW];
#pragma unroll
                            for (int k = 0; k < NW; ++k) w[k] = raw[k][u];
                            uint32_t kk = ((unsigned)(u / EPV) < left) ? 0xFFFFFFFFu : 0u;  // the row exists ...
                            if (MASK != kMaskNone) kk &= keep_of(w);                         // ... and is not masked
                            if (XFORM) E::transform(w);
                            const uint32_t m = kk & pk[e];                                  // the pair (row above, this row) counts
                            nvalid += m & 1u;
#pragma unroll
                            for (int k = 0; k < NW; ++k) { wa[k][u] = prev[e][k] & m; wb[k][u] = w[k] & m; }
                            // (past the end of the column nothing follows: prev may take anything there)
#pragma unroll
                            for (int k = 0; k < NW; ++k) prev[e][k] = w[k];
                            pk[e] = kk;
                        }
                    }
                    fold(wa, wb);
                }
                if constexpr (RING) slot_r = slot_w;  // the kAheadW groups committed past the end of the column were empty
            }
        };
        constexpr int kVecCols = (RING && !LASTAX) ? 16 / BYTES : 1;
        if (kVecCols > 1 && g.vec16 && box.e[kD - 1] % kVecCols == 0) walk(std::integral_constant<int, kVecCols>{});
        else walk(std::integral_constant<int, 1>{});
    }
    if (cur_chunk != 0xFFFFFFFFu) flush_chunk(cur_chunk);
}


// ---------------------------------------------------------------------------------------------------------------
// The walkers for 4- and 8-byte types, second formulation: TWO word streams and corrections at validity boundaries
// (xbi_mask.cuh), 16 elements per lane and step.
//
// With x' = transform(x), zeroed where x is masked, S = sum of x' over every element of a column and rows -1 and
// `ea` of a column taken as masked rows:
//     AB = sum x'[r-1] & x'[r]                                   (no validity needed)
//     A  = S - sum over {r-1 valid, r masked} of x'[r-1]         (includes the column's last row)
//     B  = S - sum over {r-1 masked, r valid} of x'[r]           (includes the column's first row)
// A step whose 16 x 32 elements all exist, follow a valid row and pass the masked-element probe (x * 0 on the FMA
// pipe, may_contain_masked) is the unmasked two-stream step: transform, one AND, two ladders -- about 10 issued
// instructions per element against 28 for the three explicit streams under keep-masks above.  Every other step
// (the first and a ragged last step of a column, steps with masked elements) takes the exact per-element test and
// feeds the correction words, where some lane has one, into small register bit planes (EventPlanes).  Lanes past
// the last column of a tile read the value whose transform is 0 from cells they fill themselves, so partial tiles
// (small chunks) stay on the fast step.
// ---------------------------------------------------------------------------------------------------------------
constexpr int kLogR2 = 4;
constexpr int kRows2 = 1 << kLogR2;  // elements per lane and step
constexpr int kEvP = 6;              // planes of the correction counters: 63 events per lane between flushes

template <bool LASTAX>
struct Ring2 {
    static constexpr int kSlots = 6;
};
template <int BYTES, bool LASTAX>
constexpr size_t ring2_bytes() {
    return 16 + (size_t)kWWarps * Ring2<LASTAX>::kSlots * 32 * (kRows2 * BYTES);
}
// LASTAX: a lane's cell is its 16 elements (64 / 128 bytes) with the 16-byte pieces XOR-swizzled by the lane number, so that
// the 128-bit reads of a quarter warp (8 lanes, same piece) hit 8 different bank groups without padding the cells -- the
// padding (+16 bytes per lane) cost the ring its sixth slot, and with 4 instead of 5 steps in flight the walkers along the
// innermost axis waited on memory (`long_scoreboard` 3.06 warps per issue in ncu)
template <int BYTES>
__device__ __forceinline__ uint32_t ring2_swz16(uint32_t run) {  // 16 x the swizzle term of a run (a lane) of the tile
    return BYTES == 4 ? ((run >> 1) & 3u) << 4 : (run & 7u) << 4;
}

// The correction planes live in LOCAL memory on purpose: they are touched at validity boundaries only, and the dozen
// registers they would occupy are what the hot step needs to stay free of spills.  Handing their address to these
// out-of-line functions is what pins them there.
template <int NW>
static __device__ __noinline__ void ev_add_words(EventPlanes<NW, kEvP>* ev, uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1) {
    ev->ecnt += (a0 | a1 | b0 | b1) ? 1u : 0u;
    ev->template add<0, 0>(a0);
    ev->template add<1, 0>(b0);
    if (NW == 2) {
        ev->template add<0, NW - 1>(a1);
        ev->template add<1, NW - 1>(b1);
    }
}
static __device__ __noinline__ void acc_bump(unsigned long long* p, uint32_t v) { *p += v; }
// warp-collective: planes -> per-lane totals (lane L: bit L) added to acc[stream][word]
template <int NW>
static __device__ __noinline__ void ev_flush_planes(EventPlanes<NW, kEvP>* ev, unsigned long long* acc, uint32_t lane) {
#pragma unroll
    for (int s = 0; s < 2; ++s) {
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            uint32_t n[kEvP + 6];
#pragma unroll
            for (int j = 0; j < kEvP; ++j) { n[j] = ev->r[s][k][j]; ev->r[s][k][j] = 0u; }
#pragma unroll
            for (int j = kEvP; j < kEvP + 6; ++j) n[j] = 0u;
            acc[s * NW + k] += butterfly_totals(n, kEvP, lane);
        }
    }
    ev->ecnt = 0u;
}

template <int KIND, int MASK, bool XFORM, bool LASTAX>
__global__ void __launch_bounds__(kWThreads, Elem<KIND>::kWords == 1 ? 2 : 1)
k_pairs_chunked_walk2(const void* __restrict__ x, const __grid_constant__ ChunkGrid g, uint32_t mask_lo, uint32_t mask_hi,
                      unsigned long long* __restrict__ counts, unsigned long long items, unsigned groups,
                      const unsigned long long* dens, unsigned cblock) {
    if (dens != nullptr && walk_density_is_scattered(dens)) return;  // the three-stream walkers take this array
    using E = Elem<KIND>;
    constexpr int BYTES = E::kBytes;
    constexpr int NW = E::kWords;
    constexpr int NBITS = E::kBits;
    static_assert(BYTES >= 4, "4- and 8-byte element types");
    constexpr int kSlots = Ring2<LASTAX>::kSlots;
    constexpr int kAhead = kSlots - 1;
    constexpr uint32_t kLanePitch = (uint32_t)(kRows2 * BYTES);  // LASTAX: a lane's 16 elements, pieces swizzled (ring2_swz16)
    constexpr uint32_t kSlotBytes = LASTAX ? 32u * kLanePitch : (uint32_t)kRows2 * 32u * (uint32_t)BYTES;
    using L = Ladder<3, 8, kLogR2>;
    using EP = EventPlanes<NW, kEvP>;

    extern __shared__ uint8_t smem_ring[];
    uint32_t lane = threadIdx.x & 31u;
    uint32_t warp = threadIdx.x >> 5;
    uint32_t ring0 = ((smem_addr_w(smem_ring) + 15u) & ~15u) + warp * (kSlots * kSlotBytes);
    // (made opaque: under register pressure ptxas otherwise re-derives these in every step)
    asm volatile("" : "+r"(lane), "+r"(warp), "+r"(ring0));
    uint32_t slot_w = 0, slot_r = 0;
    // LASTAX, cooperative copies: addresses as 32-bit offsets in 16-byte units from `xa` (arrays below 64 GB)
    const char* const xa = reinterpret_cast<const char*>(reinterpret_cast<uintptr_t>(x) & ~static_cast<uintptr_t>(15));
    unsigned long long total_bytes = (unsigned long long)BYTES;
#pragma unroll
    for (int d = 0; d < kD; ++d) total_bytes *= (unsigned long long)g.shape[d];
    const bool offsets_fit = ((total_bytes + 16ull) >> 36) == 0ull;

    const unsigned long long item_lo = (unsigned long long)blockIdx.x * items / gridDim.x;
    const unsigned long long item_hi = (unsigned long long)(blockIdx.x + 1) * items / gridDim.x;
    const int axis = g.axis;
    long long astride = 0;
#pragma unroll
    for (int d = 0; d < kD; ++d) if (d == axis) astride = g.stride[d];

    L S[NW], AB[NW];
    EP ev;
    ev.init();
#pragma unroll
    for (int k = 0; k < NW; ++k) { S[k].clear(); AB[k].clear(); }
    uint32_t it = 0, nvalid = 0;
    // per-lane totals since the last chunk change, in LOCAL memory (pinned there by acc_bump / ev_flush_planes):
    // [0, 2NW) boundary corrections of streams a and b, then S, then AB, then the valid pairs
    unsigned long long acc[4 * NW + 1];
    unsigned long long* const accC = acc, * const accS = acc + 2 * NW, * const accAB = acc + 3 * NW;
#pragma unroll
    for (int k = 0; k < 4 * NW + 1; ++k) acc[k] = 0;

    auto drain = [&]() {  // ladders -> per-lane totals (lane L: bit L)
        if (it == 0) return;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            acc_bump(accS + k, S[k].warp_totals_bounded(lane, it));
            acc_bump(accAB + k, AB[k].warp_totals_bounded(lane, it));
            S[k].clear(); AB[k].clear();
        }
        acc_bump(acc + 4 * NW, nvalid);
        nvalid = 0;
        it = 0;
    };
    auto ev_flush = [&]() { ev_flush_planes<NW>(&ev, accC, lane); };  // warp-collective
    auto ev_reserve = [&](uint32_t upcoming) {
        if (__any_sync(0xFFFFFFFFu, ev.ecnt + upcoming > EP::kCap)) ev_flush();
    };
    auto flush_chunk = [&](unsigned cid) {  // per-lane totals -> the chunk's raw table (see k_chunk_finalize)
        drain();
        if (__any_sync(0xFFFFFFFFu, ev.ecnt != 0u)) ev_flush();
        unsigned long long* dst = counts + (size_t)cid * NBITS * 4;
#pragma unroll
        for (int k = 0; k < NW; ++k) {
            const int j = k * 32 + (int)lane;
            const unsigned long long b = accS[k] - accC[NW + k], a = accS[k] - accC[k];
            if (b) atomicAdd(dst + 4 * (NBITS - 1 - j) + 1, b);
            if (a) atomicAdd(dst + 4 * (NBITS - 1 - j) + 2, a);
            if (accAB[k]) atomicAdd(dst + 4 * (NBITS - 1 - j) + 3, accAB[k]);
            accS[k] = 0; accAB[k] = 0; accC[k] = 0; accC[NW + k] = 0;
        }
        unsigned long long nvalid_total = acc[4 * NW];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) nvalid_total += __shfl_xor_sync(0xFFFFFFFFu, nvalid_total, o);
        if (lane == 0 && nvalid_total) atomicAdd(dst, nvalid_total);
        acc[4 * NW] = 0;
    };
    // all-ones when the element is not masked
    auto keep_of = [&](const uint32_t (&w)[NW]) -> uint32_t {
        if (MASK == kMaskNaN) return E::is_nan(w) ? 0u : 0xFFFFFFFFu;
        if (MASK == kMaskValue) return E::equals(w, mask_lo, mask_hi) ? 0u : 0xFFFFFFFFu;
        return 0xFFFFFFFFu;
    };

    unsigned cur_chunk = 0xFFFFFFFFu;
    ChunkBox box;
    for (unsigned long long item = item_lo + warp; item < item_hi; item += kWWarps) {
        // item -> (chunk, column tile).  Chunks narrower than a tile (cblock > 1): the items of `cblock` neighbouring
        // chunks are interleaved tile by tile, so that the warps of a CTA walk the SAME rows of neighbouring chunks at
        // the same time -- their 40-byte row segments are then contiguous in memory and every line comes from DRAM
        // once (10x10-column chunks read 4.0x their bytes when a CTA walked the tiles of one chunk side by side)
        unsigned cid, grp;
        if (cblock <= 1u) {
            cid = (unsigned)(item / groups);
            grp = (unsigned)(item - (unsigned long long)cid * groups);
        } else {
            const unsigned long long per = (unsigned long long)cblock * groups;
            const unsigned long long blk = item / per, r = item - blk * per;
            const unsigned long long first = blk * cblock, nchunks = items / groups;
            const unsigned bs = (unsigned)(nchunks - first < cblock ? nchunks - first : cblock);
            grp = (unsigned)(r / bs);
            cid = (unsigned)(first + (r - (unsigned long long)grp * bs));
        }
        if (cid != cur_chunk) {
            if (cur_chunk != 0xFFFFFFFFu) flush_chunk(cur_chunk);
            cur_chunk = cid;
            box.init(g, cid);
        }
        unsigned ea = 1;
#pragma unroll
        for (int d = 0; d < kD; ++d) if (d == axis) ea = box.e[d];
        asm volatile("" : "+r"(ea));
        if (ea < 2) continue;  // no pairs in this chunk

        unsigned long long O = 1, I = 1;
#pragma unroll
        for (int d = 0; d < kD; ++d) {
            if (d < axis) O *= box.e[d];
            if (d > axis) I *= box.e[d];
        }
        long long step_bytes = astride * BYTES;
        asm volatile("" : "+l"(step_bytes));

        auto walk = [&](auto epv_tag) {
            constexpr int EPV = decltype(epv_tag)::value;   // columns per lane
            constexpr int RS = kRows2 / EPV;                 // rows per step
            constexpr bool V16 = EPV * BYTES == 16 && !LASTAX;
            constexpr uint32_t kCell = LASTAX ? kLanePitch : (uint32_t)(EPV * BYTES);              // a lane's cell
            constexpr uint32_t kRowP = LASTAX ? (uint32_t)BYTES : 32u * (uint32_t)(EPV * BYTES);   // a row of the step
            const unsigned long long ncols = O * I / EPV, tiles = (ncols + 31) / 32;
            const unsigned long long t_lo = tiles * grp / groups, t_hi = tiles * (grp + 1) / groups;
            for (unsigned long long tile = t_lo; tile < t_hi; ++tile) {
                const unsigned long long col = tile * 32 + lane;
                unsigned live_u = col < ncols ? 1u : 0u;
                asm volatile("" : "+r"(live_u));
                const bool live = live_u != 0u;
                uint32_t cell = ring0 + lane * kCell;
                asm volatile("" : "+r"(cell));
                if (!live) {
                    // no column: the lane's cells hold the value whose transform is 0 (nothing is ever copied there
                    // during this tile; what earlier tiles copied has landed, see the wait at the end of a tile)
                    const uint32_t nw0 = neutral_word<KIND, XFORM>(0), nw1 = neutral_word<KIND, XFORM>(NW - 1);
#pragma unroll 1
                    for (int sl = 0; sl < kSlots; ++sl) {
#pragma unroll
                        for (int u = 0; u < kRows2; ++u) {
                            const uint32_t a = cell + sl * kSlotBytes + (LASTAX ? (uint32_t)(u * BYTES) : (uint32_t)(u / EPV) * kRowP + (uint32_t)((u % EPV) * BYTES));
                            if (NW == 2) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(a), "r"(nw0), "r"(nw1) : "memory");
                            else asm volatile("st.shared.u32 [%0], %1;" ::"r"(a), "r"(nw0) : "memory");
                        }
                    }
                }
                // offset of this lane's (first) column at row 0 of the chunk
                long long off = box.base;
                if ((O * I) >> 32 == 0) {
                    // (the chunk has fewer than 2^32 columns: 32-bit divisions, a fifth of the 64-bit routine's cost --
                    // with short columns the set-up of a tile is a good part of its walk)
                    const unsigned cc = (live ? (unsigned)col : 0u) * (unsigned)EPV, I32 = (unsigned)I;
                    unsigned ro = cc / I32, ri = cc - ro * I32;
#pragma unroll
                    for (int d = kD - 1; d >= 0; --d) {
                        if (d > axis) { off += (long long)(ri % box.e[d]) * g.stride[d]; ri /= box.e[d]; }
                        if (d < axis) { off += (long long)(ro % box.e[d]) * g.stride[d]; ro /= box.e[d]; }
                    }
                } else {
                    const unsigned long long cc = (live ? col : 0) * EPV;
                    unsigned long long ro = cc / I, ri = cc - ro * I;
#pragma unroll
                    for (int d = kD - 1; d >= 0; --d) {
                        if (d > axis) { off += (long long)(ri % box.e[d]) * g.stride[d]; ri /= box.e[d]; }
                        if (d < axis) { off += (long long)(ro % box.e[d]) * g.stride[d]; ro /= box.e[d]; }
                    }
                }
                const char* q = static_cast<const char*>(x) + off * BYTES;  // advanced by the copies
                // previous row: transformed, zero where masked or absent; its keep-masks.  The walk starts at row 0
                // with "row -1 masked"
                uint32_t prev[EPV][NW], pkb = 0u;  // pkb bit e: the row above in column e exists and is not masked
                constexpr uint32_t kAllCols = (1u << EPV) - 1u;
#pragma unroll
                for (int e = 0; e < EPV; ++e) {
#pragma unroll
                    for (int k = 0; k < NW; ++k) prev[e][k] = 0;
                }
                bool pv_all = false;  // warp-uniform: every lane's previous row(s) valid
                unsigned nsteps = (ea + RS - 1) / RS;
                asm volatile("" : "+r"(nsteps));
                const uint32_t nv_fast = live ? (uint32_t)kRows2 : 0u;
                // LASTAX: 16-byte copies when every lane's run starts on a 16-byte boundary (warp-uniform)
                const bool vec = LASTAX && __all_sync(0xFFFFFFFFu, !live || (reinterpret_cast<uintptr_t>(q) & 15u) == 0u);
                // LASTAX with aligned runs: the warp copies COOPERATIVELY.  A lane copying its own run's 16-byte pieces
                // makes every copy instruction touch 32 different lines (one per run): the L1 looks up one line per
                // cycle, which capped the kernel near 2.8 TB/s.  Instead NP neighbouring lanes take the NP pieces of
                // one run's step (64 / 128 contiguous bytes), so an instruction touches 32 / NP lines.  A lane then
                // reads cells other lanes copied: __syncwarp after the wait and before a slot is overwritten.
                constexpr int NP = kRows2 * BYTES / 16;  // 16-byte pieces of one run per step
                constexpr int RPI = 32 / NP;             // runs per copy instruction of the warp
                uint32_t co[LASTAX ? NP : 1];            // (start of run j * RPI + lane / NP, minus xa) / 16 + lane % NP
                uint32_t colive = 0u;                    // bit j: that run exists
                bool coop = false;
                // (cell of run j * RPI + lane / NP, piece lane % NP swizzled by that run's number: float32 -- the term
                // does not depend on j; float64 -- it alternates with the parity of j)
                const uint32_t cbase = ring0 + (lane / NP) * kLanePitch;
                const uint32_t cpiece0 = ((lane % NP) << 4) ^ ring2_swz16<BYTES>(lane / NP);
                const uint32_t cpiece1 = ((lane % NP) << 4) ^ ring2_swz16<BYTES>((uint32_t)RPI + lane / NP);
                const uint32_t myswz = LASTAX ? ring2_swz16<BYTES>(lane) : 0u;
                if constexpr (LASTAX) {
                    coop = vec && offsets_fit;
                    if (coop) {
                        const uint32_t my16 = (uint32_t)((unsigned long long)(q - xa) >> 4);
                        const uint32_t lives = __ballot_sync(0xFFFFFFFFu, live);
#pragma unroll
                        for (int j = 0; j < NP; ++j) {
                            const uint32_t r = (uint32_t)(j * RPI) + lane / NP;
                            co[j] = __shfl_sync(0xFFFFFFFFu, my16, (int)r) + lane % NP;
                            colive |= ((lives >> r) & 1u) << j;
                        }
                    }
                    __syncwarp();  // every lane is done with the previous tile's cells
                }
                unsigned issued = 0;
                auto issue = [&]() {
                    if (issued < nsteps) {
                        const unsigned left = ea - RS * issued;  // rows from this step on (uniform)
                        if constexpr (LASTAX) {
                            if (coop && left >= (unsigned)RS) {
                                const uint32_t dstw = cbase + slot_w * kSlotBytes;
#pragma unroll
                                for (int j = 0; j < NP; ++j) {
                                    if ((colive >> j) & 1u) {
                                        const char* src = xa + 16ull * (unsigned long long)(co[j] + issued * (unsigned)NP);
                                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dstw + (uint32_t)(j * RPI) * kLanePitch + ((j & 1) ? cpiece1 : cpiece0)), "l"(src) : "memory");
                                    }
                                }
                            } else if (live) {
                                const uint32_t dst = cell + slot_w * kSlotBytes;
                                const char* qs = q + (size_t)issued * (size_t)(kRows2 * BYTES);
                                if (vec && left >= (unsigned)RS) {
#pragma unroll
                                    for (int v = 0; v < NP; ++v)
                                        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + ((16u * v) ^ myswz)), "l"(qs + 16 * v) : "memory");
                                } else {
#pragma unroll
                                    for (int u = 0; u < RS; ++u)
                                        if ((unsigned)u < left)
                                            cpa_elem<BYTES>(dst + (((uint32_t)(u * BYTES) & ~15u) ^ myswz) + ((uint32_t)(u * BYTES) & 15u), qs + u * BYTES);
                                }
                            }
                        } else if (live) {
                            const uint32_t dst = cell + slot_w * kSlotBytes;
                            if (left >= (unsigned)RS) {
#pragma unroll
                                for (int u = 0; u < RS; ++u) {
                                    if constexpr (V16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * kRowP), "l"(q) : "memory");
                                    else cpa_elem<BYTES>(dst + u * kRowP, q);
                                    q += step_bytes;
                                }
                            } else {
#pragma unroll
                                for (int u = 0; u < RS; ++u) {
                                    if ((unsigned)u < left) {
                                        if constexpr (V16) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + u * kRowP), "l"(q) : "memory");
                                        else cpa_elem<BYTES>(dst + u * kRowP, q);
                                        q += step_bytes;
                                    }
                                }
                            }
                        }
                    }
                    cpa_commit_w();
                    slot_w = (slot_w + 1 == kSlots) ? 0 : slot_w + 1;
                    ++issued;
                };
#pragma unroll 1
                for (int k = 0; k < kAhead; ++k) issue();
#pragma unroll 1
                for (unsigned s = 0; s < nsteps; ++s) {
                    if (LASTAX) __syncwarp();  // the slot about to be refilled has been read by every lane
                    issue();
                    cpa_wait_w<kAhead>();
                    if (LASTAX) __syncwarp();  // ... and the copies other lanes made for this one have landed
                    const unsigned left = ea - RS * s;  // rows from this step on (uniform)
                    const uint32_t src = cell + slot_r * kSlotBytes;
                    slot_r = (slot_r + 1 == kSlots) ? 0 : slot_r + 1;
                    uint32_t flat[kRows2 * NW];  // memory order: element u = row * EPV + column, NW words each
                    if constexpr (LASTAX || V16) {
#pragma unroll
                        for (int v = 0; v < kRows2 * BYTES / 16; ++v) {
                            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                                         : "=r"(flat[4 * v]), "=r"(flat[4 * v + 1]), "=r"(flat[4 * v + 2]), "=r"(flat[4 * v + 3])
                                         : "r"(src + (LASTAX ? ((16u * v) ^ myswz) : (uint32_t)v * kRowP)));
                        }
                    } else {
#pragma unroll
                        for (int u = 0; u < kRows2; ++u) {
                            uint32_t w[NW];
                            lds_elem_w<NW>(src + u * kRowP, w);
#pragma unroll
                            for (int k = 0; k < NW; ++k) flat[u * NW + k] = w[k];
                        }
                    }
                    uint32_t sw[NW][kRows2], abw[NW][kRows2];
                    const bool first = (s == 0), ragged = left < (unsigned)RS;  // uniform
                    if (ragged) {
                        // rows past the end of the column: stale cells -> the value whose transform is 0
#pragma unroll
                        for (int u = 0; u < kRows2; ++u) {
                            if ((unsigned)(u / EPV) >= left) {
#pragma unroll
                                for (int k = 0; k < NW; ++k) flat[u * NW + k] = neutral_word<KIND, XFORM>(k == 0 ? 0 : NW - 1);
                            }
                        }
                    }
                    // Which step?  clean: no lane holds a masked element and every lane's row above is valid (or this is
                    // row 0) -- the unmasked code.  lanes: some lanes hold nothing but masked elements, above and here
                    // (columns inside a land mask); they are zeroed as a whole, no validity changes anywhere.  Anything
                    // else: the exact per-element test.
                    bool dirty = false, anyd = false, fast;
                    uint32_t lm = 0xFFFFFFFFu;  // all-ones: this lane's elements count
                    if (MASK != kMaskNone) {
                        dirty = may_contain_masked<KIND, MASK>(flat, mask_lo, mask_hi);
                        anyd = __any_sync(0xFFFFFFFFu, dirty);
                    }
                    if (!anyd) {
                        fast = first || pv_all;
                    } else {
                        const bool allm = dirty && !ragged && all_masked_quick<KIND, MASK>(flat, mask_lo, mask_hi);
                        const bool c = dirty ? (allm && (first || pkb == 0u)) : (first || pkb == kAllCols);
                        fast = __all_sync(0xFFFFFFFFu, c);
                        lm = dirty ? 0u : 0xFFFFFFFFu;
                        if (fast && __all_sync(0xFFFFFFFFu, dirty)) continue;  // nothing but masked elements in the whole step
                    }
                    if (fast) {
                        if (!anyd && !ragged) {
#pragma unroll
                            for (int u = 0; u < kRows2; ++u) {
                                const int e = u % EPV;
                                uint32_t w[NW];
#pragma unroll
                                for (int k = 0; k < NW; ++k) w[k] = flat[u * NW + k];
                                if (XFORM) E::transform(w);
#pragma unroll
                                for (int k = 0; k < NW; ++k) { sw[k][u] = w[k]; abw[k][u] = prev[e][k] & w[k]; prev[e][k] = w[k]; }
                            }
                        } else {
#pragma unroll
                            for (int u = 0; u < kRows2; ++u) {
                                const int e = u % EPV;
                                const bool exists = (unsigned)(u / EPV) < left;  // uniform
                                uint32_t w[NW];
#pragma unroll
                                for (int k = 0; k < NW; ++k) w[k] = flat[u * NW + k];
                                if (XFORM) E::transform(w);
#pragma unroll
                                for (int k = 0; k < NW; ++k) {
                                    w[k] &= lm;  // (absent rows: transform of the neutral value, zero already)
                                    sw[k][u] = w[k];
                                    abw[k][u] = prev[e][k] & w[k];
                                    prev[e][k] = exists ? w[k] : prev[e][k];  // the column's last row stays for the end of the walk
                                }
                            }
                        }
                        uint32_t rows_here = ragged ? left : (uint32_t)RS;
                        if (first) {
                            // row -1 is "masked": the b-stream loses row 0 of every column
                            ev_reserve((uint32_t)EPV);
#pragma unroll
                            for (int e = 0; e < EPV; ++e) {
                                ev_add_words<NW>(&ev, 0u, 0u, sw[0][e], NW == 2 ? sw[NW - 1][e] : 0u);
                            }
                            pkb = lm & kAllCols;
                            pv_all = !anyd;
                            rows_here -= 1u;
                        }
                        nvalid += (live && !dirty) ? rows_here * (uint32_t)EPV : 0u;
                    } else {
                        ev_reserve((uint32_t)kRows2);
                        uint32_t vbits = 0;
#pragma unroll
                        for (int u = 0; u < kRows2; ++u) {
                            const int e = u % EPV;
                            uint32_t w[NW];
#pragma unroll
                            for (int k = 0; k < NW; ++k) w[k] = flat[u * NW + k];
                            uint32_t kk = ((unsigned)(u / EPV) < left) ? 0xFFFFFFFFu : 0u;  // the row exists ...
                            if (MASK != kMaskNone) kk &= keep_of(w);                         // ... and is not masked
                            const uint32_t pke = 0u - ((pkb >> e) & 1u);                     // the row above: all-ones / 0
                            if (XFORM) E::transform(w);
                            uint32_t ca[NW], cb[NW], any_c = 0u;
#pragma unroll
                            for (int k = 0; k < NW; ++k) {
                                w[k] &= kk;
                                ca[k] = prev[e][k] & ~kk;    // row above valid, this one masked / absent
                                cb[k] = w[k] & ~pke;         // this row valid, the one above masked / absent
                                any_c |= ca[k] | cb[k];
                                sw[k][u] = w[k];
                                abw[k][u] = prev[e][k] & w[k];
                                prev[e][k] = w[k];
                            }
                            vbits |= (kk & pke) & (1u << u);  // the pair (row above, this row) counts
                            pkb = (pkb & ~(1u << e)) | (kk & (1u << e));
                            if (__any_sync(0xFFFFFFFFu, any_c != 0u))
                                ev_add_words<NW>(&ev, ca[0], NW == 2 ? ca[NW - 1] : 0u, cb[0], NW == 2 ? cb[NW - 1] : 0u);
                        }
                        nvalid += live ? (uint32_t)__popc(vbits) : 0u;
                        pv_all = __all_sync(0xFFFFFFFFu, pkb == kAllCols);
                    }
#pragma unroll
                    for (int k = 0; k < NW; ++k) {
                        S[k].add_chunk(sw[k], it);
                        AB[k].add_chunk(abw[k], it);
                    }
                    ++it;
                    if (it == L::kFlushChunks) drain();
                }
                // the row after the last one is "masked": the a-stream loses the last valid row of every column
                // (prev is zero where a ragged last step has already dealt with it)
                {
                    uint32_t any_c = 0u;
#pragma unroll
                    for (int e = 0; e < EPV; ++e)
#pragma unroll
                        for (int k = 0; k < NW; ++k) any_c |= prev[e][k];
                    if (__any_sync(0xFFFFFFFFu, any_c != 0u)) {
                        ev_reserve((uint32_t)EPV);
#pragma unroll
                        for (int e = 0; e < EPV; ++e) ev_add_words<NW>(&ev, prev[e][0], NW == 2 ? prev[e][NW - 1] : 0u, 0u, 0u);
                    }
                }
                slot_r = slot_w;  // the kAhead groups committed past the end of the column were empty
            }
        };
        constexpr int kVecCols = LASTAX ? 1 : 16 / BYTES;
        if (kVecCols > 1 && g.vec16 && box.e[kD - 1] % kVecCols == 0) walk(std::integral_constant<int, kVecCols>{});
        else walk(std::integral_constant<int, 1>{});
    }
    if (cur_chunk != 0xFFFFFFFFu) flush_chunk(cur_chunk);
}

template <typename K>
cudaError_t set_smem_once(K kern, size_t smem, std::atomic<unsigned long long>& attr_set) {
    if (smem <= 48 * 1024) return cudaSuccess;
    int dev = 0;
    cudaGetDevice(&dev);
    const unsigned long long dev_bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!(attr_set.load(std::memory_order_acquire) & dev_bit)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr_set.fetch_or(dev_bit, std::memory_order_release);
    }
    return cudaSuccess;
}

template <int KIND, int MASK, bool XFORM, bool LASTAX>
cudaError_t launch_walk_ax(const void* x, const ChunkGrid& g, uint32_t mlo, uint32_t mhi, unsigned long long* counts,
                           unsigned long long items, unsigned groups, unsigned grid, cudaStream_t stream, bool three,
                           const unsigned long long* dens) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    constexpr bool RING = BYTES >= 4;
    if constexpr (RING) {
        if (!three) {  // two streams + boundary corrections
            constexpr size_t smem2 = ring2_bytes<BYTES, LASTAX>();
            auto kern2 = k_pairs_chunked_walk2<KIND, MASK, XFORM, LASTAX>;
            static std::atomic<unsigned long long> attr_set2{0ull};
            if (cudaError_t e = set_smem_once(kern2, smem2, attr_set2)) return e;
            // chunks narrower than a tile of 32 lanes along an outer axis: interleave 8 neighbouring chunks (see the kernel)
            const unsigned lanes_per_row = g.chunk[kD - 1] / (g.vec16 ? (unsigned)(16 / BYTES) : 1u);
            // (along the innermost axis a tile is 32 runs of one chunk and the runs of the chunks next to it continue the
            // same rows: measured, float64 31x64x64 chunks 1.57 -> 2.54 TB/s, 56 chunks unchanged, but 32-element runs
            // 1.76 -> 1.08 -- interleaved from 64-element runs on)
            const unsigned cblock = (g.bounds[kD - 1] == nullptr &&
                                     (LASTAX ? (g.nchunk[kD - 1] > 1u && g.chunk[kD - 1] >= 64u) : lanes_per_row < 32u)) ? 8u : 1u;
            kern2<<<grid, kWThreads, smem2, stream>>>(x, g, mlo, mhi, counts, items, groups, dens, cblock);
            return cudaGetLastError();
        }
    }
    constexpr size_t smem = RING ? 16 + (size_t)kWWarps * kSlots * 32 * (LASTAX ? kRows * BYTES + 16 : kRows * BYTES) : 0;
    auto kern = k_pairs_chunked_walk<KIND, MASK, XFORM, RING, LASTAX>;
    static std::atomic<unsigned long long> attr_set{0ull};
    if (cudaError_t e = set_smem_once(kern, smem, attr_set)) return e;
    // chunks narrower than a tile along an outer axis: eight neighbouring chunks interleaved, as for the two-stream walkers
    const unsigned lanes_per_row3 = g.chunk[kD - 1] / (g.vec16 ? (unsigned)(16 / BYTES) : 1u);
    const unsigned cblock3 = (!LASTAX && g.axis != kD - 1 && lanes_per_row3 < 32u && g.bounds[kD - 1] == nullptr) ? 8u : 1u;
    kern<<<grid, kWThreads, smem, stream>>>(x, g, mlo, mhi, counts, items, groups, dens, cblock3);
    return cudaGetLastError();
}

template <int KIND, int MASK, bool XFORM>
cudaError_t launch_walk(const void* x, const ChunkGrid& g, uint32_t mlo, uint32_t mhi, unsigned long long* counts,
                        unsigned long long items, unsigned groups, unsigned grid, cudaStream_t stream, bool three,
                        const unsigned long long* dens) {
    if constexpr (Elem<KIND>::kBytes >= 4) {
        if (g.axis == kD - 1) return launch_walk_ax<KIND, MASK, XFORM, true>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
    }
    return launch_walk_ax<KIND, MASK, XFORM, false>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
}

template <int KIND>
cudaError_t dispatch_walk(const void* x, const ChunkGrid& g, bool transform, int mask_mode, uint32_t mlo, uint32_t mhi,
                          unsigned long long* counts, unsigned long long items, unsigned groups, unsigned grid,
                          cudaStream_t stream, bool three,
                        const unsigned long long* dens) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    if (transform) {
        if constexpr (isf) {
            switch (mask_mode) {
                case kMaskNone: return launch_walk<KIND, kMaskNone, true>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
                case kMaskNaN: return launch_walk<KIND, kMaskNaN, true>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
                default: return launch_walk<KIND, kMaskValue, true>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            }
        }
        return cudaErrorInvalidValue;
    }
    switch (mask_mode) {
        case kMaskNone: return launch_walk<KIND, kMaskNone, false>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
        case kMaskNaN:
            if constexpr (isf) return launch_walk<KIND, kMaskNaN, false>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            return cudaErrorInvalidValue;
        default: return launch_walk<KIND, kMaskValue, false>(x, g, mlo, mhi, counts, items, groups, grid, stream, three, dens);
    }
}

}  // namespace

// Counts into the raw per-chunk tables (the caller runs k_chunk_finalize) when the grid suits the walkers:
// *taken = false leaves everything to the odometer kernel of xbi_chunked.cu.
cudaError_t launch_pairs_chunked_walk(const void* x, int dtype, const ChunkGrid& g, unsigned nchunks, bool transform,
                                      int mask_mode, uint64_t mask_bits, unsigned long long* counts, cudaStream_t stream,
                                      bool force, int flavour, bool* taken) {
    *taken = false;
    if (g.axis < 0 || nchunks == 0) return cudaSuccess;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int per_sm = dtype_bytes(dtype) == 8 ? 1 : 2;
    const unsigned long long warps = (unsigned long long)sms * per_sm * kWWarps;
    const int bytes = dtype_bytes(dtype);
    // A plan per walker kernel: how many columns a lane takes, hence how many tiles (items) a chunk has.
    // Outer axes, 4- / 8-byte types: a lane takes 16 bytes of a row when every row of every chunk is made of whole,
    // aligned 16-byte vectors -- base pointer, every stride and the (regular) chunk length along the innermost
    // dimension multiples of 16 bytes (ragged edge chunks fall back to one column per lane inside the kernel) -- and a
    // chunk has enough columns to fill some lanes of a tile that way.  The two-stream walkers keep lanes without a
    // column on the fast step, so 8 live lanes are enough (float64, 10x10-column chunks: 1.87 TB/s with two columns
    // per lane, 0.92 with one); the three-stream walkers want 64 vector columns (4x4-column chunks: 4 live lanes of
    // 32, 0.3 TB/s instead of 1.1).
    struct Plan {
        ChunkGrid gg;
        unsigned long long units, items;
    };
    auto make_plan = [&](bool three) {
        Plan pl;
        pl.gg = g;
        pl.gg.vec16 = 0;
        if (bytes >= 4 && g.axis != kD - 1 && g.bounds[kD - 1] == nullptr && reinterpret_cast<uintptr_t>(x) % 16 == 0) {
            const long long epv = 16 / bytes;
            bool ok = g.chunk[kD - 1] % epv == 0;
            for (int d = g.lo; d < kD - 1; ++d) ok = ok && (g.stride[d] % epv == 0);
            unsigned long long c = 1;
            for (int d = 0; d < kD; ++d)
                if (d != g.axis) c *= g.chunk[d];
            ok = ok && (force || c / (unsigned long long)epv >= (three ? 64ull : 8ull));  // (force: the tests pin this path on small grids)
            pl.gg.vec16 = ok ? 1 : 0;
        }
        // the work inside a (full) chunk that the warps share: tiles of 32 lanes, one item each (edge chunks: some
        // items stay empty)
        unsigned long long cols = 1;
        for (int d = 0; d < kD; ++d)
            if (d != g.axis) cols *= g.chunk[d];
        if (pl.gg.vec16) cols /= (unsigned long long)(16 / bytes);
        pl.units = (cols + 31) / 32;
        pl.items = pl.units * nchunks;
        return pl;
    };
    const Plan plan2 = make_plan(false), plan3 = make_plan(true);
    const Plan& first = (bytes < 4 || flavour == 2) ? plan3 : plan2;  // the kernel that (most likely) runs
    if (first.units * nchunks < 2 * warps && !force) return cudaSuccess;  // too few columns to occupy the GPU this way
    // the innermost dimension as analysed axis: lanes sit on neighbouring RUNS, each copy instruction touches 32
    // different sectors; worth it only while a lane uses up its sectors, i.e. runs of at least a sector or two
    if (g.axis == kD - 1 && ((unsigned long long)g.chunk[kD - 1] * dtype_bytes(dtype) < 64 || dtype_bytes(dtype) < 4) && !force)
        return cudaSuccess;
    if (plan2.units > 0xFFFFFFFFull || plan3.units > 0xFFFFFFFFull) return cudaSuccess;
    const unsigned grid = (unsigned)(sms * per_sm);
    const uint32_t mlo = (uint32_t)(mask_bits & 0xFFFFFFFFu), mhi = (uint32_t)(mask_bits >> 32);
    auto run = [&](bool three, const unsigned long long* dens) -> cudaError_t {
        const Plan& pl = three ? plan3 : plan2;
        const ChunkGrid& gg = pl.gg;
        const unsigned long long items = pl.items;
        const unsigned groups = (unsigned)pl.units;
        switch (dtype) {
            case XBI_U8: return dispatch_walk<kU8>(x, gg, transform, mask_mode, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            case XBI_U16: return dispatch_walk<kU16>(x, gg, transform, mask_mode, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            case XBI_U32: return dispatch_walk<kU32>(x, gg, transform, mask_mode, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            case XBI_U64: return dispatch_walk<kU64>(x, gg, transform, mask_mode, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            case XBI_F16: return dispatch_walk<kF16>(x, gg, transform, mask_mode, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            case XBI_F32: return dispatch_walk<kF32>(x, gg, transform, mask_mode, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            case XBI_F64: return dispatch_walk<kF64>(x, gg, transform, mask_mode, mlo, mhi, counts, items, groups, grid, stream, three, dens);
            default: return cudaErrorInvalidValue;
        }
    };
    // short runs along the innermost axis (two steps of 16 elements or fewer per run): the per-column work of the
    // two-stream walkers (first-row / last-row corrections) outweighs their cheaper steps -- 1.4 against 1.7 TB/s at
    // 32-element runs, 1.6 against 1.9 at 64
    if (flavour == 0 && bytes == 4 && g.axis == kD - 1 && g.chunk[kD - 1] <= 64) flavour = 2;
    // ... and runs that do not start on 16-byte boundaries (element-sized copies, 32 lines per copy instruction):
    // 0.66 against 1.76 TB/s at 131-element runs
    if (flavour == 0 && bytes >= 4 && g.axis == kD - 1) {
        const long long epv = 16 / bytes;
        bool aligned = reinterpret_cast<uintptr_t>(x) % 16 == 0 && g.bounds[kD - 1] == nullptr && g.chunk[kD - 1] % epv == 0;
        for (int d = g.lo; d < kD - 1; ++d) aligned = aligned && (g.stride[d] % epv == 0);
        if (!aligned) flavour = 2;
    }
    cudaError_t e;
    if (bytes < 4 || flavour == 2) {
        e = run(true, nullptr);  // 1- and 2-byte types: three explicit streams, loads into registers
    } else if (flavour == 1 || mask_mode == kMaskNone) {
        e = run(false, nullptr);
    } else {
        // masked: sample the boundary density along the analysed axis, launch both, one returns at once
        CountJob job = {};
        job.x = x;
        job.dtype = dtype;
        job.transform = transform;
        job.mask_mode = mask_mode;
        job.mask_bits = mask_bits;
        long long numel = 1;
        for (int d = 0; d < kD; ++d) numel *= g.shape[d];
        job.inner = g.stride[g.axis];
        job.n = g.shape[g.axis];
        job.outer = numel / (job.n * job.inner);
        e = launch_mask_density_to(job, numel - job.inner, counts + kDensBnd, counts + kDensPairs, stream);
        if (e != cudaSuccess) return e;
        e = run(false, counts);
        if (e != cudaSuccess) return e;
        e = run(true, counts);
    }
    if (e != cudaSuccess) return e;
    note_fast_launch();
    *taken = true;
    return cudaSuccess;
}

}  // namespace xbi
