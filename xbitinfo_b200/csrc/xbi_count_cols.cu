// K1 fast path for adjacency along a NON-last axis (inner > 1): column walkers.
//
// The C-contiguous array is a matrix of R = outer*n rows of `inner` elements; the partner
// of an element is the one directly below it.  A thread owns one 16-byte column chunk and
// walks down a segment of rows, keeping the previous row in registers -- every element is
// read from HBM exactly once and needs no halo.  Rows reach the thread through a per-thread
// cp.async pipeline in shared memory (a warp copies 512 contiguous bytes per row), five
// 4-row groups ahead of the fold; four rows (16 words) are folded per step into the
// bit-sliced counters of xbi_common.cuh.
//
// Only two word streams are needed: b (every row below the first; masked elements zeroed)
// and a&b.  The a-stream total is the b-stream total + bits(first row of the matrix) -
// bits(its last row); two small row-plane passes (xbi_count_generic.cu) add those to the
// plus / minus sums.  Masked modes add the boundary corrections of xbi_mask.cuh.
// Pairs that straddle two blocks of the analysed axis are subtracted by k_block_edges.
#include <atomic>

#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"
#include "xbi_mask.cuh"
#include "xbi_tma.cuh"

// (constexpr members that only some instantiations use)
#pragma nv_diag_suppress 177

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
    if (Elem<KIND>::kBytes == 8) {
        k[0] = k[1] = keep_elem64<KIND, MASK>(w[0], w[1], mlo, mhi);
        k[2] = k[3] = keep_elem64<KIND, MASK>(w[2], w[3], mlo, mhi);
    } else {
#pragma unroll
        for (int i = 0; i < 4; ++i) k[i] = keep_word<KIND, MASK>(w[i], mlo);
    }
}
template <int BYTES>
__device__ __forceinline__ uint32_t count_chunk(const uint32_t (&v)[4]) {
    if (BYTES == 8) return (v[0] & 1u) + (v[2] & 1u);
    if (BYTES == 4) return (v[0] & 1u) + (v[1] & 1u) + (v[2] & 1u) + (v[3] & 1u);
    constexpr uint32_t m = (BYTES == 2) ? 0x00010001u : 0x01010101u;
    return __popc(v[0] & m) + __popc(v[1] & m) + __popc(v[2] & m) + __popc(v[3] & m);
}

// ---------------------------------------------------------------------------------
// Per-thread state of a column walker: the counters, the previous row of its 16-byte
// column chunk (x': transformed, masked elements zeroed) and, in masked modes, that row's
// keep-mask.
// ---------------------------------------------------------------------------------
template <int KIND, int MASK, bool XFORM>
struct ColWalker {
    static constexpr int BYTES = Elem<KIND>::kBytes;
    static constexpr bool K64 = (BYTES == 8);
    static constexpr bool MASKED = (MASK != kMaskNone);
    static constexpr bool DENSE = MaskMode<MASK>::kDense;  // three explicit streams under keep-masks (scattered masks)
    static constexpr int MT = MaskMode<MASK>::kTest;       // kMaskNone / kMaskNaN / kMaskValue
    static constexpr int NL = K64 ? 2 : 1;    // ladders per stream (low / high words)
    static constexpr int LOGC = K64 ? 3 : 4;  // 4 rows give 8 low + 8 high words, or 16 words
    static constexpr bool WHOLE = (BYTES >= 4);  // an element is one or two whole words
    static constexpr int WPE = K64 ? 2 : 1;      // words per element (whole-word types)
    static constexpr int EPC = 16 / BYTES;       // elements per 16-byte chunk
    static constexpr int kNumLadders = NL * (DENSE ? 3 : 2);
    using L = Ladder<3, (kNumLadders >= 6 ? 4 : kNumLadders >= 4 ? 5 : 8), LOGC>;

    L Bl[NL], ABl[NL], Al[DENSE ? NL : 1];
    unsigned long long accB[NL], accAB[NL], accA[NL];
    uint32_t prev[4];
    uint32_t pm;            // masked modes, whole-word types: bit e = element e of prev is masked
    uint32_t prev_keep[4];  // masked modes, packed types: keep-mask of prev ...
    bool prev_clean;        // ... which may be stale while prev_clean (no masked element in prev)
    bool lane_live;     // false while this lane shadows a chunk that does not exist (ragged warp item)
    uint32_t vrep;      // whole-word types: bit EPC*r + e set = element e of the chunk exists (last chunk
                        // of a row whose length is not a multiple of 16 bytes), replicated for 4 rows
    uint32_t it;
    uint32_t ninvalid;  // sparse flavour: pairs with a masked element, dense flavour: valid pairs (since the last flush)
    unsigned long long ninvalid_total;
    static constexpr int kEvWords = (MASKED && !DENSE) ? Elem<KIND>::kWords : 1;
    static constexpr int kEvPlanes = K64 ? 3 : 5;  // 64-bit types (two words per element): 3 planes, one row per reserve
    EventPlanes<kEvWords, kEvPlanes> er;            // boundary corrections of the masked modes
    unsigned long long (*s_minus)[64];              // the block's shared correction totals (flush target)

    __device__ __forceinline__ void init(unsigned long long (*shared_minus)[64]) {
        s_minus = shared_minus;
#pragma unroll
        for (int i = 0; i < NL; ++i) { Bl[i].clear(); ABl[i].clear(); accB[i] = 0; accAB[i] = 0; accA[i] = 0; }
        if (DENSE) {
#pragma unroll
            for (int i = 0; i < NL; ++i) Al[i].clear();
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) { prev[i] = 0u; prev_keep[i] = 0u; }
        prev_clean = true;
        pm = 0u;
        vrep = 0xFFFFu;
        lane_live = true;
        it = 0;
        ninvalid = 0;
        ninvalid_total = 0;
        if (MASKED && !DENSE) er.init();
    }

    __device__ __forceinline__ void flush(uint32_t lane) {
        if (!lane_live) {  // whatever an idle lane gathered since the last flush is discarded
#pragma unroll
            for (int i = 0; i < NL; ++i) { Bl[i].clear(); ABl[i].clear(); if (DENSE) Al[i].clear(); }
            ninvalid = 0;
        }
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            accB[i] += Bl[i].warp_totals(lane, it);
            accAB[i] += ABl[i].warp_totals(lane, it);
            Bl[i].clear(); ABl[i].clear();
            if (DENSE) { accA[i] += Al[i].warp_totals(lane, it); Al[i].clear(); }
        }
        ninvalid_total += ninvalid;
        ninvalid = 0;
        it = 0;
    }

    // first row of a work item (raw words; `live` = the chunk exists)
    __device__ __forceinline__ void start(const uint4& v, bool live, uint32_t mlo, uint32_t mhi) {
        prev[0] = v.x; prev[1] = v.y; prev[2] = v.z; prev[3] = v.w;
        if constexpr (DENSE) {
            keep_chunk<KIND, MT>(prev, mlo, mhi, prev_keep);
            xform_chunk<KIND, XFORM>(prev);
            if (!live) {
#pragma unroll
                for (int i = 0; i < 4; ++i) { prev[i] = 0u; prev_keep[i] = 0u; }
            }
            return;
        }
        if constexpr (MASKED && WHOLE) {
            pm = 0u;
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
                if (masked_elem<KIND, MT>(prev[WPE * e], prev[WPE * e + WPE - 1], mlo, mhi)) {
                    pm |= 1u << e;
#pragma unroll
                    for (int k = 0; k < WPE; ++k) prev[WPE * e + k] = neutral_word<KIND, XFORM>(k);
                }
            }
        } else if constexpr (MASKED) {
            keep_chunk<KIND, MT>(prev, mlo, mhi, prev_keep);
        }
        xform_chunk<KIND, XFORM>(prev);
        if constexpr (MASKED && !WHOLE) {
#pragma unroll
            for (int i = 0; i < 4; ++i) prev[i] &= prev_keep[i];
            prev_clean = (prev_keep[0] & prev_keep[1] & prev_keep[2] & prev_keep[3]) == 0xFFFFFFFFu;
        }
        if (!live) {  // a lane without a chunk: behaves like a fully masked column
#pragma unroll
            for (int i = 0; i < 4; ++i) { prev[i] = 0u; prev_keep[i] = 0u; }
            prev_clean = false;
            pm = (1u << EPC) - 1u;
        }
    }

    // fold 4 rows; rows >= nrows are absent (tail group only)
    __device__ __forceinline__ void process(const uint4 (&buf)[4], int nrows, uint32_t mlo, uint32_t mhi, uint32_t lane) {
        if constexpr (DENSE) {
            // ---- dense flavour: a, b and a&b formed explicitly under the pair's keep-mask ----
            uint32_t wa[4][4], wb[4][4], wab[4][4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                uint32_t c[4] = {buf[r].x, buf[r].y, buf[r].z, buf[r].w};
                uint32_t keep[4], v[4];
                keep_chunk<KIND, MT>(c, mlo, mhi, keep);
                xform_chunk<KIND, XFORM>(c);
                const bool present = r < nrows;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    v[i] = present ? (keep[i] & prev_keep[i]) : 0u;
                    wa[r][i] = prev[i] & v[i];
                    wb[r][i] = c[i] & v[i];
                    wab[r][i] = wa[r][i] & wb[r][i];
                    if (present) { prev_keep[i] = keep[i]; prev[i] = c[i]; }
                }
                ninvalid += count_chunk<BYTES>(v);
            }
            if constexpr (!K64) {
                uint32_t f[16];
#pragma unroll
                for (int r = 0; r < 4; ++r)
#pragma unroll
                    for (int i = 0; i < 4; ++i) f[4 * r + i] = wa[r][i];
                Al[0].add_chunk(f, it);
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
            } else {
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t f[1 << LOGC];
#pragma unroll
                    for (int r = 0; r < 4; ++r) { f[2 * r] = wa[r][h]; f[2 * r + 1] = wa[r][h + 2]; }
                    Al[DENSE ? h % NL : 0].add_chunk(f, it);
#pragma unroll
                    for (int r = 0; r < 4; ++r) { f[2 * r] = wb[r][h]; f[2 * r + 1] = wb[r][h + 2]; }
                    Bl[h % NL].add_chunk(f, it);
#pragma unroll
                    for (int r = 0; r < 4; ++r) { f[2 * r] = wab[r][h]; f[2 * r + 1] = wab[r][h + 2]; }
                    ABl[h % NL].add_chunk(f, it);
                }
            }
            ++it;
            if (it == L::kFlushChunks) flush(lane);
            return;
        }
        uint32_t cur[4][4];
        bool skip = false;  // warp-uniform: nothing in this group (nor the row above it) is valid
        if constexpr (MASKED && WHOLE) {
            // ---- whole-word elements: per-thread bit masks, only the dirty threads pay ----
#pragma unroll
            for (int r = 0; r < 4; ++r) { cur[r][0] = buf[r].x; cur[r][1] = buf[r].y; cur[r][2] = buf[r].z; cur[r][3] = buf[r].w; }
            constexpr uint32_t kRowMask = (1u << EPC) - 1u;
            uint32_t m = 0u;  // bit EPC*r + e: element e of row r is masked
            bool warp_dirty = false;  // warp-uniform
            {
                uint32_t raw[16];
#pragma unroll
                for (int j = 0; j < 16; ++j) raw[j] = cur[j >> 2][j & 3];
                const bool tdirty = (pm != 0u) || may_contain_masked<KIND, MT>(raw, mlo, mhi);
                if (__any_sync(0xFFFFFFFFu, tdirty)) {
                    warp_dirty = true;
                    const bool allm = (pm == kRowMask) && all_masked_quick<KIND, MT>(raw, mlo, mhi);
                    if (__all_sync(0xFFFFFFFFu, allm)) {
                        ninvalid += (uint32_t)nrows * (uint32_t)__popc(vrep & kRowMask);
                        return;  // nothing valid in the whole warp: prev stays zero, pm stays all-masked
                    }
                    if (tdirty) {
#pragma unroll
                        for (int j = 0; j < 4 * EPC; ++j) {
                            if (masked_elem<KIND, MT>(raw[WPE * j], raw[WPE * j + WPE - 1], mlo, mhi)) {
                                m |= 1u << j;
                                // overwritten with the value whose transform is 0: drops out of every sum
#pragma unroll
                                for (int k = 0; k < WPE; ++k) cur[(WPE * j + k) >> 2][(WPE * j + k) & 3] = neutral_word<KIND, XFORM>(k);
                            }
                        }
                    }
                }
            }
#pragma unroll
            for (int r = 0; r < 4; ++r) xform_chunk<KIND, XFORM>(cur[r]);
            if (warp_dirty) {
                uint32_t bnd = 0u;
                if (m | pm) {  // this thread's chunk has masked elements (or sits under one)
                    const uint32_t present = ((nrows >= 4) ? (0xFFFFu >> (16 - 4 * EPC)) : ((1u << (EPC * nrows)) - 1u)) & vrep;
                    const uint32_t above = (m << EPC) | pm;  // the element over each element
                    ninvalid += __popc((m | above) & present);
                    if (lane_live) bnd = (m ^ above) & present;
                }
                // validity changes between two rows of this chunk: the valid element of the pair goes
                // into the correction planes
                constexpr int kHalves = K64 ? 4 : 1, kRowsPer = 4 / kHalves;
#pragma unroll
                for (int h = 0; h < kHalves; ++h) {
                    er.reserve((uint32_t)(kRowsPer * EPC), lane, s_minus, 8 * BYTES);
                    const uint32_t hb = (bnd >> (EPC * kRowsPer * h)) & ((1u << (EPC * kRowsPer)) - 1u);
                    if (hb) {
#pragma unroll
                        for (int r = kRowsPer * h; r < kRowsPer * (h + 1); ++r) {
#pragma unroll
                            for (int e = 0; e < EPC; ++e) {
                                if ((bnd >> (EPC * r + e)) & 1u) {
                                    if ((m >> (EPC * r + e)) & 1u) {  // a valid, b masked
                                        er.template add<0, 0>((r == 0) ? prev[WPE * e] : cur[r > 0 ? r - 1 : 0][WPE * e]);
                                        if constexpr (K64)
                                            er.template add<0, kEvWords - 1>((r == 0) ? prev[WPE * e + 1] : cur[r > 0 ? r - 1 : 0][WPE * e + 1]);
                                    } else {  // a masked, b valid
                                        er.template add<1, 0>(cur[r][WPE * e]);
                                        if constexpr (K64) er.template add<1, kEvWords - 1>(cur[r][WPE * e + 1]);
                                    }
                                }
                            }
                        }
                        er.ecnt += (uint32_t)__popc(hb);
                    }
                }
                if ((m | pm) && nrows > 0) pm = (m >> (EPC * (nrows - 1))) & kRowMask;
            }
        } else if constexpr (MASKED) {
            // ---- packed 8/16-bit elements: warp-uniform general path on keep-masks ----
            uint32_t raw[16];
#pragma unroll
            for (int r = 0; r < 4; ++r) { raw[4 * r] = buf[r].x; raw[4 * r + 1] = buf[r].y; raw[4 * r + 2] = buf[r].z; raw[4 * r + 3] = buf[r].w; }
            const bool dirty = __any_sync(0xFFFFFFFFu, !prev_clean || may_contain_masked<KIND, MT>(raw, mlo, mhi));
#pragma unroll
            for (int r = 0; r < 4; ++r) {
#pragma unroll
                for (int i = 0; i < 4; ++i) cur[r][i] = raw[4 * r + i];
                xform_chunk<KIND, XFORM>(cur[r]);
            }
            if (dirty) {
                if (prev_clean) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) prev_keep[i] = 0xFFFFFFFFu;
                }
                uint32_t keep[4][4];
                uint32_t anyk = prev_keep[0] | prev_keep[1] | prev_keep[2] | prev_keep[3];
#pragma unroll
                for (int r = 0; r < 4; ++r) {
                    const uint32_t rw[4] = {raw[4 * r], raw[4 * r + 1], raw[4 * r + 2], raw[4 * r + 3]};
                    keep_chunk<KIND, MT>(rw, mlo, mhi, keep[r]);
                    if (r < nrows) anyk |= keep[r][0] | keep[r][1] | keep[r][2] | keep[r][3];
                }
                if (__all_sync(0xFFFFFFFFu, anyk == 0u)) {
                    ninvalid += (uint32_t)nrows * (16u / BYTES);
                    skip = true;  // prev stays zero, prev_keep stays zero
                } else {
                    uint32_t bnd = 0u;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) cur[r][i] &= keep[r][i];
                        if (r < nrows) {
                            uint32_t inval[4];
#pragma unroll
                            for (int i = 0; i < 4; ++i) {
                                const uint32_t pk = (r == 0) ? prev_keep[i] : keep[r > 0 ? r - 1 : 0][i];
                                inval[i] = ~(pk & keep[r][i]);
                                bnd |= pk ^ keep[r][i];
                            }
                            ninvalid += count_chunk<BYTES>(inval);
                        }
                    }
                    er.reserve(16u, lane, s_minus, 8 * BYTES);
                    if (bnd && lane_live) {  // validity changes between two rows of this chunk
#pragma unroll
                        for (int r = 0; r < 4; ++r) {
                            if (r < nrows) {
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t pk = (r == 0) ? prev_keep[i] : keep[r > 0 ? r - 1 : 0][i];
                                    const uint32_t pa = (r == 0) ? prev[i] : cur[r > 0 ? r - 1 : 0][i];
                                    const uint32_t va = pk & ~keep[r][i], vb = ~pk & keep[r][i];
                                    if (va) er.template add<0, 0>(pa & va);
                                    if (vb) er.template add<1, 0>(cur[r][i] & vb);
                                    er.ecnt += (va | vb) ? 1u : 0u;
                                }
                            }
                        }
                    }
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        if (r < nrows) {
#pragma unroll
                            for (int i = 0; i < 4; ++i) prev_keep[i] = keep[r][i];
                        }
                    }
                    prev_clean = (prev_keep[0] & prev_keep[1] & prev_keep[2] & prev_keep[3]) == 0xFFFFFFFFu;
                }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                cur[r][0] = buf[r].x; cur[r][1] = buf[r].y; cur[r][2] = buf[r].z; cur[r][3] = buf[r].w;
                xform_chunk<KIND, XFORM>(cur[r]);
            }
        }
        if (skip) return;

        uint32_t wb[4][4], wab[4][4];
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const bool present = r < nrows;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                wb[r][i] = present ? cur[r][i] : 0u;
                wab[r][i] = wb[r][i] & prev[i];
            }
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (present) prev[i] = cur[r][i];
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
        } else {
            // words 0,2 of a chunk are low halves, words 1,3 high halves
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                uint32_t f[1 << LOGC];
#pragma unroll
                for (int r = 0; r < 4; ++r) { f[2 * r] = wb[r][h]; f[2 * r + 1] = wb[r][h + 2]; }
                Bl[h % NL].add_chunk(f, it);
#pragma unroll
                for (int r = 0; r < 4; ++r) { f[2 * r] = wab[r][h]; f[2 * r + 1] = wab[r][h + 2]; }
                ABl[h % NL].add_chunk(f, it);
            }
        }
        ++it;
        if (it == L::kFlushChunks) flush(lane);
    }

    // end of kernel: block reduction, one atomic per non-zero counter and block.  s_sum / s_n are
    // shared scratch of the calling kernel (s_minus was handed to init()); all threads of the block call.
    __device__ __forceinline__ void finish(uint32_t lane, int nthreads, unsigned long long (*s_sum)[64],
                                           unsigned long long* s_n, Workspace* ws, unsigned long long flat_pairs) {
        flush(lane);
        for (int i = threadIdx.x; i < 3 * 64; i += nthreads) (&s_sum[0][0])[i] = 0;
        if (MASKED) {
            if (threadIdx.x == 0) *s_n = 0;
        }
        __syncthreads();
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            const int j = K64 ? (i * 32 + (int)lane) : ((int)lane % (8 * BYTES));
            // the a-stream is the b-stream shifted up one row; the launcher adds the first row of
            // the matrix and subtracts the last one (row-plane passes)
            const unsigned long long a = DENSE ? accA[i] : accB[i];
            if (a) atomicAdd(&s_sum[S_A][j], a);
            if (accB[i]) atomicAdd(&s_sum[S_B][j], accB[i]);
            if (accAB[i]) atomicAdd(&s_sum[S_AB][j], accAB[i]);
        }
        if (MASKED) {
            if (ninvalid_total) atomicAdd(s_n, ninvalid_total);
            if (!DENSE) er.flush(lane, s_minus, 8 * BYTES);
        }
        __syncthreads();
        for (int i = threadIdx.x; i < 3 * 64; i += nthreads) {
            const unsigned long long v = (&s_sum[0][0])[i];
            if (v) atomicAdd(&ws->plus.sums[0][0] + i, v);
        }
        if (MASKED && !DENSE) {
            for (int i = threadIdx.x; i < 2 * 64; i += nthreads) {
                const unsigned long long v = (&s_minus[0][0])[i];
                if (v) atomicAdd(&ws->minus.sums[0][0] + i, v);  // [S_A], [S_B] are the first two rows
            }
            if (threadIdx.x == 0 && *s_n) atomicAdd(&ws->minus.npairs, *s_n);
        }
        if (DENSE) {
            if (threadIdx.x == 0 && *s_n) atomicAdd(&ws->plus.npairs, *s_n);  // valid pairs, counted directly
        } else if (blockIdx.x == 0 && threadIdx.x == 0) {
            atomicAdd(&ws->plus.npairs, flat_pairs);
        }
    }
};

// the column walkers take the dense flavour when more than one sampled pair in this many is a
// validity boundary (measured break-even of the two flavours: about 1.4 % boundaries)
constexpr unsigned long long kColsDenseOneIn = 96ull;

constexpr int kStripBytes = 512;               // a warp's slice of a row: 32 lanes x 16 bytes
constexpr int kTileBytes = 4 * kStripBytes;     // one 4-row group of a warp

// ---------------------------------------------------------------------------------
// The column-walker kernel.  Work item = (segment, chunk) = divmod(item, chunks_per_row): one
// 16-byte column chunk walked down one segment of rows.  The pair rows
// [0, nseg*seg_min + seg_extra) are split into nseg segments; the first seg_extra segments are
// one row longer.  A warp takes 32 consecutive items at a time (grid-stride over "warp items"),
// so all its lanes run the same number of full 4-row groups -- the flush shuffles rely on that
// -- plus one predicated tail group for the remaining 0..4 rows of each lane's segment.
// A lane's rows travel through its own slice of a shared-memory ring with 16-byte cp.async
// (LDGSTS, L1 bypass) kAhead 4-row groups ahead of the fold.  A lane only ever reads back the bytes it copied itself, so the pipeline is
// per thread: cp.async.commit_group / wait_group, no mbarrier, no warp or block barrier and
// no registers held by loads in flight.  160 KB per SM are in flight.
// ---------------------------------------------------------------------------------
constexpr int kAsyncThreads = 512;
constexpr int kAsyncWarps = kAsyncThreads / 32;
constexpr int kAsyncRing = 6;             // groups of 4 rows x 512 B per warp
constexpr int kAhead = kAsyncRing - 1;    // groups in flight
constexpr size_t kAsyncSmem = 128 + (size_t)kAsyncWarps * kAsyncRing * kTileBytes;

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

__device__ __forceinline__ void cp_async4(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
// Rows are 16-byte aligned and a multiple of 16 bytes long (k_pairs_cols_elem below takes the others).
template <int KIND, int MASK, bool XFORM>
__global__ void __launch_bounds__(kAsyncThreads, 1)
k_pairs_cols(const char* __restrict__ x, long long row_bytes, long long chunks_per_row, long long nseg,
             long long seg_min, long long seg_extra, uint32_t mask_lo, uint32_t mask_hi,
             unsigned long long flat_pairs, Workspace* __restrict__ ws) {
    using W = ColWalker<KIND, MASK, XFORM>;
    // sparse and dense flavours are both launched; the sampled boundary density picks one
    if (W::MASKED && sampled_mask_is_dense(&ws->plus, kColsDenseOneIn) != W::DENSE) return;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_addr(smem_raw) + 127u) & ~127u;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t my0 = smem0 + warp * (kAsyncRing * kTileBytes) + lane * 16;  // + slot * kTileBytes + r * 512

    const long long total = chunks_per_row * nseg;
    const long long n_witems = (total + 31) >> 5;
    const long long warps_total = (long long)gridDim.x * kAsyncWarps;
    const long long groups = seg_min >> 2;  // full 4-row groups, the same for every item
    const long long stride = row_bytes;     // bytes between rows

    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_minus[2][64];
    __shared__ unsigned long long s_n;
    for (int i = threadIdx.x; i < 2 * 64; i += kAsyncThreads) (&s_minus[0][0])[i] = 0;
    __syncthreads();

    W cw;
    cw.init(s_minus);
    uint32_t slot_w = 0, slot_r = 0;  // ring positions (the same on every lane)

    for (long long wi = (long long)blockIdx.x * kAsyncWarps + warp; wi < n_witems; wi += warps_total) {
        long long item = (wi << 5) + lane;
        const bool active = item < total;
        // only the very last warp item can have idle lanes.  Idle lanes shadow a valid item;
        // what they add is discarded at the flush.
        const bool ragged = __any_sync(0xFFFFFFFFu, !active);
        if (ragged) cw.flush(lane);
        cw.lane_live = active;
        if (!active) item = total - 1;
        long long seg = 0, chunk = item;
        if (nseg > 1) { seg = item / chunks_per_row; chunk = item - seg * chunks_per_row; }
        const long long row_begin = seg * seg_min + (seg < seg_extra ? seg : seg_extra);
        const int tail_rows = (int)(seg_min + (seg < seg_extra ? 1 : 0) - (groups << 2));  // 0..4
        const bool any_tail = __any_sync(0xFFFFFFFFu, tail_rows > 0);
        const long long ngroups = groups + (any_tail ? 1 : 0);

        const char* q0 = x + row_begin * row_bytes + chunk * 16;
        const char* q = q0 + stride;  // first b-row; advanced by issue()
        long long issued = 0;
        // issue group `issued` (4 rows; the tail group re-reads an in-bounds row for absent rows)
        auto issue = [&]() {
            if (issued < ngroups) {
                const uint32_t dst = my0 + slot_w * kTileBytes;
                if (issued < groups) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) cp_async16(dst + r * kStripBytes, q + r * stride);
                    q += 4 * stride;
                } else {
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const long long rr = (r < tail_rows) ? r : (long long)tail_rows - 1;  // -1: last full-group row
                        cp_async16(dst + r * kStripBytes, q + rr * stride);
                    }
                }
            }
            cp_async_commit();  // (possibly empty: keeps the group count uniform)
            slot_w = (slot_w + 1 == kAsyncRing) ? 0 : slot_w + 1;
            ++issued;
        };
#pragma unroll 1
        for (int k = 0; k < kAhead; ++k) issue();
        cw.start(ldg_stream(reinterpret_cast<const uint4*>(q0)), true, mask_lo, mask_hi);
#pragma unroll 1
        for (long long g = 0; g < ngroups; ++g) {
            issue();
            cp_async_wait<kAhead>();  // group g has landed (kAhead younger ones may be in flight)
            const uint32_t src = my0 + slot_r * kTileBytes;
            uint4 buf[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) buf[r] = lds128(src + r * kStripBytes);
            slot_r = (slot_r + 1 == kAsyncRing) ? 0 : slot_r + 1;
            if (g < groups) cw.process(buf, 4, mask_lo, mask_hi, lane);
            else cw.process(buf, tail_rows, mask_lo, mask_hi, lane);
        }
        // the kAhead groups committed past the end of the item are empty; the ring position
        // moved with them, keep reader and writer aligned
        slot_r = slot_w;
        if (ragged) {
            cw.flush(lane);
            cw.lane_live = true;
        }
    }
    cp_async_wait<0>();

    cw.finish(lane, kAsyncThreads, s_sum, &s_n, ws, flat_pairs);
}


// ---------------------------------------------------------------------------------
// Rows that are only ELEMENT aligned (row length not a multiple of 16 bytes, or a base address that is not):
// 4- and 8-byte element types.  No 16-byte piece of a row keeps its alignment from one row to the next, so a
// lane owns single columns instead of a 16-byte chunk: a warp takes a strip of 32*EPC columns of one row
// segment, lane l the columns l, l+32, (l+64, l+96) of it.  Every element-sized cp.async of the warp then
// covers 32 consecutive elements (whole sectors but for the two ends), lands conflict-free in the lane's
// cells of the ring, and comes back with conflict-free element-sized shared loads; the walker sees the same
// "four rows of a 16-byte chunk" as in the aligned kernel -- which columns they are does not matter to it.
//
// The last strip of a row may hold few columns (513 = 4 x 128 + 1).  When half of a lane's element slots or
// more would be idle there, the slots walk different ROW ranges of the segment side by side instead
// (`split` sub-segments of Ls pair rows each; the < split rows that remain are walked by the last
// sub-segment alone in one more group).  Slots without a column read the neutral value (transform 0) and
// are excluded from the masked modes' bookkeeping through ColWalker::vrep.
//
// Measured alternatives: strips do not begin on sector boundaries, so neighbouring strips
// fetch the 64 bytes around their common edge twice, and as warps drift apart the second fetch misses the L2 --
// DRAM delivers 1.17x the array at 513 columns.  Walking the 16 warps of a block in lock step (a block barrier
// every other group, no side-by-side ranges) brings that to 1.00x but costs more issue slots than it saves: this
// kernel is bound by its instruction count (16 copies + 16 shared loads per group instead of 4 + 4).
// ---------------------------------------------------------------------------------
template <int KIND, int MASK, bool XFORM>
__global__ void __launch_bounds__(kAsyncThreads, 1)
k_pairs_cols_elem(const char* __restrict__ x, long long inner, long long strips, long long nseg, long long seg_min,
                  long long seg_extra, uint32_t mask_lo, uint32_t mask_hi, unsigned long long flat_pairs,
                  Workspace* __restrict__ ws) {
    using W = ColWalker<KIND, MASK, XFORM>;
    constexpr int BYTES = W::BYTES, EPC = W::EPC, WPE = W::WPE;
    static_assert(BYTES >= 4 && !W::DENSE, "whole-word element types, sparse flavour");
    constexpr int CW = 32 * EPC;           // columns of a strip
    constexpr int kCell = 32 * BYTES;      // bytes of one element slot of a strip row (all 32 lanes)
    constexpr uint32_t kAll = (1u << EPC) - 1u;
    constexpr uint32_t kRep = ((1u << (4 * EPC)) - 1u) / kAll;  // replicate a row's bits for 4 rows

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_addr(smem_raw) + 127u) & ~127u;
    const uint32_t warp = __shfl_sync(0xFFFFFFFFu, threadIdx.x >> 5, 0);
    const uint32_t lane = threadIdx.x & 31u;
    const uint32_t my0 = smem0 + warp * (kAsyncRing * kTileBytes) + lane * BYTES;  // + slot*kTileBytes + r*512 + e*kCell

    const long long n_witems = strips * nseg;
    const long long warps_total = (long long)gridDim.x * kAsyncWarps;
    const long long rounds = (n_witems + warps_total - 1) / warps_total;
    const long long gw = (long long)blockIdx.x * kAsyncWarps + warp;
    const long long stride = inner * BYTES;  // bytes between rows
    const int rem = (int)(inner - (strips - 1) * CW);  // columns of a row's last strip: 1..CW
    const int tail_w = (rem + 31) >> 5;                // element slots a lane needs there
    const int tail_split = EPC / tail_w;               // 4, 2 or 1 sub-segments side by side

    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned long long s_minus[2][64];
    __shared__ unsigned long long s_n;
    for (int i = threadIdx.x; i < 2 * 64; i += kAsyncThreads) (&s_minus[0][0])[i] = 0;
    __syncthreads();

    W cw;
    cw.init(s_minus);
    uint32_t slot_w = 0, slot_r = 0;  // ring positions (the same on every lane)
    auto copy_elem = [](uint32_t dst, const char* src) {
        if constexpr (BYTES == 8) cp_async8(dst, src);
        else cp_async4(dst, src);
    };

#pragma unroll 1
    for (long long round = 0; round < rounds; ++round) {
        // a block takes 16 consecutive items per round; the start moves by one warp from round to round, so that
        // no warp is dealt the (cheaper or dearer) last strip of a row every time
        const long long wi = round * warps_total + (gw + round) % warps_total;
        if (wi >= n_witems) continue;
        const long long seg = wi / strips, strip = wi - seg * strips;
        const long long row_begin = seg * seg_min + (seg < seg_extra ? seg : seg_extra);
        const long long len = seg_min + (seg < seg_extra ? 1 : 0);  // pair rows of the segment
        const bool last_strip = strip == strips - 1;
        const int split = last_strip ? tail_split : 1;
        const int sh = (split == 1) ? 31 : (split == EPC ? 0 : 1);  // slot e belongs to sub-segment e >> sh ...
        const int wmask = (split == 1) ? (EPC - 1) : (EPC / split - 1);  // ... and to column block e & wmask
        const long long Ls = len / split;                 // pair rows of a sub-segment
        const int left = (int)(len - Ls * split);         // < split: walked by the last sub-segment alone
        const long long full = Ls >> 2;
        const int tail_rows = (int)(Ls & 3);
        const long long ngroups = full + (tail_rows ? 1 : 0) + (left ? 1 : 0);

        // slot e: its column, whether it exists, and the address of the first row of its walk
        const char* p[EPC];
        uint32_t em = 0u;       // bit e: slot e has a column
        uint32_t em_last = 0u;  // ... and belongs to the last sub-segment
#pragma unroll
        for (int e = 0; e < EPC; ++e) {
            const long long col = strip * CW + (long long)(e & wmask) * 32 + lane;
            const int sub = e >> sh;
            const bool have = col < inner;
            if (have) em |= 1u << e;
            if (have && sub == split - 1) em_last |= 1u << e;
            p[e] = x + ((row_begin + sub * Ls) * inner + (have ? col : 0)) * BYTES;
        }
        cw.vrep = em * kRep;
        if (em != kAll) {
            // nothing of the previous item is in flight any more: the cells of the slots without a column are
            // never written by a copy -- fill them with the neutral value in every slot of this lane's ring
#pragma unroll 1
            for (int c = 0; c < kAsyncRing * 4; ++c) {
                const uint32_t cell = my0 + (c >> 2) * kTileBytes + (c & 3) * kStripBytes;
#pragma unroll
                for (int e = 0; e < EPC; ++e) {
                    if (!((em >> e) & 1u)) {
#pragma unroll
                        for (int k = 0; k < WPE; ++k)
                            asm volatile("st.shared.u32 [%0], %1;" ::"r"(cell + e * kCell + 4 * k), "r"(neutral_word<KIND, XFORM>(k)) : "memory");
                    }
                }
            }
        }
        long long issued = 0;
        long long roff = stride;  // byte offset of the first row of group `issued` from a slot's first row
        // (warp-uniform) every lane has all its columns and walks one row range: no predicates, one address per row
        const bool plain = !last_strip || rem == CW;
        // issue group `issued` (4 rows; absent rows of a short group re-read an in-bounds row)
        auto issue = [&]() {
            if (issued < ngroups) {
                const uint32_t dst = my0 + slot_w * kTileBytes;
                if (plain && issued < full) {
                    const char* q = p[0] + roff;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
#pragma unroll
                        for (int e = 0; e < EPC; ++e) copy_elem(dst + r * kStripBytes + e * kCell, q + e * kCell);
                        q += stride;
                    }
                    roff += 4 * stride;
                } else {
                    const bool is_left = issued >= full && !(tail_rows && issued == full);
                    const int nr = issued < full ? 4 : (is_left ? left : tail_rows);
                    const uint32_t who = is_left ? em_last : em;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        const long long at = roff + ((r < nr) ? r : nr - 1) * stride;
#pragma unroll
                        for (int e = 0; e < EPC; ++e) {
                            const char* a = p[e] + at;
                            asm volatile("" : "+l"(a));  // (formed once, outside the predicate)
                            if ((who >> e) & 1u) copy_elem(dst + r * kStripBytes + e * kCell, a);
                        }
                    }
                    roff += (long long)nr * stride;
                }
            }
            cp_async_commit();  // (possibly empty: keeps the group count uniform)
            slot_w = (slot_w + 1 == kAsyncRing) ? 0 : slot_w + 1;
            ++issued;
        };
#pragma unroll 1
        for (int k = 0; k < kAhead; ++k) issue();
        {
            uint32_t fw[4];
#pragma unroll
            for (int e = 0; e < EPC; ++e) {
#pragma unroll
                for (int k = 0; k < WPE; ++k) {
                    fw[WPE * e + k] = neutral_word<KIND, XFORM>(k);
                    if ((em >> e) & 1u) fw[WPE * e + k] = __ldg(reinterpret_cast<const uint32_t*>(p[e]) + k);
                }
            }
            cw.start(make_uint4(fw[0], fw[1], fw[2], fw[3]), true, mask_lo, mask_hi);
        }
#pragma unroll 1
        for (long long g = 0; g < ngroups; ++g) {
            issue();
            cp_async_wait<kAhead>();  // group g has landed (kAhead younger ones may be in flight)
            const uint32_t src = my0 + slot_r * kTileBytes;
            uint4 buf[4];
#pragma unroll
            for (int r = 0; r < 4; ++r) {
                if constexpr (BYTES == 8) {
                    const uint2 e0 = lds64(src + r * kStripBytes), e1 = lds64(src + r * kStripBytes + kCell);
                    buf[r] = make_uint4(e0.x, e0.y, e1.x, e1.y);
                } else {
                    buf[r] = make_uint4(lds32(src + r * kStripBytes), lds32(src + r * kStripBytes + kCell),
                                        lds32(src + r * kStripBytes + 2 * kCell), lds32(src + r * kStripBytes + 3 * kCell));
                }
            }
            slot_r = (slot_r + 1 == kAsyncRing) ? 0 : slot_r + 1;
            if (g < full) {
                cw.process(buf, 4, mask_lo, mask_hi, lane);
            } else {
                int nr = tail_rows;
                if (!(tail_rows && g == full)) {
                    // the rows left over by the split: the slots of the other sub-segments are done (their cells
                    // hold stale rows) -- neutral values, and out of the masked modes' bookkeeping
                    nr = left;
#pragma unroll
                    for (int r = 0; r < 4; ++r) {
                        uint32_t w[4] = {buf[r].x, buf[r].y, buf[r].z, buf[r].w};
#pragma unroll
                        for (int e = 0; e < EPC; ++e)
#pragma unroll
                            for (int k = 0; k < WPE; ++k)
                                if (!((em_last >> e) & 1u)) w[WPE * e + k] = neutral_word<KIND, XFORM>(k);
                        buf[r] = make_uint4(w[0], w[1], w[2], w[3]);
                    }
                    cw.vrep = em_last * kRep;
                }
                cw.process(buf, nr, mask_lo, mask_hi, lane);
            }
        }
        // the kAhead groups committed past the end of the item are empty; the ring position
        // moved with them, keep reader and writer aligned
        slot_r = slot_w;
    }
    cp_async_wait<0>();

    cw.finish(lane, kAsyncThreads, s_sum, &s_n, ws, flat_pairs);
}

// `chunks` = work items per row segment: 16-byte column chunks (32 of them make a warp item), or -- ELEM, rows that
// are only element aligned -- strips of 32*EPC columns (a warp item each)
template <int KIND, int MASK, bool XFORM, bool ELEM>
cudaError_t launch_cols_async(const CountJob& job, Workspace* ws, cudaStream_t stream, int64_t row_bytes, int64_t chunks,
                              int64_t pair_rows, int sms) {
    auto kern = [] {
        if constexpr (ELEM) return k_pairs_cols_elem<KIND, MASK, XFORM>;
        else return k_pairs_cols<KIND, MASK, XFORM>;
    }();
    int dev = 0;
    cudaGetDevice(&dev);
    static std::atomic<unsigned long long> attr_set{0ull};  // per instantiation; bit = device
    const unsigned long long dev_bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!(attr_set.load(std::memory_order_acquire) & dev_bit)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kAsyncSmem);
        if (e != cudaSuccess) return e;
        attr_set.fetch_or(dev_bit, std::memory_order_release);
    }
    // Work split: warp items = 32-chunk strips x row segments, dealt round-robin to the
    // sms*16 warps.  All items cost the same (segments differ by one row at most), so the
    // kernel ends when the warps with one item more than the others finish: pick the number
    // of segments (4 to 8 rounds of items per warp, segments of at least 64 rows) that wastes
    // the least of the last round.  (With 4.3 items per warp 13 % of the SM time was idle.)
    const int64_t warps = (int64_t)sms * kAsyncWarps;
    const int64_t max_seg = pair_rows / 64 > 0 ? pair_rows / 64 : 1;
    constexpr int64_t per_warp = ELEM ? 1 : 32;  // items a warp takes at a time
    auto witems = [&](int64_t ns) { return (chunks * ns + per_warp - 1) / per_warp; };
    int64_t lo_seg = (4 * warps * per_warp + chunks - 1) / chunks, hi_seg = (8 * warps * per_warp) / chunks;
    if (lo_seg < 1) lo_seg = 1;
    if (hi_seg < lo_seg) hi_seg = lo_seg;
    if (lo_seg > max_seg) lo_seg = max_seg;
    if (hi_seg > max_seg) hi_seg = max_seg;
    if (hi_seg - lo_seg > 4096) hi_seg = lo_seg + 4096;  // (narrow rows: bound the search)
    int64_t nseg = lo_seg;
    double best = -1.0;
    for (int64_t ns = lo_seg; ns <= hi_seg; ++ns) {
        const int64_t it = witems(ns), rounds = (it + warps - 1) / warps;
        const double eff = (double)it / (double)(rounds * warps);
        if (eff > best + 1e-9) { best = eff; nseg = ns; }
    }
    const int64_t seg_min = pair_rows / nseg;
    const int64_t seg_extra = pair_rows - seg_min * nseg;
    const int64_t total = witems(nseg) * 32;  // threads' worth of work
    int64_t blocks = (total + kAsyncThreads - 1) / kAsyncThreads;
    if (blocks > sms) blocks = sms;
    if constexpr (ELEM) {
        kern<<<(unsigned)blocks, kAsyncThreads, kAsyncSmem, stream>>>(
            static_cast<const char*>(job.x), (long long)job.inner, (long long)chunks, (long long)nseg,
            (long long)seg_min, (long long)seg_extra, (uint32_t)job.mask_bits, (uint32_t)(job.mask_bits >> 32),
            (unsigned long long)(pair_rows * job.inner), ws);
    } else {
        kern<<<(unsigned)blocks, kAsyncThreads, kAsyncSmem, stream>>>(
            static_cast<const char*>(job.x), (long long)row_bytes, (long long)chunks, (long long)nseg, (long long)seg_min,
            (long long)seg_extra, (uint32_t)job.mask_bits, (uint32_t)(job.mask_bits >> 32),
            (unsigned long long)(pair_rows * job.inner), ws);
    }
    return cudaGetLastError();
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

// What follows the walkers: the a-stream correction (row planes) and, for 16-byte aligned rows, the pairs that
// straddle two blocks of the analysed axis.
template <int KIND, int MASK, bool XFORM>
cudaError_t cols_followups(const CountJob& job, Workspace* ws, cudaStream_t stream, bool aligned, int64_t chunks,
                           int64_t pair_rows, int sms) {
    {
        // a-stream correction: + planes(first row) - planes(row after the last pair row)
        // (masked elements contribute nothing, as in the walkers; the dense flavour forms its
        // a-stream explicitly, so these passes return at once when it is the one that ran)
        const RawSums* flavour = (MASK != kMaskNone && aligned) ? &ws->plus : nullptr;
        cudaError_t e = launch_elem_planes(job, 0, job.inner, &ws->plus, stream, flavour, kColsDenseOneIn);
        if (e != cudaSuccess) return e;
        e = launch_elem_planes(job, pair_rows * job.inner, job.inner, &ws->minus, stream, flavour, kColsDenseOneIn);
        if (e != cudaSuccess) return e;
    }
    if (job.outer > 1 && aligned) {
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
    return cudaGetLastError();
}

template <int KIND, int MASK, bool XFORM>
cudaError_t launch_cols(const CountJob& job, Workspace* ws, cudaStream_t stream, int64_t* done_lo, int64_t* done_hi,
                        bool* edges_done) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    const int64_t row_bytes = job.inner * BYTES;
    const bool aligned = row_bytes % 16 == 0 && reinterpret_cast<uintptr_t>(job.x) % 16 == 0;
    // rows that are only element aligned: whole-word element types take the short-copy variant,
    // packed 8/16-bit types stay on the generic kernel
    if (!aligned && BYTES < 4) return cudaSuccess;
    const int64_t rows = job.outer * job.n;
    const int64_t pair_rows = rows - 1;
    if (pair_rows < 1) return cudaSuccess;
    const int64_t chunks = (row_bytes + 15) / 16;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    {
        cudaError_t e;
        if (aligned) {
            e = launch_cols_async<KIND, MASK, XFORM, false>(job, ws, stream, row_bytes, chunks, pair_rows, sms);
            if constexpr (MASK != kMaskNone) {  // the other flavour; one of the two returns at once
                if (e == cudaSuccess)
                    e = launch_cols_async<KIND, MASK + 2, XFORM, false>(job, ws, stream, row_bytes, chunks, pair_rows, sms);
            }
        } else {
            constexpr int64_t kStripCols = 32 * (16 / BYTES);
            if constexpr (BYTES >= 4)
                e = launch_cols_async<KIND, MASK, XFORM, true>(job, ws, stream, row_bytes, (job.inner + kStripCols - 1) / kStripCols,
                                                               pair_rows, sms);
            else e = cudaErrorInvalidValue;
        }
        if (e != cudaSuccess) return e;
    }
    if (cudaError_t e = cols_followups<KIND, MASK, XFORM>(job, ws, stream, aligned, chunks, pair_rows, sms)) return e;
    *edges_done = aligned;  // otherwise the caller runs the generic edge pass
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

cudaError_t launch_cols_followups(const CountJob& job, Workspace* ws, cudaStream_t stream) {
    // 4-byte element types, unmasked, 16-byte aligned rows (the two-axis kernel of xbi_count_xy.cu)
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int64_t chunks = job.inner * 4 / 16, pair_rows = job.outer * job.n - 1;
    if (job.transform) return cols_followups<kF32, kMaskNone, true>(job, ws, stream, true, chunks, pair_rows, sms);
    return cols_followups<kU32, kMaskNone, false>(job, ws, stream, true, chunks, pair_rows, sms);
}

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
