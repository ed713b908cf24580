// K1 fast path for adjacency along the LAST axis (inner == 1): TMA-staged rows.
//
// The flat array is viewed as rows of 128 bytes.  A persistent CTA (1 producer warp +
// 15 counting warps) streams stages of 240 or 480 rows through a ring of shared-memory
// buffers: one elected producer thread issues a 2-D tiled TMA load with the 128-byte
// swizzle per stage (plus a 16-byte bulk copy of the halo: the first bytes of the row
// after the stage) and the counting warps wait on the stage's mbarrier.  Each counting
// thread owns a contiguous 64-byte (128-byte for 64-bit types) segment of a row, so the
// neighbour of every element but the last is already in its registers; the swizzle makes
// the per-thread 128-bit shared-memory reads bank-conflict free.  Words go through the
// signed-exponent transform (xbitinfo/_py_bitinfo.py:42-94), the optional mask test and
// into the bit-sliced counters of xbi_common.cuh.  HBM traffic: every byte is read once.
//
// Unmasked pairs need only two word streams: a (every element) and a&b.  The b-stream
// total equals the a-stream total shifted by one element, i.e. A - bits(first element)
// + bits(halo element of the last stage); block 0 applies that correction.
#include <atomic>

#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"
#include "xbi_mask.cuh"
#include "xbi_tma.cuh"
#include "xbi_tmap.h"

namespace xbi {

namespace {

constexpr int kCountThreads = 480;  // 15 counting warps + 1 producer warp = 512 threads -> 128 registers each
constexpr int kCountWarps = kCountThreads / 32;
constexpr int kThreads = kCountThreads + 32;  // + producer warp
constexpr int kRowBytes = 128;
constexpr int kBoxRows = 240;  // rows per TMA box (<= 256)
constexpr long long kRowsFringeMax = 32768;  // leftover pairs (and row ends) the producer warps take along

template <int KIND>
struct Seg {
    static constexpr int kBytes = Elem<KIND>::kBytes;
    static constexpr bool k64 = (kBytes == 8);
    static constexpr int kSegWords = k64 ? 32 : 16;     // words a thread consumes per stage
    static constexpr int kSegChunks = kSegWords / 4;    // 16-byte chunks
    static constexpr int kExtra = k64 ? 2 : 1;          // neighbour words
    static constexpr int kThreadsPerRow = 32 / kSegWords;  // 2 or 1
    static constexpr int kRowsPerStage = kCountThreads / kThreadsPerRow;                 // 240 or 480
    static constexpr int kStageBytes = kRowsPerStage * kRowBytes;                        // 30 KB or 60 KB
    static constexpr int kStages = k64 ? 3 : 5;
    static constexpr int kElemsPerStage = kStageBytes / kBytes;
    static constexpr int kLaddersPerStream = k64 ? 2 : 1;
    static constexpr int kBits = 8 * kBytes;
};

// shared memory: [stages][kStageBytes] (1024-aligned) | halo [stages][16] | full[stages] | empty[stages]
template <int KIND>
constexpr size_t smem_bytes() {
    return 1024 + (size_t)Seg<KIND>::kStages * Seg<KIND>::kStageBytes + Seg<KIND>::kStages * 16 +
           2 * Seg<KIND>::kStages * 8 + 64 + 16 * 32 * 2 * (Seg<KIND>::k64 ? 2 : 1) * 4;
}

// ---- per-word helpers for types that fit one word -------------------------------------
template <int KIND, bool XFORM>
__device__ __forceinline__ uint32_t xform_word(uint32_t w) {
    if (!XFORM) return w;
    if (KIND == kF32) return sexp_f32(w);
    if (KIND == kF16) return sexp_f16x2(w);
    return w;
}
// the word holding, field by field, the successor element of every element of `cur`
template <int BYTES>
__device__ __forceinline__ uint32_t successor_word(uint32_t cur, uint32_t next) {
    if (BYTES == 4) return next;
    return __funnelshift_r(cur, next, 8 * BYTES);
}
// one element of BYTES bytes from shared memory, zero-extended into 32-bit words
template <int BYTES>
__device__ __forceinline__ void lds_elem(uint32_t addr, uint32_t (&w)[BYTES == 8 ? 2 : 1]) {
    if constexpr (BYTES == 8) {
        const uint2 v = lds64(addr);
        w[0] = v.x;
        w[1] = v.y;
    } else if constexpr (BYTES == 4) {
        w[0] = lds32(addr);
    } else if constexpr (BYTES == 2) {
        uint16_t v;
        asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
        w[0] = v;
    } else {
        uint32_t v;
        asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(addr));
        w[0] = v;
    }
}

// What the TMA ranges leave over -- the elements before the first 16-byte boundary, the ragged tail shorter
// than one stage, and the row ends among them -- is a few thousand pairs at most.  Counted after the kernel by a
// single CTA they cost 25-40 us per call (round 1: a (1,1,1)-grid generic kernel; then the finish kernel); here
// the PRODUCER warps take them, 32 pairs per CTA and round, once all their TMA loads are issued and while the
// counting warps are still busy with the last stages: per-bit __ballot_sync + __popc on the three explicit
// streams (a, b, a&b) of the pairs' transformed words, lane k keeping the totals of bit k.  Nothing is left for
// the critical path.
struct RowsFringe {
    const void* x;  // element 0 of the array
    long long n;    // length of the analysed (last) axis
    // [0], [1]: flat pairs (f, f+1), f in [lo, lo+cnt) -> plus sums;
    // [2], [3]: row ends t in [lo, lo+cnt), pair (t*n + n-1, t*n + n) -> minus sums
    long long lo[4], cnt[4];
};

template <int KIND, int MT, bool XFORM>
static __device__ __noinline__ void count_fringe(const RowsFringe& fr, uint32_t mask_lo, uint32_t mask_hi, uint32_t lane,
                                                 RawSums* __restrict__ plus, RawSums* __restrict__ minus) {
    using E = Elem<KIND>;
    constexpr int NW = E::kWords;
#pragma unroll 1
    for (int kind = 0; kind < 2; ++kind) {
        const long long c0 = fr.cnt[2 * kind], total = c0 + fr.cnt[2 * kind + 1];
        if (total <= 0) continue;
        unsigned long long accA[NW], accB[NW], accAB[NW], nvalid = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { accA[w] = 0; accB[w] = 0; accAB[w] = 0; }
#pragma unroll 1
        for (long long base = (long long)blockIdx.x * 32; base < total; base += (long long)gridDim.x * 32) {
            const long long idx = base + lane;
            uint32_t a[NW], b[NW];
#pragma unroll
            for (int w = 0; w < NW; ++w) { a[w] = 0; b[w] = 0; }
            bool ok = false;
            if (idx < total) {
                long long f = idx < c0 ? fr.lo[2 * kind] + idx : fr.lo[2 * kind + 1] + (idx - c0);
                if (kind == 1) f = f * fr.n + (fr.n - 1);
                E::load(fr.x, f, a);
                E::load(fr.x, f + 1, b);
                ok = true;
                if (MT == kMaskNaN) ok = !(E::is_nan(a) || E::is_nan(b));
                if (MT == kMaskValue) ok = !(E::equals(a, mask_lo, mask_hi) || E::equals(b, mask_lo, mask_hi));
                if (XFORM) { E::transform(a); E::transform(b); }
                if (!ok) {
#pragma unroll
                    for (int w = 0; w < NW; ++w) { a[w] = 0; b[w] = 0; }
                }
            }
            nvalid += (unsigned long long)__popc(__ballot_sync(0xFFFFFFFFu, ok));
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const uint32_t ab = a[w] & b[w];
#pragma unroll
                for (int bit = 0; bit < 32; ++bit) {
                    const uint32_t ca = __popc(__ballot_sync(0xFFFFFFFFu, (a[w] >> bit) & 1u));
                    const uint32_t cb = __popc(__ballot_sync(0xFFFFFFFFu, (b[w] >> bit) & 1u));
                    const uint32_t cab = __popc(__ballot_sync(0xFFFFFFFFu, (ab >> bit) & 1u));
                    if (lane == (uint32_t)bit) { accA[w] += ca; accB[w] += cb; accAB[w] += cab; }
                }
            }
        }
        RawSums* dst = kind == 0 ? plus : minus;
#pragma unroll
        for (int w = 0; w < NW; ++w) {
            const int j = w * 32 + (int)lane;
            if (accA[w]) atomicAdd(&dst->sums[S_A][j], accA[w]);
            if (accB[w]) atomicAdd(&dst->sums[S_B][j], accB[w]);
            if (accAB[w]) atomicAdd(&dst->sums[S_AB][j], accAB[w]);
        }
        if (lane == 0 && nvalid) atomicAdd(&dst->npairs, nvalid);
    }
}

template <int KIND, int MASK, bool XFORM>
__global__ void __launch_bounds__(kThreads, 1)
k_pairs_rows_tma(const __grid_constant__ CUtensorMap tmap, const char* __restrict__ region, long long nstages,
                 uint32_t mask_lo, uint32_t mask_hi, RawSums* __restrict__ dst, RawSums* __restrict__ dst_minus,
                 long long row_len, long long region_phase, const __grid_constant__ RowsFringe fringe) {
    using S = Seg<KIND>;
    constexpr int BYTES = S::kBytes;
    constexpr int NL = S::kLaddersPerStream;
    constexpr bool MASKED = (MASK != kMaskNone);
    constexpr bool DENSE = MaskMode<MASK>::kDense;  // three explicit streams (scattered masks)
    constexpr int MT = MaskMode<MASK>::kTest;       // kMaskNone / kMaskNaN / kMaskValue
    constexpr int kNumLadders = NL * (DENSE ? 3 : 2);
    using L = Ladder<3, (kNumLadders >= 6 ? 3 : kNumLadders >= 4 ? 5 : 8)>;
    // sparse and dense flavours are both launched; the sampled boundary density picks one
    // (break-even of the two flavours: about one boundary in 512 pairs; the dense flavour of the
    // 64-bit types is slower, so they stay sparse up to one in 256)
    if (MASKED && sampled_mask_is_dense(dst, S::k64 ? 256ull : 512ull) != DENSE) return;

    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_addr(smem_raw) + 1023u) & ~1023u;
    const uint32_t halo0 = smem0 + S::kStages * S::kStageBytes;
    const uint32_t full0 = halo0 + S::kStages * 16;
    const uint32_t empty0 = full0 + S::kStages * 8;
    const uint32_t ebuf0 = (empty0 + S::kStages * 8 + 15u) & ~15u;  // [16 stages][32 lanes][a,b words]

    const uint32_t warp = threadIdx.x >> 5;
    const uint32_t lane = threadIdx.x & 31u;

    // block totals; the boundary corrections of the masked modes are added to s_minus as they occur
    __shared__ unsigned long long s_sum[3][64];
    __shared__ unsigned int s_minus[2][64];  // boundary corrections of the a and b streams (32-bit: see warp_event_push)
    __shared__ unsigned long long s_ninvalid;
    for (int i = threadIdx.x; i < 3 * 64; i += kThreads) (&s_sum[0][0])[i] = 0;
    for (int i = threadIdx.x; i < 2 * 64; i += kThreads) (&s_minus[0][0])[i] = 0;
    if (threadIdx.x == 0) {
        s_ninvalid = 0;
        for (int s = 0; s < S::kStages; ++s) {
            mbar_init(full0 + 8 * s, 1);
            mbar_init(empty0 + 8 * s, kCountWarps + ((!S::k64 && row_len > 0) ? 1 : 0));
        }
        mbar_fence_init();
        tma_prefetch_descriptor(&tmap);
    }
    __syncthreads();

    // the producer is the LAST warp: the scheduler favours higher warp ids, so its few
    // instructions (TMA issue, row edges) are never starved by the 15 counting warps
    if (warp == kCountWarps) {
        // ================= producer (+ row edges) =================
        // Lane 0 keeps the ring full.  When row_len > 0 the whole warp additionally removes,
        // for every landed stage, the flat pairs that straddle two rows of the analysed
        // axis (last element of a row, first of the next): the reference never forms them
        // (xbitinfo/_py_bitinfo.py:155-160).  They are read from shared memory -- no extra
        // HBM traffic -- and accumulated into the `minus` sums.
        using E = Elem<KIND>;
        constexpr int NW = E::kWords;
        constexpr int LAG = S::kStages - 1;
        constexpr long long EPS = S::kElemsPerStage;
        // (64-bit types keep the separate edge pass: the extra state costs them more than it saves)
        const bool edges = (NW == 1) && row_len > 0;  // fewer than 32 row ends per stage: lane L takes the L-th
        // Lane L parks the raw (last, first) words of its row end in a private shared-memory
        // slot; every 16 stages the parked words are transformed and folded 16 at a time.
        using EL = Ladder<2, 4>;
        EL eA[NW], eB[NW], eAB[NW];
#pragma unroll
        for (int w = 0; w < NW; ++w) { eA[w].clear(); eB[w].clear(); eAB[w].clear(); }
        unsigned long long accA[NW], accB[NW], accAB[NW], acc_n = 0;
#pragma unroll
        for (int w = 0; w < NW; ++w) { accA[w] = 0; accB[w] = 0; accAB[w] = 0; }
        uint32_t nedge = 0, eit = 0, okmask = 0, gcount = 0;
        const uint32_t my_slots = ebuf0 + lane * (2 * NW * 4);  // + k * 32 * (2*NW*4)
        auto eflush = [&]() {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                accA[w] += eA[w].warp_totals(lane, eit);
                accB[w] += eB[w].warp_totals(lane, eit);
                accAB[w] += eAB[w].warp_totals(lane, eit);
                eA[w].clear(); eB[w].clear(); eAB[w].clear();
            }
            acc_n += nedge;
            nedge = 0;
            eit = 0;
        };
        auto fold_group = [&]() {
            uint32_t wa[NW][16], wb[NW][16];
#pragma unroll
            for (int k = 0; k < 16; ++k) {
                uint32_t a[NW], b[NW];
                const uint32_t slot = my_slots + k * (32 * 2 * NW * 4);
                if constexpr (NW == 2) {
                    const uint4 v = lds128(slot);
                    a[0] = v.x; a[1] = v.y; b[0] = v.z; b[1] = v.w;
                } else {
                    const uint2 v = lds64(slot);
                    a[0] = v.x; b[0] = v.y;
                }
                bool ok = (okmask >> k) & 1u;
                if (MT == kMaskNaN) ok = ok && !(E::is_nan(a) || E::is_nan(b));
                if (MT == kMaskValue) ok = ok && !(E::equals(a, mask_lo, mask_hi) || E::equals(b, mask_lo, mask_hi));
                if (XFORM) { E::transform(a); E::transform(b); }
#pragma unroll
                for (int w = 0; w < NW; ++w) { wa[w][k] = ok ? a[w] : 0u; wb[w][k] = ok ? b[w] : 0u; }
                nedge += ok ? 1u : 0u;
            }
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                uint32_t wab[16];
#pragma unroll
                for (int k = 0; k < 16; ++k) wab[k] = wa[w][k] & wb[w][k];
                eA[w].add16(wa[w], eit);
                eB[w].add16(wb[w], eit);
                eAB[w].add16(wab, eit);
            }
            ++eit;
            if (eit == EL::kFlushChunks) eflush();
            okmask = 0;
            gcount = 0;
        };
        // position of the stage's first element inside its row, tracked without divisions
        long long phase = 0, phase_step = 0;
        if (edges) {
            phase = (region_phase + (long long)blockIdx.x * EPS) % row_len;
            phase_step = ((long long)gridDim.x * EPS) % row_len;
        }
        auto stage_edges = [&](uint32_t j) {
            const uint32_t s = j % S::kStages;
            mbar_wait(full0 + 8 * s, (j / S::kStages) & 1u);
            const uint32_t sbase = smem0 + s * S::kStageBytes;
            const long long d = (row_len - 1 - phase) + (long long)lane * row_len;  // this lane's row end
            if (d < EPS) {
                uint32_t a[NW], b[NW];
                const uint32_t oa = (uint32_t)d * BYTES, ob = oa + BYTES;
                const uint32_t ra = oa >> 7, rb = ob >> 7;
                const uint32_t pa = sbase + (ra << 7) + (((((oa >> 4) & 7u) ^ (ra & 7u)) << 4) | (oa & 15u));
                const uint32_t pb = (ob == (uint32_t)S::kStageBytes)
                                        ? (halo0 + 16 * s)
                                        : sbase + (rb << 7) + (((((ob >> 4) & 7u) ^ (rb & 7u)) << 4) | (ob & 15u));
                lds_elem<BYTES>(pa, a);
                lds_elem<BYTES>(pb, b);
                const uint32_t slot = my_slots + gcount * (32 * 2 * NW * 4);
                if constexpr (NW == 2) {
                    asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(slot), "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(b[1]) : "memory");
                } else {
                    asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slot), "r"(a[0]), "r"(b[0]) : "memory");
                }
                okmask |= 1u << gcount;
            }
            phase += phase_step;
            if (phase >= row_len) phase -= row_len;
            __syncwarp();
            if (lane == 0) mbar_arrive(empty0 + 8 * s);
            if (++gcount == 16) fold_group();
        };

        uint32_t iter = 0;
        for (long long g = blockIdx.x; g < nstages; g += gridDim.x, ++iter) {
            if (lane == 0) {
                const uint32_t s = iter % S::kStages;
                const uint32_t ph = (iter / S::kStages) & 1u;
                mbar_wait(empty0 + 8 * s, ph ^ 1u);
                mbar_arrive_expect_tx(full0 + 8 * s, S::kStageBytes + 16);
                const long long row0 = g * S::kRowsPerStage;
#pragma unroll
                for (int b = 0; b < S::kRowsPerStage / kBoxRows; ++b)
                    tma_load_2d(smem0 + s * S::kStageBytes + b * kBoxRows * kRowBytes, &tmap, 0, (int32_t)(row0 + kBoxRows * b),
                                full0 + 8 * s);
                bulk_load_1d(halo0 + 16 * s, region + (row0 + S::kRowsPerStage) * kRowBytes, 16, full0 + 8 * s);
            }
            __syncwarp();
            if (edges && iter >= (uint32_t)LAG) stage_edges(iter - LAG);
        }
        if (edges) {
            for (uint32_t j = (iter > (uint32_t)LAG ? iter - LAG : 0); j < iter; ++j) stage_edges(j);
            if (gcount > 0) fold_group();
            eflush();
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                const int jbit = w * 32 + (int)lane;
                if (accA[w]) atomicAdd(&dst_minus->sums[S_A][jbit], accA[w]);
                if (accB[w]) atomicAdd(&dst_minus->sums[S_B][jbit], accB[w]);
                if (accAB[w]) atomicAdd(&dst_minus->sums[S_AB][jbit], accAB[w]);
            }
            // acc_n is per lane: reduce over the warp
            unsigned long long nsum = acc_n;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) nsum += __shfl_xor_sync(0xFFFFFFFFu, nsum, o);
            if (lane == 0 && nsum) atomicAdd(&dst_minus->npairs, nsum);
        }
        // the leftovers of the TMA ranges (see RowsFringe), while the counting warps finish the last stages
        count_fringe<KIND, MT, XFORM>(fringe, mask_lo, mask_hi, lane, dst, dst_minus);
        return;
    }

    // ================= counting warps =================
    // Two word streams in every mode: a (x', masked elements zeroed) and a&b.  Masked modes
    // add the boundary corrections and the invalid-pair count of xbi_mask.cuh.
    const uint32_t t = threadIdx.x;
    const uint32_t row = t / S::kThreadsPerRow;
    const uint32_t cbase = (t % S::kThreadsPerRow) * S::kSegChunks;
    const uint32_t swz = row & 7u;
    // where the neighbour words (first words of the next segment) live
    const bool nb_same_row = (cbase + S::kSegChunks) < 8;
    const bool nb_halo = (!nb_same_row) && (row == S::kRowsPerStage - 1);
    const uint32_t nb_off = nb_same_row ? (row * kRowBytes + (((cbase + S::kSegChunks) ^ swz) << 4))
                                        : ((row + 1) * kRowBytes + (((row + 1) & 7u) << 4));

    L A[NL], AB[NL], B[DENSE ? NL : 1];
#pragma unroll
    for (int i = 0; i < NL; ++i) { A[i].clear(); AB[i].clear(); }
    if (DENSE) {
#pragma unroll
        for (int i = 0; i < NL; ++i) B[i].clear();
    }
    unsigned long long accA[NL], accAB[NL], accB[NL];
#pragma unroll
    for (int i = 0; i < NL; ++i) { accA[i] = 0; accAB[i] = 0; accB[i] = 0; }
    // sparse flavour: pairs with a masked element; dense flavour: valid pairs (this thread, since the last flush)
    uint32_t ninvalid = 0;
    unsigned long long ninvalid_total = 0;
    constexpr uint32_t kEvFold = S::kBits < 32 ? S::kBits : 32;  // (boundary corrections: see warp_event_push)
    uint32_t it = 0;

    auto flush = [&]() {
#pragma unroll
        for (int i = 0; i < NL; ++i) {
            accA[i] += A[i].warp_totals(lane, it);
            accAB[i] += AB[i].warp_totals(lane, it);
            A[i].clear(); AB[i].clear();
            if (DENSE) { accB[i] += B[i].warp_totals(lane, it); B[i].clear(); }
        }
        ninvalid_total += ninvalid;
        ninvalid = 0;
        it = 0;
    };

    uint32_t iter = 0;
    for (long long g = blockIdx.x; g < nstages; g += gridDim.x, ++iter) {
        const uint32_t s = iter % S::kStages;
        const uint32_t ph = (iter / S::kStages) & 1u;
        mbar_wait(full0 + 8 * s, ph);
        const uint32_t sbase = smem0 + s * S::kStageBytes;
        uint32_t w[S::kSegWords + S::kExtra];
#pragma unroll
        for (int k = 0; k < S::kSegChunks; ++k) {
            const uint4 v = lds128(sbase + row * kRowBytes + (((cbase + k) ^ swz) << 4));
            w[4 * k + 0] = v.x; w[4 * k + 1] = v.y; w[4 * k + 2] = v.z; w[4 * k + 3] = v.w;
        }
        {
            const uint32_t a = nb_halo ? (halo0 + 16 * s) : (sbase + nb_off);
            if (S::k64) {
                const uint2 v = lds64(a);
                w[S::kSegWords] = v.x;
                w[S::kSegWords + 1] = v.y;
            } else {
                w[S::kSegWords] = lds32(a);
            }
        }
        bool skip = false;  // warp-uniform: every element this warp holds is masked

        if constexpr (DENSE) {
            // ---- dense flavour: a, b and a&b formed explicitly under the pair's keep-mask ----
            if constexpr (!S::k64) {
                uint32_t tw[17], keep[17];
#pragma unroll
                for (int i = 0; i < 17; ++i) { tw[i] = xform_word<KIND, XFORM>(w[i]); keep[i] = keep_word<KIND, MT>(w[i], mask_lo); }
                uint32_t a[16], b[16], ab[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    const uint32_t v = keep[i] & successor_word<BYTES>(keep[i], keep[i + 1]);
                    a[i] = tw[i] & v;
                    b[i] = successor_word<BYTES>(tw[i], tw[i + 1]) & v;
                    ab[i] = a[i] & b[i];
                    ninvalid += count_fields<BYTES>(v);
                }
                A[0].add16(a, it);
                B[0].add16(b, it);
                AB[0].add16(ab, it);
            } else {
                uint32_t lo[17], hi[17], keep[17];
#pragma unroll
                for (int e = 0; e < 17; ++e) {
                    lo[e] = w[2 * e];
                    hi[e] = (XFORM && KIND == kF64) ? sexp_f64_hi(w[2 * e + 1]) : w[2 * e + 1];
                    keep[e] = keep_elem64<KIND, MT>(w[2 * e], w[2 * e + 1], mask_lo, mask_hi);
                }
                uint32_t v[16], x0[16], x1[16], x2[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) { v[e] = keep[e] & keep[e + 1]; ninvalid += v[e] & 1u; }
#pragma unroll
                for (int e = 0; e < 16; ++e) { x0[e] = lo[e] & v[e]; x1[e] = lo[e + 1] & v[e]; x2[e] = x0[e] & x1[e]; }
                A[0].add16(x0, it); B[0].add16(x1, it); AB[0].add16(x2, it);
#pragma unroll
                for (int e = 0; e < 16; ++e) { x0[e] = hi[e] & v[e]; x1[e] = hi[e + 1] & v[e]; x2[e] = x0[e] & x1[e]; }
                A[NL - 1].add16(x0, it); B[DENSE ? NL - 1 : 0].add16(x1, it); AB[NL - 1].add16(x2, it);
            }
        } else if constexpr (BYTES < 4) {
            // ---- packed 8/16-bit elements (a NaN / value mask is decided on keep-masks) ----
            uint32_t tw[17];
#pragma unroll
            for (int i = 0; i < 17; ++i) tw[i] = xform_word<KIND, XFORM>(w[i]);
            if (MASKED && __any_sync(0xFFFFFFFFu, may_contain_masked<KIND, MT>(w, mask_lo, mask_hi))) {
                uint32_t keep[17], anyk = 0u;
#pragma unroll
                for (int i = 0; i < 17; ++i) { keep[i] = keep_word_fast<KIND, MT>(w[i], mask_lo); anyk |= keep[i]; }
                if (__all_sync(0xFFFFFFFFu, anyk == 0u)) {
                    ninvalid += 16u * (4u / BYTES);
                    skip = true;
                } else {
                    // invalid pairs: counted per field (at most 16 per field and stage), folded once; a word slot
                    // goes through the boundary corrections only when some lane's validity changes inside it (round 1
                    // pushed both streams of all 16 slots -- 32 ballots -- whenever any lane had any boundary: with 32
                    // elements per thread nearly every warp of a land-masked float16 array holds a coast line)
                    constexpr uint32_t kOnes = (BYTES == 2) ? 0x00010001u : 0x01010101u;
                    uint32_t inv = 0u;
#pragma unroll
                    for (int i = 0; i < 17; ++i) tw[i] &= keep[i];
#pragma unroll
                    for (int i = 0; i < 16; ++i) {
                        const uint32_t ks = successor_word<BYTES>(keep[i], keep[i + 1]);
                        inv += ~(keep[i] & ks) & kOnes;
                        if (__any_sync(0xFFFFFFFFu, keep[i] != ks)) {  // rare
                            const uint32_t va[1] = {tw[i] & ~ks};                                            // a valid, b masked
                            const uint32_t vb[1] = {successor_word<BYTES>(tw[i], tw[i + 1]) & ~keep[i]};   // a masked, b valid
                            warp_event_push<0>((keep[i] & ~ks) != 0u, va, lane, s_minus, kEvFold);
                            warp_event_push<1>((~keep[i] & ks) != 0u, vb, lane, s_minus, kEvFold);
                        }
                    }
                    ninvalid += (BYTES == 2) ? ((inv & 0xFFFFu) + (inv >> 16)) : ((inv * 0x01010101u) >> 24);
                }
            }
            if (!skip) {
                uint32_t a[16], ab[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    a[i] = tw[i];
                    ab[i] = tw[i] & successor_word<BYTES>(tw[i], tw[i + 1]);
                }
                A[0].add16(a, it);
                AB[0].add16(ab, it);
            }
        } else {
            // ---- whole-word elements (4 / 8 bytes): 17 elements, the last one is the neighbour.
            // Masked modes: a per-thread bit mask of the masked elements; only the threads that
            // hold one pay for the exact test, the zeroing and the corrections. ----
            constexpr int WPE = S::k64 ? 2 : 1;
            if (MASKED) {
                const bool tdirty = may_contain_masked<KIND, MT>(w, mask_lo, mask_hi);
                if (__any_sync(0xFFFFFFFFu, tdirty)) {
                    if (__all_sync(0xFFFFFFFFu, all_masked_quick<KIND, MT>(w, mask_lo, mask_hi))) {
                        ninvalid += 16u;
                        skip = true;
                    } else {
                        uint32_t m = 0u, bnd = 0u;  // bit e: element e (0..16) is masked / differs in validity from e+1
                        if (tdirty) {
#pragma unroll
                            for (int e = 0; e < 17; ++e) {
                                if (masked_elem<KIND, MT>(w[WPE * e], w[WPE * e + WPE - 1], mask_lo, mask_hi)) {
                                    m |= 1u << e;
                                    // overwritten with the value whose transform is 0: drops out of every sum
#pragma unroll
                                    for (int k = 0; k < WPE; ++k) w[WPE * e + k] = neutral_word<KIND, XFORM>(k);
                                }
                            }
                            ninvalid += __popc((m | (m >> 1)) & 0xFFFFu);
                            bnd = (m ^ (m >> 1)) & 0xFFFFu;
                        }
                        // validity changes between elements e and e+1: the valid one of the two is re-read from
                        // the stage buffer (still ours), transformed, and handed to the warp (WarpEvents), one
                        // boundary per lane and round
#pragma unroll 1
                        while (__any_sync(0xFFFFFFFFu, bnd != 0u)) {
                            const bool have = bnd != 0u;
                            bool a_masked = false;
                            uint32_t v[Elem<KIND>::kWords];
#pragma unroll
                            for (int k = 0; k < WPE; ++k) v[k] = 0u;
                            if (have) {
                                const uint32_t e = (uint32_t)__ffs((int)bnd) - 1u;
                                bnd &= bnd - 1u;
                                a_masked = (m >> e) & 1u;
                                const uint32_t ve = a_masked ? e + 1u : e;
                                const uint32_t wi = ve * WPE;  // word index inside the thread's segment
                                const uint32_t addr = (ve == 16u)
                                                          ? (nb_halo ? (halo0 + 16 * s) : (sbase + nb_off))
                                                          : (sbase + row * kRowBytes + (((cbase + (wi >> 2)) ^ swz) << 4) + ((wi & 3u) << 2));
                                if constexpr (S::k64) {
                                    const uint2 q = lds64(addr);
                                    v[0] = q.x;
                                    v[1] = q.y;
                                } else {
                                    v[0] = lds32(addr);
                                }
                                if (XFORM) Elem<KIND>::transform(v);
                            }
                            warp_event_push<0>(have && !a_masked, v, lane, s_minus, kEvFold);
                            warp_event_push<1>(have && a_masked, v, lane, s_minus, kEvFold);
                        }
                    }
                }
            }
            if (!skip) {
                uint32_t lo[17], hi[S::k64 ? 17 : 1];
#pragma unroll
                for (int e = 0; e < 17; ++e) {
                    if constexpr (S::k64) {
                        lo[e] = w[2 * e];
                        hi[e] = (XFORM && KIND == kF64) ? sexp_f64_hi(w[2 * e + 1]) : w[2 * e + 1];
                    } else {
                        lo[e] = xform_word<KIND, XFORM>(w[e]);
                    }
                }
                uint32_t x0[16], x1[16];
#pragma unroll
                for (int e = 0; e < 16; ++e) { x0[e] = lo[e]; x1[e] = lo[e] & lo[e + 1]; }
                A[0].add16(x0, it);
                AB[0].add16(x1, it);
                if constexpr (S::k64) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) { x0[e] = hi[e]; x1[e] = hi[e] & hi[e + 1]; }
                    A[NL - 1].add16(x0, it);
                    AB[NL - 1].add16(x1, it);
                }
            }
        }
        // the stage's words are in registers and consumed: hand the buffer back
        __syncwarp();
        if (lane == 0) mbar_arrive(empty0 + 8 * s);
        if (!skip) {
            ++it;
            if (it == L::kFlushChunks) flush();
        }
    }
    flush();

    // ---- block reduction over the 15 counting warps (named barrier 1, producer not involved)
    // (s_sum, s_minus, s_ninvalid were zeroed before the first barrier of the kernel)
#pragma unroll
    for (int i = 0; i < NL; ++i) {
        const int j = S::k64 ? (i * 32 + (int)lane) : ((int)lane % S::kBits);
        if (accA[i]) atomicAdd(&s_sum[S_A][j], accA[i]);
        if (accAB[i]) atomicAdd(&s_sum[S_AB][j], accAB[i]);
        if (DENSE) {
            if (accB[i]) atomicAdd(&s_sum[S_B][j], accB[i]);
        } else {
            if (accA[i]) atomicAdd(&s_sum[S_B][j], accA[i]);  // shifted by one element: corrected below (block 0)
        }
    }
    if (MASKED) {
        if (ninvalid_total) atomicAdd(&s_ninvalid, ninvalid_total);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kCountThreads) : "memory");
    for (int i = t; i < 3 * 64; i += kCountThreads) {
        const unsigned long long v = (&s_sum[0][0])[i];
        if (v) atomicAdd(&dst->sums[0][0] + i, v);
    }
    if (MASKED && !DENSE) {
        for (int i = t; i < 2 * 64; i += kCountThreads) {
            const unsigned long long v = (&s_minus[0][0])[i];
            if (v) atomicAdd(&dst_minus->sums[0][0] + i, v);  // [S_A], [S_B] are the first two rows
        }
        if (t == 0 && s_ninvalid) atomicAdd(&dst_minus->npairs, s_ninvalid);
    }
    if (DENSE && t == 0 && s_ninvalid) atomicAdd(&dst->npairs, s_ninvalid);  // valid pairs, counted directly

    if (!DENSE && blockIdx.x == 0 && t < 64) {
        // b-stream = a-stream shifted by one element: + bits(element after the region) - bits(first element)
        using E = Elem<KIND>;
        const long long last = nstages * (long long)S::kElemsPerStage;
        uint32_t first_w[E::kWords], last_w[E::kWords];
        E::load(region, 0, first_w);
        E::load(region, last, last_w);
        bool first_ok = true, last_ok = true;
        if (MT == kMaskNaN) { first_ok = !E::is_nan(first_w); last_ok = !E::is_nan(last_w); }
        if (MT == kMaskValue) { first_ok = !E::equals(first_w, mask_lo, mask_hi); last_ok = !E::equals(last_w, mask_lo, mask_hi); }
        if (XFORM) { E::transform(first_w); E::transform(last_w); }
        if ((int)t < S::kBits) {
            const uint32_t wsel = t >> 5, bsel = t & 31u;
            const unsigned long long up = last_ok ? ((last_w[E::kWords == 2 ? wsel : 0] >> bsel) & 1u) : 0u;
            const unsigned long long down = first_ok ? ((first_w[E::kWords == 2 ? wsel : 0] >> bsel) & 1u) : 0u;
            if (up != down) atomicAdd(&dst->sums[S_B][t], up - down);
        }
        if (t == 0) atomicAdd(&dst->npairs, (unsigned long long)last);
    }
}

// ---------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------
template <int KIND, int MASK, bool XFORM>
cudaError_t launch_rows(const CountJob& job, int64_t npairs_flat, Workspace* ws, cudaStream_t stream, int64_t* done_lo,
                        int64_t* done_hi, bool* edges_done, RowsLeftovers* left) {
    using S = Seg<KIND>;
    const int bytes = S::kBytes;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(job.x);
    const int64_t head = (int64_t)(((16 - (addr & 15)) & 15) / bytes);
    const int64_t numel = job.outer * job.n * job.inner;
    if (npairs_flat <= head) return cudaSuccess;
    int64_t nstages = (npairs_flat - head) / S::kElemsPerStage;
    // the 16-byte halo copy after the last stage must stay inside the array
    while (nstages > 0 && (head + nstages * S::kElemsPerStage) * bytes + 16 > numel * bytes) --nstages;
    if (nstages < 1) return cudaSuccess;
    const uint64_t nrows = (uint64_t)nstages * S::kRowsPerStage;
    // TMA coordinates are 32-bit: 2^31 rows of 128 bytes are 256 GiB, more than one B200 holds.  An error, not a
    // silent detour through the generic kernel
    if (nrows >= (1ull << 31)) return cudaErrorNotSupported;
    EncodeFn encode = get_encode_fn();
    if (!encode) return cudaErrorSymbolNotFound;  // reported by the caller (xbi_api.cu), never silently

    const char* region = static_cast<const char*>(job.x) + head * bytes;
    CUtensorMap tmap;
    const cuuint64_t gdim[2] = {32, nrows};
    const cuuint64_t gstride[1] = {kRowBytes};
    const cuuint32_t box[2] = {32, kBoxRows};
    const cuuint32_t estride[2] = {1, 1};
    const CUresult r = encode(&tmap, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, const_cast<char*>(region), gdim, gstride, box,
                              estride, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                              CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return cudaErrorInvalidValue;

    auto kern = k_pairs_rows_tma<KIND, MASK, XFORM>;
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    // per instantiation and device (the opt-in is per device); a bit set lets concurrent host threads race benignly
    static std::atomic<unsigned long long> attr_set{0ull};
    const unsigned long long dev_bit = (dev >= 0 && dev < 64) ? (1ull << dev) : 0ull;
    if (!(attr_set.load(std::memory_order_acquire) & dev_bit)) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes<KIND>());
        if (e != cudaSuccess) return e;
        attr_set.fetch_or(dev_bit, std::memory_order_release);
    }
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const unsigned grid = (unsigned)(nstages < sms ? nstages : sms);
    // rows long enough that a stage holds fewer than 32 row ends: the producer warp removes the
    // row-straddling pairs in shared memory (one per lane); shorter rows go to the edge pass
    const bool in_kernel_edges = !S::k64 && job.n > S::kElemsPerStage / 32 && job.outer > 1;
    // the leftovers the producer warps take along (both flavours of a masked mode get the same plan; only the
    // one that runs counts them)
    const int64_t dlo = head, dhi = head + nstages * S::kElemsPerStage;
    RowsFringe fr = {};
    fr.x = job.x;
    fr.n = job.n;
    RowsLeftovers lv = {};
    if (dlo + (npairs_flat - dhi) <= kRowsFringeMax) {
        fr.lo[0] = 0;   fr.cnt[0] = dlo;
        fr.lo[1] = dhi; fr.cnt[1] = npairs_flat - dhi;
        lv.flat = true;
    }
    {
        // row ends: t in [0, outer-1) sits at flat index t*n + n-1; with in-kernel edges those inside [dlo, dhi) are
        // done by the producer warps from shared memory
        const int64_t nedges = job.outer - 1;
        int64_t e_lo[2] = {0, 0}, e_cnt[2] = {nedges, 0};
        if (in_kernel_edges) {
            e_cnt[0] = dlo / job.n;
            e_lo[1] = dhi / job.n;
            e_cnt[1] = nedges - dhi / job.n;
        }
        for (int i = 0; i < 2; ++i)
            if (e_cnt[i] < 0) e_cnt[i] = 0;
        if (e_cnt[0] + e_cnt[1] <= kRowsFringeMax) {
            for (int i = 0; i < 2; ++i) { fr.lo[2 + i] = e_lo[i]; fr.cnt[2 + i] = e_cnt[i]; }
            lv.edges = true;
        }
    }
    kern<<<grid, kThreads, smem_bytes<KIND>(), stream>>>(tmap, region, (long long)nstages, (uint32_t)job.mask_bits,
                                                          (uint32_t)(job.mask_bits >> 32), &ws->plus, &ws->minus,
                                                          in_kernel_edges ? (long long)job.n : 0ll,
                                                          (long long)(head % job.n), fr);
    *edges_done = in_kernel_edges;
    *left = lv;
    note_fast_launch();
    *done_lo = dlo;
    *done_hi = dhi;
    return cudaGetLastError();
}

template <int KIND>
cudaError_t dispatch_rows(const CountJob& job, int64_t npairs_flat, Workspace* dst, cudaStream_t stream, int64_t* lo,
                          int64_t* hi, bool* ed, RowsLeftovers* lv) {
    constexpr bool isf = (KIND == kF16 || KIND == kF32 || KIND == kF64);
    if (job.transform) {
        if constexpr (isf) {
            switch (job.mask_mode) {
                case kMaskNone: return launch_rows<KIND, kMaskNone, true>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
                // masked: both flavours are launched, one of them returns at once (see xbi_internal.h)
                case kMaskNaN: {
                    const cudaError_t e = launch_rows<KIND, kMaskNaN, true>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
                    if (e != cudaSuccess || *hi <= *lo) return e;
                    return launch_rows<KIND, kMaskNaNDense, true>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
                }
                default: {
                    const cudaError_t e = launch_rows<KIND, kMaskValue, true>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
                    if (e != cudaSuccess || *hi <= *lo) return e;
                    return launch_rows<KIND, kMaskValueDense, true>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
                }
            }
        }
        return cudaSuccess;
    }
    // untransformed data of any type is counted as the unsigned integer of its width; a NaN
    // mask without the transform is rare enough to stay on the generic kernel
    if (job.mask_mode == kMaskNaN) return cudaSuccess;
    constexpr int UK = (KIND == kF16) ? kU16 : (KIND == kF32) ? kU32 : (KIND == kF64) ? kU64 : KIND;
    if (job.mask_mode == kMaskNone) return launch_rows<UK, kMaskNone, false>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
    const cudaError_t e = launch_rows<UK, kMaskValue, false>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
    if (e != cudaSuccess || *hi <= *lo) return e;
    return launch_rows<UK, kMaskValueDense, false>(job, npairs_flat, dst, stream, lo, hi, ed, lv);
}

}  // namespace

cudaError_t launch_pairs_tma(const CountJob& job, int64_t npairs_flat, Workspace* dst, cudaStream_t stream,
                             int64_t* done_lo, int64_t* done_hi, bool* edges_done, RowsLeftovers* leftovers) {
    *done_lo = 0;
    *done_hi = 0;
    *edges_done = false;
    *leftovers = RowsLeftovers{};
    if (job.inner != 1) return cudaSuccess;  // strided adjacency: see xbi_count_cols.cu
    switch (job.dtype) {
        case XBI_U8: return dispatch_rows<kU8>(job, npairs_flat, dst, stream, done_lo, done_hi, edges_done, leftovers);
        case XBI_U16: return dispatch_rows<kU16>(job, npairs_flat, dst, stream, done_lo, done_hi, edges_done, leftovers);
        case XBI_U32: return dispatch_rows<kU32>(job, npairs_flat, dst, stream, done_lo, done_hi, edges_done, leftovers);
        case XBI_U64: return dispatch_rows<kU64>(job, npairs_flat, dst, stream, done_lo, done_hi, edges_done, leftovers);
        case XBI_F16: return dispatch_rows<kF16>(job, npairs_flat, dst, stream, done_lo, done_hi, edges_done, leftovers);
        case XBI_F32: return dispatch_rows<kF32>(job, npairs_flat, dst, stream, done_lo, done_hi, edges_done, leftovers);
        case XBI_F64: return dispatch_rows<kF64>(job, npairs_flat, dst, stream, done_lo, done_hi, edges_done, leftovers);
        default: return cudaErrorInvalidValue;
    }
}

}  // namespace xbi
