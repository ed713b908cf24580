// Shared device-side building blocks of the bit-pair counting kernels (sm_100a).
//
// What is counted (reference: xbitinfo/_py_bitinfo.py:97-160): for every bit position k
// of the analysis word and every adjacent pair (a, b) = (X[j], X[j+1]) along one axis,
// the 2x2 histogram of (bit k of a, bit k of b).  The histogram is recovered from three
// per-bit population counts and the pair count N:
//     A  = #pairs with a_k = 1      B = #pairs with b_k = 1      AB = #pairs with both
//     [[N-A-B+AB, B-AB], [A-AB, AB]]
// so the kernels only ever have to add up bit planes of three word streams.
//
// How it is counted: "vertical" (bit-sliced) counters.  Each thread keeps, per stream, a
// small stack of 32-bit registers; bit p of register r is binary digit r of the running
// count for bit position p.  Sixteen input words are folded in with a tree of 15
// carry-save adders (2 LOP3 each), higher digits are reached through a short ladder of
// deferred carry-save adders and a ripple counter.  Cost ~2 LOP3 per word and stream,
// independent of the data.  Population counts are only extracted at flush time, by a
// 5-step butterfly across the warp that leaves lane L holding the count of bit position L.
#pragma once
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace xbi {

// ---------------------------------------------------------------------------------
// raw accumulator block in device memory (part of the caller-provided workspace)
// ---------------------------------------------------------------------------------
// sums[s][j]: stream s in {A, B, AB}, j = bit number inside the element counted from the
// least significant bit.  "plus" collects every flat pair (f, f+inner); "minus" collects
// the flat pairs that straddle two rows/blocks of the analysed axis and therefore are
// not pairs of the reference (they are subtracted in the combine kernel).
struct RawSums {
    unsigned long long sums[3][64];
    unsigned long long npairs;
    unsigned long long pad[7];
};
struct Workspace {
    RawSums plus;
    RawSums minus;
};

enum Stream { S_A = 0, S_B = 1, S_AB = 2 };

// plus.pad[0] / pad[1]: validity boundaries / pairs seen by the sampling kernel.  The masked
// fast kernels take the dense flavour when more than one sampled pair in `one_in` is a boundary.
__device__ __forceinline__ bool sampled_mask_is_dense(const RawSums* plus, unsigned long long one_in) {
    const volatile unsigned long long* p = plus->pad;
    return p[0] * one_in > p[1];
}

// ---------------------------------------------------------------------------------
// carry-save adder on bit planes: (hi, lo) = a + b + c per bit position
// ---------------------------------------------------------------------------------
__device__ __forceinline__ void csa(uint32_t& hi, uint32_t& lo, uint32_t a, uint32_t b, uint32_t c) {
    // written as explicit LOP3s: left to the optimiser, the xor/and chains of a whole adder
    // tree get re-associated into ~60% more two-input operations
    uint32_t h, l;
    asm("lop3.b32 %0, %1, %2, %3, 0xE8;" : "=r"(h) : "r"(a), "r"(b), "r"(c));  // majority
    asm("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(l) : "r"(a), "r"(b), "r"(c));  // parity
    hi = h;
    lo = l;
}

// ---------------------------------------------------------------------------------
// flush helpers (out of line on purpose, see Ladder::warp_totals)
// ---------------------------------------------------------------------------------
// n[from..to) += carry, as a ripple of half adders on bit planes
static __device__ __noinline__ void ripple_add(uint32_t* n, int from, int to, uint32_t carry) {
    for (int j = from; j < to; ++j) {
        const uint32_t t = n[j] & carry;
        n[j] ^= carry;
        carry = t;
    }
}
// n[0..planes) are the binary digits (as bit planes) of 32 per-lane counters per bit
// position.  5-step butterfly: at distance d a lane keeps the bit positions whose bit
// log2(d) equals its own, receives the partner's digits for those positions and adds.
// Afterwards lane L holds the warp-wide count of bit position L; n needs planes+5 slots.
static __device__ __noinline__ uint32_t butterfly_totals(uint32_t* n, int planes, uint32_t lane) {
    for (int step = 0; step < 5; ++step) {
        const int d = 16 >> step;
        const uint32_t hi_sel = (d == 16) ? 0xFFFF0000u : (d == 8) ? 0xFF00FF00u : (d == 4) ? 0xF0F0F0F0u
                               : (d == 2) ? 0xCCCCCCCCu : 0xAAAAAAAAu;
        const uint32_t keep = (lane & d) ? hi_sel : ~hi_sel;
        uint32_t carry = 0;
        const int np = planes + step;
        for (int j = 0; j < np; ++j) {
            const uint32_t mine = n[j] & keep;
            const uint32_t give = n[j] & ~keep;
            const uint32_t got = __shfl_xor_sync(0xFFFFFFFFu, give, d);
            uint32_t hi, lo;
            csa(hi, lo, mine, got, carry);
            n[j] = lo;
            carry = hi;
        }
        n[np] = carry;
    }
    uint32_t total = 0;
    for (int j = 0; j < planes + 5; ++j) total |= ((n[j] >> lane) & 1u) << j;
    return total;
}

// ---------------------------------------------------------------------------------
// Vertical counter.  A "chunk" of 2^LOGC words is folded in with a static tree of
// carry-save adders (digits 0..LOGC-1); its carry enters DYN deferred carry-save levels
// (each holds a plane and a pending carry) and finally UP ripple planes.  Capacity between
// flushes: kFlushChunks chunks (chosen so that no digit can overflow).
// ---------------------------------------------------------------------------------
template <int DYN, int UP, int LOGC = 4>
struct Ladder {
    static constexpr int kChunk = 1 << LOGC;
    static constexpr int kPlanes = LOGC + DYN + UP;
    static constexpr uint32_t kFlushChunks = 1u << (DYN + UP - 1);

    uint32_t p[LOGC];         // digits 0..LOGC-1
    uint32_t P[DYN], Q[DYN];  // digit LOGC+l and its pending carry of the same weight
    uint32_t U[UP];           // digits LOGC+DYN ...

    __device__ __forceinline__ void clear() {
#pragma unroll
        for (int i = 0; i < LOGC; ++i) p[i] = 0;
#pragma unroll
        for (int i = 0; i < DYN; ++i) { P[i] = 0; Q[i] = 0; }
#pragma unroll
        for (int i = 0; i < UP; ++i) U[i] = 0;
    }

    // fold N = 2^k words into digits 0..k-1, return the carry of weight N
    template <int N>
    __device__ __forceinline__ uint32_t reduce(const uint32_t* w) {
        if constexpr (N == 2) {
            uint32_t c;
            csa(c, p[0], p[0], w[0], w[1]);
            return c;
        } else {
            constexpr int D = (N == 4) ? 1 : (N == 8) ? 2 : (N == 16) ? 3 : 4;
            const uint32_t a = reduce<N / 2>(w);
            const uint32_t b = reduce<N / 2>(w + N / 2);
            uint32_t c;
            csa(c, p[D], p[D], a, b);
            return c;
        }
    }

    // fold one chunk in; `it` is the number of chunks folded in since the last clear()
    __device__ __forceinline__ void add_chunk(const uint32_t (&w)[kChunk], uint32_t it) {
        push_level<0>(reduce<kChunk>(w), it);
    }
    __device__ __forceinline__ void add16(const uint32_t (&w)[16], uint32_t it) {
        static_assert(LOGC == 4, "add16 needs a 16-word chunk");
        add_chunk(w, it);
    }

    // the chunk's carry word enters the deferred levels; which levels fire depends
    // only on `it`, so the branches are uniform across the whole grid.  Level l holds a
    // pending carry in Q[l] exactly when bit l of the number of chunks pushed so far is set.
    template <int LVL>
    __device__ __forceinline__ void push_level(uint32_t carry, uint32_t it) {
        if constexpr (LVL == DYN) {
#pragma unroll
            for (int j = 0; j < UP; ++j) {
                const uint32_t t = U[j] & carry;
                U[j] ^= carry;
                carry = t;
            }
        } else {
            if (((it >> LVL) & 1u) == 0u) {
                Q[LVL] = carry;
            } else {
                uint32_t up;
                csa(up, P[LVL], P[LVL], Q[LVL], carry);
                push_level<LVL + 1>(up, it);
            }
        }
    }
    // add a single word of weight 1 (rare events only: costs 2 ops per plane)
    __device__ __forceinline__ void add1_slow(uint32_t w) {
        uint32_t carry = w;
#pragma unroll
        for (int j = 0; j < LOGC; ++j) { uint32_t t = p[j] & carry; p[j] ^= carry; carry = t; }
#pragma unroll
        for (int j = 0; j < DYN; ++j) { uint32_t t = P[j] & carry; P[j] ^= carry; carry = t; }
#pragma unroll
        for (int j = 0; j < UP; ++j) { uint32_t t = U[j] & carry; U[j] ^= carry; carry = t; }
    }

    // Extract: after the call, lane L of the warp returns the total (over the 32 lanes)
    // count of bit position L.  `it` = chunks pushed since clear().  Must be called by all
    // 32 lanes.  Does not clear.
    __device__ __forceinline__ uint32_t warp_totals(uint32_t lane, uint32_t it) const {
        // The digits go through a small local-memory array and two out-of-line helpers:
        // flushes are rare, and keeping ~600 unrolled instructions per call site out of
        // the kernels keeps their hot loops compact in the instruction cache.
        constexpr int W = kPlanes + 6;
        uint32_t n[W];
#pragma unroll
        for (int i = 0; i < LOGC; ++i) n[i] = p[i];
#pragma unroll
        for (int i = 0; i < DYN; ++i) n[LOGC + i] = P[i];
#pragma unroll
        for (int i = 0; i < UP; ++i) n[LOGC + DYN + i] = U[i];
#pragma unroll
        for (int i = kPlanes; i < W; ++i) n[i] = 0;
        // pending carries have the weight of their level
#pragma unroll
        for (int l = 0; l < DYN; ++l)
            if ((it >> l) & 1u) ripple_add(n, LOGC + l, kPlanes + 1, Q[l]);
        return butterfly_totals(n, kPlanes + 1, lane);
    }
    // The same for a counter that has seen only `it` >= 1 chunks (uniform across the warp): a lane's count per
    // bit position is at most it * 2^LOGC < 2^(LOGC + bits(it)), so only that many digit planes can be non-zero
    // and the butterfly (whose cost is proportional to the number of planes) runs on those.  For work items
    // of a tile or two (small chunks of a chunk grid) this is 2-3 times cheaper than the full extraction.
    __device__ __forceinline__ uint32_t warp_totals_bounded(uint32_t lane, uint32_t it) const {
        constexpr int W = kPlanes + 6;
        uint32_t n[W];
#pragma unroll
        for (int i = 0; i < LOGC; ++i) n[i] = p[i];
#pragma unroll
        for (int i = 0; i < DYN; ++i) n[LOGC + i] = P[i];
#pragma unroll
        for (int i = 0; i < UP; ++i) n[LOGC + DYN + i] = U[i];
#pragma unroll
        for (int i = kPlanes; i < W; ++i) n[i] = 0;
        int planes = LOGC + (32 - __clz((int)it));
        if (planes > kPlanes + 1) planes = kPlanes + 1;
#pragma unroll
        for (int l = 0; l < DYN; ++l)
            if ((it >> l) & 1u) ripple_add(n, LOGC + l, planes, Q[l]);
        return butterfly_totals(n, planes, lane);
    }
};

// ---------------------------------------------------------------------------------
// signed-exponent transform on raw IEEE words (reference: _py_bitinfo.py:42-94)
//   field F -> F - bias if F >= bias, else ~F (within the field); sign, mantissa kept.
// Closed form verified against the reference on every float16 pattern and on random
// float32/float64 patterns (tests/test_oracle.py).
// ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sexp_f32(uint32_t w) {
    // |x| >= 1.0 or NaN  <=>  exponent field >= 127   (one FSETP with |.| modifier)
    const bool big = !(fabsf(__uint_as_float(w)) < 1.0f);
    return big ? (w - 0x3F800000u) : (w ^ 0x7F800000u);
}
__device__ __forceinline__ uint32_t sexp_f64_hi(uint32_t hi) {
    // the high word of a double compares like a float: field >= 1023 <=> |hi| >= 0x3FF00000
    const bool big = !(fabsf(__uint_as_float(hi)) < __uint_as_float(0x3FF00000u));
    return big ? (hi - 0x3FF00000u) : (hi ^ 0x7FF00000u);
}
__device__ __forceinline__ uint32_t sexp_f16_single(uint32_t w) {  // one half in the low 16 bits
    return ((w & 0x7C00u) >= 0x3C00u) ? (w - 0x3C00u) : (w ^ 0x7C00u);
}
__device__ __forceinline__ uint32_t sexp_f16x2(uint32_t w) {  // two halves packed in one word
    // per-half 0xFFFF where |x| >= 1 or x is NaN, i.e. exponent field >= 15: one packed
    // half-precision compare (unordered >=) on |w|
    const __half2 h = *reinterpret_cast<const __half2*>(&w);
    const uint32_t m = __hgeu2_mask(__habs2(h), __half2half2(__ushort_as_half((unsigned short)0x3C00)));
    // selected halves: subtract the bias (no borrow leaves a selected half: its magnitude is
    // >= 0x3C00); the others: flip the exponent field
    return (w ^ (0x7C007C00u & ~m)) - (m & 0x3C003C00u);
}

}  // namespace xbi
