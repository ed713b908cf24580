// Masked counting (masked_value / NaN pair exclusion) -- device-side building blocks.
//
// A pair is dropped when either element is masked.  With k_f = 1 for
// an unmasked element, x'_f = k_f ? T(x_f) : 0 (T = signed-exponent transform) and pairs
// (f, g = f + inner), the three plane sums and the pair count of the valid pairs are
//
//     AB = sum x'_f & x'_g                                (no validity needed at all)
//     A  = sum x'_f  -  sum_{k_f=1, k_g=0} x'_f           ("a valid, b masked" boundaries)
//     B  = sum x'_g  -  sum_{k_f=0, k_g=1} x'_g           ("a masked, b valid" boundaries)
//     N  = #flat pairs - #{pairs with k_f & k_g = 0}
//
// so a masked pass is the unmasked two-stream pass over x' plus corrections that are only
// non-zero where validity CHANGES between neighbours.  The fast kernels therefore
//   1. probe each thread's words for masked elements with instructions that stay off the
//      ALU pipe where possible (x*0 accumulated with FFMA/HFMA2: NaN iff a NaN/Inf went in);
//   2. take the plain unmasked code when a whole warp is clean;
//   3. otherwise build the keep-masks, zero the masked elements, count the invalid pairs
//      and, at validity boundaries only, add the correction words to small counters (rows kernel:
//      broadcast over the warp, WarpEvents; column walkers: register bit planes, EventPlanes)
//      (-> the `minus` sums of the workspace).
// The probe may report false positives (Inf, doubles >= 2^1017, ...), never false negatives:
// step 3 always uses the exact test.
#pragma once
#include "xbi_common.cuh"
#include "xbi_elem.cuh"
#include "xbi_internal.h"

namespace xbi {

// ---------------------------------------------------------------------------------
// exact keep-masks: all-ones in the field of every element that is NOT masked
// ---------------------------------------------------------------------------------
template <int KIND, int MASK>
__device__ __forceinline__ uint32_t keep_word(uint32_t w, uint32_t mask_lo) {  // types of <= 32 bits
    if (MASK == kMaskNone) return 0xFFFFFFFFu;
    constexpr int BYTES = Elem<KIND>::kBytes;
    if (BYTES == 4) {
        if (MASK == kMaskNaN) return ((w & 0x7FFFFFFFu) > 0x7F800000u) ? 0u : 0xFFFFFFFFu;
        return (w == mask_lo) ? 0u : 0xFFFFFFFFu;
    } else if (BYTES == 2) {
        if (MASK == kMaskNaN) {
            const uint32_t mag = w & 0x7FFF7FFFu;  // NaN <=> magnitude > 0x7C00
            const uint32_t nan = ((mag + 0x03FF03FFu) >> 15) & 0x00010001u;
            return ~(nan * 0xFFFFu);
        }
        const uint32_t x = w ^ (mask_lo * 0x00010001u);
        const uint32_t nz = ((((x & 0x7FFF7FFFu) + 0x7FFF7FFFu) | x) >> 15) & 0x00010001u;
        return nz * 0xFFFFu;
    } else {
        const uint32_t x = w ^ (mask_lo * 0x01010101u);
        const uint32_t nz = ((((x & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | x) >> 7) & 0x01010101u;
        return nz * 0xFFu;
    }
}
// the same, with the one-instruction form where the hardware has one: float16 NaN mask = packed "a == a" (HSET2)
template <int KIND, int MASK>
__device__ __forceinline__ uint32_t keep_word_fast(uint32_t w, uint32_t mask_lo) {
    if constexpr (KIND == kF16 && MASK == kMaskNaN) {
        const __half2 h = *reinterpret_cast<const __half2*>(&w);
        return __heq2_mask(h, h);  // ordered compare: 0xFFFF per half unless NaN
    } else {
        return keep_word<KIND, MASK>(w, mask_lo);
    }
}
template <int KIND, int MASK>
__device__ __forceinline__ uint32_t keep_elem64(uint32_t lo, uint32_t hi, uint32_t mask_lo, uint32_t mask_hi) {
    if (MASK == kMaskNone) return 0xFFFFFFFFu;
    bool masked;
    if (MASK == kMaskNaN) {
        const uint32_t m = hi & 0x7FFFFFFFu;
        masked = m > 0x7FF00000u || (m == 0x7FF00000u && lo != 0u);
    } else {
        masked = (lo == mask_lo) && (hi == mask_hi);
    }
    return masked ? 0u : 0xFFFFFFFFu;
}
// number of elements whose field in v is all-ones (v holds whole fields of ones / zeros)
template <int BYTES>
__device__ __forceinline__ uint32_t count_fields(uint32_t v) {
    if (BYTES >= 4) return v & 1u;
    if (BYTES == 2) return __popc(v & 0x00010001u);
    return __popc(v & 0x01010101u);
}

// ---------------------------------------------------------------------------------
// probe: may these raw words contain a masked element?  (64-bit types: w = lo,hi,lo,hi,...)
// ---------------------------------------------------------------------------------
__device__ __forceinline__ float fma_zero(uint32_t w, float acc) {  // acc + w*0, IEEE (NaN/Inf -> NaN)
    float r;
    asm("fma.rn.f32 %0, %1, 0f00000000, %2;" : "=f"(r) : "f"(__uint_as_float(w)), "f"(acc));
    return r;
}
__device__ __forceinline__ uint32_t hfma2_zero(uint32_t w, uint32_t acc) {  // per half: acc + w*0
    uint32_t r;
    asm("{.reg .b32 z; mov.b32 z, 0; fma.rn.f16x2 %0, %1, z, %2;}" : "=r"(r) : "r"(w), "r"(acc));
    return r;
}

// two words at a time (sm_100 packed fp32 FMA): halves the issue slots the probe takes
__device__ __forceinline__ unsigned long long fma2_zero(uint32_t w0, uint32_t w1, unsigned long long acc) {
    unsigned long long r;
    asm("{.reg .b64 t, z; mov.b64 t, {%1, %2}; mov.b64 z, 0; fma.rn.f32x2 %0, t, z, %3;}"
        : "=l"(r)
        : "r"(w0), "r"(w1), "l"(acc));
    return r;
}

// {a0, a1} * {b0, b1} + {c0, c1}, IEEE
__device__ __forceinline__ unsigned long long fma2_abc(uint32_t a0, uint32_t a1, uint32_t b0, uint32_t b1, uint32_t c0, uint32_t c1) {
    unsigned long long r;
    asm("{.reg .b64 a, b, c; mov.b64 a, {%1, %2}; mov.b64 b, {%3, %4}; mov.b64 c, {%5, %6}; fma.rn.f32x2 %0, a, b, c;}"
        : "=l"(r)
        : "r"(a0), "r"(a1), "r"(b0), "r"(b1), "r"(c0), "r"(c1));
    return r;
}
__device__ __forceinline__ unsigned long long fma2_pairs(unsigned long long a, unsigned long long b, unsigned long long c) {
    unsigned long long r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}

template <int KIND, int MASK, int N>
__device__ __forceinline__ bool may_contain_masked(const uint32_t (&w)[N], uint32_t mask_lo, uint32_t mask_hi) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    if (MASK == kMaskNone) return false;
    if (MASK == kMaskNaN) {
        if (BYTES == 4 || BYTES == 8) {
            // (the high word of a double read as a float32 is NaN/Inf exactly when the top eight exponent bits are ones:
            // every double NaN/Inf, plus finite |x| >= 2^1017)
            // float32: the word itself.  THREE values per FMA lane: a * b + c
            // is NaN whenever one of the three is; without a NaN among them it can only become one through an infinity
            // (in the data, or a product beyond 3.4e38) meeting a zero factor or an opposite infinity -- a false
            // positive that costs the exact test, never a wrong count.  17 float32 words: 4 FFMA2 + 1 FADD where x * 0
            // accumulated pair by pair took 8 + 4 -- issue slots are what the probe costs (the FMA pipe is idle in
            // these kernels).
            if (BYTES == 8) {
                // (float64 keeps x * 0 accumulated over the high words: the column walkers' 8 high words per group gain
                // nothing from the packed form -- configs[2] measured 1 % slower with it -- and the rows kernel is
                // DRAM-bound in this mode)
                float s[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int i = 1, c = 0; i < N; i += 2, ++c) s[c & 3] = fma_zero(w[i], s[c & 3]);
                const float t = (s[0] + s[1]) + (s[2] + s[3]);
                return !(t == t);
            }
            constexpr int STEP = 1;
            constexpr int M = N / STEP;  // values to probe
            constexpr int G = (M + 5) / 6;
            constexpr uint32_t kOne = 0x3F800000u;
            auto val = [&](int i, uint32_t missing) -> uint32_t { return i < M ? w[i * STEP + STEP - 1] : missing; };
            unsigned long long g[G];
#pragma unroll
            for (int k = 0; k < G; ++k) {
                const int b = 6 * k;
                g[k] = fma2_abc(val(b, 0u), val(b + 1, 0u), val(b + 2, kOne), val(b + 3, kOne), val(b + 4, 0u), val(b + 5, 0u));
            }
            int n = G;
#pragma unroll
            while (n > 1) {
                const int m = (n + 2) / 3;
#pragma unroll
                for (int k = 0; k < m; ++k) {
                    const int b = 3 * k;
                    const unsigned long long one2 = ((unsigned long long)kOne << 32) | kOne;
                    g[k] = fma2_pairs(g[b], b + 1 < n ? g[b + 1] : one2, b + 2 < n ? g[b + 2] : 0ull);
                }
                n = m;
            }
            const float t = __uint_as_float((uint32_t)g[0]) + __uint_as_float((uint32_t)(g[0] >> 32));
            return !(t == t);
        } else {
            uint32_t s[2] = {0u, 0u};
#pragma unroll
            for (int i = 0; i < N; ++i) s[i & 1] = hfma2_zero(w[i], s[i & 1]);
            return ((s[0] | s[1]) & 0x7FFF7FFFu) != 0u;  // sums are +-0 or NaN per half
        }
    } else {
        if (BYTES == 4) {
            bool hit = false;
#pragma unroll
            for (int i = 0; i < N; ++i) hit |= (w[i] == mask_lo);
            return hit;
        } else if (BYTES == 8) {
            bool hit = false;
#pragma unroll
            for (int i = 1; i < N; i += 2) hit |= (w[i] == mask_hi);  // high words only: conservative
            return hit;
        } else {
            // any zero field in w ^ pattern  ("haszero": exact for the any-question)
            constexpr uint32_t ones = (BYTES == 2) ? 0x00010001u : 0x01010101u;
            constexpr uint32_t highs = (BYTES == 2) ? 0x80008000u : 0x80808080u;
            const uint32_t pat = mask_lo * ones;
            uint32_t acc = 0u;
#pragma unroll
            for (int i = 0; i < N; ++i) {
                const uint32_t x = w[i] ^ pat;
                acc |= (x - ones) & ~x;
            }
            return (acc & highs) != 0u;
        }
    }
}

// ---------------------------------------------------------------------------------
// whole-word element types (4 and 8 bytes): exact per-element test, and a cheap sufficient
// test for "every element of these words is masked" (NaN: all quiet NaNs; value: all equal)
// ---------------------------------------------------------------------------------
template <int KIND, int MASK>
__device__ __forceinline__ bool masked_elem(uint32_t lo, uint32_t hi, uint32_t mask_lo, uint32_t mask_hi) {
    // 4-byte types: the element is `lo`, `hi` is ignored
    if (MASK == kMaskNone) return false;
    if (Elem<KIND>::kBytes == 8) {
        if (MASK == kMaskNaN) {
            const double d = __hiloint2double((int)hi, (int)lo);
            return d != d;  // one DSETP; only dirty threads get here
        }
        return lo == mask_lo && hi == mask_hi;
    }
    if (MASK == kMaskNaN) {
        const float f = __uint_as_float(lo);
        return f != f;
    }
    return lo == mask_lo;
}
// Raw word(s) whose transformed image is all-zero bits: masked elements are overwritten with
// it so that they drop out of every plane sum (1.0 under the signed-exponent transform).
template <int KIND, bool XFORM>
__device__ __forceinline__ uint32_t neutral_word(int word_in_elem) {
    if (!XFORM) return 0u;
    if (KIND == kF32) return 0x3F800000u;
    if (KIND == kF64) return word_in_elem == 1 ? 0x3FF00000u : 0u;
    return 0u;
}
template <int KIND, int MASK, int N>
__device__ __forceinline__ bool all_masked_quick(const uint32_t (&w)[N], uint32_t mask_lo, uint32_t mask_hi) {
    constexpr int BYTES = Elem<KIND>::kBytes;
    static_assert(BYTES >= 4, "whole-word element types only");
    if (MASK == kMaskNaN) {
        constexpr int STEP = (BYTES == 8) ? 2 : 1;
        constexpr uint32_t quiet = (BYTES == 8) ? 0x7FF80000u : 0x7FC00000u;  // exponent ones + quiet bit
        uint32_t acc = 0xFFFFFFFFu;
#pragma unroll
        for (int i = STEP - 1; i < N; i += STEP) acc &= w[i];
        return (acc & quiet) == quiet;
    } else {
        uint32_t acc = 0u;
        if (BYTES == 8) {
#pragma unroll
            for (int i = 0; i < N; i += 2) acc |= (w[i] ^ mask_lo) | (w[i + 1] ^ mask_hi);
        } else {
#pragma unroll
            for (int i = 0; i < N; ++i) acc |= w[i] ^ mask_lo;
        }
        return acc == 0u;
    }
}

// ---------------------------------------------------------------------------------
// Boundary corrections of the rows kernel, warp-cooperative: a correction word is broadcast to the warp
// (__shfl_sync) and every lane L whose bit L of it is set bumps the block's shared total of bit position L.
// No bit planes, no registers, no local memory, no butterfly at the end; an event costs the whole warp a
// shuffle, a test and (for about half the lanes) one shared-memory atomic.  (The per-thread rippled planes in
// local memory this replaces cost land-masked arrays a tenth to a third of their throughput: every event was a
// chain of dependent local loads and stores that 31 lanes waited for.)  All calls warp-converged.
//   s_minus: shared [2][64] totals (stream a, stream b), bit number from the least significant bit, zeroed
//   before the first event; fold = min(bits per element, 32): packed types fold their fields onto one another.
//   The totals are 32-BIT words: a block never sees 2^32 elements (180 GB over 148 blocks), and a 32-bit shared
//   atomic add is one fire-and-forget ATOMS.ADD, where the 64-bit one is a load / compare-and-swap spin loop
//   (LDS.64 + ATOMS.CAST.SPIN.64 + branch) whose latency every event of a land-masked array paid in full.
// ---------------------------------------------------------------------------------
template <int STREAM, int NW>
__device__ __forceinline__ void warp_event_push(bool have, const uint32_t (&v)[NW], uint32_t lane,
                                                unsigned int (*s_minus)[64], uint32_t fold) {
    // STREAM 0: a valid / b masked (a-stream correction), 1: a masked / b valid
    unsigned hv = __ballot_sync(0xFFFFFFFFu, have);
    while (hv) {
        const int src = __ffs((int)hv) - 1;
        hv &= hv - 1u;
#pragma unroll
        for (int k = 0; k < NW; ++k)
            if ((__shfl_sync(0xFFFFFFFFu, v[k], src) >> lane) & 1u) atomicAdd(&s_minus[STREAM][k * 32 + (int)(lane % fold)], 1u);
    }
}

// ---------------------------------------------------------------------------------
// Boundary-correction counters: P bit planes per stream and element word, in registers.  A
// correction word is added with 2P logic operations.  Before a warp adds up to `upcoming`
// more events per lane it calls reserve(): if some lane could overflow, the whole warp
// flushes -- the planes are summed over the warp with the butterfly of xbi_common.cuh (lane L
// ends up with the count of bit position L) and added to the block's shared `minus` totals.
// No local memory, no per-event memory traffic; the cost of a flush (~400 instructions per
// stream) is paid once per 2^P - 1 - upcoming events of the busiest lane.
// ---------------------------------------------------------------------------------
template <int NW, int P>
struct EventPlanes {
    static constexpr uint32_t kCap = (1u << P) - 1u;
    uint32_t r[2][NW][P];
    uint32_t ecnt;  // events added by this lane since the last flush (bounds every bit position's count)

    __device__ __forceinline__ void init() {
#pragma unroll
        for (int s = 0; s < 2; ++s)
#pragma unroll
            for (int w = 0; w < NW; ++w)
#pragma unroll
                for (int j = 0; j < P; ++j) r[s][w][j] = 0u;
        ecnt = 0u;
    }
    // stream 0: a-stream correction ("a valid, b masked"), stream 1: b-stream correction
    template <int STREAM, int WI>
    __device__ __forceinline__ void add(uint32_t word) {
#pragma unroll
        for (int j = 0; j < P; ++j) {
            const uint32_t t = r[STREAM][WI][j] & word;
            r[STREAM][WI][j] ^= word;
            word = t;
        }
    }
    // warp-collective (all 32 lanes, converged): s_minus = shared [2][64] totals, bit number from the LSB
    __device__ __forceinline__ void flush(uint32_t lane, unsigned long long (*s_minus)[64], int nbits) {
#pragma unroll
        for (int s = 0; s < 2; ++s) {
#pragma unroll
            for (int w = 0; w < NW; ++w) {
                uint32_t n[P + 6];
#pragma unroll
                for (int j = 0; j < P; ++j) { n[j] = r[s][w][j]; r[s][w][j] = 0u; }
#pragma unroll
                for (int j = P; j < P + 6; ++j) n[j] = 0u;
                const uint32_t total = butterfly_totals(n, P, lane);
                if (total) atomicAdd(&s_minus[s][w * 32 + (int)(lane % (uint32_t)(nbits < 32 ? nbits : 32))], (unsigned long long)total);
            }
        }
        ecnt = 0u;
    }
    // warp-collective: make room for `upcoming` more events in every lane
    __device__ __forceinline__ void reserve(uint32_t upcoming, uint32_t lane, unsigned long long (*s_minus)[64], int nbits) {
        if (__any_sync(0xFFFFFFFFu, ecnt + upcoming > kCap)) flush(lane, s_minus, nbits);
    }
};

}  // namespace xbi
