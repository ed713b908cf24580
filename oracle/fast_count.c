/*
 * oracle/fast_count.c -- TEST INFRASTRUCTURE ONLY (see oracle/bitinfo_oracle.py).
 *
 * A plain-C, multi-threaded restatement of the reference's pair counting for inputs too large for the numpy
 * oracle (it needs ~90 bytes of temporaries per pair and byte plane): the full-size configurations of
 * BASELINE.json are checked BIT-EXACTLY against this on the GPU box's host cores.
 *
 * What it computes (reference: xbitinfo/_py_bitinfo.py):
 *   - the analysis word of an element: floats go through signed_exponent (42-94) -- the biased exponent field F
 *     becomes F - bias when F >= bias and ~F (within the field) otherwise, sign and mantissa untouched --
 *     integers are reinterpreted as unsigned (xbitinfo/xbitinfo.py:352-355);
 *   - pairs (X[j], X[j+1]) along one axis of a C-contiguous (outer, n, inner) array (155-160);
 *   - counts[k][a][b] = number of pairs whose bit k (k = 0 the most significant) is a in the first and b in the
 *     second element: the (nbits,2,2) layout of bitpaircount (97-140);
 *   - mask_mode 1 / 2: a pair with a NaN element / an element bit-identical to mask_bits is not counted (the
 *     BitInformation.jl semantics restated in bitinfo_oracle.py, parity-unpinned like there).
 * It is itself pinned: tests/test_oracle.py holds it to the numpy oracle (and through it to the reference's
 * golden vectors) on every dtype, axis and mask mode.
 *
 * Formulation: 64 pairs at a time, the analysis words of the first and of the second elements (zeroed for
 * masked pairs) are transposed into bit planes (64 x 64 bit-matrix transpose) and three popcounts per bit give
 * A (first element set), B (second set) and AB (both); n = valid pairs; the 2x2 table follows from A, B, AB and
 * n.  Nothing here is shared with the CUDA kernels or the numpy oracle.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

typedef struct {
    int nbits, nmant, is_float, mask_mode; /* mask_mode: 0 none, 1 NaN, 2 value */
    uint64_t mask_bits, field, one, mant_mask;
} spec_t;

static inline uint64_t load_elem(const unsigned char* p, int bytes) {
    switch (bytes) {
        case 1: return *p;
        case 2: { uint16_t v; memcpy(&v, p, 2); return v; }
        case 4: { uint32_t v; memcpy(&v, p, 4); return v; }
        default: { uint64_t v; memcpy(&v, p, 8); return v; }
    }
}

static inline int is_masked(uint64_t raw, const spec_t* s) {
    if (s->mask_mode == 1) return ((raw & s->field) == s->field) && ((raw & s->mant_mask) != 0);
    if (s->mask_mode == 2) return raw == s->mask_bits;
    return 0;
}

static inline uint64_t analysis_word(uint64_t raw, const spec_t* s) {
    if (!s->is_float) return raw;
    return ((raw & s->field) >= s->one) ? raw - s->one : raw ^ s->field;
}

/* 64 x 64 bit-matrix transpose in place (recursive block swaps): afterwards bit i of w[j] is what bit j of w[i]
 * was, i.e. w[j] collects bit j of the 64 words. */
static void transpose64(uint64_t w[64]) {
    uint64_t m = 0x00000000FFFFFFFFull;
    for (int j = 32; j != 0; j >>= 1, m ^= (m << j)) {
        for (int k = 0; k < 64; k = (k + j + 1) & ~j) {
            const uint64_t t = ((w[k] >> j) ^ w[k + j]) & m;
            w[k] ^= t << j;
            w[k + j] ^= t;
        }
    }
}

/* pairs (base + f, base + f + inner) for f in [lo, hi): 64 pairs at a time are turned into bit planes and
 * counted with popcounts (A: first elements, B: second elements, AB: both) */
static void count_range(const unsigned char* x, int bytes, int64_t base, int64_t lo, int64_t hi, int64_t inner,
                        const spec_t* s, uint64_t* A, uint64_t* B, uint64_t* AB, uint64_t* N) {
    const int nbits = s->nbits;
    uint64_t nv = 0;
    for (int64_t f0 = lo; f0 < hi; f0 += 64) {
        const int cnt = hi - f0 < 64 ? (int)(hi - f0) : 64;
        uint64_t wa[64], wb[64];
        for (int i = 0; i < cnt; ++i) {
            const uint64_t ra = load_elem(x + (size_t)(base + f0 + i) * bytes, bytes);
            const uint64_t rb = load_elem(x + (size_t)(base + f0 + i + inner) * bytes, bytes);
            const int ok = !(is_masked(ra, s) || is_masked(rb, s));
            const uint64_t keep = ok ? ~(uint64_t)0 : 0;
            wa[i] = analysis_word(ra, s) & keep;
            wb[i] = analysis_word(rb, s) & keep;
            nv += (uint64_t)ok;
        }
        for (int i = cnt; i < 64; ++i) { wa[i] = 0; wb[i] = 0; }
        transpose64(wa);
        transpose64(wb);
        for (int j = 0; j < nbits; ++j) {
            A[j] += (uint64_t)__builtin_popcountll(wa[j]);
            B[j] += (uint64_t)__builtin_popcountll(wb[j]);
            AB[j] += (uint64_t)__builtin_popcountll(wa[j] & wb[j]);
        }
    }
    *N += nv;
}

/*
 * counts: nbits*4 values, counts[4*k + 2*a + b], k = 0 the most significant bit.  Overwritten.
 * bytes: 1, 2, 4 or 8; is_float: IEEE half / single / double for bytes 2 / 4 / 8.
 * Returns 0, or 1 for bad arguments.
 */
int fast_count_pairs(const void* data, int bytes, int is_float, int64_t outer, int64_t n, int64_t inner, int mask_mode,
                     uint64_t mask_bits, uint64_t* counts, int threads) {
    if (!(bytes == 1 || bytes == 2 || bytes == 4 || bytes == 8) || (is_float && bytes == 1)) return 1;
    if (outer < 0 || n < 0 || inner < 0 || mask_mode < 0 || mask_mode > 2 || (mask_mode == 1 && !is_float)) return 1;
    spec_t s;
    s.nbits = 8 * bytes;
    s.is_float = is_float;
    s.mask_mode = mask_mode;
    s.nmant = bytes == 2 ? 10 : bytes == 4 ? 23 : 52;
    const int nexp = s.nbits - 1 - s.nmant;
    s.field = is_float ? ((((uint64_t)1 << nexp) - 1) << s.nmant) : 0;
    s.one = is_float ? ((((uint64_t)1 << (nexp - 1)) - 1) << s.nmant) : 0; /* bias << nmant */
    s.mant_mask = is_float ? (((uint64_t)1 << s.nmant) - 1) : 0;
    s.mask_bits = bytes == 8 ? mask_bits : (mask_bits & (((uint64_t)1 << s.nbits) - 1));
    const int nbits = s.nbits;
    uint64_t A[64] = {0}, B[64] = {0}, AB[64] = {0}, N = 0;
    if (outer > 0 && n >= 2 && inner > 0) {
        const int64_t per_block = (n - 1) * inner;           /* pairs per outer index */
        const int64_t piece = 1 << 22;                       /* pairs per task */
        const int64_t pieces_per_block = (per_block + piece - 1) / piece;
        const int64_t tasks = outer * pieces_per_block;
        const unsigned char* x = (const unsigned char*)data;
#ifdef _OPENMP
        if (threads > 0) omp_set_num_threads(threads);
#pragma omp parallel
#endif
        {
            uint64_t tA[64] = {0}, tB[64] = {0}, tAB[64] = {0}, tN = 0;
#ifdef _OPENMP
#pragma omp for schedule(dynamic, 4)
#endif
            for (int64_t t = 0; t < tasks; ++t) {
                const int64_t o = t / pieces_per_block, pc = t - o * pieces_per_block;
                const int64_t lo = pc * piece, hi = lo + piece < per_block ? lo + piece : per_block;
                count_range(x, bytes, o * n * inner, lo, hi, inner, &s, tA, tB, tAB, &tN);
            }
#ifdef _OPENMP
#pragma omp critical
#endif
            {
                for (int j = 0; j < nbits; ++j) { A[j] += tA[j]; B[j] += tB[j]; AB[j] += tAB[j]; }
                N += tN;
            }
        }
    }
    for (int j = 0; j < nbits; ++j) {
        const int k = nbits - 1 - j;
        counts[4 * k + 0] = N - A[j] - B[j] + AB[j];
        counts[4 * k + 1] = B[j] - AB[j];
        counts[4 * k + 2] = A[j] - AB[j];
        counts[4 * k + 3] = AB[j];
    }
    return 0;
}
