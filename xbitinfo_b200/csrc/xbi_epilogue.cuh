// K2 arithmetic shared by the stand-alone epilogue kernel (xbi_epilogue.cu) and the fused finish
// kernels (xbi_finish.cu): mutual information of one bit from its 2x2 pair counts, plus the binomial
// significance filter.
#pragma once
#include <cmath>
#include <cuda_runtime.h>

namespace xbi {

// Follows xbitinfo/_py_bitinfo.py:143-152 operation by operation:
//   p = counts / size                      (line 147)
//   zero cells masked out                  (line 148)
//   pr = row sums, ps = column sums        (lines 149-150)
//   sum p * log(p / (pr * ps)) / log(2)    (line 151), summed in the order 00,01,10,11
// Round-to-nearest intrinsics keep ptxas from contracting mul+add into FMA.  0 pairs -> NaN.
__device__ __forceinline__ double mutual_information_bit(unsigned long long c00, unsigned long long c01,
                                                         unsigned long long c10, unsigned long long c11, int set_zero,
                                                         double z) {
    const unsigned long long n = c00 + c01 + c10 + c11;
    if (n == 0) return nan("");
    const double size = (double)n;
    const double p00 = __ddiv_rn((double)c00, size), p01 = __ddiv_rn((double)c01, size);
    const double p10 = __ddiv_rn((double)c10, size), p11 = __ddiv_rn((double)c11, size);
    const double pr0 = __dadd_rn(p00, p01), pr1 = __dadd_rn(p10, p11);
    const double ps0 = __dadd_rn(p00, p10), ps1 = __dadd_rn(p01, p11);
    double acc = 0.0;
    if (c00) acc = __dadd_rn(acc, __dmul_rn(p00, log(__ddiv_rn(p00, __dmul_rn(pr0, ps0)))));
    if (c01) acc = __dadd_rn(acc, __dmul_rn(p01, log(__ddiv_rn(p01, __dmul_rn(pr0, ps1)))));
    if (c10) acc = __dadd_rn(acc, __dmul_rn(p10, log(__ddiv_rn(p10, __dmul_rn(pr1, ps0)))));
    if (c11) acc = __dadd_rn(acc, __dmul_rn(p11, log(__ddiv_rn(p11, __dmul_rn(pr1, ps1)))));
    double v = __ddiv_rn(acc, 0.6931471805599453);  // np.log(2)
    if (set_zero) {
        // BitInformation.jl binom_confidence / binom_free_entropy:
        //   p = min(1, 1/2 + z / (2 sqrt(N))),  Hfree = 1 - H2(p),  M <= Hfree -> 0
        const double p = __dadd_rn(0.5, __ddiv_rn(z, __dmul_rn(2.0, sqrt(size))));
        double hfree = 1.0;
        if (p < 1.0) {
            const double q = __dadd_rn(1.0, -p);
            const double h = -__dadd_rn(__dmul_rn(p, log2(p)), __dmul_rn(q, log2(q)));
            hfree = __dadd_rn(1.0, -h);
        }
        if (v <= hfree) v = 0.0;
    }
    return v;
}

}  // namespace xbi
