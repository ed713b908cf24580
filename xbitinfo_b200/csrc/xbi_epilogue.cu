// Combine kernel (raw plane sums -> (nbits,2,2) counts) and K2 (mutual information +
// binomial significance filter on the nbits x 4 counts).
#include <cmath>

#include "xbi_common.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

// counts[4k + 2a + b], k = 0 is the most significant bit (bitpaircount layout,
// xbitinfo/_py_bitinfo.py:128-140).  Flat pairs minus the pairs that straddle two blocks
// of the analysed axis are exactly the pairs of _py_bitinfo.py:155-160.
__global__ void k_combine(const Workspace* __restrict__ ws, int nbits, unsigned long long* __restrict__ counts) {
    const int j = threadIdx.x;  // bit number from the least significant bit
    if (j >= nbits) return;
    const unsigned long long A = ws->plus.sums[S_A][j] - ws->minus.sums[S_A][j];
    const unsigned long long B = ws->plus.sums[S_B][j] - ws->minus.sums[S_B][j];
    const unsigned long long AB = ws->plus.sums[S_AB][j] - ws->minus.sums[S_AB][j];
    const unsigned long long N = ws->plus.npairs - ws->minus.npairs;
    const int k = nbits - 1 - j;
    // atomics: several streams may add slabs of the same variable into one counts buffer
    atomicAdd(&counts[4 * k + 0], N - A - B + AB);
    atomicAdd(&counts[4 * k + 1], B - AB);
    atomicAdd(&counts[4 * k + 2], A - AB);
    atomicAdd(&counts[4 * k + 3], AB);
}

// One thread per bit.  Follows xbitinfo/_py_bitinfo.py:143-152 operation by operation:
//   p = counts / size                      (line 147)
//   zero cells masked out                  (line 148)
//   pr = row sums, ps = column sums        (lines 149-150)
//   sum p * log(p / (pr * ps)) / log(2)    (line 151), summed in the order 00,01,10,11
// Round-to-nearest intrinsics keep ptxas from contracting mul+add into FMA.
// One thread per (table, bit): `total` = number of tables x nbits.
__global__ void k_mutual_information(const unsigned long long* __restrict__ counts, long long total, int set_zero,
                                     double z, double* __restrict__ mi) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    const unsigned long long c00 = counts[4 * k + 0], c01 = counts[4 * k + 1];
    const unsigned long long c10 = counts[4 * k + 2], c11 = counts[4 * k + 3];
    const unsigned long long n = c00 + c01 + c10 + c11;
    if (n == 0) {
        mi[k] = nan("");
        return;
    }
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
        double p = __dadd_rn(0.5, __ddiv_rn(z, __dmul_rn(2.0, sqrt(size))));
        double hfree = 1.0;
        if (p < 1.0) {
            const double q = __dadd_rn(1.0, -p);
            const double h = -__dadd_rn(__dmul_rn(p, log2(p)), __dmul_rn(q, log2(q)));
            hfree = __dadd_rn(1.0, -h);
        }
        if (v <= hfree) v = 0.0;
    }
    mi[k] = v;
}

}  // namespace

cudaError_t launch_combine(const Workspace* ws, int nbits, unsigned long long* counts, cudaStream_t stream) {
    k_combine<<<1, 64, 0, stream>>>(ws, nbits, counts);
    note_launch();
    return cudaGetLastError();
}

cudaError_t launch_mutual_information(const unsigned long long* counts, int nbits, long long nbatch, int set_zero,
                                      double z_quantile, double* mi, cudaStream_t stream) {
    const long long total = nbatch * nbits;
    if (total <= 0) return cudaSuccess;
    const int threads = total < 256 ? 64 : 256;
    k_mutual_information<<<(unsigned)((total + threads - 1) / threads), threads, 0, stream>>>(counts, total, set_zero,
                                                                                             z_quantile, mi);
    note_launch();
    return cudaGetLastError();
}

// Wichura's algorithm AS241 (PPND16): normal quantile, relative accuracy ~1e-16.
double normal_quantile(double p) {
    const double q = p - 0.5;
    if (std::fabs(q) <= 0.425) {
        const double r = 0.180625 - q * q;
        const double num = (((((((2.5090809287301226727e3 * r + 3.3430575583588128105e4) * r + 6.7265770927008700853e4) * r +
                                4.5921953931549871457e4) * r + 1.3731693765509461125e4) * r + 1.9715909503065514427e3) * r +
                             1.3314166789178437745e2) * r + 3.3871328727963666080e0) * q;
        const double den = ((((((5.2264952788528545610e3 * r + 2.8729085735721942674e4) * r + 3.9307895800092710610e4) * r +
                               2.1213794301586595867e4) * r + 5.3941960214247511077e3) * r + 6.8718700749205790830e2) * r +
                            4.2313330701600911252e1) * r + 1.0;
        return num / den;
    }
    double r = q <= 0.0 ? p : 1.0 - p;
    r = std::sqrt(-std::log(r));
    double x;
    if (r <= 5.0) {
        r -= 1.6;
        const double num = ((((((7.74545014278341407640e-4 * r + 2.27238449892691845833e-2) * r + 2.41780725177450611770e-1) * r +
                               1.27045825245236838258e0) * r + 3.64784832476320460504e0) * r + 5.76949722146069140550e0) * r +
                            4.63033784615654529590e0) * r + 1.42343711074968357734e0;
        const double den = ((((((1.05075007164441684324e-9 * r + 5.47593808499534494600e-4) * r + 1.51986665636164571966e-2) * r +
                               1.48103976427480074590e-1) * r + 6.89767334985100004550e-1) * r + 1.67638483018380384940e0) * r +
                            2.05319162663775882187e0) * r + 1.0;
        x = num / den;
    } else {
        r -= 5.0;
        const double num = ((((((2.01033439929228813265e-7 * r + 2.71155556874348757815e-5) * r + 1.24266094738807843860e-3) * r +
                               2.65321895265761230930e-2) * r + 2.96560571828504891230e-1) * r + 1.78482653991729133580e0) * r +
                            5.46378491116411436990e0) * r + 6.65790464350110377720e0;
        const double den = ((((((2.04426310338993978564e-15 * r + 1.42151175831644588870e-7) * r + 1.84631831751005468180e-5) * r +
                               7.86869131145613259100e-4) * r + 1.48753612908506148525e-2) * r + 1.36929880922735805310e-1) * r +
                            5.99832206555887937690e-1) * r + 1.0;
        x = num / den;
    }
    return q < 0.0 ? -x : x;
}

}  // namespace xbi
