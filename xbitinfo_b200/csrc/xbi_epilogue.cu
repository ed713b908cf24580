// K2 as a stand-alone kernel (mutual information + binomial significance filter on stacks of nbits x 4 counts;
// the single-table case normally runs inside the finish kernel, xbi_finish.cu) and the normal quantile.
#include <cmath>

#include "xbi_common.cuh"
#include "xbi_epilogue.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

// One thread per (table, bit): `total` = number of tables x nbits.  Arithmetic: xbi_epilogue.cuh.
__global__ void k_mutual_information(const unsigned long long* __restrict__ counts, long long total, int set_zero,
                                     double z, double* __restrict__ mi) {
    const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= total) return;
    mi[k] = mutual_information_bit(counts[4 * k + 0], counts[4 * k + 1], counts[4 * k + 2], counts[4 * k + 3], set_zero, z);
}

}  // namespace

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
