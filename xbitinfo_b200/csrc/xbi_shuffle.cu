// K4: byte shuffle and bit shuffle -- the lossless-compression pre-filters that follow
// bitround on the reference's save path (xbitinfo/save_compressed.py:44-80: netCDF
// `shuffle=True` = the HDF5 shuffle filter; :155-209: zarr Blosc with SHUFFLE / BITSHUFFLE).
// Keeping them on the device lets the rounded array leave HBM already in the layout the
// entropy coder wants.
//
//   byte shuffle   dst[j*N + i]            = src[i*S + j]          (N elements of S bytes)
//   bit shuffle    bit (i%8) of dst[p*N/8 + i/8] = bit p of element i   (p = 8*byte + bit,
//                  least significant bit first; N a multiple of 8)
//
// Both are pure streaming permutations: 2 x N x S bytes of HBM traffic.
//   * byte shuffle: a thread takes 16 consecutive elements (S 128-bit loads), permutes the
//     bytes with PRMT and writes 16 bytes to each of the S planes (a warp writes 512
//     contiguous bytes per plane).
//   * bit shuffle: a warp takes a tile of 1024 elements; the words are loaded coalesced,
//     turned through padded shared memory so that a lane holds 32 consecutive elements,
//     transposed as a 32x32 bit matrix in registers (5 x 16 masked swaps) and each lane
//     writes one word per bit plane (a warp writes 128 contiguous bytes per plane).
#include "xbi_common.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

constexpr int kThreads = 256;

// ---------------------------------------------------------------------------------
// optional fused bitround (K3's arithmetic, xbi_bitround.cu) applied to the words as they are
// loaded by the forward kernels: round-then-shuffle in one pass over HBM instead of two
// ---------------------------------------------------------------------------------
struct RoundSpec {
    uint32_t drop;                // mantissa bits to drop; 0 = no rounding
    unsigned long long half_m1;   // 2^(drop-1) - 1
    unsigned long long keep;      // ~(2^drop - 1) within the element width
};
template <int S>
__device__ __forceinline__ void round_words(uint32_t* w, int nwords, const RoundSpec& r) {
    // w: consecutive little-endian words of whole elements of S bytes
    if (S == 4) {
#pragma unroll
        for (int i = 0; i < nwords; ++i) w[i] = (w[i] + ((w[i] >> r.drop) & 1u) + (uint32_t)r.half_m1) & (uint32_t)r.keep;
    } else if (S == 2) {
#pragma unroll
        for (int i = 0; i < nwords; ++i) {
            const uint32_t lo = w[i] & 0xFFFFu, hi = w[i] >> 16;
            const uint32_t rl = (lo + ((lo >> r.drop) & 1u) + (uint32_t)r.half_m1) & (uint32_t)r.keep;
            const uint32_t rh = (hi + ((hi >> r.drop) & 1u) + (uint32_t)r.half_m1) & (uint32_t)r.keep;
            w[i] = (rl & 0xFFFFu) | (rh << 16);
        }
    } else {
#pragma unroll
        for (int i = 0; i + 1 < nwords; i += 2) {
            unsigned long long b = ((unsigned long long)w[i + 1] << 32) | w[i];
            b = (b + ((b >> r.drop) & 1ull) + r.half_m1) & r.keep;
            w[i] = (uint32_t)b;
            w[i + 1] = (uint32_t)(b >> 32);
        }
    }
}
// one element given as S bytes (byte-wise tails)
template <int S>
__device__ __forceinline__ void round_bytes(uint8_t (&e)[S], const RoundSpec& r) {
    unsigned long long b = 0;
#pragma unroll
    for (int j = 0; j < S; ++j) b |= (unsigned long long)e[j] << (8 * j);
    b = (b + ((b >> r.drop) & 1ull) + r.half_m1) & r.keep;
#pragma unroll
    for (int j = 0; j < S; ++j) e[j] = (uint8_t)(b >> (8 * j));
}

// ---------------------------------------------------------------------------------
// byte shuffle / unshuffle
// ---------------------------------------------------------------------------------
// 16 elements of S bytes held as 4*S words  <->  S planes of 16 bytes (4 words each)
template <int S>
__device__ __forceinline__ void bytes_to_planes(const uint32_t (&in)[4 * S], uint32_t (&out)[S][4]) {
    if constexpr (S == 2) {
        // word = (e0.b0 e0.b1 e1.b0 e1.b1); plane j word q = byte j of elements 4q..4q+3
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            out[0][q] = __byte_perm(in[2 * q], in[2 * q + 1], 0x6420);
            out[1][q] = __byte_perm(in[2 * q], in[2 * q + 1], 0x7531);
        }
    } else if constexpr (S == 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t a = in[4 * q], b = in[4 * q + 1], c = in[4 * q + 2], d = in[4 * q + 3];
            const uint32_t ab02 = __byte_perm(a, b, 0x6240);  // a0 b0 a2 b2
            const uint32_t ab13 = __byte_perm(a, b, 0x7351);  // a1 b1 a3 b3
            const uint32_t cd02 = __byte_perm(c, d, 0x6240);
            const uint32_t cd13 = __byte_perm(c, d, 0x7351);
            out[0][q] = __byte_perm(ab02, cd02, 0x5410);  // a0 b0 c0 d0
            out[2][q] = __byte_perm(ab02, cd02, 0x7632);  // a2 b2 c2 d2
            out[1][q] = __byte_perm(ab13, cd13, 0x5410);
            out[3][q] = __byte_perm(ab13, cd13, 0x7632);
        }
    } else {  // S == 8: element e = words 2e (bytes 0..3), 2e+1 (bytes 4..7)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t a = in[8 * q + h], b = in[8 * q + 2 + h], c = in[8 * q + 4 + h], d = in[8 * q + 6 + h];
                const uint32_t ab02 = __byte_perm(a, b, 0x6240);
                const uint32_t ab13 = __byte_perm(a, b, 0x7351);
                const uint32_t cd02 = __byte_perm(c, d, 0x6240);
                const uint32_t cd13 = __byte_perm(c, d, 0x7351);
                out[4 * h + 0][q] = __byte_perm(ab02, cd02, 0x5410);
                out[4 * h + 2][q] = __byte_perm(ab02, cd02, 0x7632);
                out[4 * h + 1][q] = __byte_perm(ab13, cd13, 0x5410);
                out[4 * h + 3][q] = __byte_perm(ab13, cd13, 0x7632);
            }
        }
    }
}
template <int S>
__device__ __forceinline__ void planes_to_bytes(const uint32_t (&in)[S][4], uint32_t (&out)[4 * S]) {
    if constexpr (S == 2) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            out[2 * q] = __byte_perm(in[0][q], in[1][q], 0x5140);      // p0.0 p1.0 p0.1 p1.1
            out[2 * q + 1] = __byte_perm(in[0][q], in[1][q], 0x7362);  // p0.2 p1.2 p0.3 p1.3
        }
    } else if constexpr (S == 4) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const uint32_t p0 = in[0][q], p1 = in[1][q], p2 = in[2][q], p3 = in[3][q];
            const uint32_t lo01 = __byte_perm(p0, p1, 0x5140);  // a0 a1 b0 b1
            const uint32_t hi01 = __byte_perm(p0, p1, 0x7362);  // c0 c1 d0 d1
            const uint32_t lo23 = __byte_perm(p2, p3, 0x5140);  // a2 a3 b2 b3
            const uint32_t hi23 = __byte_perm(p2, p3, 0x7362);
            out[4 * q + 0] = __byte_perm(lo01, lo23, 0x5410);  // a0 a1 a2 a3
            out[4 * q + 1] = __byte_perm(lo01, lo23, 0x7632);  // b
            out[4 * q + 2] = __byte_perm(hi01, hi23, 0x5410);  // c
            out[4 * q + 3] = __byte_perm(hi01, hi23, 0x7632);  // d
        }
    } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint32_t p0 = in[4 * h][q], p1 = in[4 * h + 1][q], p2 = in[4 * h + 2][q], p3 = in[4 * h + 3][q];
                const uint32_t lo01 = __byte_perm(p0, p1, 0x5140);
                const uint32_t hi01 = __byte_perm(p0, p1, 0x7362);
                const uint32_t lo23 = __byte_perm(p2, p3, 0x5140);
                const uint32_t hi23 = __byte_perm(p2, p3, 0x7362);
                out[8 * q + h] = __byte_perm(lo01, lo23, 0x5410);
                out[8 * q + 2 + h] = __byte_perm(lo01, lo23, 0x7632);
                out[8 * q + 4 + h] = __byte_perm(hi01, hi23, 0x5410);
                out[8 * q + 6 + h] = __byte_perm(hi01, hi23, 0x7632);
            }
        }
    }
}

// groups of 16 elements; N = 16*ngroups + tail, the tail is moved byte by byte
template <int S, bool FORWARD>
__global__ void __launch_bounds__(kThreads)
k_byte_shuffle(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long n, long long ngroups, RoundSpec rs) {
    const long long stride = (long long)gridDim.x * kThreads;
    if (FORWARD) {
        for (long long g = (long long)blockIdx.x * kThreads + threadIdx.x; g < ngroups; g += stride) {
            uint32_t in[4 * S], out[S][4];
            const uint4* p = reinterpret_cast<const uint4*>(src) + g * S;
#pragma unroll
            for (int k = 0; k < S; ++k) {
                const uint4 v = __ldg(p + k);
                in[4 * k] = v.x; in[4 * k + 1] = v.y; in[4 * k + 2] = v.z; in[4 * k + 3] = v.w;
            }
            if (rs.drop) {
#pragma unroll
                for (int k = 0; k < 4 * S; k += 2) round_words<S>(in + k, 2, rs);
            }
            bytes_to_planes<S>(in, out);
#pragma unroll
            for (int j = 0; j < S; ++j)
                *reinterpret_cast<uint4*>(dst + j * n + 16 * g) = make_uint4(out[j][0], out[j][1], out[j][2], out[j][3]);
        }
    } else {
        // A lane rebuilds 16 elements = S consecutive 16-byte vectors; written directly, a warp's store
        // instruction would scatter 32 vectors 16*S bytes apart (half-filled sectors for S >= 4).  The
        // vectors are turned through shared memory (rows of S vectors, pitch S+1: conflict free both ways)
        // so that every store instruction of the warp covers 512 contiguous bytes.
        __shared__ uint4 turn[kThreads / 32][32 * (S + 1)];
        const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
        uint4* tw = turn[warp];
        for (long long g0 = ((long long)blockIdx.x * kThreads + threadIdx.x) & ~31ll; g0 < ngroups; g0 += stride) {
            const long long g = g0 + lane;
            if (g < ngroups) {
                uint32_t in[S][4], out[4 * S];
#pragma unroll
                for (int j = 0; j < S; ++j) {
                    const uint4 v = __ldg(reinterpret_cast<const uint4*>(src + j * n + 16 * g));
                    in[j][0] = v.x; in[j][1] = v.y; in[j][2] = v.z; in[j][3] = v.w;
                }
                planes_to_bytes<S>(in, out);
#pragma unroll
                for (int k = 0; k < S; ++k)
                    tw[lane * (S + 1) + k] = make_uint4(out[4 * k], out[4 * k + 1], out[4 * k + 2], out[4 * k + 3]);
            }
            __syncwarp();
            const long long nvec = (ngroups - g0 < 32 ? ngroups - g0 : 32) * S;  // vectors this warp holds
            uint4* p = reinterpret_cast<uint4*>(dst) + g0 * S;
#pragma unroll
            for (int k = 0; k < S; ++k) {
                const uint32_t f = k * 32 + lane;  // vector index in the warp's output
                if (f < nvec) p[f] = tw[(f / S) * (S + 1) + (f % S)];
            }
            __syncwarp();
        }
    }
    // elements outside the 16-element groups (unaligned buffers take this path entirely)
    for (long long i = 16 * ngroups + (long long)blockIdx.x * kThreads + threadIdx.x; i < n; i += stride) {
        if (FORWARD) {
            uint8_t e[S];
#pragma unroll
            for (int j = 0; j < S; ++j) e[j] = src[i * S + j];
            if (rs.drop) round_bytes<S>(e, rs);
#pragma unroll
            for (int j = 0; j < S; ++j) dst[j * n + i] = e[j];
        } else {
#pragma unroll
            for (int j = 0; j < S; ++j) dst[i * S + j] = src[j * n + i];
        }
    }
}

// ---------------------------------------------------------------------------------
// bit shuffle / unshuffle
// ---------------------------------------------------------------------------------
// 32x32 bit-matrix transpose in registers, least significant bit = column 0:
// afterwards a[c] bit r == (before) a[r] bit c.
__device__ __forceinline__ void transpose32(uint32_t (&a)[32]) {
#pragma unroll
    for (int j = 16; j >= 1; j >>= 1) {
        const uint32_t m = (j == 16) ? 0x0000FFFFu : (j == 8) ? 0x00FF00FFu : (j == 4) ? 0x0F0F0F0Fu
                          : (j == 2) ? 0x33333333u : 0x55555555u;
#pragma unroll
        for (int k = 0; k < 32; ++k) {
            if ((k & j) == 0) {
                const uint32_t t = ((a[k] >> j) ^ a[k + j]) & m;
                a[k + j] ^= t;
                a[k] ^= t << j;
            }
        }
    }
}

constexpr int kBsWarps = kThreads / 32;
constexpr int kPitch = 36;  // words per staged row of 32 words: 128-bit accesses stay conflict free

// One warp-tile = 1024 elements of 4 bytes' worth of words: W = S/4 word planes for S >= 4
// (low / high words of doubles), two 16-bit elements per word for S == 2.
//   S == 4: tile = 1024 elements; lane t ends up with elements 32t..32t+31 -> word t of 32 planes
//   S == 8: tile = 1024 elements; low words -> planes 0..31, high words -> planes 32..63
//   S == 2: tile = 2048 elements; lane t holds elements 64t..64t+63 -> words 2t, 2t+1 of 16 planes
template <int S, bool FORWARD>
__global__ void __launch_bounds__(kThreads)
k_bit_shuffle(const uint32_t* __restrict__ src, uint32_t* __restrict__ dst, long long n, long long ntiles, RoundSpec rs) {
    constexpr int EPT = (S == 2) ? 2048 : 1024;    // elements per warp tile
    constexpr int WPT = EPT * S / 4;               // words per warp tile (1024 or 2048)
    constexpr int NSUB = WPT / 1024;               // 32x32-word sub-tiles per warp tile
    __shared__ __align__(16) uint32_t stage[kBsWarps][32 * kPitch];
    const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31u;
    uint32_t* st = stage[warp];
    const long long plane_words = n / 32;  // words per bit plane
    const long long wstride = (long long)gridDim.x * kBsWarps;
    for (long long tile = (long long)blockIdx.x * kBsWarps + warp; tile < ntiles; tile += wstride) {
        if (FORWARD) {
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                // S == 8: sub 0 gathers the low words, sub 1 the high words of the tile's elements
                uint32_t a[32];
                if constexpr (S == 8) {
                    // coalesced: lane reads uint4 = (lo0 hi0 lo1 hi1) of elements 2*(lane+32k), +1
                    const uint4* p = reinterpret_cast<const uint4*>(src + tile * WPT);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        uint4 v = __ldg(p + lane + 32 * k);  // elements 2*(lane+32k), 2*(lane+32k)+1
                        if (rs.drop) round_words<S>(&v.x, 4, rs);
                        const uint32_t e = 2 * (lane + 32 * k);    // element index in the tile
                        const uint32_t w0 = sub == 0 ? v.x : v.y, w1 = sub == 0 ? v.z : v.w;
                        // element e lives in staged row e/32, column e%32
                        *reinterpret_cast<uint2*>(&st[(e >> 5) * kPitch + (e & 31u)]) = make_uint2(w0, w1);
                    }
                } else {
                    const uint4* p = reinterpret_cast<const uint4*>(src + tile * WPT + sub * 1024);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        uint4 v = __ldg(p + lane + 32 * k);
                        if (rs.drop) round_words<S>(&v.x, 4, rs);
                        const uint32_t w = 4 * (lane + 32 * k);  // word index in the sub-tile
                        *reinterpret_cast<uint4*>(&st[(w >> 5) * kPitch + (w & 31u)]) = v;
                    }
                }
                __syncwarp();
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint4 v = *reinterpret_cast<const uint4*>(&st[lane * kPitch + 4 * k]);
                    a[4 * k] = v.x; a[4 * k + 1] = v.y; a[4 * k + 2] = v.z; a[4 * k + 3] = v.w;
                }
                __syncwarp();
                if constexpr (S == 2) {
                    // 64 elements in 32 words (e[2r] | e[2r+1] << 16) -> word r = e[r] | e[r+32] << 16
                    uint32_t b[32];
#pragma unroll
                    for (int r = 0; r < 32; ++r) {
                        const uint32_t lo = a[r >> 1], hi = a[16 + (r >> 1)];
                        b[r] = (r & 1) ? __byte_perm(lo, hi, 0x7632) : __byte_perm(lo, hi, 0x5410);
                    }
                    transpose32(b);
                    // b[c], c < 16: bit c of elements 0..31 of this lane's 64; b[c+16]: of elements 32..63
                    const long long w0 = tile * (EPT / 32) + sub * 0 + 2 * lane;
#pragma unroll
                    for (int c = 0; c < 16; ++c)
                        *reinterpret_cast<uint2*>(&dst[c * plane_words + w0]) = make_uint2(b[c], b[c + 16]);
                } else {
                    transpose32(a);
                    const long long w0 = tile * (EPT / 32) + lane;
                    const int pbase = (S == 8) ? 32 * sub : 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) dst[(pbase + c) * plane_words + w0] = a[c];
                }
            }
        } else {
            uint32_t lo64[S == 8 ? 32 : 1];  // S == 8: low words of this lane's 16 output vectors
            (void)lo64;
#pragma unroll
            for (int sub = 0; sub < NSUB; ++sub) {
                uint32_t a[32];
                if constexpr (S == 2) {
                    uint32_t b[32];
                    const long long w0 = tile * (EPT / 32) + 2 * lane;
#pragma unroll
                    for (int c = 0; c < 16; ++c) {
                        const uint2 v = *reinterpret_cast<const uint2*>(&src[c * plane_words + w0]);
                        b[c] = v.x;
                        b[c + 16] = v.y;
                    }
                    transpose32(b);  // b[r] = e[r] | e[r+32] << 16
#pragma unroll
                    for (int q = 0; q < 16; ++q) {
                        a[q] = __byte_perm(b[2 * q], b[2 * q + 1], 0x5410);       // e[2q] | e[2q+1] << 16
                        a[16 + q] = __byte_perm(b[2 * q], b[2 * q + 1], 0x7632);  // e[32+2q] | e[32+2q+1] << 16
                    }
                } else {
                    const long long w0 = tile * (EPT / 32) + lane;
                    const int pbase = (S == 8) ? 32 * sub : 0;
#pragma unroll
                    for (int c = 0; c < 32; ++c) a[c] = __ldg(&src[(pbase + c) * plane_words + w0]);
                    transpose32(a);
                }
                // lane holds 32 consecutive words of the (sub-)tile: back through shared memory
#pragma unroll
                for (int k = 0; k < 8; ++k)
                    *reinterpret_cast<uint4*>(&st[lane * kPitch + 4 * k]) = make_uint4(a[4 * k], a[4 * k + 1], a[4 * k + 2], a[4 * k + 3]);
                __syncwarp();
                if constexpr (S == 8) {
                    // the low words of the tile's elements wait in registers for the high words; then
                    // 16-byte vectors (lo hi lo hi) of elements 2m, 2m+1 go out, 512 contiguous bytes per
                    // store instruction of the warp
                    uint4* q = reinterpret_cast<uint4*>(dst + tile * WPT);
#pragma unroll
                    for (int k = 0; k < 16; ++k) {
                        const uint32_t m = lane + 32 * k, e = 2 * m;  // vector / first element in the tile
                        const uint2 v = *reinterpret_cast<const uint2*>(&st[(e >> 5) * kPitch + (e & 31u)]);
                        if (sub == 0) { lo64[2 * k] = v.x; lo64[2 * k + 1] = v.y; }
                        else q[m] = make_uint4(lo64[2 * k], v.x, lo64[2 * k + 1], v.y);
                    }
                } else {
                    uint4* p = reinterpret_cast<uint4*>(dst + tile * WPT + sub * 1024);
#pragma unroll
                    for (int k = 0; k < 8; ++k) {
                        const uint32_t w = 4 * (lane + 32 * k);
                        p[lane + 32 * k] = *reinterpret_cast<const uint4*>(&st[(w >> 5) * kPitch + (w & 31u)]);
                    }
                }
                __syncwarp();
            }
        }
    }
}

// elements [lo, n) of a bit shuffle that do not fill a warp tile: one thread per 8 elements
template <int S, bool FORWARD>
__global__ void __launch_bounds__(kThreads)
k_bit_shuffle_tail(const uint8_t* __restrict__ src, uint8_t* __restrict__ dst, long long n, long long lo, RoundSpec rs) {
    const long long nb = n / 8;  // bytes per bit plane
    const long long stride = (long long)gridDim.x * kThreads;
    for (long long g = lo / 8 + (long long)blockIdx.x * kThreads + threadIdx.x; g < nb; g += stride) {
        if (FORWARD) {
            uint8_t e[8][S];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < S; ++j) e[i][j] = src[(8 * g + i) * S + j];
            if (rs.drop) {
#pragma unroll
                for (int i = 0; i < 8; ++i) round_bytes<S>(e[i], rs);
            }
#pragma unroll
            for (int j = 0; j < S; ++j)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    uint32_t b = 0;
#pragma unroll
                    for (int i = 0; i < 8; ++i) b |= ((e[i][j] >> k) & 1u) << i;
                    dst[(8 * j + k) * nb + g] = (uint8_t)b;
                }
        } else {
            uint8_t e[8][S];
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < S; ++j) e[i][j] = 0;
#pragma unroll
            for (int j = 0; j < S; ++j)
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    const uint32_t b = src[(8 * j + k) * nb + g];
#pragma unroll
                    for (int i = 0; i < 8; ++i) e[i][j] |= (uint8_t)(((b >> i) & 1u) << k);
                }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < S; ++j) dst[(8 * g + i) * S + j] = e[i][j];
        }
    }
}

template <int S, bool FORWARD>
cudaError_t launch_byte(const void* src, void* dst, int64_t n, int sms, cudaStream_t stream, RoundSpec rs) {
    // the 128-bit accesses need 16-byte aligned planes: n % 16 == 0 and aligned bases,
    // otherwise everything goes through the byte-wise tail loop
    const bool aligned = (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0) && (n % 16 == 0);
    const long long ngroups = aligned ? n / 16 : 0;
    long long blocks = ((aligned ? ngroups : n) + kThreads - 1) / kThreads;
    if (blocks > (long long)sms * 32) blocks = (long long)sms * 32;
    if (blocks < 1) blocks = 1;
    k_byte_shuffle<S, FORWARD><<<(unsigned)blocks, kThreads, 0, stream>>>(static_cast<const uint8_t*>(src),
                                                                          static_cast<uint8_t*>(dst), n, ngroups, rs);
    note_launch();
    return cudaGetLastError();
}

template <int S, bool FORWARD>
cudaError_t launch_bit(const void* src, void* dst, int64_t n, int sms, cudaStream_t stream, RoundSpec rs) {
    constexpr int EPT = (S == 2) ? 2048 : 1024;
    // plane rows must be word aligned for the tile kernel: n % 32 == 0 (and 16-byte aligned
    // bases; S == 2 writes 8-byte pairs: n % 64 == 0)
    const bool aligned = (reinterpret_cast<uintptr_t>(src) % 16 == 0) && (reinterpret_cast<uintptr_t>(dst) % 16 == 0) &&
                         (n % (S == 2 ? 64 : 32) == 0);
    const long long ntiles = aligned ? n / EPT : 0;
    if (ntiles > 0) {
        long long blocks = (ntiles + kBsWarps - 1) / kBsWarps;
        if (blocks > (long long)sms * 16) blocks = (long long)sms * 16;
        k_bit_shuffle<S, FORWARD><<<(unsigned)blocks, kThreads, 0, stream>>>(static_cast<const uint32_t*>(src),
                                                                             static_cast<uint32_t*>(dst), n, ntiles, rs);
        note_launch();
    }
    const long long lo = ntiles * EPT;
    if (lo < n) {
        long long blocks = ((n - lo) / 8 + kThreads - 1) / kThreads;
        if (blocks > (long long)sms * 8) blocks = (long long)sms * 8;
        if (blocks < 1) blocks = 1;
        k_bit_shuffle_tail<S, FORWARD><<<(unsigned)blocks, kThreads, 0, stream>>>(static_cast<const uint8_t*>(src),
                                                                                   static_cast<uint8_t*>(dst), n, lo, rs);
        note_launch();
    }
    return cudaGetLastError();
}

}  // namespace

cudaError_t launch_shuffle(const void* src, void* dst, int itemsize, int64_t numel, bool bits, bool forward,
                           int drop_bits, cudaStream_t stream) {
    if (numel <= 0) return cudaSuccess;
    RoundSpec rs{0u, 0ull, ~0ull};
    if (drop_bits > 0 && forward) {
        const int width = 8 * itemsize;
        const unsigned long long full = width == 64 ? ~0ull : ((1ull << width) - 1ull);
        rs.drop = (uint32_t)drop_bits;
        rs.keep = (full >> drop_bits) << drop_bits;
        rs.half_m1 = (1ull << (drop_bits - 1)) - 1ull;
    }
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
#define XBI_SHUF(S)                                                                                       \
    (bits ? (forward ? launch_bit<S, true>(src, dst, numel, sms, stream, rs) : launch_bit<S, false>(src, dst, numel, sms, stream, rs)) \
          : (forward ? launch_byte<S, true>(src, dst, numel, sms, stream, rs) : launch_byte<S, false>(src, dst, numel, sms, stream, rs)))
    switch (itemsize) {
        case 2: return XBI_SHUF(2);
        case 4: return XBI_SHUF(4);
        case 8: return XBI_SHUF(8);
        default: return cudaErrorInvalidValue;
    }
#undef XBI_SHUF
}

}  // namespace xbi
