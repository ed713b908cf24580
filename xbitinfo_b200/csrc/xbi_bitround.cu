// K3: bitround -- round to nearest, ties to even, keeping `keepbits` explicit mantissa
// bits, in place or into a second array, on the integer view of the IEEE bits.  Replaces the arithmetic behind
// _np_bitround (xbitinfo/bitround.py:7-12, numcodecs BitRound): with drop = M - keepbits
//     b += ((b >> drop) & 1) + (2^(drop-1) - 1);   b &= ~(2^drop - 1)
// (wrapping add).  Pure streaming kernel: one 128-bit load and one 128-bit store per
// 16 bytes, 2 x numel x itemsize bytes of HBM traffic.
#include <cstdlib>

#include "xbi_common.cuh"
#include "xbi_internal.h"

namespace xbi {

namespace {

__device__ __forceinline__ uint32_t round32(uint32_t b, uint32_t drop, uint32_t half_m1, uint32_t keep) {
    return (b + ((b >> drop) & 1u) + half_m1) & keep;
}
__device__ __forceinline__ uint32_t round16x2(uint32_t w, uint32_t drop, uint32_t half_m1, uint32_t keep) {
    // the two halves must not carry into each other: round them in separate registers
    const uint32_t lo = ((w & 0xFFFFu) + (((w & 0xFFFFu) >> drop) & 1u) + half_m1) & keep;
    const uint32_t hi = ((w >> 16) + (((w >> 16) >> drop) & 1u) + half_m1) & keep;
    return (lo & 0xFFFFu) | (hi << 16);
}
__device__ __forceinline__ unsigned long long round64(unsigned long long b, uint32_t drop, unsigned long long half_m1,
                                                      unsigned long long keep) {
    return (b + ((b >> drop) & 1ull) + half_m1) & keep;
}

template <int BYTES>
__device__ __forceinline__ uint4 round_vec(uint4 v, uint32_t drop, unsigned long long half_m1, unsigned long long keep) {
    if (BYTES == 4) {
        v.x = round32(v.x, drop, (uint32_t)half_m1, (uint32_t)keep);
        v.y = round32(v.y, drop, (uint32_t)half_m1, (uint32_t)keep);
        v.z = round32(v.z, drop, (uint32_t)half_m1, (uint32_t)keep);
        v.w = round32(v.w, drop, (uint32_t)half_m1, (uint32_t)keep);
    } else if (BYTES == 2) {
        v.x = round16x2(v.x, drop, (uint32_t)half_m1, (uint32_t)keep);
        v.y = round16x2(v.y, drop, (uint32_t)half_m1, (uint32_t)keep);
        v.z = round16x2(v.z, drop, (uint32_t)half_m1, (uint32_t)keep);
        v.w = round16x2(v.w, drop, (uint32_t)half_m1, (uint32_t)keep);
    } else {
        unsigned long long a = ((unsigned long long)v.y << 32) | v.x;
        unsigned long long b = ((unsigned long long)v.w << 32) | v.z;
        a = round64(a, drop, half_m1, keep);
        b = round64(b, drop, half_m1, keep);
        v.x = (uint32_t)a; v.y = (uint32_t)(a >> 32);
        v.z = (uint32_t)b; v.w = (uint32_t)(b >> 32);
    }
    return v;
}

template <int BYTES>
__device__ __forceinline__ void round_scalar(const void* src, void* dst, long long i, uint32_t drop,
                                             unsigned long long half_m1, unsigned long long keep) {
    if (BYTES == 2) {
        const uint32_t b = static_cast<const uint16_t*>(src)[i];
        static_cast<uint16_t*>(dst)[i] = (uint16_t)((b + ((b >> drop) & 1u) + (uint32_t)half_m1) & (uint32_t)keep);
    } else if (BYTES == 4) {
        static_cast<uint32_t*>(dst)[i] = round32(static_cast<const uint32_t*>(src)[i], drop, (uint32_t)half_m1, (uint32_t)keep);
    } else {
        static_cast<unsigned long long*>(dst)[i] = round64(static_cast<const unsigned long long*>(src)[i], drop, half_m1, keep);
    }
}

constexpr int kThreads = 256;

__device__ __forceinline__ uint4 ld_stream(const uint4* p) {
    uint4 v;
    asm volatile("ld.global.L1::no_allocate.v4.u32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(uint4* p, const uint4& v) {
    asm volatile("st.global.cs.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

// head: elements before the first 16-byte boundary; nvec: number of whole 16-byte vectors.
// VPT independent 128-bit loads are in flight per thread before the first store.
// src == dst: in place; otherwise the two have the same alignment within 16 bytes and do not overlap.
template <int BYTES, int VPT, bool HINT>
__global__ void __launch_bounds__(kThreads)
k_bitround(const void* src, void* x, long long numel, long long head, long long nvec, uint32_t drop,
           unsigned long long half_m1, unsigned long long keep) {
    constexpr int EPV = 16 / BYTES;
    uint4* vbase = reinterpret_cast<uint4*>(static_cast<char*>(x) + head * BYTES);
    const uint4* sbase = reinterpret_cast<const uint4*>(static_cast<const char*>(src) + head * BYTES);
    const long long stride = (long long)gridDim.x * kThreads * VPT;
    for (long long base = (long long)blockIdx.x * kThreads * VPT; base < nvec; base += stride) {
        uint4 v[VPT];
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const long long i = base + (long long)u * kThreads + threadIdx.x;
            if (i < nvec) v[u] = HINT ? ld_stream(sbase + i) : sbase[i];
        }
#pragma unroll
        for (int u = 0; u < VPT; ++u) {
            const long long i = base + (long long)u * kThreads + threadIdx.x;
            if (i < nvec) {
                const uint4 r = round_vec<BYTES>(v[u], drop, half_m1, keep);
                if (HINT) st_stream(vbase + i, r); else vbase[i] = r;
            }
        }
    }
    // ragged ends (fewer than 2 vectors' worth of elements in total)
    if (blockIdx.x == 0) {
        const long long tail0 = head + nvec * EPV;
        for (long long i = threadIdx.x; i < head; i += kThreads) round_scalar<BYTES>(src, x, i, drop, half_m1, keep);
        for (long long i = tail0 + threadIdx.x; i < numel; i += kThreads) round_scalar<BYTES>(src, x, i, drop, half_m1, keep);
    }
}

template <int BYTES, int VPT, bool HINT>
void launch_variant(const void* src, void* x, long long numel, long long head, long long nvec, uint32_t drop, unsigned long long half_m1,
                    unsigned long long keep, int sms, int waves, cudaStream_t stream) {
    long long blocks = (nvec + (long long)kThreads * VPT - 1) / ((long long)kThreads * VPT);
    const long long cap = (long long)sms * waves;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    k_bitround<BYTES, VPT, HINT><<<(unsigned)blocks, kThreads, 0, stream>>>(src, x, numel, head, nvec, drop, half_m1, keep);
}

template <int BYTES>
void launch_bytes(const void* src, void* x, long long numel, long long head, long long nvec, uint32_t drop, unsigned long long half_m1,
                  unsigned long long keep, int sms, cudaStream_t stream) {
    // development switches (defaults are the measured best)
    static const int vpt = std::getenv("XBI_BR_VPT") ? std::atoi(std::getenv("XBI_BR_VPT")) : 8;
    static const int hint = std::getenv("XBI_BR_HINT") ? std::atoi(std::getenv("XBI_BR_HINT")) : 0;
    static const int waves = std::getenv("XBI_BR_WAVES") ? std::atoi(std::getenv("XBI_BR_WAVES")) : 64;
    if (vpt == 8) {
        if (hint) launch_variant<BYTES, 8, true>(src, x, numel, head, nvec, drop, half_m1, keep, sms, waves, stream);
        else launch_variant<BYTES, 8, false>(src, x, numel, head, nvec, drop, half_m1, keep, sms, waves, stream);
    } else if (vpt == 2) {
        if (hint) launch_variant<BYTES, 2, true>(src, x, numel, head, nvec, drop, half_m1, keep, sms, waves, stream);
        else launch_variant<BYTES, 2, false>(src, x, numel, head, nvec, drop, half_m1, keep, sms, waves, stream);
    } else {
        if (hint) launch_variant<BYTES, 4, true>(src, x, numel, head, nvec, drop, half_m1, keep, sms, waves, stream);
        else launch_variant<BYTES, 4, false>(src, x, numel, head, nvec, drop, half_m1, keep, sms, waves, stream);
    }
}

}  // namespace

cudaError_t launch_bitround(void* x, int dtype, int64_t numel, int keepbits, cudaStream_t stream) {
    return launch_bitround_to(x, x, dtype, numel, keepbits, stream);
}

cudaError_t launch_bitround_to(const void* src, void* x, int dtype, int64_t numel, int keepbits, cudaStream_t stream) {
    if (numel <= 0) return cudaSuccess;
    const int bytes = dtype_bytes(dtype);
    const int mant = dtype == XBI_F16 ? 10 : dtype == XBI_F32 ? 23 : 52;
    const uint32_t drop = (uint32_t)(mant - keepbits);
    if (src != x && (drop == 0 || (reinterpret_cast<uintptr_t>(src) & 15) != (reinterpret_cast<uintptr_t>(x) & 15))) {
        // identity, or source and destination sit differently inside their 16-byte vectors (never the case for
        // whole allocations): plain copy, then round in place
        const cudaError_t e = cudaMemcpyAsync(x, src, (size_t)numel * bytes, cudaMemcpyDeviceToDevice, stream);
        if (e != cudaSuccess) return e;
        src = x;
    }
    if (drop == 0) return cudaSuccess;  // identity
    const int width = bytes * 8;
    const unsigned long long full = width == 64 ? ~0ull : ((1ull << width) - 1ull);
    const unsigned long long keep = (full >> drop) << drop;
    const unsigned long long half_m1 = (1ull << (drop - 1)) - 1ull;

    const uintptr_t addr = reinterpret_cast<uintptr_t>(x);
    long long head = (long long)(((16 - (addr & 15)) & 15) / bytes);
    if (head > numel) head = numel;
    const long long nvec = (numel - head) / (16 / bytes);

    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    switch (bytes) {
        case 2: launch_bytes<2>(src, x, numel, head, nvec, drop, half_m1, keep, sms, stream); break;
        case 4: launch_bytes<4>(src, x, numel, head, nvec, drop, half_m1, keep, sms, stream); break;
        case 8: launch_bytes<8>(src, x, numel, head, nvec, drop, half_m1, keep, sms, stream); break;
        default: return cudaErrorInvalidValue;
    }
    note_launch();
    return cudaGetLastError();
}

}  // namespace xbi
