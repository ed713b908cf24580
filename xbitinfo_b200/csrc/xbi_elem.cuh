// Element traits: how one array element of each supported type maps onto 32-bit words.
// Narrow types are zero-extended into one word here (the generic kernel processes one
// element per thread and word); 64-bit types use a (lo, hi) word pair.
#pragma once
#include "xbi_common.cuh"
#include "xbi_internal.h"

namespace xbi {

template <int KIND>
struct Elem;

template <>
struct Elem<kU8> {
    static constexpr int kWords = 1, kBits = 8, kBytes = 1;
    static __device__ __forceinline__ void load(const void* x, long long f, uint32_t (&w)[1]) {
        w[0] = __ldg(static_cast<const uint8_t*>(x) + f);
    }
    static __device__ __forceinline__ void transform(uint32_t (&)[1]) {}
    static __device__ __forceinline__ bool is_nan(const uint32_t (&)[1]) { return false; }
    static __device__ __forceinline__ bool equals(const uint32_t (&w)[1], uint32_t lo, uint32_t) {
        return w[0] == (lo & 0xFFu);
    }
};

template <>
struct Elem<kU16> {
    static constexpr int kWords = 1, kBits = 16, kBytes = 2;
    static __device__ __forceinline__ void load(const void* x, long long f, uint32_t (&w)[1]) {
        w[0] = __ldg(static_cast<const uint16_t*>(x) + f);
    }
    static __device__ __forceinline__ void transform(uint32_t (&)[1]) {}
    static __device__ __forceinline__ bool is_nan(const uint32_t (&)[1]) { return false; }
    static __device__ __forceinline__ bool equals(const uint32_t (&w)[1], uint32_t lo, uint32_t) {
        return w[0] == (lo & 0xFFFFu);
    }
};

template <>
struct Elem<kU32> {
    static constexpr int kWords = 1, kBits = 32, kBytes = 4;
    static __device__ __forceinline__ void load(const void* x, long long f, uint32_t (&w)[1]) {
        w[0] = __ldg(static_cast<const uint32_t*>(x) + f);
    }
    static __device__ __forceinline__ void transform(uint32_t (&)[1]) {}
    static __device__ __forceinline__ bool is_nan(const uint32_t (&)[1]) { return false; }
    static __device__ __forceinline__ bool equals(const uint32_t (&w)[1], uint32_t lo, uint32_t) {
        return w[0] == lo;
    }
};

template <>
struct Elem<kU64> {
    static constexpr int kWords = 2, kBits = 64, kBytes = 8;
    static __device__ __forceinline__ void load(const void* x, long long f, uint32_t (&w)[2]) {
        const uint2 v = __ldg(static_cast<const uint2*>(x) + f);
        w[0] = v.x;
        w[1] = v.y;
    }
    static __device__ __forceinline__ void transform(uint32_t (&)[2]) {}
    static __device__ __forceinline__ bool is_nan(const uint32_t (&)[2]) { return false; }
    static __device__ __forceinline__ bool equals(const uint32_t (&w)[2], uint32_t lo, uint32_t hi) {
        return w[0] == lo && w[1] == hi;
    }
};

template <>
struct Elem<kF16> : Elem<kU16> {
    static __device__ __forceinline__ void transform(uint32_t (&w)[1]) { w[0] = sexp_f16_single(w[0]); }
    static __device__ __forceinline__ bool is_nan(const uint32_t (&w)[1]) { return (w[0] & 0x7FFFu) > 0x7C00u; }
};

template <>
struct Elem<kF32> : Elem<kU32> {
    static __device__ __forceinline__ void transform(uint32_t (&w)[1]) { w[0] = sexp_f32(w[0]); }
    static __device__ __forceinline__ bool is_nan(const uint32_t (&w)[1]) { return (w[0] & 0x7FFFFFFFu) > 0x7F800000u; }
};

template <>
struct Elem<kF64> : Elem<kU64> {
    static __device__ __forceinline__ void transform(uint32_t (&w)[2]) { w[1] = sexp_f64_hi(w[1]); }
    static __device__ __forceinline__ bool is_nan(const uint32_t (&w)[2]) {
        const uint32_t m = w[1] & 0x7FFFFFFFu;
        return m > 0x7FF00000u || (m == 0x7FF00000u && w[0] != 0u);
    }
};

}  // namespace xbi
