// Shared helpers for the sm_100a kernels of the DADD hot path.
#pragma once

#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdio>
#include <cstring>
#include <type_traits>

#include "../../include/dadd_b200.h"

namespace daddk {

extern thread_local char g_last_error[512];
extern std::atomic<int64_t> g_launches;

inline int fail(const char* fmt, const char* a = "", long long x = 0, long long y = 0) {
    snprintf(g_last_error, sizeof(g_last_error), fmt, a, x, y);
    return 1;
}

#define DADD_REQUIRE(cond, what)                                              \
    do {                                                                      \
        if (!(cond)) return ::daddk::fail("%s: requirement failed: " #cond, what); \
    } while (0)

// Run `...` with `T` bound to the element type of a 16-bit dtype code / of any dtype code.
#define DADD_DISPATCH_16(dtype, T, ...)                                          \
    do {                                                                         \
        if ((dtype) == DADD_BF16) { using T = __nv_bfloat16; __VA_ARGS__; }      \
        else { using T = __half; __VA_ARGS__; }                                  \
    } while (0)
#define DADD_DISPATCH_ANY(dtype, T, ...)                                         \
    do {                                                                         \
        if ((dtype) == DADD_F32) { using T = float; __VA_ARGS__; }               \
        else if ((dtype) == DADD_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
        else { using T = __half; __VA_ARGS__; }                                  \
    } while (0)

inline bool dtype_ok(int dtype) { return dtype == DADD_F32 || dtype == DADD_BF16 || dtype == DADD_F16; }
inline bool dtype16_ok(int dtype) { return dtype == DADD_BF16 || dtype == DADD_F16; }

// Call after every launch: counts it and turns a launch error into a return code.
inline int launched(const char* what) {
    g_launches.fetch_add(1, std::memory_order_relaxed);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
        return 2;
    }
    return 0;
}

inline int cuda_ok(cudaError_t e, const char* what) {
    if (e == cudaSuccess) return 0;
    snprintf(g_last_error, sizeof(g_last_error), "%s: %s", what, cudaGetErrorString(e));
    return 2;
}

__device__ __forceinline__ float to_f(float v) { return v; }
__device__ __forceinline__ float to_f(__nv_bfloat16 v) { return __bfloat162float(v); }
__device__ __forceinline__ float to_f(__half v) { return __half2float(v); }

// two floats -> one 32-bit register of two 16-bit elements (lo in the low half)
template <typename T>
__device__ __forceinline__ uint32_t pack2(float lo, float hi);
template <>
__device__ __forceinline__ uint32_t pack2<__nv_bfloat16>(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
template <>
__device__ __forceinline__ uint32_t pack2<__half>(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---- 8-element vectors (16 B of bf16/fp16, 32 B of fp32) ---------------------------------------------
template <typename T>
struct Vec8 {
    uint4 raw;
    __device__ __forceinline__ void load(const T* p) { raw = *reinterpret_cast<const uint4*>(p); }
    __device__ __forceinline__ void store(T* p) const { *reinterpret_cast<uint4*>(p) = raw; }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        const uint32_t w[4] = {raw.x, raw.y, raw.z, raw.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            if constexpr (std::is_same_v<T, __nv_bfloat16>) {
                f[2 * i] = __uint_as_float(w[i] << 16);
                f[2 * i + 1] = __uint_as_float(w[i] & 0xffff0000u);
            } else {
                const float2 v = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                f[2 * i] = v.x;
                f[2 * i + 1] = v.y;
            }
        }
    }
    __device__ __forceinline__ void pack(const float (&f)[8]) {
        raw = make_uint4(pack2<T>(f[0], f[1]), pack2<T>(f[2], f[3]), pack2<T>(f[4], f[5]), pack2<T>(f[6], f[7]));
    }
};

template <>
struct Vec8<float> {
    float4 a, b;
    __device__ __forceinline__ void load(const float* p) {
        a = *reinterpret_cast<const float4*>(p);
        b = *reinterpret_cast<const float4*>(p + 4);
    }
    __device__ __forceinline__ void store(float* p) const {
        *reinterpret_cast<float4*>(p) = a;
        *reinterpret_cast<float4*>(p + 4) = b;
    }
    __device__ __forceinline__ void unpack(float (&f)[8]) const {
        f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
        f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
    }
    __device__ __forceinline__ void pack(const float (&f)[8]) {
        a = make_float4(f[0], f[1], f[2], f[3]);
        b = make_float4(f[4], f[5], f[6], f[7]);
    }
};

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ float silu(float y) { return __fdividef(y, 1.0f + __expf(-y)); }

inline int num_sms() {
    static int n = 0;
    if (n == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

}  // namespace daddk
