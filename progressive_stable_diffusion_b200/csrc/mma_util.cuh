// Warp-level 16-bit MMA fragments (mma.sync.m16n8k16, bf16 or fp16 operands, fp32 accumulate) read straight from
// padded shared-memory tiles.
#pragma once

#include "common.cuh"

namespace daddk {

template <typename T>
__device__ __forceinline__ void mma_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
    if constexpr (std::is_same_v<T, __nv_bfloat16>) {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    } else {
        asm volatile(
            "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
    }
}

// A (16 x 16, row-major [row][k], `stride` elements per row): lane (g = lane/4, t = lane%4) holds
// a0 = (g, k0+2t..), a1 = (g+8, k0+2t..), a2 = (g, k0+8+2t..), a3 = (g+8, k0+8+2t..).
template <typename T>
__device__ __forceinline__ void load_a_frag(uint32_t (&a)[4], const T* base, int stride, int k0, int g, int t) {
    const T* p = base + g * stride + k0 + 2 * t;
    a[0] = *reinterpret_cast<const uint32_t*>(p);
    a[1] = *reinterpret_cast<const uint32_t*>(p + 8 * stride);
    a[2] = *reinterpret_cast<const uint32_t*>(p + 8);
    a[3] = *reinterpret_cast<const uint32_t*>(p + 8 * stride + 8);
}

// B (16 x 8) given as [n][k] rows (k contiguous): b0 = (k0+2t.., n = g), b1 = (k0+8+2t.., n = g).
template <typename T>
__device__ __forceinline__ void load_b_frag(uint32_t& b0, uint32_t& b1, const T* base, int stride, int k0, int g, int t) {
    const T* p = base + g * stride + k0 + 2 * t;
    b0 = *reinterpret_cast<const uint32_t*>(p);
    b1 = *reinterpret_cast<const uint32_t*>(p + 8);
}

// 2^x on the MUFU pipe, flush-to-zero (inputs are <= 0 in every softmax here).
__device__ __forceinline__ float fast_exp2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

__device__ __forceinline__ float quad_max(float v) {
    v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 1));
    return fmaxf(v, __shfl_xor_sync(0xffffffffu, v, 2));
}
__device__ __forceinline__ float quad_sum(float v) {
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v + __shfl_xor_sync(0xffffffffu, v, 2);
}

}  // namespace daddk
