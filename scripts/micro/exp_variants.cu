// Micro-benchmark (design input for the attention softmax): elements per clock per SM of the exponential variants
// available on sm_100a: MUFU ex2 f32 / f16x2 / bf16x2, a Cody-Waite + degree-3 polynomial on packed f32x2 FMA ops, and
// mixes of the two.  One CTA of W warps on one SM; clock64() around an unrolled loop of independent chains.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_fp16.h>
#include <cuda_bf16.h>

__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2h2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2b2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint64_t pk(float a, float b) { uint64_t r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ void upk(uint64_t v, float& a, float& b) { asm("mov.b64 {%0, %1}, %2;" : "=f"(a), "=f"(b) : "l"(v)); }
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) { uint64_t d; asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ uint64_t add2(uint64_t a, uint64_t b) { uint64_t d; asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// 2^x for x <= 0 (x >= -126), two elements: t = x + 1.5*2^23 (round to nearest integer in the low mantissa bits),
// f = x - (t - magic) in [-0.5, 0.5], p(f) degree 3, exponent inserted with an integer shift-add.
__device__ __forceinline__ uint64_t poly_exp2_x2(uint64_t x) {
    const uint64_t MAGIC = pk(12582912.0f, 12582912.0f), NMAGIC = pk(-12582912.0f, -12582912.0f);
    const uint64_t C3 = pk(0.05550410866f, 0.05550410866f), C2 = pk(0.2402265070f, 0.2402265070f), C1 = pk(0.6931471806f, 0.6931471806f),
                   C0 = pk(1.0f, 1.0f);
    uint64_t t = add2(x, MAGIC);
    uint64_t n = add2(t, NMAGIC);
    uint64_t nn; asm("{\n\t.reg .f32 a, b;\n\tmov.b64 {a, b}, %1;\n\tneg.f32 a, a;\n\tneg.f32 b, b;\n\tmov.b64 %0, {a, b};\n\t}" : "=l"(nn) : "l"(n));
    uint64_t f = add2(x, nn);
    uint64_t p = fma2(C3, f, C2);
    p = fma2(p, f, C1);
    p = fma2(p, f, C0);
    float p0, p1, t0, t1;
    upk(p, p0, p1);
    upk(t, t0, t1);
    p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
    return pk(p0, p1);
}

template <int MODE, int ILP>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = -0.001f * (threadIdx.x + i) - 0.01f;
    uint32_t h[ILP];
    for (int i = 0; i < ILP; ++i) { __half2 v = __floats2half2_rn(a[i], a[i] * 0.5f); h[i] = *reinterpret_cast<uint32_t*>(&v); }
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) a[i] = ex2f(a[i]) - 1.5f;
        } else if constexpr (MODE == 1) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) h[i] = ex2h2(h[i]) ^ 0x80008000u;
        } else if constexpr (MODE == 2) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) h[i] = ex2b2(h[i]) ^ 0x80008000u;
        } else if constexpr (MODE == 3) {          // polynomial only
#pragma unroll
            for (int i = 0; i < ILP; i += 2) { float x0, x1; upk(poly_exp2_x2(pk(a[i], a[i + 1])), x0, x1); a[i] = x0 - 1.5f; a[i + 1] = x1 - 1.5f; }
        } else if constexpr (MODE == 4) {          // half MUFU, half polynomial
#pragma unroll
            for (int i = 0; i < ILP; i += 4) {
                float x0, x1; upk(poly_exp2_x2(pk(a[i], a[i + 1])), x0, x1); a[i] = x0 - 1.5f; a[i + 1] = x1 - 1.5f;
                a[i + 2] = ex2f(a[i + 2]) - 1.5f; a[i + 3] = ex2f(a[i + 3]) - 1.5f;
            }
        } else if constexpr (MODE == 5) {          // 3/4 MUFU, 1/4 polynomial
#pragma unroll
            for (int i = 0; i < ILP; i += 8) {
                float x0, x1; upk(poly_exp2_x2(pk(a[i], a[i + 1])), x0, x1); a[i] = x0 - 1.5f; a[i + 1] = x1 - 1.5f;
#pragma unroll
                for (int j = 2; j < 8; ++j) a[i + j] = ex2f(a[i + j]) - 1.5f;
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < ILP; ++i) s += a[i] + __uint_as_float(h[i]);
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

__global__ void accuracy(float* maxrel) {
    float worst = 0.0f;
    for (int i = threadIdx.x; i < 1 << 20; i += blockDim.x) {
        const float x = -40.0f * (float)i / (float)(1 << 20);
        float p0, p1; upk(poly_exp2_x2(pk(x, x - 0.37f)), p0, p1);
        const float r0 = exp2f(x), r1 = exp2f(x - 0.37f);
        worst = fmaxf(worst, fmaxf(fabsf(p0 - r0) / r0, fabsf(p1 - r1) / r1));
    }
    atomicMax(reinterpret_cast<int*>(maxrel), __float_as_int(worst));
}

int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    const int iters = 256;
    const char* names[] = {"mufu ex2 f32", "mufu ex2 f16x2", "mufu ex2 bf16x2", "poly f32x2", "1/2 mufu + 1/2 poly", "3/4 mufu + 1/4 poly"};
    for (int warps : {4, 8, 16}) {
        long long h;
#define RUN(MODE, ILP, EPI) k<MODE, ILP><<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("warps/SM=%2d %-22s ILP=%2d: %6.2f elements/clk/SM\n", warps, names[MODE], ILP, (double)iters * ILP * EPI * warps * 32 / (double)h);
        RUN(0, 16, 1) RUN(1, 16, 2) RUN(2, 16, 2) RUN(3, 16, 1) RUN(4, 16, 1) RUN(5, 16, 1)
    }
    float* mr; cudaMalloc(&mr, 4); cudaMemset(mr, 0, 4);
    accuracy<<<1, 1024>>>(mr);
    float hm; cudaMemcpy(&hm, mr, 4, cudaMemcpyDeviceToHost);
    printf("poly exp2 max relative error on [-40, 0]: %.3e (%s)\n", hm, cudaGetErrorString(cudaGetLastError()));
    return 0;
}
