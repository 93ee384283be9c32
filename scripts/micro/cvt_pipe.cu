// Micro-benchmark: does the fp32 -> bf16x2 pack (F2FP.BF16.F32.PACK_AB) share the XU pipe with MUFU.EX2 on sm_100a?
// One CTA of 16 warps on one SM; elements per clock per SM of: ex2 alone, pack alone, ex2 + pack in the attention kernel's 2:1
// ratio, ex2 + an integer round-and-permute pack (IADD + IADD + PRMT on the ALU pipe).
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
__device__ __forceinline__ float ex2f(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t pack_cvt(float lo, float hi) { uint32_t r; asm volatile("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo)); return r; }
__device__ __forceinline__ uint32_t pack_int(float lo, float hi) {       // round-half-up on the bits, then take the upper halves
    uint32_t a = __float_as_uint(lo) + 0x8000u, b = __float_as_uint(hi) + 0x8000u, r;
    asm volatile("prmt.b32 %0, %1, %2, 0x7632;" : "=r"(r) : "r"(a), "r"(b));
    return r;
}
template <int MODE>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[16];
    for (int i = 0; i < 16; ++i) a[i] = -0.001f * (threadIdx.x + i) - 0.01f;
    uint32_t acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if constexpr (MODE == 0) {
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = ex2f(a[i]) - 1.5f;
        } else if constexpr (MODE == 1) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) { uint32_t p = pack_cvt(a[i], a[i + 1]); acc ^= p; a[i] = __uint_as_float(p) * 0.5f; }
        } else if constexpr (MODE == 2) {
#pragma unroll
            for (int i = 0; i < 16; i += 2) { float x = ex2f(a[i]), y = ex2f(a[i + 1]); acc ^= pack_cvt(x, y); a[i] = x - 1.5f; a[i + 1] = y - 1.5f; }
        } else {
#pragma unroll
            for (int i = 0; i < 16; i += 2) { float x = ex2f(a[i]), y = ex2f(a[i + 1]); acc ^= pack_int(x, y); a[i] = x - 1.5f; a[i + 1] = y - 1.5f; }
        }
    }
    long long t1 = clock64();
    float s = __uint_as_float(acc);
    for (int i = 0; i < 16; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    const int iters = 256, warps = 16;
    long long h;
    const char* names[] = {"ex2 only (16 per iter)", "cvt.rn.bf16x2 only (8 packs per iter)", "16 ex2 + 8 cvt packs", "16 ex2 + 8 integer packs"};
#define RUN(MODE) k<MODE><<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("%-40s: %8.1f clk per iteration per warp-scheduler slot, %6.2f elements/clk/SM\n", names[MODE], (double)h / iters / (warps / 4), (double)iters * 16 * warps * 32 / (double)h);
    RUN(0) RUN(1) RUN(2) RUN(3)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
