// Micro-benchmark: MUFU.EX2 issue rate of one vs two warps per SM sub-partition (design input for the attention softmax).
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ float ex2(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <int ILP>
__global__ void k(float* out, long long* cyc, int iters) {
    float a[ILP];
    for (int i = 0; i < ILP; ++i) a[i] = -0.001f * (threadIdx.x + i);
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < ILP; ++i) a[i] = ex2(a[i]);
    }
    long long t1 = clock64();
    float s = 0; for (int i = 0; i < ILP; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    float* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    const int iters = 256;
    for (int warps : {4, 8, 12, 16}) {
        long long h;
#define RUN(ILP) k<ILP><<<1, warps * 32>>>(out, cyc, iters); cudaDeviceSynchronize(); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        printf("warps/SM=%2d ILP=%2d: %.2f cycles per MUFU warp-instr per SMSP-warp (%.2f cycles per instr per SMSP)\n", warps, ILP, (double)h / (iters * ILP), (double)h / (iters * ILP) / (warps / 4.0));
        RUN(1) RUN(4) RUN(8) RUN(16)
    }
    return 0;
}
