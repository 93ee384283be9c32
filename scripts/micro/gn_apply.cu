// Micro-benchmark: what bounds the normalise pass of the streaming GroupNorm kernel?  One CTA per SM, 480 (or 960) threads, a 40 KB
// fp16 slab in shared memory, thread = fixed 8-channel column; per 8-element vector: LDS.128 -> 8 x (cvt, FFMA, [tanh], FFMA) -> pack
// -> [STG.128].  Prints clk per slab for: tanh + store, tanh only, store only (identity), neither.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o scripts/micro/gn_apply scripts/micro/gn_apply.cu
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
constexpr int NPIX = 64, C = 320, V = C / 8;
__device__ __forceinline__ float tanh_fast(float x) { float y; asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
template <bool TANH, int STORE, int UNROLL>      // STORE: 0 none (results xor-ed into a register), 1 STG.128, 2 STS.128 in place
__global__ void __launch_bounds__(1024, 1) k(__half* y, long long* cyc, int iters, int PH) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int tid = threadIdx.x, v = tid % V, ph = tid / V;
    for (int i = tid; i < NPIX * C / 2; i += blockDim.x) reinterpret_cast<__half2*>(smem)[i] = __floats2half2_rn(0.01f * (i & 63), -0.02f * (i & 31));
    float sa[8], sb[8];
    uint32_t acc = 0;
    for (int i = 0; i < 8; ++i) { sa[i] = 0.5f + 0.01f * i; sb[i] = 0.1f * i - 0.3f; }
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        __half* dst = y + ((size_t)blockIdx.x * 16 + (it & 15)) * NPIX * C + v * 8;
        for (int px = ph; px < NPIX; px += UNROLL * PH) {
            uint4 q[UNROLL];
#pragma unroll
            for (int u = 0; u < UNROLL; ++u)
                if (px + u * PH < NPIX) q[u] = *reinterpret_cast<const uint4*>(smem + ((size_t)(px + u * PH) * C + v * 8) * 2);
#pragma unroll
            for (int u = 0; u < UNROLL; ++u) {
                if (px + u * PH >= NPIX) continue;
                const uint32_t w[4] = {q[u].x, q[u].y, q[u].z, q[u].w};
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&w[i]));
                    float h0 = fmaf(f.x, sa[2 * i], sb[2 * i]), h1 = fmaf(f.y, sa[2 * i + 1], sb[2 * i + 1]);
                    if (TANH) { h0 = fmaf(h0, tanh_fast(h0), h0); h1 = fmaf(h1, tanh_fast(h1), h1); }
                    __half2 r = __floats2half2_rn(h0, h1);
                    o[i] = *reinterpret_cast<uint32_t*>(&r);
                }
                if (STORE == 1) *reinterpret_cast<uint4*>(dst + (size_t)(px + u * PH) * C) = make_uint4(o[0], o[1], o[2], o[3]);
                else if (STORE == 2) *reinterpret_cast<uint4*>(smem + ((size_t)(px + u * PH) * C + v * 8) * 2) = make_uint4(o[0], o[1], o[2], o[3]);
                else acc ^= o[0] ^ o[1] ^ o[2] ^ o[3];
            }
        }
        __syncthreads();
    }
    const long long t1 = clock64();
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
    if (acc == 0x12345678u) y[0] = __float2half(sa[0]);
}
int main() {
    __half* y; long long* cyc; static long long h[148];
    cudaMalloc(&y, (size_t)148 * 16 * NPIX * C * 2); cudaMalloc(&cyc, sizeof(h));
    const int iters = 64;
#define RUN(T, S, U, THREADS, name)                                                                                   \
    {                                                                                                                  \
        cudaFuncSetAttribute(k<T, S, U>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);                      \
        k<T, S, U><<<148, THREADS, 48 * 1024>>>(y, cyc, iters, THREADS / V);                                           \
        cudaError_t e = cudaDeviceSynchronize();                                                                       \
        cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);                                                         \
        printf("%-34s threads=%4d unroll=%d: %7.0f clk per 40 KB slab (%s)\n", name, THREADS, U, (double)h[0] / iters, cudaGetErrorString(e)); \
    }
    for (int rep = 0; rep < 2; ++rep) {
        RUN(true, 1, 2, 480, "tanh + STG") RUN(true, 0, 2, 480, "tanh, no store") RUN(true, 2, 2, 480, "tanh + STS in place") RUN(true, 2, 4, 480, "tanh + STS in place")
        RUN(false, 0, 2, 480, "identity, no store") RUN(false, 2, 2, 480, "identity + STS in place") RUN(false, 1, 2, 480, "identity + STG")
        RUN(true, 0, 2, 960, "tanh, no store") RUN(true, 2, 2, 960, "tanh + STS in place") RUN(true, 0, 2, 320, "tanh, no store")
    }
    return 0;
}
