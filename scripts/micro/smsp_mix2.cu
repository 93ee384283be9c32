// Micro-benchmark: which of the softmax loop's instruction classes contend inside one SM sub-partition?
// 16 warps (4 per sub-partition; warp & 3 = sub-partition = TMEM lane quarter).  Each of the 4 warps of a sub-partition
// gets a role from a 4-letter pattern and loops for a fixed time; the table prints clk per iteration for every role.
//   M  64 x ex2.approx (MUFU)          L  tcgen05.ld 32x32b.x64 + wait (8 KB per warp)      l  2 x tcgen05.ld .x32 + wait
//   S  tcgen05.st 32x32b.x32 + wait    F  512 FFMA (8 per element)                           P  64 x (FMUL + ex2 + FADD) + 32 F2FP
//   X  ld.x64 -> 64 x (FFMA, ex2, FADD), 32 F2FP -> st.x32 (the softmax inner loop)          -  idle
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I progressive_stable_diffusion_b200/csrc -o scripts/micro/smsp_mix scripts/micro/smsp_mix.cu -lcuda
#include <cstdio>
#include <cstring>
#include <cuda_fp16.h>
#include "tc_util.cuh"

namespace daddk { thread_local char g_last_error[512] = ""; std::atomic<int64_t> g_launches{0}; }
using namespace daddk::tc;


// half2 polynomial exp2 (FMA pipe): 2^x for a pair of scores, x <= 0 relative to the tile's row maximum.  hc = {c0,c1,c2,c3} (already
// scaled by 2^frac of the tile offset), magic = 1551 + integer tile offset, clampv = -(15 + integer tile offset).
struct PolyC { uint32_t c0, c1, c2, c3, magic, clampv; };
__device__ __forceinline__ uint32_t h2u(__half2 h) { return *reinterpret_cast<uint32_t*>(&h); }
__device__ __forceinline__ PolyC make_polyc() {
    PolyC k;
    k.c0 = h2u(__floats2half2_rn(1.0f, 1.0f));
    k.c1 = h2u(__floats2half2_rn(0.6951786f, 0.6951786f));
    k.c2 = h2u(__floats2half2_rn(0.2402265f, 0.2402265f));
    k.c3 = h2u(__floats2half2_rn(0.0555041f, 0.0555041f));
    k.magic = h2u(__floats2half2_rn(1551.0f, 1551.0f));
    k.clampv = h2u(__floats2half2_rn(-15.0f, -15.0f));
    return k;
}
__device__ __forceinline__ uint32_t poly_ex2_h2(float x0, float x1, const PolyC& k) {
    uint32_t xh, t, nf, f, p, e, r;
    asm("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(xh) : "f"(x1), "f"(x0));
    asm("max.f16x2 %0, %1, %2;" : "=r"(xh) : "r"(xh), "r"(k.clampv));
    asm("add.rn.f16x2 %0, %1, %2;" : "=r"(t) : "r"(xh), "r"(k.magic));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(nf) : "r"(t), "r"(k.magic));
    asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(f) : "r"(xh), "r"(nf));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(k.c3), "r"(f), "r"(k.c2));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(k.c1));
    asm("fma.rn.f16x2 %0, %1, %2, %3;" : "=r"(p) : "r"(p), "r"(f), "r"(k.c0));
    e = (t << 10) & 0x7C007C00u;
    asm("mul.rn.f16x2 %0, %1, %2;" : "=r"(r) : "r"(p), "r"(e));
    return r;
}
template <int NPOLY>      // NPOLY of every 8 pairs go through the polynomial
__device__ __forceinline__ void mix_loop(uint32_t (&sr)[64], const PolyC& k) {
    uint32_t pk[32];
#pragma unroll
    for (int e = 0; e < 64; e += 2) {
        const float a0 = fmaf(__uint_as_float(sr[e]), 0.125f, -1.0f), a1 = fmaf(__uint_as_float(sr[e + 1]), 0.125f, -1.0f);
        if (((e >> 1) & 7) < NPOLY) {
            pk[e >> 1] = poly_ex2_h2(a0, a1, k);
        } else {
            float p0, p1;
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p0) : "f"(a0));
            asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(p1) : "f"(a1));
            __half2 h = __floats2half2_rn(p0, p1);
            pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
        }
    }
#pragma unroll
    for (int e = 0; e < 32; ++e) sr[e] ^= pk[e] & 1u;
}

struct Res { long long iters[16]; long long clk[16]; };
struct Pattern { char role[4]; };

__global__ void __launch_bounds__(512, 1) k(Res* out, Pattern pat, long long budget) {
    __shared__ uint32_t tmem_base;
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int quarter = warp & 3, slot = warp >> 2;
    const char role = pat.role[slot];
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t tS = tmem + slot * 64 + lane_base, tP = tmem + 256 + slot * 32 + lane_base;
    uint32_t sr[64];
#pragma unroll
    for (int e = 0; e < 64; ++e) sr[e] = __float_as_uint(-(float)((threadIdx.x + e) & 15));
    long long iters = 0;
    float sink = 0.f;
    const long long t0 = clock64();
    long long t1 = t0;
    if (role != '-') {
        while (true) {
            t1 = clock64();
            if (__shfl_sync(0xffffffffu, (int)(t1 - t0 >= budget), 0)) break;
            if (role == 'M') {
#pragma unroll
                for (int e = 0; e < 64; ++e) sr[e] = __float_as_uint(ex2(__uint_as_float(sr[e])));
            } else if (role == 'L') {
                tmem_ld64(tS, sr);
                tmem_wait_ld();
            } else if (role == 'l') {
                tmem_ld32(tS, *reinterpret_cast<uint32_t(*)[32]>(sr));
                tmem_ld32(tS + 32, *reinterpret_cast<uint32_t(*)[32]>(sr + 32));
                tmem_wait_ld();
            } else if (role == 'S') {
                tmem_st32(tP, *reinterpret_cast<uint32_t(*)[32]>(sr));
                tmem_wait_st();
            } else if (role == 'F') {
#pragma unroll
                for (int r = 0; r < 8; ++r)
#pragma unroll
                    for (int e = 0; e < 64; ++e) sr[e] = __float_as_uint(fmaf(__uint_as_float(sr[e]), 0.999f, 0.001f));
            } else if (role == 'P' || role == 'X') {
                if (role == 'X') {
                    tmem_ld64(tS, sr);
                    tmem_wait_ld();
                }
                uint32_t pk[32];
                float acc = 0.f;
#pragma unroll
                for (int e = 0; e < 64; e += 2) {
                    const float p0 = ex2(fmaf(__uint_as_float(sr[e]), 0.125f, -1.0f)), p1 = ex2(fmaf(__uint_as_float(sr[e + 1]), 0.125f, -1.0f));
                    acc += p0 + p1;
                    __half2 h = __floats2half2_rn(p0, p1);
                    pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
                }
                sink += acc;
                if (role == 'X') {
                    tmem_st32(tP, pk);
                    tmem_wait_st();
                } else {
#pragma unroll
                    for (int e = 0; e < 32; ++e) sr[e] ^= pk[e] & 1u;
                }
            } else if (role == 'Q' || role == 'R') {          // Q: FFMA + ex2 + F2FP     R: ex2 + F2FP
                uint32_t pk[32];
#pragma unroll
                for (int e = 0; e < 64; e += 2) {
                    float a0 = __uint_as_float(sr[e]), a1 = __uint_as_float(sr[e + 1]);
                    if (role == 'Q') a0 = fmaf(a0, 0.125f, -1.0f), a1 = fmaf(a1, 0.125f, -1.0f);
                    const float p0 = ex2(a0), p1 = ex2(a1);
                    __half2 h = __floats2half2_rn(p0, p1);
                    pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) sr[e] ^= pk[e] & 1u;
            } else if (role == 'T' || role == 'U') {          // packed f32x2 as in the kernel: T with row sums, U without
                uint32_t pk[32];
                unsigned long long acc0 = 0, acc1 = 0, cc, nb;
                asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(0.125f));
                asm("mov.b64 %0, {%1, %1};" : "=l"(nb) : "f"(-1.0f));
#pragma unroll
                for (int e = 0; e < 64; e += 2) {
                    unsigned long long v, a;
                    asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "r"(sr[e]), "r"(sr[e + 1]));
                    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(a) : "l"(v), "l"(cc), "l"(nb));
                    float a0, a1;
                    asm("mov.b64 {%0, %1}, %2;" : "=f"(a0), "=f"(a1) : "l"(a));
                    const float p0 = ex2(a0), p1 = ex2(a1);
                    if (role == 'T') {
                        unsigned long long pp;
                        asm("mov.b64 %0, {%1, %2};" : "=l"(pp) : "f"(p0), "f"(p1));
                        if (e & 2) asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc1) : "l"(pp));
                        else asm("add.rn.f32x2 %0, %0, %1;" : "+l"(acc0) : "l"(pp));
                    }
                    __half2 h = __floats2half2_rn(p0, p1);
                    pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
                }
                sink += __uint_as_float((uint32_t)(acc0 ^ acc1));
#pragma unroll
                for (int e = 0; e < 32; ++e) sr[e] ^= pk[e] & 1u;
            } else if (role >= 'a' && role <= 'f') {
                const PolyC kc = make_polyc();
                if (role == 'a') mix_loop<1>(sr, kc);
                else if (role == 'b') mix_loop<2>(sr, kc);
                else if (role == 'c') mix_loop<3>(sr, kc);
                else if (role == 'd') mix_loop<4>(sr, kc);
                else if (role == 'e') mix_loop<8>(sr, kc);
                else mix_loop<0>(sr, kc);
            } else if (role == 'V') {                         // phase-separated: 64 FFMA, then 64 ex2, then 32 F2FP
#pragma unroll
                for (int e = 0; e < 64; ++e) sr[e] = __float_as_uint(fmaf(__uint_as_float(sr[e]), 0.125f, -1.0f));
#pragma unroll
                for (int e = 0; e < 64; ++e) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+r"(sr[e]));
                uint32_t pk[32];
#pragma unroll
                for (int e = 0; e < 64; e += 2) {
                    __half2 h = __floats2half2_rn(__uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
                    pk[e >> 1] = *reinterpret_cast<uint32_t*>(&h);
                }
#pragma unroll
                for (int e = 0; e < 32; ++e) sr[e] ^= pk[e] & 1u;
            }
            ++iters;
        }
    }
#pragma unroll
    for (int e = 0; e < 64; ++e) sink += __uint_as_float(sr[e]);
    if (sink == 123.456f) out[0].iters[0] = -1;
    if (lane == 0) { out[blockIdx.x].iters[warp] = iters; out[blockIdx.x].clk[warp] = t1 - t0; }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

int main() {
    Res* d; static Res h[148];
    cudaMalloc(&d, sizeof(h));
    const char* pats[] = {"QQ--", "ff--", "aa--", "bb--", "cc--", "dd--", "ee--", "QQQQ", "ffff", "aaaa", "bbbb", "cccc", "dddd", "eeee", "e---", "bbb-", "ccc-"};
    for (const char* p : pats) {
        Pattern pat; memcpy(pat.role, p, 4);
        cudaMemset(d, 0, sizeof(h));
        k<<<148, 512>>>(d, pat, 300000);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        printf("%s :", p);
        for (int slot = 0; slot < 4; ++slot) {
            const int w = slot * 4;          // sub-partition 0's warp of this slot
            if (p[slot] == '-') printf("      -   ");
            else printf("  %c %7.1f", p[slot], h[0].iters[w] ? (double)h[0].clk[w] / (double)h[0].iters[w] : 0.0);
        }
        printf("   clk per iteration (%s)\n", cudaGetErrorString(e));
    }
    return 0;
}
