// Micro-benchmark: does TMEM / MUFU traffic of the softmax warps slow the tcgen05.mma stream of the attention kernel?
// Warp 0 issues the per-tile MMA pattern of self_attn_tc3 (3 x SS N=128 + 8 x TS N=48, K = 16 each) back to back while 16
// background warps (lane quarter = warp & 3, as in the kernel) run one of: nothing, tcgen05.ld only, ex2 only, tcgen05.st only,
// ld + ex2 + st (the softmax loop without its barriers).  Prints clk per tile for the MMA stream and the background warps'
// iteration rate, per mode, plus the same with the PV MMAs in SS form (P from shared memory, N = 48).
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I progressive_stable_diffusion_b200/csrc -o scripts/micro/mma_contend scripts/micro/mma_contend.cu -lcuda
#include <cstdio>
#include "tc_util.cuh"

namespace daddk { thread_local char g_last_error[512] = ""; std::atomic<int64_t> g_launches{0}; }
using namespace daddk::tc;

struct Res { long long mma_clk; long long bg_iters; long long bg_clk; };

__global__ void __launch_bounds__(640, 1) k(Res* out, int reps, int mode, int pv_form, int mma_on) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    __shared__ volatile int done;
    __shared__ long long iters_sh[20];
    for (int i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { done = 0; mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        if (lane == 0) {
            const uint64_t da = smem_desc(smem_u32(smem), 16, 1024);
            const uint64_t db = smem_desc(smem_u32(smem + 16384), 16, 1024);
            const uint64_t dv = smem_desc(smem_u32(smem + 32768), 16384, 1024);
            const uint64_t dp = smem_desc(smem_u32(smem + 49152), 16, 1024);       // P as a K-major smem tile (2 panels of 64 keys)
            const uint32_t idq = instr_desc(0, 128, 0), idp = instr_desc(0, 48, 1);
            uint32_t phase = 0;
            auto tile = [&](int i) {
                const uint32_t tS = tmem + (i % 3) * 128;
                for (int ks = 0; ks < 3; ++ks) mma_ss(tS, da + (uint64_t)((ks * 32) >> 4), db + (uint64_t)((ks * 32) >> 4), idq, ks > 0);
                if (pv_form == 0) {
                    for (int kk = 0; kk < 8; ++kk) mma_ts(tmem + 384 + (i & 1) * 64, tS + kk * 8, dv + (uint64_t)((kk * 2048) >> 4), idp, kk > 0);
                } else {
                    for (int kk = 0; kk < 8; ++kk)
                        mma_ss(tmem + 384 + (i & 1) * 64, dp + (uint64_t)(((kk >> 2) * 16384 + (kk & 3) * 32) >> 4), dv + (uint64_t)((kk * 2048) >> 4), idp, kk > 0);
                }
            };
            if (mma_on) {
                tile(0);
                mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1;
                const long long t0 = clock64();
                for (int r = 0; r < reps; ++r) tile(r);
                mma_commit(&bar); mbar_wait(&bar, phase); phase ^= 1;
                out[blockIdx.x].mma_clk = clock64() - t0;
            } else {
                const long long t0 = clock64();
                while (clock64() - t0 < 400000) {}
                out[blockIdx.x].mma_clk = 0;
            }
            done = 1;
        }
    } else if (warp >= 4) {
        const int w = warp - 4, quarter = warp & 3, g = w >> 3, half = (w >> 2) & 1;
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        long long iters = 0;
        float sink = 0.f;
        uint32_t sr[64];
#pragma unroll
        for (int e = 0; e < 64; ++e) sr[e] = threadIdx.x + e;
        const long long t0 = clock64();
        while (mode != 0) {
            int d = done;
            d = __shfl_sync(0xffffffffu, d, 0);
            if (d) break;
            const uint32_t tS = tmem + ((iters * 2 + g) % 3) * 128 + lane_base;
            if (mode == 1 || mode == 4) {
                tmem_ld64(tS + half * 64, sr);
                tmem_wait_ld();
            }
            if (mode == 2 || mode == 4) {
#pragma unroll
                for (int e = 0; e < 64; ++e) sr[e] = __float_as_uint(ex2(__uint_as_float(sr[e]) * 1e-30f));
            }
            if (mode == 3 || mode == 4) {
                uint32_t pk[32];
#pragma unroll
                for (int e = 0; e < 32; ++e) pk[e] = sr[2 * e] ^ sr[2 * e + 1];
                tmem_st32(tmem + 256 + g * 64 + half * 32 + lane_base, pk);   // a region the MMAs do not touch
                tmem_wait_st();
            }
#pragma unroll
            for (int e = 0; e < 64; ++e) sink += __uint_as_float(sr[e]);
            ++iters;
        }
        const long long t1 = clock64();
        if (sink == 123.456f) out[0].bg_iters = -1;
        if (lane == 0) iters_sh[warp] = iters;
        if (warp == 4 && lane == 0) out[blockIdx.x].bg_clk = t1 - t0;
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x == 0) {
        long long s = 0;
        for (int w2 = 4; w2 < 20; ++w2) s += iters_sh[w2];
        out[blockIdx.x].bg_iters = mode ? s : 0;
    }
    if (threadIdx.x < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

int main() {
    Res* d; Res h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    const int reps = 400;
    const char* modes[] = {"idle", "tcgen05.ld only", "ex2 only", "tcgen05.st only", "ld + ex2 + st"};
    for (int mma_on : {1, 0})
        for (int pv_form : {0, 1}) {
            if (!mma_on && pv_form) continue;
            for (int mode = 0; mode < 5; ++mode) {
                if (!mma_on && mode == 0) continue;
                cudaMemset(d, 0, sizeof(h));
                k<<<148, 640, 100 * 1024>>>(d, reps, mode, pv_form, mma_on);
                cudaError_t e = cudaDeviceSynchronize();
                cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
                // one background iteration = 16 warps x (8 KB ld, 2048 ex2, 4 KB st) -> per "tile" (8 warps) figures
                const double it_clk = h[0].bg_iters ? (double)h[0].bg_clk / ((double)h[0].bg_iters / 16.0) : 0.0;
                printf("mma %s pv=%s bg=%-16s : %8.1f clk per MMA tile | bg: %8.1f clk per iteration per warp (%s)\n", mma_on ? "on " : "off",
                       pv_form ? "SS" : "TS", modes[mode], mma_on ? (double)h[0].mma_clk / reps : 0.0, it_clk, cudaGetErrorString(e));
            }
        }
    return 0;
}
