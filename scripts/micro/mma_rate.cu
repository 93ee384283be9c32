// Micro-benchmark: sustained cost of short tcgen05.mma instructions (cta_group::1, kind::f16, M = 128, K = 16) issued by one
// thread - SS form (A, B from shared memory) and TS form (A from TMEM) for several N - and of the pattern the attention
// kernel issues per 128 x 128 tile (3 x SS N=128 + 8 x TS N=48/64).  Operands are zeros; only timing matters.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -I progressive_stable_diffusion_b200/csrc -o scripts/micro/mma_rate scripts/micro/mma_rate.cu -lcuda
#include <cstdio>
#include "tc_util.cuh"

namespace daddk { thread_local char g_last_error[512] = ""; std::atomic<int64_t> g_launches{0}; }
using namespace daddk::tc;

struct Res { long long clk[16]; };

__device__ __forceinline__ void commit_wait(uint64_t* bar, uint32_t& phase) {
    mma_commit(bar);
    mbar_wait(bar, phase);
    phase ^= 1;
}

__global__ void __launch_bounds__(128, 1) k(Res* out, int reps) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t bar;
    __shared__ uint32_t tmem_base;
    for (int i = threadIdx.x; i < 64 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (threadIdx.x == 0) { mbar_init(&bar, 1); asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
    if (threadIdx.x < 32) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tmem_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = tmem_base;
    if (threadIdx.x == 0) {
        const uint64_t da = smem_desc(smem_u32(smem), 16, 1024);                 // K-major A / B tiles
        const uint64_t db = smem_desc(smem_u32(smem + 16384), 16, 1024);
        const uint64_t dv = smem_desc(smem_u32(smem + 32768), 16384, 1024);      // MN-major V tile
        uint32_t phase = 0;
        int slot = 0;
        auto run = [&](auto body) {
            body(); commit_wait(&bar, phase);                                    // warm
            const long long t0 = clock64();
            for (int r = 0; r < reps; ++r) body();
            commit_wait(&bar, phase);
            out[blockIdx.x].clk[slot++] = clock64() - t0;
        };
        // 0..3: 16 x SS with N = 256, 128, 64, 48
        for (int n : {256, 128, 64, 48}) {
            const uint32_t id = instr_desc(0, n, 0);
            run([&] { for (int i = 0; i < 16; ++i) mma_ss(tmem, da, db, id, i > 0); });
        }
        // 4..7: 16 x TS (A = TMEM cols 256.., 16-bit) with N = 256 (K-major B), then MN-major B with N = 128?, 64, 48
        for (int n : {128, 64, 48, 16}) {
            const uint32_t id = instr_desc(0, n <= 64 ? n : 64, 1);
            run([&] { for (int i = 0; i < 16; ++i) mma_ts(tmem, tmem + 256 + (i & 7) * 8, dv + (uint64_t)(((i & 7) * 2048) >> 4), id, i > 0); });
        }
        // 8: attention tile pattern, 3 x SS N=128 + 8 x TS N=48 (v3)    9: ... TS N=64 (v2)
        for (int n : {48, 64}) {
            const uint32_t idq = instr_desc(0, 128, 0), idp = instr_desc(0, n, 1);
            run([&] {
                for (int ks = 0; ks < 3; ++ks) mma_ss(tmem, da + (uint64_t)((ks * 32) >> 4), db + (uint64_t)((ks * 32) >> 4), idq, ks > 0);
                for (int kk = 0; kk < 8; ++kk) mma_ts(tmem + 384, tmem + 128 + kk * 8, dv + (uint64_t)((kk * 2048) >> 4), idp, kk > 0);
            });
        }
        // 10: same pattern with a commit after each half (as the kernel does: 3 commits per tile)
        {
            const uint32_t idq = instr_desc(0, 128, 0), idp = instr_desc(0, 48, 1);
            __shared__ uint64_t dummy[2];
            if (true) { mbar_init(&dummy[0], 1); mbar_init(&dummy[1], 1); }
            run([&] {
                for (int ks = 0; ks < 3; ++ks) mma_ss(tmem, da + (uint64_t)((ks * 32) >> 4), db + (uint64_t)((ks * 32) >> 4), idq, ks > 0);
                mma_commit(&dummy[0]);
                for (int kk = 0; kk < 8; ++kk) mma_ts(tmem + 384, tmem + 128 + kk * 8, dv + (uint64_t)((kk * 2048) >> 4), idp, kk > 0);
                mma_commit(&dummy[1]);
            });
        }
        // 11: 16 x SS N=128 to DIFFERENT accumulators (no accumulate dependency)
        {
            const uint32_t id = instr_desc(0, 128, 0);
            run([&] { for (int i = 0; i < 16; ++i) mma_ss(tmem + (i & 3) * 128, da, db, id, 0); });
        }
        // 12: 16 x TS N=48 to different accumulators
        {
            const uint32_t id = instr_desc(0, 48, 1);
            run([&] { for (int i = 0; i < 16; ++i) mma_ts(tmem + 384 + (i & 1) * 64, tmem + 128 + (i & 7) * 8, dv + (uint64_t)(((i & 7) * 2048) >> 4), id, 0); });
        }
        // 13..15: is the ISSUE of tcgen05.mma asynchronous?  clocks until the issuing thread is past 1 / 4 / 16 TS MMAs (N = 48),
        // measured on an idle pipe, without waiting for completion (slot 13, 14, 15) - compare with 62 clk of execution each
        {
            const uint32_t id = instr_desc(0, 48, 1);
            for (int n : {1, 4, 16}) {
                commit_wait(&bar, phase);
                const long long t0 = clock64();
                for (int i = 0; i < n; ++i) mma_ts(tmem + 384, tmem + 128 + (i & 7) * 8, dv + (uint64_t)(((i & 7) * 2048) >> 4), id, i > 0);
                out[blockIdx.x].clk[slot++] = (clock64() - t0) * reps;       // (printed per rep)
                commit_wait(&bar, phase);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (threadIdx.x < 32) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem));
    }
}

int main() {
    Res* d; Res h[148];
    cudaMalloc(&d, sizeof(h));
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 66 * 1024);
    const int reps = 200;
    const char* names[] = {"16 x SS N=256", "16 x SS N=128", "16 x SS N=64", "16 x SS N=48", "16 x TS N=64(req 128)", "16 x TS N=64", "16 x TS N=48",
                           "16 x TS N=16", "tile: 3 SS N=128 + 8 TS N=48", "tile: 3 SS N=128 + 8 TS N=64", "tile (N=48) + 2 commits",
                           "16 x SS N=128, 4 accumulators", "16 x TS N=48, 2 accumulators", "issue only: 1 TS MMA", "issue only: 4 TS MMAs", "issue only: 16 TS MMAs"};
    for (int grid : {1, 148}) {
        cudaMemset(d, 0, sizeof(h));
        k<<<grid, 128, 66 * 1024>>>(d, reps);
        cudaError_t e = cudaDeviceSynchronize();
        printf("grid %d: %s\n", grid, cudaGetErrorString(e));
        cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
        for (int s = 0; s < 16; ++s) {
            const double per_rep = (double)h[0].clk[s] / reps;
            const int n_mma = s == 13 ? 1 : s == 14 ? 4 : (s < 8 || s >= 11) ? 16 : 11;
            printf("  %-34s %9.1f clk per rep, %7.1f clk per MMA\n", names[s], per_rep, per_rep / n_mma);
        }
    }
    return 0;
}
