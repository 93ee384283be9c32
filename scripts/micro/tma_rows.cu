// Micro-benchmark: how fast does the TMA unit of one SM deliver box rows when the rows are NARROW (a head of d = 40 is 80
// bytes per token inside the fused (B, N, 3C) projection)?  148 CTAs, each streams [128 rows x 64 cols] 128-byte-swizzled boxes
// of a (d, N, H, B) view through a 4-deep mbarrier ring; reports clocks per box row per SM and the equivalent GB/s.
#include <atomic>
#include <cstdio>
#include <cuda_runtime.h>
#include "../../progressive_stable_diffusion_b200/csrc/tc_util.cuh"
namespace daddk { thread_local char g_last_error[512]; std::atomic<int64_t> g_launches{0}; }
using namespace daddk::tc;

template <int DEPTH>
__global__ void __launch_bounds__(128, 1) k(const __grid_constant__ CUtensorMap tm, int iters, int N, int H, int B, int box_rows, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    __shared__ uint64_t full[DEPTH];
    if (threadIdx.x == 0) {
        for (int s = 0; s < DEPTH; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        const int tiles_n = N / box_rows;
        long long t0 = clock64();
        for (int it = 0; it < iters + DEPTH; ++it) {
            const int st = it % DEPTH;
            if (it >= DEPTH) mbar_wait(&full[st], ((it / DEPTH) - 1) & 1);
            if (it < iters) {
                const int id = blockIdx.x + it * gridDim.x;
                const int t = id % tiles_n, h = (id / tiles_n) % H, b = (id / (tiles_n * H)) % B;
                mbar_expect_tx(&full[st], box_rows * 128);
                tma_load_4d(smem_u32(smem + st * 16384), &tm, &full[st], 0, t * box_rows, h, b);
            }
        }
        cyc[blockIdx.x] = clock64() - t0;
    }
}

// the same tiles fetched by the LSU: 128 threads x 16-byte cp.async into the 128-byte-swizzled layout, DEPTH groups in flight
template <int DEPTH>
__global__ void __launch_bounds__(128, 1) k_lsu(const __nv_bfloat16* base, int iters, int N, int H, int B, int d, long long* cyc) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int tiles_n = N / 128, C3 = 3 * H * d, chunks = d / 8;
    long long t0 = clock64();
    for (int it = 0; it < iters + DEPTH - 1; ++it) {
        if (it < iters) {
            const int id = blockIdx.x + it * gridDim.x;
            const int t = id % tiles_n, h = (id / tiles_n) % H, b = (id / (tiles_n * H)) % B;
            const __nv_bfloat16* src = base + ((size_t)b * N + t * 128) * C3 + h * d;
            unsigned char* dst = smem + (it % DEPTH) * 16384;
            for (int i = threadIdx.x; i < 128 * chunks; i += 128) {
                const int r = i / chunks, c = i % chunks;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst + r * 128 + ((c ^ (r & 7)) << 4))),
                             "l"(src + (size_t)r * C3 + c * 8) : "memory");
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
        asm volatile("cp.async.wait_group %0;" ::"n"(DEPTH - 1) : "memory");
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    if (threadIdx.x == 0) cyc[blockIdx.x] = clock64() - t0;
}

int main() {
    const int B = 26, N = 1024, H = 8;
    long long* cyc; cudaMalloc(&cyc, 148 * 8);
    for (int d : {40, 64, 80}) {
        const int C = H * d;
        void* buf; cudaMalloc(&buf, (size_t)B * N * 3 * C * 2); cudaMemset(buf, 0, (size_t)B * N * 3 * C * 2);
        auto report = [&](const char* what, int depth, int box_rows, int iters, float ms, cudaError_t e) {
            long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
            double avg = 0; for (int i = 0; i < 148; ++i) avg += h[i]; avg /= 148;
            printf("%-4s d=%3d (%3d B rows) box_rows=%3d depth=%2d: %5.2f clk per row per SM, %6.1f us, %5.0f GB/s useful (%s)\n", what, d, d * 2, box_rows,
                   depth, avg / ((double)iters * box_rows), ms * 1e3, 148.0 * iters * box_rows * d * 2 / (ms * 1e-3) / 1e9, cudaGetErrorString(e));
        };
        const int iters = 400;
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        float ms;
#define RUN_TMA(DEPTH, BOX) { CUtensorMap tm; if (make_map(&tm, buf, 3 * C, B, H, N, d, DADD_BF16, BOX)) { printf("map failed\n"); return 1; } \
            cudaFuncSetAttribute(k<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * 16384 + 1024); \
            for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k<DEPTH><<<148, 128, DEPTH * 16384 + 1024>>>(tm, iters, N, H, B, BOX, cyc); cudaEventRecord(e1); } \
            cudaError_t e = cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1); report("TMA", DEPTH, BOX, iters, ms, e); }
#define RUN_LSU(DEPTH) { cudaFuncSetAttribute(k_lsu<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEPTH * 16384 + 1024); \
            for (int rep = 0; rep < 2; ++rep) { cudaEventRecord(e0); k_lsu<DEPTH><<<148, 128, DEPTH * 16384 + 1024>>>((const __nv_bfloat16*)buf, iters, N, H, B, d, cyc); cudaEventRecord(e1); } \
            cudaError_t e = cudaDeviceSynchronize(); cudaEventElapsedTime(&ms, e0, e1); report("LSU", DEPTH, 128, iters, ms, e); }
        RUN_TMA(4, 128) RUN_TMA(8, 128) RUN_TMA(12, 128) RUN_TMA(12, 64)
        RUN_LSU(2) RUN_LSU(4) RUN_LSU(8)
        cudaFree(buf);
    }
    return 0;
}
