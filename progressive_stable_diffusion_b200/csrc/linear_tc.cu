// Linear layers of the transformer blocks as one persistent tcgen05 GEMM with the neighbouring elementwise work fused in:
//   Y[m][n] = sum_k X[m][k] W[n][k] (+ bias[n]) (+ R[m][n])          X: [M][K], W: [N][K] (nn.Linear layout), Y, R: [M][N]
// Serves, inside every Transformer2DModel / BasicTransformerBlock of the UNet reached through src/models/unet/unet.py:140-146
// (diffusers graph): proj_in, the fused QKV projection, to_q of the cross-attention,
// the attention output projections, the feed-forward output projection fused with its residual add (`ff(x) + x`) and
// proj_out fused with the block residual.  The shapes are short-K (K = 320 .. 1280 for most of them), where a GEMM is bound
// by its epilogue: the accumulator is double-buffered in TMEM so that tile i + 1 is multiplied while tile i drains.
//
// CTAs run as PAIRS (cluster of two, tcgen05 cta_group::2) on one 256-row x BN-column tile: each CTA loads its own 128 rows of X and
// HALF of the W tile (BN / 2 rows); the pair's leader issues M = 256 MMAs that read both CTAs' shared memory and write both CTAs'
// TMEM.  What bounds these short-K GEMMs on B200 is the bytes an SM takes in per flop (~45 B/clk per SM measured: the one-CTA form
// ingests 16 + BN / 8 KB per K-block and lost 10-25 % to cuBLAS; TMA-multicasting W to both CTAs changed nothing - every SM still
// ingests the whole W tile; profiles/r01_linear_gemm.txt, r02_linear_gemm.txt); the pair form ingests 16 + BN / 16 KB.
//   warp 8     TMA producer: ring of K-blocks {X 128 x 64, W BN x 64}, 128-byte swizzle, zero-filled edges;
//   warp 9     one elected thread issues tcgen05.mma (SS, M128 x N=BN x K16) into accumulator (i & 1);
//   warps 0-7  epilogue, two groups of four warps (TMEM lane == row): group g takes the 64-column chunks g, g + 2 of the tile:
//              tcgen05.ld -> + bias -> 16 bits (+ residual, read straight from global memory) -> swizzled staging panel -> TMA store.
//              The result is rounded to 16 bits before the residual is added, exactly like the unfused GEMM + add it replaces.
// BN = 256 / 192 when N is a multiple, else 128 (N % 64 == 0 required; edge tiles are clipped by the tensor maps).
#include <cstdlib>

#include "tc_util.cuh"

namespace daddk {
namespace lin {

using namespace daddk::tc;

constexpr int NTHREADS = 320;
constexpr int BMR = 128, BK = 64;
constexpr uint32_t A_BYTES = BMR * BK * 2;

template <int STAGES>
struct Bars {
    uint64_t full[STAGES], empty[STAGES];
    uint64_t acc_full[2], acc_empty[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_rank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// ---- CTA-pair (cta_group::2) forms: the leader's MMA reads both CTAs' shared memory and writes both CTAs' TMEM
__device__ __forceinline__ uint32_t map_to_rank(uint32_t local_smem_addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_smem_addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void tma_load_2d_pair(uint32_t dst, const CUtensorMap* map, uint32_t leader_bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(leader_bar), "r"(c0), "r"(c1)
                 : "memory");
}
__device__ __forceinline__ void mma_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {      // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}

template <typename T, int BN, int STAGES, bool PAIR>
__global__ void __launch_bounds__(NTHREADS, 1)
linear_kernel(const __grid_constant__ CUtensorMap tx, const __grid_constant__ CUtensorMap tw, const __grid_constant__ CUtensorMap ty,
              const float* __restrict__ bias, const T* __restrict__ res, int M, int N, int K) {
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t FMT = std::is_same_v<T, __nv_bfloat16> ? 1u : 0u;
    // M = 256 over the CTA pair (bits 24-28 of the instruction descriptor hold M >> 4), N = BN: each CTA supplies 128 rows of X and BN / 2 rows of W
    // (PAIR = false: every CTA is its own tile of 128 rows, cta_group::1 - the form short K = 320 GEMMs run fastest in.)
    constexpr uint32_t IDESC = PAIR ? ((instr_desc(FMT, BN, 0) & ~(0x1Fu << 24)) | ((256u >> 4) << 24)) : instr_desc(FMT, BN, 0);
    constexpr uint32_t B_BYTES = (PAIR ? BN / 2 : BN) * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int NCH = BN / 64;                                  // 64-column output chunks per tile
    constexpr uint32_t OUT_PANEL = 128 * 128;                     // 128 rows x 64 16-bit columns, 128-byte swizzle
    static_assert(BN % 64 == 0 && BN <= 256, "tile width");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sStage = smem;                                 // [STAGES]{X, W}
    unsigned char* sOut = sStage + STAGES * STAGE_BYTES;          // [2 warp groups][2 panels]
    float* sBias = reinterpret_cast<float*>(sOut + 4 * OUT_PANEL);  // [2 accumulators][BN]
    Bars<STAGES>* bars = reinterpret_cast<Bars<STAGES>*>(sBias + 2 * BN);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    // PAIR: cluster = CTA pair on two adjacent row blocks; else every CTA stands alone ("cluster" = CTA, rank 0, one row block per tile)
    const int rank = PAIR ? (int)cluster_rank() : 0, cluster = PAIR ? (int)blockIdx.x >> 1 : (int)blockIdx.x;
    const int nclusters = PAIR ? (int)gridDim.x >> 1 : (int)gridDim.x;
    constexpr int TROWS = PAIR ? 2 * BMR : BMR;
    const int nt = (N + BN - 1) / BN;
    const int pairs = ((M + TROWS - 1) / TROWS) * nt;
    const int my_tiles = (pairs - cluster + nclusters - 1) / nclusters;
    const int kblocks = (K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);                            // the leader's commit, multicast to both CTAs
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->acc_full[a], 1);
            mbar_init(&bars->acc_empty[a], PAIR ? 4 : 2);             // one arrival per epilogue warp group (PAIR: of both CTAs, on the leader's barrier)
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        if constexpr (PAIR) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync();                               // the peer's barriers exist before anything is signalled across
    fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer (n-tile fastest: the X rows of a row
            // block are read by neighbouring CTAs at the same time and hit L2; W is L2-resident throughout)
            uint32_t it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int pair = cluster + i * nclusters;
                const int m0 = (pair / nt) * TROWS + rank * BMR, n0 = (pair % nt) * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t st = it % STAGES;
                    mbar_wait(&bars->empty[st], ((it / STAGES) & 1) ^ 1);
                    const uint32_t base = smem_u32(sStage + st * STAGE_BYTES);
                    if constexpr (PAIR) {
                        if (rank == 0) mbar_expect_tx(&bars->full[st], 2 * STAGE_BYTES);   // both CTAs' loads complete on the leader's barrier
                        const uint32_t lbar = map_to_rank(smem_u32(&bars->full[st]), 0);
                        tma_load_2d_pair(base, &tx, lbar, kb * BK, m0);
                        tma_load_2d_pair(base + A_BYTES, &tw, lbar, kb * BK, n0 + rank * (BN / 2));
                    } else {
                        mbar_expect_tx(&bars->full[st], STAGE_BYTES);
                        tma_load_2d(base, &tx, &bars->full[st], kb * BK, m0);
                        tma_load_2d(base + A_BYTES, &tw, &bars->full[st], kb * BK, n0);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ---------------------------------------------------------------------- MMA issuer
        const bool leader = elect_one();
        uint32_t it = 0;
        for (int i = 0; i < (rank == 0 ? my_tiles : 0); ++i) {                   // the pair's leader issues for both CTAs
            const int a = i & 1;
            if (i >= 2) mbar_wait(&bars->acc_empty[a], ((i >> 1) - 1) & 1);      // both epilogues have drained this accumulator
            fence_after();
            for (int kb = 0; kb < kblocks; ++kb, ++it) {
                const uint32_t st = it % STAGES;
                mbar_wait(&bars->full[st], (it / STAGES) & 1);
                fence_after();
                if (leader) {
                    const uint32_t base = smem_u32(sStage + st * STAGE_BYTES);
                    const uint64_t da = smem_desc(base, 16, 1024), db = smem_desc(base + A_BYTES, 16, 1024);
#pragma unroll
                    for (int k = 0; k < BK / 16; ++k)
                        if constexpr (PAIR) mma_ss_pair(tmem + a * 256, desc_add(da, k * 32), desc_add(db, k * 32), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                        else mma_ss(tmem + a * 256, desc_add(da, k * 32), desc_add(db, k * 32), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                    if constexpr (PAIR) {
                        mma_commit_pair(&bars->empty[st]);
                        if (kb + 1 == kblocks) mma_commit_pair(&bars->acc_full[a]);
                    } else {
                        mma_commit(&bars->empty[st]);
                        if (kb + 1 == kblocks) mma_commit(&bars->acc_full[a]);
                    }
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue (two groups of four warps)
        // Group g takes the 64-column chunks g, g + 2 of the tile: TMEM row -> + bias -> 16 bits (+ residual) -> one of the group's two
        // swizzled staging panels -> TMA store (rows beyond M and columns beyond N are clipped by the tensor map).  A panel is
        // rewritten once the store issued from it two chunks ago has been read out of shared memory.
        const int group = warp >> 2;
        const int row = (warp & 3) * 32 + lane;                       // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool store_leader = (tid & 127) == 0;
        uint32_t pc = 0;                                              // chunks this group has staged so far
        for (int i = 0; i < my_tiles; ++i) {
            const int pair = cluster + i * nclusters;
            const int m0 = (pair / nt) * TROWS + rank * BMR, n0 = (pair % nt) * BN;
            const int a = i & 1;
            float* bs = sBias + a * BN;
            if (tid < BN) bs[tid] = (bias && n0 + tid < N) ? bias[n0 + tid] : 0.0f;      // (slot last read two tiles ago)
            mbar_wait(&bars->acc_full[a], (i >> 1) & 1);
            fence_after();
            named_sync(1, 256);
            const bool row_ok = m0 + row < M;
#pragma unroll 1
            for (int c = group; c < NCH; c += 2, ++pc) {
                unsigned char* panel = sOut + (group * 2 + (pc & 1)) * OUT_PANEL;
                if (store_leader) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
                named_sync(2 + group, 128);
                uint32_t v[64];
                tmem_ld32(tmem + a * 256 + c * 64 + lane_base, *reinterpret_cast<uint32_t(*)[32]>(v));
                tmem_ld32(tmem + a * 256 + c * 64 + 32 + lane_base, *reinterpret_cast<uint32_t(*)[32]>(v + 32));
                uint4 rr[8];
                const bool col_ok = n0 + c * 64 < N;                  // N % 64 == 0: a chunk is inside or outside as a whole
                if (res && row_ok && col_ok) {
                    const uint4* rp = reinterpret_cast<const uint4*>(res + (size_t)(m0 + row) * N + n0 + c * 64);
#pragma unroll
                    for (int j = 0; j < 8; ++j) rr[j] = rp[j];
                }
                tmem_wait_ld();
                const bool last_read = c + 2 >= NCH;                  // the group's last read of this accumulator
                if (last_read) fence_before();
                const float* bc = bs + c * 64;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const float4 b0 = *reinterpret_cast<const float4*>(bc + j * 8), b1 = *reinterpret_cast<const float4*>(bc + j * 8 + 4);
                    Vec8<T> t;
                    t.raw.x = pack2<T>(__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y);
                    t.raw.y = pack2<T>(__uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w);
                    t.raw.z = pack2<T>(__uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y);
                    t.raw.w = pack2<T>(__uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w);
                    if (res && row_ok && col_ok) {                    // (product rounded to 16 bits first, like the unfused GEMM + add)
                        Vec8<T> q;
                        q.raw = rr[j];
                        float f[8], g[8];
                        t.unpack(f);
                        q.unpack(g);
#pragma unroll
                        for (int e = 0; e < 8; ++e) f[e] += g[e];
                        t.pack(f);
                    }
                    *reinterpret_cast<uint4*>(panel + row * 128 + ((j ^ (row & 7)) << 4)) = t.raw;      // 128-byte swizzle
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                named_sync(2 + group, 128);
                if (store_leader) {
                    if (last_read) {                                  // hand the accumulator back to the MMA warp: ONE (remote) arrival per group
                        if (rank == 0) mbar_arrive(&bars->acc_empty[a]);
                        else mbar_arrive_remote(map_to_rank(smem_u32(&bars->acc_empty[a]), 0));
                    }
                    tma_store_2d(&ty, smem_u32(panel), n0 + c * 64, m0);
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            }
        }
        if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    if constexpr (PAIR) cluster_sync();                               // no CTA leaves while its peer may still signal its barriers
    if (warp == 9) {
        fence_after();
        if constexpr (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

// 2-D row-major (rows, cols) 16-bit tensor, box = 64 columns x box_rows, 128-byte swizzle
static int make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int dtype, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail("%s: cuTensorMapEncodeTiled is unavailable", "dadd_linear_fwd");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype == DADD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("%s: cuTensorMapEncodeTiled failed (CUresult %lld)", "dadd_linear_fwd", (long long)r);
    return 0;
}

template <typename T, int BN, int STAGES, bool PAIR>
static int launch(const void* x, const void* w, const float* bias, const void* res, void* y, int64_t M, int N, int K, int dtype,
                  cudaStream_t s) {
    CUtensorMap tx, tw, ty;
    if (make_map_2d(&tx, x, M, K, dtype, BMR) || make_map_2d(&tw, w, N, K, dtype, PAIR ? BN / 2 : BN) || make_map_2d(&ty, y, M, N, dtype, BMR)) return 1;
    const size_t smem = (size_t)STAGES * (A_BYTES + (PAIR ? BN / 2 : BN) * BK * 2) + (size_t)4 * 128 * 128 + 2 * BN * sizeof(float) + sizeof(Bars<STAGES>) + 1024;
    const int64_t pairs = ((M + (PAIR ? 2 : 1) * BMR - 1) / ((PAIR ? 2 : 1) * BMR)) * ((N + BN - 1) / BN);
    auto kern = linear_kernel<T, BN, STAGES, PAIR>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "linear smem")) return 2;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = PAIR ? 2 : 1;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(NTHREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    static int max_clusters = 0;                       // co-resident CTA pairs / CTAs (one CTA per SM): the persistent grid (per instance)
    if (max_clusters == 0) {
        if (PAIR) {
            cfg.gridDim = dim3(num_sms() & ~1, 1, 1);
            if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters <= 0) max_clusters = num_sms() / 2;
        } else {
            max_clusters = num_sms();
        }
    }
    const int clusters = (int)(pairs < max_clusters ? pairs : max_clusters);
    cfg.gridDim = dim3((PAIR ? 2 : 1) * clusters, 1, 1);
    const int Mi = (int)M;
    const T* resT = (const T*)res;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tx, tw, ty, bias, resT, Mi, N, K);
    if (e != cudaSuccess) return cuda_ok(e, "dadd_linear_fwd launch");
    return launched("dadd_linear_fwd");
}

}  // namespace lin
}  // namespace daddk

using namespace daddk;

extern "C" int dadd_linear_supported(int64_t M, int N, int K) { return (M > 0 && K > 0 && K % 8 == 0 && N > 0 && N % 64 == 0) ? 1 : 0; }

extern "C" int dadd_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y, int64_t M, int N, int K,
                               int dtype, void* stream) {
    DADD_REQUIRE(x && w && y, "dadd_linear_fwd");
    DADD_REQUIRE(dtype16_ok(dtype), "dadd_linear_fwd");
    DADD_REQUIRE(M >= 0 && M < (1ll << 31) - 128, "dadd_linear_fwd");
    if (M == 0) return 0;
    if (!dadd_linear_supported(M, N, K))
        return fail("%s: needs K %% 8 == 0 and N %% 64 == 0 (N = %lld, K = %lld)", "dadd_linear_fwd", (long long)N, (long long)K);
    DADD_REQUIRE(((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)residual) % 16 == 0, "dadd_linear_fwd");
    cudaStream_t s = (cudaStream_t)stream;
    // tile width: 256 / 192 where N is a multiple, else 128 (N = 320: the third tile is half empty; the tensor maps clip it)
    static const int force_bn = [] { const char* e = getenv("DADD_LIN_BN"); return e ? atoi(e) : 0; }();
    static const int force_pair = [] { const char* e = getenv("DADD_LIN_PAIR"); return e ? atoi(e) : -1; }();
    const int bn = force_bn ? force_bn : (N % 256 == 0 ? 256 : (N % 192 == 0 ? 192 : 128));
    // CTA pairs (cta_group::2) pay a cross-CTA hand-off per tile: they win from K = 640 on, short K = 320 tiles run faster alone
    const bool pair = force_pair >= 0 ? force_pair != 0 : K >= 512;
    if (pair) {
        if (bn == 256) DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 256, 4, true>(x, w, bias, residual, y, M, N, K, dtype, s)));
        if (bn == 192) DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 192, 4, true>(x, w, bias, residual, y, M, N, K, dtype, s)));
        DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 128, 6, true>(x, w, bias, residual, y, M, N, K, dtype, s)));
    }
    if (bn == 256) DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 256, 3, false>(x, w, bias, residual, y, M, N, K, dtype, s)));
    if (bn == 192) DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 192, 3, false>(x, w, bias, residual, y, M, N, K, dtype, s)));
    DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 128, 4, false>(x, w, bias, residual, y, M, N, K, dtype, s)));
    return 1;
}
