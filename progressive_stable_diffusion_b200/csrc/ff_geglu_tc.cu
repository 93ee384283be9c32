// Feed-forward input projection fused with its GEGLU gate on the 5th-generation tensor cores:
//   Y[m][j] = (X W_v^T + b_v)[m][j] * gelu_erf((X W_g^T + b_g)[m][j]),   W = [W_v ; W_g]  (2 inner x K, nn.Linear layout)
// Reference: diffusers FeedForward.net[0] = GEGLU(dim, 4 dim): `hidden, gate = proj(x).chunk(2, -1); hidden * F.gelu(gate)`
// inside every BasicTransformerBlock of the UNet reached through src/models/unet/unet.py:140-146 (SURVEY.md A.5, section
// 8f row f4).  The library path writes the (M, 2 inner) projection to HBM and reads it back in a separate GEGLU pass - the
// largest activation of the network (545 MB at 104 x 1024 x 2560); here it never leaves the SM.
//
// Persistent CTA per SM, tile = 128 rows x (128 value + 128 gate) columns.  CTAs run as clusters of two that work on the same
// columns of two adjacent 128-row blocks: each CTA fetches ONE half of the shared weight tile (rank 0: W_v, rank 1: W_g) and
// TMA-multicasts it into both CTAs' shared memory, which cuts the L2 -> SM operand traffic of the pair from 96 to 64 KB per
// K-block (the one-CTA version of this tile is bound by exactly that traffic); the MMAs stay per-CTA (cta_group::1).
//   warp 8   TMA producer: ring of K-blocks {X 128x64, W_v 128x64, W_g 128x64} (128-byte swizzle, zero-filled edges); a
//            stage is reused once BOTH CTAs' MMAs have read it (their commits are multicast to both `empty` barriers);
//   warp 9   one elected thread issues tcgen05.mma (SS, M128 x N256 x K16) into one of two 256-column TMEM accumulators;
//   warps 0-7  epilogue of the other accumulator: warps 0-3 own the 128 rows for columns [0,64), warps 4-7 for [64,128):
//            tcgen05.ld value + gate, bias, GELU, 16-bit pack, swizzled staging panel, TMA store (rows beyond M clipped).
// GELU is the exact (erf) form: Phi(g) = 1/2 erfc(-g / sqrt 2) with erfc(z) = 2^(z R(z)) on [0, 4.25], R a degree-6 minimax
// polynomial (relative error 6.4e-6 in erfc, <= 7.1e-7 absolute in gelu: below half an ulp of the 16-bit output) - one
// MUFU.EX2 per element instead of erff's divergent branches, Horner on packed f32x2 operations, so the epilogue stays under
// the MMA time of a K = 320 tile.
#include <cstdlib>

#include "tc_util.cuh"

namespace daddk {
namespace ffg {

using namespace daddk::tc;

constexpr int NTHREADS = 320;
constexpr int BMR = 128, BNH = 128, BK = 64;                 // rows, value (= gate) columns, K elements per block
constexpr uint32_t A_BYTES = BMR * BK * 2, B_BYTES = 2 * BNH * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
constexpr uint32_t OUT_PANEL = 128 * 128;                     // 128 rows x 64 16-bit columns

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
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
        ::"r"(dst), "l"((uint64_t)map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void mma_commit_mc(uint64_t* bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"(mask)
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
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];" ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1)
                 : "memory");
}

__device__ __forceinline__ uint64_t pack_f2(float lo, float hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
__device__ __forceinline__ void unpack_f2(uint64_t v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ uint64_t fma_f2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}
__device__ __forceinline__ uint64_t add_f2(uint64_t a, uint64_t b) {
    uint64_t d;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
    return d;
}

// (v0 * gelu(g0), v1 * gelu(g1)), gelu(x) = x * Phi(x) in its exact (erf) form: Phi(x) = 1/2 erfc(-x / sqrt 2) with
// erfc(z) = 2^(z R(z)) on [0, 4.25], R a degree-6 minimax polynomial.  In terms of a = |x| (clamped to 4.25 sqrt 2):
// 1/2 erfc = 2^(a R'(a) - 1), R'(a) = R(a / sqrt 2) / sqrt 2 (coefficients folded).  |error| <= 7.1e-7 in gelu.  The Horner chain
// runs on packed f32x2 operations (11 instead of 17.5 instructions per element; measured equal in time on B200 - the tile is
// bound by the MMAs and the power cap, not by the epilogue's issue slots - kept for the headroom).
__device__ __forceinline__ void geglu_pair(float v0, float v1, float g0, float g1, float& o0, float& o1) {
    const float a0 = fminf(fabsf(g0), 6.0104076f), a1 = fminf(fabsf(g1), 6.0104076f);
    const uint64_t a = pack_f2(a0, a1);
    uint64_t r = pack_f2(-1.9058701354879304e-06f, -1.9058701354879304e-06f);
    r = fma_f2(r, a, pack_f2(6.299046071944758e-05f, 6.299046071944758e-05f));
    r = fma_f2(r, a, pack_f2(-0.0009410823695361614f, -0.0009410823695361614f));
    r = fma_f2(r, a, pack_f2(0.008546620607376099f, 0.008546620607376099f));
    r = fma_f2(r, a, pack_f2(-0.054029423743486404f, -0.054029423743486404f));
    r = fma_f2(r, a, pack_f2(-0.45841315388679504f, -0.45841315388679504f));
    r = fma_f2(r, a, pack_f2(-1.1512556076049805f, -1.1512556076049805f));
    float e0, e1;
    unpack_f2(fma_f2(r, a, pack_f2(-1.0f, -1.0f)), e0, e1);
    const float h0 = ex2(e0), h1 = ex2(e1);                  // 1/2 erfc(|x| / sqrt 2)
    o0 = v0 * (g0 * (g0 < 0.0f ? h0 : 1.0f - h0));
    o1 = v1 * (g1 * (g1 < 0.0f ? h1 : 1.0f - h1));
}

template <typename T, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
ff_geglu_kernel(const __grid_constant__ CUtensorMap tx, const __grid_constant__ CUtensorMap tw, const __grid_constant__ CUtensorMap ty,
                const float* __restrict__ bias, int M, int K, int inner) {
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t FMT = std::is_same_v<T, __nv_bfloat16> ? 1u : 0u;
    constexpr uint32_t IDESC = instr_desc(FMT, 2 * BNH, 0);

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sStage = smem;                                 // [STAGES]{X, W_v, W_g}
    unsigned char* sOut = sStage + STAGES * STAGE_BYTES;          // [2 accumulators][2 panels]
    float* sBias = reinterpret_cast<float*>(sOut + 4 * OUT_PANEL);     // [2 accumulators][256]
    Bars<STAGES>* bars = reinterpret_cast<Bars<STAGES>*>(sBias + 512);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int rank = (int)cluster_rank(), cluster = (int)blockIdx.x >> 1, nclusters = (int)gridDim.x >> 1;
    const int nt = inner / BNH;
    const int pairs = ((M + 2 * BMR - 1) / (2 * BMR)) * nt;         // pair = two adjacent row blocks x one column block
    const int my_tiles = (pairs - cluster + nclusters - 1) / nclusters;
    const int kblocks = (K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 2);                            // this CTA's and its peer's MMA commits
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(&bars->acc_full[a], 1);
            mbar_init(&bars->acc_empty[a], 256);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 9) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_before();
    __syncthreads();
    cluster_sync();                                                   // the peer's barriers exist before anything is multicast
    fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer (n-tile fastest: the X rows of a
            // row block are shared by the CTAs working on its neighbours and stay in L2; W is L2-resident throughout)
            uint32_t it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int pair = cluster + i * nclusters;
                const int m0 = (pair / nt) * 2 * BMR + rank * BMR, n0 = (pair % nt) * BNH;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t st = it % STAGES;
                    mbar_wait(&bars->empty[st], ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&bars->full[st], STAGE_BYTES);          // X + my half of W + the peer's half
                    const uint32_t base = smem_u32(sStage + st * STAGE_BYTES);
                    tma_load_2d(base, &tx, &bars->full[st], kb * BK, m0);
                    tma_load_2d_mc(base + A_BYTES + rank * (B_BYTES / 2), &tw, &bars->full[st], kb * BK, rank * inner + n0, (uint16_t)3);
                }
            }
        }
    } else if (warp == 9) {
        // ---------------------------------------------------------------------- MMA issuer
        const bool leader = elect_one();
        uint32_t it = 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int a = i & 1;
            if (i >= 2) mbar_wait(&bars->acc_empty[a], ((i >> 1) - 1) & 1);      // the epilogue has drained this accumulator
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
                        mma_ss(tmem + a * 256, desc_add(da, k * 32), desc_add(db, k * 32), IDESC, (kb > 0 || k > 0) ? 1u : 0u);
                    mma_commit_mc(&bars->empty[st], (uint16_t)3);
                    if (kb + 1 == kblocks) mma_commit(&bars->acc_full[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue (8 warps, 2 column halves)
        const int half = warp >> 2;                                   // columns [64 half, 64 half + 64) of the tile
        const int row = (warp & 3) * 32 + lane;                       // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const bool store_leader = tid == 0;
        for (int i = 0; i < my_tiles; ++i) {
            const int pair = cluster + i * nclusters;
            const int m0 = (pair / nt) * 2 * BMR + rank * BMR, n0 = (pair % nt) * BNH;
            const int a = i & 1;
            float* bs = sBias + a * 256;
            // the staging panels and bias slots of accumulator a were last used by tile i - 2: its store must have left them
            if (store_leader && i >= 2) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
            named_sync(1, 256);
            bs[tid] = bias[tid < 128 ? n0 + tid : inner + n0 + tid - 128];
            mbar_wait(&bars->acc_full[a], (i >> 1) & 1);
            fence_after();
            named_sync(1, 256);
            const uint32_t tv = tmem + a * 256 + half * 64 + lane_base, tg = tv + 128;
            unsigned char* stage = sOut + (a * 2 + half) * OUT_PANEL + row * 128;
#pragma unroll
            for (int c = 0; c < 2; ++c) {                             // 32 columns per round trip to TMEM
                uint32_t v[32], g[32];
                tmem_ld32(tv + c * 32, v);
                tmem_ld32(tg + c * 32, g);
                tmem_wait_ld();
                if (c == 1) {                                         // accumulator fully read: hand it back to the MMA warp
                    fence_before();
                    mbar_arrive(&bars->acc_empty[a]);
                }
                const float* bv = bs + half * 64 + c * 32;
                const float* bg = bv + 128;
#pragma unroll
                for (int j = 0; j < 4; ++j) {                         // 8 columns -> one 16-byte chunk of the staging row
                    const float4 bv0 = *reinterpret_cast<const float4*>(bv + j * 8), bv1 = *reinterpret_cast<const float4*>(bv + j * 8 + 4);
                    const float4 bg0 = *reinterpret_cast<const float4*>(bg + j * 8), bg1 = *reinterpret_cast<const float4*>(bg + j * 8 + 4);
                    const float bvv[8] = {bv0.x, bv0.y, bv0.z, bv0.w, bv1.x, bv1.y, bv1.z, bv1.w};
                    const float bgg[8] = {bg0.x, bg0.y, bg0.z, bg0.w, bg1.x, bg1.y, bg1.z, bg1.w};
                    float o[8];
#pragma unroll
                    for (int e = 0; e < 8; e += 2) {
                        float x0, x1, y0, y1;
                        unpack_f2(add_f2(pack_f2(__uint_as_float(v[j * 8 + e]), __uint_as_float(v[j * 8 + e + 1])), pack_f2(bvv[e], bvv[e + 1])), x0, x1);
                        unpack_f2(add_f2(pack_f2(__uint_as_float(g[j * 8 + e]), __uint_as_float(g[j * 8 + e + 1])), pack_f2(bgg[e], bgg[e + 1])), y0, y1);
                        geglu_pair(x0, x1, y0, y1, o[e], o[e + 1]);
                    }
                    uint4 out;
                    out.x = pack2<T>(o[0], o[1]);
                    out.y = pack2<T>(o[2], o[3]);
                    out.z = pack2<T>(o[4], o[5]);
                    out.w = pack2<T>(o[6], o[7]);
                    const int chunk = c * 4 + j;
                    *reinterpret_cast<uint4*>(stage + ((chunk ^ (row & 7)) << 4)) = out;      // 128-byte swizzle
                }
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            named_sync(1, 256);
            if (store_leader) {
                tma_store_2d(&ty, smem_u32(sOut + (a * 2) * OUT_PANEL), n0, m0);
                tma_store_2d(&ty, smem_u32(sOut + (a * 2 + 1) * OUT_PANEL), n0 + 64, m0);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
        if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    cluster_sync();                                                   // no CTA leaves while its peer may still signal its barriers
    if (warp == 9) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

// 2-D row-major (rows, cols) 16-bit tensor, box = 64 columns x box_rows, 128-byte swizzle
static int make_map_2d(CUtensorMap* map, const void* base, int64_t rows, int64_t cols, int64_t row_stride, int dtype, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail("%s: cuTensorMapEncodeTiled is unavailable", "dadd_ff_geglu_fwd");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)row_stride * 2};
    const cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, dtype == DADD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("%s: cuTensorMapEncodeTiled failed (CUresult %lld)", "dadd_ff_geglu_fwd", (long long)r);
    return 0;
}

template <typename T>
static int launch(const CUtensorMap& tx, const CUtensorMap& tw, const CUtensorMap& ty, const float* bias, int64_t M, int K, int inner,
                  cudaStream_t s) {
    constexpr int STAGES = 3;
    const size_t smem = (size_t)STAGES * STAGE_BYTES + 4 * OUT_PANEL + 512 * sizeof(float) + sizeof(Bars<STAGES>) + 1024;
    const int64_t pairs = ((M + 2 * BMR - 1) / (2 * BMR)) * (inner / BNH);
    auto kern = ff_geglu_kernel<T, STAGES>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "ff_geglu smem")) return 2;
    cudaLaunchConfig_t cfg{};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cfg.blockDim = dim3(NTHREADS, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    static int max_clusters = 0;                       // co-resident CTA pairs (one CTA per SM): the persistent grid
    if (max_clusters == 0) {
        cfg.gridDim = dim3(num_sms() & ~1, 1, 1);
        if (cudaOccupancyMaxActiveClusters(&max_clusters, kern, &cfg) != cudaSuccess || max_clusters <= 0) max_clusters = num_sms() / 2;
    }
    const int clusters = (int)(pairs < max_clusters ? pairs : max_clusters);
    cfg.gridDim = dim3(2 * clusters, 1, 1);
    const int Mi = (int)M;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, tx, tw, ty, bias, Mi, K, inner);
    if (e != cudaSuccess) return cuda_ok(e, "dadd_ff_geglu_fwd launch");
    return launched("dadd_ff_geglu_fwd");
}

}  // namespace ffg
}  // namespace daddk

using namespace daddk;

extern "C" int dadd_ff_geglu_fwd(const void* x, const void* w, const float* bias, void* y, int64_t M, int K, int inner, int dtype,
                                 void* stream) {
    DADD_REQUIRE(x && w && bias && y, "dadd_ff_geglu_fwd");
    DADD_REQUIRE(dtype16_ok(dtype), "dadd_ff_geglu_fwd");
    DADD_REQUIRE(M >= 0 && M < (1ll << 31) - 128 && K > 0 && K % 8 == 0 && inner > 0 && inner % 128 == 0, "dadd_ff_geglu_fwd");
    DADD_REQUIRE(((uintptr_t)x | (uintptr_t)w | (uintptr_t)y) % 16 == 0, "dadd_ff_geglu_fwd");
    if (M == 0) return 0;
    CUtensorMap tx, tw, ty;
    if (ffg::make_map_2d(&tx, x, M, K, K, dtype, ffg::BMR) || ffg::make_map_2d(&tw, w, 2 * (int64_t)inner, K, K, dtype, ffg::BNH) ||
        ffg::make_map_2d(&ty, y, M, inner, inner, dtype, ffg::BMR))
        return 1;
    DADD_DISPATCH_16(dtype, T, return ffg::launch<T>(tx, tw, ty, bias, M, K, inner, (cudaStream_t)stream));
    return 1;
}
