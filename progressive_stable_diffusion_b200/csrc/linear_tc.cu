// Linear layers of the transformer blocks as one persistent tcgen05 GEMM with the neighbouring elementwise work fused in:
//   Y[m][n] = sum_k X[m][k] W[n][k] (+ bias[n]) (+ R[m][n])          X: [M][K], W: [N][K] (nn.Linear layout), Y, R: [M][N]
// Serves, inside every Transformer2DModel / BasicTransformerBlock of the UNet reached through src/models/unet/unet.py:140-146
// (diffusers graph): proj_in, the fused QKV projection, to_q of the cross-attention,
// the attention output projections, the feed-forward output projection fused with its residual add (`ff(x) + x`) and
// proj_out fused with the block residual.  The shapes are short-K (K = 320 .. 1280 for most of them), where a GEMM is bound
// by its epilogue: the accumulator is double-buffered in TMEM so that tile i + 1 is multiplied while tile i drains.
//
//   warp 8     TMA producer: ring of K-blocks {X 128 x 64, W BN x 64}, 128-byte swizzle, zero-filled edges;
//   warp 9     one elected thread issues tcgen05.mma (SS, M128 x N=BN x K16) into accumulator (i & 1);
//   warps 0-7  epilogue in two phases.  (1) row owners (TMEM lane == row; warps 0-3 / 4-7 take the two column halves):
//              tcgen05.ld -> + bias -> 16-bit -> padded staging tile in shared memory;  (2) all 256 threads walk the staging
//              tile in row-major 16-byte chunks: (+ residual chunk, coalesced) -> coalesced 16-byte global stores.  The
//              result is rounded to 16 bits before the residual is added, exactly like the unfused GEMM + add it replaces.
// BN = 256 when N % 256 == 0, else 160 (N = 320, 640, 960, 1920 are multiples of 160).
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
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : DADD_R8(r, 0), DADD_R8(r, 8)
        : "r"(taddr));
}

template <typename T, int BN, int STAGES>
__global__ void __launch_bounds__(NTHREADS, 1)
linear_kernel(const __grid_constant__ CUtensorMap tx, const __grid_constant__ CUtensorMap tw, const float* __restrict__ bias,
              const T* __restrict__ res, T* __restrict__ y, int M, int N, int K) {
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t FMT = std::is_same_v<T, __nv_bfloat16> ? 1u : 0u;
    constexpr uint32_t IDESC = instr_desc(FMT, BN, 0);
    constexpr uint32_t B_BYTES = BN * BK * 2, STAGE_BYTES = A_BYTES + B_BYTES;
    constexpr int PITCH = BN * 2 + 16;                            // staging row pitch (bytes): conflict-free 16-byte row-owner stores
    constexpr int HALF = BN / 2;                                  // columns per epilogue warp group
    constexpr int CHUNKS = BN / 8;                                // 16-byte chunks per staging row
    static_assert(BN % 32 == 0 && BN <= 256, "tile width");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sStage = smem;                                 // [STAGES]{X, W}
    unsigned char* sOut = sStage + STAGES * STAGE_BYTES;          // [128][PITCH]
    float* sBias = reinterpret_cast<float*>(sOut + BMR * PITCH);  // [BN]
    Bars<STAGES>* bars = reinterpret_cast<Bars<STAGES>*>(sBias + BN);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int mt = (M + BMR - 1) / BMR, nt = N / BN;
    const int tiles = mt * nt;
    const int my_tiles = (tiles - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const int kblocks = (K + BK - 1) / BK;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
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
    fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp == 8) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer (n-tile fastest: the X rows of a row
            // block are read by neighbouring CTAs at the same time and hit L2; W is L2-resident throughout)
            uint32_t it = 0;
            for (int i = 0; i < my_tiles; ++i) {
                const int tile = (int)blockIdx.x + i * (int)gridDim.x;
                const int m0 = (tile / nt) * BMR, n0 = (tile % nt) * BN;
                for (int kb = 0; kb < kblocks; ++kb, ++it) {
                    const uint32_t st = it % STAGES;
                    mbar_wait(&bars->empty[st], ((it / STAGES) & 1) ^ 1);
                    mbar_expect_tx(&bars->full[st], STAGE_BYTES);
                    const uint32_t base = smem_u32(sStage + st * STAGE_BYTES);
                    tma_load_2d(base, &tx, &bars->full[st], kb * BK, m0);
                    tma_load_2d(base + A_BYTES, &tw, &bars->full[st], kb * BK, n0);
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
                    mma_commit(&bars->empty[st]);
                    if (kb + 1 == kblocks) mma_commit(&bars->acc_full[a]);
                }
                __syncwarp();
            }
        }
    } else {
        // ---------------------------------------------------------------------- epilogue (8 warps)
        const int half = warp >> 2;
        const int row = (warp & 3) * 32 + lane;                       // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        for (int i = 0; i < my_tiles; ++i) {
            const int tile = (int)blockIdx.x + i * (int)gridDim.x;
            const int m0 = (tile / nt) * BMR, n0 = (tile % nt) * BN;
            const int a = i & 1;
            if (tid < BN) sBias[tid] = bias ? bias[n0 + tid] : 0.0f;
            mbar_wait(&bars->acc_full[a], (i >> 1) & 1);
            fence_after();
            named_sync(1, 256);                                       // bias staged; phase 2 of the previous tile has left sOut
            // ---- phase 1: accumulator row -> + bias -> 16-bit -> staging
            const uint32_t tacc = tmem + a * 256 + half * HALF + lane_base;
            unsigned char* srow = sOut + row * PITCH + half * HALF * 2;
            const float* bs = sBias + half * HALF;
#pragma unroll
            for (int c = 0; c < HALF; c += 16) {
                uint32_t v[16];
                tmem_ld16(tacc + c, v);
                tmem_wait_ld();
                if (c + 16 == HALF) {                                 // accumulator fully read: hand it back to the MMA warp
                    fence_before();
                    mbar_arrive(&bars->acc_empty[a]);
                }
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float4 b0 = *reinterpret_cast<const float4*>(bs + c + j * 8), b1 = *reinterpret_cast<const float4*>(bs + c + j * 8 + 4);
                    uint4 out;
                    out.x = pack2<T>(__uint_as_float(v[j * 8 + 0]) + b0.x, __uint_as_float(v[j * 8 + 1]) + b0.y);
                    out.y = pack2<T>(__uint_as_float(v[j * 8 + 2]) + b0.z, __uint_as_float(v[j * 8 + 3]) + b0.w);
                    out.z = pack2<T>(__uint_as_float(v[j * 8 + 4]) + b1.x, __uint_as_float(v[j * 8 + 5]) + b1.y);
                    out.w = pack2<T>(__uint_as_float(v[j * 8 + 6]) + b1.z, __uint_as_float(v[j * 8 + 7]) + b1.w);
                    *reinterpret_cast<uint4*>(srow + (c + j * 8) * 2) = out;
                }
            }
            named_sync(1, 256);
            // ---- phase 2: row-major 16-byte chunks: (+ residual) -> coalesced global stores
            const int rows = min(BMR, M - m0);
            for (int idx = tid; idx < rows * CHUNKS; idx += 256) {
                const int r = idx / CHUNKS, j = idx - r * CHUNKS;
                Vec8<T> t;
                t.raw = *reinterpret_cast<const uint4*>(sOut + r * PITCH + j * 16);
                const size_t g = (size_t)(m0 + r) * N + n0 + j * 8;
                if (res) {
                    Vec8<T> rr;
                    rr.load(res + g);
                    float f[8], q[8];
                    t.unpack(f);
                    rr.unpack(q);
#pragma unroll
                    for (int e = 0; e < 8; ++e) f[e] += q[e];
                    t.pack(f);
                }
                t.store(y + g);
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 9) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
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

template <typename T, int BN, int STAGES>
static int launch(const void* x, const void* w, const float* bias, const void* res, void* y, int64_t M, int N, int K, int dtype,
                  cudaStream_t s) {
    CUtensorMap tx, tw;
    if (make_map_2d(&tx, x, M, K, dtype, BMR) || make_map_2d(&tw, w, N, K, dtype, BN)) return 1;
    const size_t smem = (size_t)STAGES * (A_BYTES + BN * BK * 2) + (size_t)BMR * (BN * 2 + 16) + BN * sizeof(float) + sizeof(Bars<STAGES>) + 1024;
    const int64_t tiles = ((M + BMR - 1) / BMR) * (N / BN);
    const int grid = (int)(tiles < num_sms() ? tiles : num_sms());
    auto kern = linear_kernel<T, BN, STAGES>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "linear smem")) return 2;
    kern<<<grid, NTHREADS, smem, s>>>(tx, tw, bias, (const T*)res, (T*)y, (int)M, N, K);
    return launched("dadd_linear_fwd");
}

}  // namespace lin
}  // namespace daddk

using namespace daddk;

extern "C" int dadd_linear_supported(int64_t M, int N, int K) { return (M > 0 && K > 0 && K % 8 == 0 && N > 0 && (N % 256 == 0 || N % 160 == 0)) ? 1 : 0; }

extern "C" int dadd_linear_fwd(const void* x, const void* w, const float* bias, const void* residual, void* y, int64_t M, int N, int K,
                               int dtype, void* stream) {
    DADD_REQUIRE(x && w && y, "dadd_linear_fwd");
    DADD_REQUIRE(dtype16_ok(dtype), "dadd_linear_fwd");
    DADD_REQUIRE(M >= 0 && M < (1ll << 31) - 128, "dadd_linear_fwd");
    if (M == 0) return 0;
    if (!dadd_linear_supported(M, N, K))
        return fail("%s: needs K %% 8 == 0 and N a multiple of 160 or 256 (N = %lld, K = %lld)", "dadd_linear_fwd", (long long)N, (long long)K);
    DADD_REQUIRE(((uintptr_t)x | (uintptr_t)w | (uintptr_t)y | (uintptr_t)residual) % 16 == 0, "dadd_linear_fwd");
    cudaStream_t s = (cudaStream_t)stream;
    if (N % 256 == 0) DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 256, 3>(x, w, bias, residual, y, M, N, K, dtype, s)));
    DADD_DISPATCH_16(dtype, T, return (lin::launch<T, 160, 4>(x, w, bias, residual, y, M, N, K, dtype, s)));
    return 1;
}
