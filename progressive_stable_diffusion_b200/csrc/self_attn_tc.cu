// K1: self-attention core on the 5th-generation tensor cores (tcgen05.mma, accumulators in TMEM, operands by TMA).
// Reference op: F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 (installed at
// src/models/attention_processor_routing_gates.py:284-286); o = softmax(q k^T / sqrt d) v per (sample, head).
//
// One CTA = one (sample, head, 128-query tile); keys/values stream through in 128-key tiles.
//   warp 4 (one lane)  TMA producer: Q once, then K_j / V_j tiles into a small smem ring (mbarrier expect_tx)
//   warp 5 (one lane)  MMA issuer:  S = Q K_j^T   (tcgen05.mma SS, M=128, N=128, K=16 per instruction)
//                                   O += P_j V_j  (tcgen05.mma TS: P read from TMEM, V MN-major from smem, N=64 per panel)
//   warps 0-3          softmax: one query row per thread (TMEM lane == row), tcgen05.ld S -> online max/sum in fp32 ->
//                      exp2 -> 16-bit P written back over S with tcgen05.st, O rescaled in TMEM when the row max moved;
//                      epilogue O / l -> global.
// Layouts: Q/K/V are read in place from the fused (B, N, 3C) projection output through 4-D tensor maps
// (d, N, H, B); a head's d = 40 / 80 / 160 columns are fetched as 64-element 128-byte-swizzled panels whose out-of-range
// columns TMA fills with zeros, so no repacking pass and no padded copy of the activations exists in HBM.
// TMEM: S/P columns [0,128), O columns [128, 128 + 64*NP).  Two CTAs share an SM (d <= 80), so one CTA's softmax
// overlaps the other's MMAs.
#include "tc_util.cuh"

namespace daddk {
namespace tc {

struct Barriers {
    uint64_t q_full, s_full, p_full, o_full;
    uint64_t k_full[2], v_full[2], kv_empty[2];
    uint32_t tmem_base;
};

template <typename T, int NP, int STAGES>
__global__ void __launch_bounds__(192, NP <= 2 ? 2 : 1)
self_attn_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                    const __grid_constant__ CUtensorMap tv, T* __restrict__ o, int64_t o_stride, int N, int d,
                    float scale_log2e) {
    constexpr uint32_t TMEM_COLS = (128 + 64 * NP) <= 256 ? 256 : 512;
    constexpr uint32_t FMT = std::is_same_v<T, __nv_bfloat16> ? 1u : 0u;
    constexpr uint32_t IDESC_QK = instr_desc(FMT, 128, 0);
    constexpr uint32_t IDESC_PV = instr_desc(FMT, 64, 1);

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // the dynamic smem base is only guaranteed 16-byte aligned: round up (the launcher reserves the slack)
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sQ = smem;
    unsigned char* sK = sQ + NP * PANEL_BYTES;
    unsigned char* sV = sK + STAGES * NP * PANEL_BYTES;
    Barriers* bars = reinterpret_cast<Barriers*>(sV + STAGES * NP * PANEL_BYTES);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int q0 = blockIdx.x * BM, h = blockIdx.y, b = blockIdx.z;
    const int nkv = (N + BN - 1) / BN;
    const int ksteps = (d + 15) >> 4;

    if (tid == 0) {
        mbar_init(&bars->q_full, 1);
        mbar_init(&bars->s_full, 1);
        mbar_init(&bars->p_full, 128);
        mbar_init(&bars->o_full, 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->k_full[s], 1);
            mbar_init(&bars->v_full[s], 1);
            mbar_init(&bars->kv_empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 5) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = bars->tmem_base;
    const uint32_t tS = tmem, tO = tmem + 128;

    if (warp == 4) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer
            mbar_expect_tx(&bars->q_full, NP * PANEL_BYTES);
            for (int p = 0; p < NP; ++p) tma_load_4d(smem_u32(sQ + p * PANEL_BYTES), &tq, &bars->q_full, p * 64, q0, h, b);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % STAGES, use = j / STAGES;
                mbar_wait(&bars->kv_empty[st], (use & 1) ^ 1);
                mbar_expect_tx(&bars->k_full[st], NP * PANEL_BYTES);
                for (int p = 0; p < NP; ++p)
                    tma_load_4d(smem_u32(sK + (st * NP + p) * PANEL_BYTES), &tk, &bars->k_full[st], p * 64, j * BN, h, b);
                mbar_expect_tx(&bars->v_full[st], NP * PANEL_BYTES);
                for (int p = 0; p < NP; ++p)
                    tma_load_4d(smem_u32(sV + (st * NP + p) * PANEL_BYTES), &tv, &bars->v_full[st], p * 64, j * BN, h, b);
            }
        }
    } else if (warp == 5) {
        if (lane == 0) {
            // ------------------------------------------------------------------ MMA issuer
            mbar_wait(&bars->q_full, 0);
            for (int j = 0; j < nkv; ++j) {
                const int st = j % STAGES, use = j / STAGES;
                mbar_wait(&bars->k_full[st], use & 1);
                fence_after();
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t off = (ks >> 2) * PANEL_BYTES + (ks & 3) * 32;
                    const uint64_t da = smem_desc(smem_u32(sQ) + off, 16, 1024);
                    const uint64_t db = smem_desc(smem_u32(sK + st * NP * PANEL_BYTES) + off, 16, 1024);
                    mma_ss(tS, da, db, IDESC_QK, ks > 0);
                }
                mma_commit(&bars->s_full);
                mbar_wait(&bars->p_full, j & 1);
                mbar_wait(&bars->v_full[st], use & 1);
                fence_after();
                for (int p = 0; p < NP; ++p) {
                    for (int kk = 0; kk < BN / 16; ++kk) {
                        const uint64_t dbv = smem_desc(smem_u32(sV + (st * NP + p) * PANEL_BYTES) + kk * 2048, PANEL_BYTES, 1024);
                        mma_ts(tO + p * 64, tS + kk * 8, dbv, IDESC_PV, (j > 0 || kk > 0) ? 1u : 0u);
                    }
                }
                mma_commit(&bars->kv_empty[st]);
            }
            mma_commit(&bars->o_full);
        }
    } else {
        // ---------------------------------------------------------------------- softmax / correction / epilogue
        const uint32_t lane_base = (uint32_t)(warp * 32) << 16;
        float m = -INFINITY, l = 0.0f;
        for (int j = 0; j < nkv; ++j) {
            mbar_wait(&bars->s_full, j & 1);
            fence_after();
            const bool ragged = (j + 1) * BN > N;
            uint32_t v[32];
            float mx = m;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(tS + lane_base + c * 32, v);
                tmem_wait_ld();
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float s = __uint_as_float(v[i]);
                    if (ragged && j * BN + c * 32 + i >= N) s = -INFINITY;
                    mx = fmaxf(mx, s);
                }
            }
            const float alpha = ex2((m - mx) * scale_log2e);
            const float mb = mx * scale_log2e;
            float rs = 0.0f;
#pragma unroll 1
            for (int c = 0; c < 4; ++c) {
                tmem_ld32(tS + lane_base + c * 32, v);
                tmem_wait_ld();
                uint32_t pk[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float s0 = __uint_as_float(v[2 * i]), s1 = __uint_as_float(v[2 * i + 1]);
                    float p0 = ex2(fmaf(s0, scale_log2e, -mb)), p1 = ex2(fmaf(s1, scale_log2e, -mb));
                    if (ragged) {
                        if (j * BN + c * 32 + 2 * i >= N) p0 = 0.0f;
                        if (j * BN + c * 32 + 2 * i + 1 >= N) p1 = 0.0f;
                    }
                    rs += p0 + p1;
                    pk[i] = pack2<T>(p0, p1);
                }
                tmem_st16(tS + lane_base + c * 16, pk);      // P (16-bit) overwrites S columns that were already consumed
            }
            l = l * alpha + rs;
            m = mx;
            if (j > 0 && __any_sync(0xffffffffu, alpha != 1.0f)) {   // O(j-1) is complete: s_full(j) was committed after PV(j-1)
#pragma unroll 1
                for (int c = 0; c < 2 * NP; ++c) {
                    tmem_ld32(tO + lane_base + c * 32, v);
                    tmem_wait_ld();
#pragma unroll
                    for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                    tmem_st32(tO + lane_base + c * 32, v);
                }
            }
            tmem_wait_st();
            fence_before();
            mbar_arrive(&bars->p_full);
        }
        mbar_wait(&bars->o_full, 0);
        fence_after();
        const float inv = 1.0f / l;
        const int row = q0 + tid;
        T* orow = o + ((int64_t)b * N + row) * o_stride + (int64_t)h * d;
        const int chunks = d >> 3;
#pragma unroll 1
        for (int c = 0; c < chunks; ++c) {
            uint32_t r[8];
            tmem_ld8(tO + lane_base + c * 8, r);
            tmem_wait_ld();
            if (row < N) {
                uint4 out;
                out.x = pack2<T>(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
                out.y = pack2<T>(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
                out.z = pack2<T>(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
                out.w = pack2<T>(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
                *reinterpret_cast<uint4*>(orow + c * 8) = out;
            }
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 5) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <typename T, int NP, int STAGES>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, void* o, int64_t os, int B, int H, int N,
                  int d, float scale, cudaStream_t s) {
    const size_t smem = (size_t)(NP + 2 * STAGES * NP) * PANEL_BYTES + sizeof(Barriers) + 1024;
    auto kern = self_attn_tc_kernel<T, NP, STAGES>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "self_attn_tc smem")) return 2;
    dim3 grid((N + BM - 1) / BM, H, B);
    kern<<<grid, 192, smem, s>>>(tq, tk, tv, (T*)o, os, N, d, scale * 1.4426950408889634f);
    return launched("dadd_self_attn_fwd(tcgen05)");
}

}  // namespace tc

bool self_attn_tc_supported(int N, int d) { return N >= 128 && d % 8 == 0 && d >= 8 && d <= 192; }

int self_attn_tc(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os, int B,
                 int H, int N, int d, float scale, int dtype, cudaStream_t s) {
    CUtensorMap tq, tk, tv;
    if (tc::make_map(&tq, q, qs, B, H, N, d, dtype) || tc::make_map(&tk, k, ks, B, H, N, d, dtype) ||
        tc::make_map(&tv, v, vs, B, H, N, d, dtype))
        return 1;
    const int np = (d + 63) / 64;
#define DADD_TC(NPV, STG) DADD_DISPATCH_16(dtype, T, return (tc::launch<T, NPV, STG>(tq, tk, tv, o, os, B, H, N, d, scale, s)))
    if (np == 1) DADD_TC(1, 2);
    if (np == 2) DADD_TC(2, 1);
    DADD_TC(3, 1);
#undef DADD_TC
    return 1;
}

}  // namespace daddk
