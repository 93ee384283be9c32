// K2 on the 5th-generation tensor cores: fused multi-pathway cross-attention core for N >= 128 query rows.
// Reference: SplitInjectionAttentionProcessor.__call__ (src/models/attention_processor_routing_gates.py:148-178):
//   z = g_d softmax(Q Kd^T/sqrt d) Vd + g_a softmax(Q Ka^T/sqrt d) Va (+ lambda softmax(Q Kx^T/sqrt d) Vx)
// and the single 32-token softmax of OrdinalIPAttnProcessor2_0.__call__ (src/models/attention_processor_base.py:96-118).
//
// The op moves 4 N C bytes (Q in, O out) for 4 N L C flop with L = 48 condition tokens: it is bound by HBM bytes, and the
// warp-level mma.sync kernel (cross_attn.cu) spends ~35 instructions per query row and head on fragment shuffles.  Here a
// query row is one thread and one TMEM lane:
//   * persistent CTA per SM; work item = (sample, head, 128 query rows); a TMA producer warp keeps a ring of STAGES items
//     (Q tile + the K_cat / V_cat of that (sample, head), read in place through 4-D tensor maps, 128-byte swizzle, columns
//     beyond d zero-filled) in flight, which is what covers the HBM latency;
//   * one elected thread issues S = Q K_cat^T (tcgen05.mma SS, M128 x N=L x K16) into TMEM and, once the row owners have
//     written P, O = P V_cat (TS form: P from TMEM, V MN-major from shared memory);
//   * two softmax groups (warps 0-3 / 4-7) alternate over the items: tcgen05.ld of the L scores of the row, an independent
//     softmax per SEG-token segment entirely in registers (no shuffles), the segment's routing gate folded into its
//     normaliser (invariant I11: sum_s g_s P_s V_s = [g_s P_s]_s V_cat), 16-bit P back to TMEM, then the epilogue:
//     O -> 16-bit -> (d <= 64) swizzled staging panel + one TMA store, (d > 64) 16-byte global stores from the row owner.
// TMEM columns: S0 S1 [0,128) (P_q, 16-bit, overwrites S_q once the row owners hold the scores in registers), O0 O1 [128, 128 +
// 128 NP): 256 columns for d <= 64, so TWO CTAs share an SM there (the per-item chain QK -> softmax -> PV -> epilogue is latency-
// bound; four row groups per SM in flight hide it), 384 -> one CTA for d <= 128.
#include <cstdlib>

#include "tc_util.cuh"

namespace daddk {
namespace xtc {

using namespace daddk::tc;

constexpr int NTHREADS = 320;      // 8 softmax warps + TMA producer + MMA issuer

template <int STAGES>
struct Bars {
    uint64_t full[STAGES], empty[STAGES];
    uint64_t s_full[2], p_full[2], o_full[2];
    uint32_t tmem_base;
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : DADD_R8(r, 0), DADD_R8(r, 8)
        : "r"(taddr));
}

// SEG tokens per segment, NSEG segments, L = SEG * NSEG keys; NP = 64-column panels per head (d <= 64 NP)
template <typename T, int NP, int SEG, int NSEG, int STAGES>
__global__ void __launch_bounds__(NTHREADS, NP == 1 ? 2 : 1)
cross_attn_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                     const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap to, T* __restrict__ o,
                     int64_t o_stride, int B, int H, int N, int d,
                     const float* __restrict__ gates, float scale_log2e, int group_walk) {
    constexpr int L = SEG * NSEG;
    constexpr uint32_t TMEM_COLS = NP == 1 ? 256 : 512;
    constexpr uint32_t FMT = std::is_same_v<T, __nv_bfloat16> ? 1u : 0u;
    constexpr uint32_t IDESC_QK = instr_desc(FMT, L, 0);
    constexpr uint32_t IDESC_PV = instr_desc(FMT, 64, 1);
    constexpr uint32_t Q_PANEL = 128 * 128, KV_PANEL = L * 128;
    constexpr uint32_t STAGE_BYTES = NP * (Q_PANEL + 2 * KV_PANEL);
    constexpr uint32_t COL_S = 0, COL_O = 128;
    static_assert(L % 16 == 0 && L <= 64, "key count");
    static_assert(COL_O + 128 * NP <= TMEM_COLS, "TMEM budget");
    static_assert(KV_PANEL % 1024 == 0, "swizzle atoms");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sStage = smem;                                // [STAGES]{Q[NP], K[NP], V[NP]}
    // d <= 64: a head's output row is 80 .. 128 bytes that straddle 128-byte lines, so it leaves through a swizzled staging
    // panel and one TMA store (1.75 L2 requests per row instead of one 16-byte request per chunk); wider heads store directly
    constexpr bool TMA_STORE = NP == 1;
    unsigned char* sO = sStage + STAGES * STAGE_BYTES;           // [2 groups] staging panel (TMA_STORE only)
    Bars<STAGES>* bars = reinterpret_cast<Bars<STAGES>*>(sO + (TMA_STORE ? 2 * Q_PANEL : 0));

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nq = (N + BM - 1) / BM;
    const int items = nq * H * B;
    // Item order.  group_walk (long sequences, many heads per CTA): a CTA walks whole (sample, head) groups - group = CTA + j *
    // CTAs, its nq tiles one after the other - so that the tiles reuse the K_cat / V_cat already sitting in a stage while the
    // neighbouring CTAs work on the other heads of the sample at the same time and share the 128-byte lines of Q / O in L2.
    // Otherwise item = CTA + i * CTAs (best balance when there are few groups or few tiles per group).
    const int groups = H * B;
    const int my_items = group_walk ? ((groups - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x) * nq
                                    : (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    auto item_of = [&](int i) {
        return group_walk ? ((int)blockIdx.x + (i / nq) * (int)gridDim.x) * nq + i % nq : (int)blockIdx.x + i * (int)gridDim.x;
    };
    const int ksteps = (d + 15) >> 4;

    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->full[s], 1);
            mbar_init(&bars->empty[s], 1);
        }
        for (int q = 0; q < 2; ++q) {
            mbar_init(&bars->s_full[q], 1);
            mbar_init(&bars->p_full[q], 128);
            mbar_init(&bars->o_full[q], 1);
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
            // ------------------------------------------------------------------ TMA producer: ring of whole items
            int held[STAGES];                                    // (sample, head) whose K_cat / V_cat a stage holds
            for (int s = 0; s < STAGES; ++s) held[s] = -1;
            for (int i = 0; i < my_items; ++i) {
                const int item = item_of(i);
                const int t = item % nq, bh = item / nq, h = bh % H, b = bh / H;
                const int st = i % STAGES;
                mbar_wait(&bars->empty[st], ((i / STAGES) & 1) ^ 1);
                const bool kv = held[st] != bh;
                held[st] = bh;
                mbar_expect_tx(&bars->full[st], kv ? STAGE_BYTES : NP * Q_PANEL);
                unsigned char* base = sStage + st * STAGE_BYTES;
                for (int p = 0; p < NP; ++p) {
                    tma_load_4d(smem_u32(base + p * Q_PANEL), &tq, &bars->full[st], p * 64, t * BM, h, b);
                    if (kv) {
                        tma_load_4d(smem_u32(base + NP * Q_PANEL + p * KV_PANEL), &tk, &bars->full[st], p * 64, 0, h, b);
                        tma_load_4d(smem_u32(base + NP * (Q_PANEL + KV_PANEL) + p * KV_PANEL), &tv, &bars->full[st], p * 64, 0, h, b);
                    }
                }
            }
        }
    } else if (warp == 9) {
        // ---------------------------------------------------------------------- MMA issuer: QK0 QK1 | PV0 QK2 | PV1 QK3 | ...
        // (S_q(i+2) overwrites P_q(i): it is issued after PV_q(i), and the tensor core executes in issue order)
        const bool leader = elect_one();
        auto issue_qk = [&](int i) {
            const int st = i % STAGES, q = i & 1;
            mbar_wait(&bars->full[st], (i / STAGES) & 1);
            fence_after();
            if (leader) {
                const uint32_t base = smem_u32(sStage + st * STAGE_BYTES);
                const uint64_t da = smem_desc(base, 16, 1024), db = smem_desc(base + NP * Q_PANEL, 16, 1024);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t offq = (ks >> 2) * Q_PANEL + (ks & 3) * 32, offk = (ks >> 2) * KV_PANEL + (ks & 3) * 32;
                    mma_ss(tmem + COL_S + q * 64, desc_add(da, offq), desc_add(db, offk), IDESC_QK, ks > 0);
                }
                mma_commit(&bars->s_full[q]);
            }
            __syncwarp();
        };
        auto issue_pv = [&](int i) {
            const int st = i % STAGES, q = i & 1, k = i >> 1;
            mbar_wait(&bars->p_full[q], k & 1);
            fence_after();
            if (leader) {
                const uint32_t vbase = smem_u32(sStage + st * STAGE_BYTES + NP * (Q_PANEL + KV_PANEL));
                const uint32_t tO = tmem + COL_O + q * (64 * NP), tP = tmem + COL_S + q * 64;
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const uint64_t dv = smem_desc(vbase + p * KV_PANEL, KV_PANEL, 1024);
#pragma unroll
                    for (int kk = 0; kk < L / 16; ++kk) mma_ts(tO + p * 64, tP + kk * 8, desc_add(dv, kk * 2048), IDESC_PV, kk > 0);
                }
                mma_commit(&bars->o_full[q]);
                mma_commit(&bars->empty[st]);                            // Q, K, V of this stage are no longer needed
            }
            __syncwarp();
        };
        if (my_items > 0) issue_qk(0);
        if (my_items > 1) issue_qk(1);
        for (int i = 0; i < my_items; ++i) {
            issue_pv(i);
            if (i + 2 < my_items) issue_qk(i + 2);
        }
    } else {
        // ---------------------------------------------------------------------- softmax groups / epilogue
        const int q = warp >> 2;
        const int wrow = (warp & 3) * 32 + lane;                         // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + COL_S + q * 64 + lane_base;
        const uint32_t tP = tS;
        const uint32_t tO = tmem + COL_O + q * (64 * NP) + lane_base;
        const bool store_leader = (warp & 3) == 0 && lane == 0;
        float gate[NSEG];
#pragma unroll
        for (int s = 0; s < NSEG; ++s) gate[s] = gates[s];
        int k = 0;
        for (int i = q; i < my_items; i += 2, ++k) {
            const int item = item_of(i);
            const int t = item % nq, h = (item / nq) % H, b = item / (nq * H);
            mbar_wait(&bars->s_full[q], k & 1);
            fence_after();
            uint32_t sr[L];
#pragma unroll
            for (int c = 0; c < L; c += 16) tmem_ld16(tS + c, sr + c);
            tmem_wait_ld();
            uint32_t pk[L / 2];
#pragma unroll
            for (int s = 0; s < NSEG; ++s) {
                float mx0 = __uint_as_float(sr[s * SEG]), mx1 = __uint_as_float(sr[s * SEG + 1]);
#pragma unroll
                for (int j = 2; j < SEG; j += 4) {
                    mx0 = max3(mx0, __uint_as_float(sr[s * SEG + j]), __uint_as_float(sr[s * SEG + j + 1]));
                    if (j + 2 < SEG) mx1 = max3(mx1, __uint_as_float(sr[s * SEG + j + 2]), __uint_as_float(sr[s * SEG + j + 3]));
                }
                const float nmb = -fmaxf(mx0, mx1) * scale_log2e;
                float p[SEG], l0 = 0.0f, l1 = 0.0f;
#pragma unroll
                for (int j = 0; j < SEG; j += 2) {
                    p[j] = ex2(fmaf(__uint_as_float(sr[s * SEG + j]), scale_log2e, nmb));
                    p[j + 1] = ex2(fmaf(__uint_as_float(sr[s * SEG + j + 1]), scale_log2e, nmb));
                    l0 += p[j];
                    l1 += p[j + 1];
                }
                const float r = __fdividef(gate[s], l0 + l1);
#pragma unroll
                for (int j = 0; j < SEG; j += 2) pk[(s * SEG + j) >> 1] = pack2<T>(p[j] * r, p[j + 1] * r);
            }
#pragma unroll
            for (int c = 0; c < L / 2; c += 8) {
                uint32_t v8[8];
#pragma unroll
                for (int e = 0; e < 8; ++e) v8[e] = pk[c + e];
                tmem_st8(tP + c, v8);
            }
            tmem_wait_st();
            fence_before();
            mbar_arrive(&bars->p_full[q]);
            // epilogue: O -> 16-bit -> global (heads merged, ready for to_out)
            mbar_wait(&bars->o_full[q], k & 1);
            fence_after();
            const int chunks = d >> 3;
            if constexpr (TMA_STORE) {
                if (store_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has left the panel
                named_sync(1 + q, 128);
                unsigned char* stage = sO + q * Q_PANEL + wrow * 128;
#pragma unroll 1
                for (int c0 = 0; c0 < chunks; c0 += 4) {
                    uint32_t r[32];
                    tmem_ld32(tO + c0 * 8, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        const int c = c0 + cc;
                        if (c < chunks) {
                            uint4 out;
                            out.x = pack2<T>(__uint_as_float(r[cc * 8 + 0]), __uint_as_float(r[cc * 8 + 1]));
                            out.y = pack2<T>(__uint_as_float(r[cc * 8 + 2]), __uint_as_float(r[cc * 8 + 3]));
                            out.z = pack2<T>(__uint_as_float(r[cc * 8 + 4]), __uint_as_float(r[cc * 8 + 5]));
                            out.w = pack2<T>(__uint_as_float(r[cc * 8 + 6]), __uint_as_float(r[cc * 8 + 7]));
                            *reinterpret_cast<uint4*>(stage + (((c & 7) ^ (wrow & 7)) << 4)) = out;   // 128-byte swizzle
                        }
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                fence_before();
                named_sync(1 + q, 128);
                if (store_leader) {
                    tma_store_4d(&to, smem_u32(sO + q * Q_PANEL), 0, t * BM, h, b);      // rows >= N, columns >= d clipped by TMA
                    asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                }
            } else {
                const int grow = t * BM + wrow;
                T* orow = o + ((int64_t)b * N + grow) * o_stride + (int64_t)h * d;
#pragma unroll 1
                for (int c0 = 0; c0 < chunks; c0 += 4) {
                    uint32_t r[32];
                    tmem_ld32(tO + c0 * 8, r);
                    tmem_wait_ld();
#pragma unroll
                    for (int cc = 0; cc < 4; ++cc) {
                        if (c0 + cc < chunks && grow < N) {
                            uint4 out;
                            out.x = pack2<T>(__uint_as_float(r[cc * 8 + 0]), __uint_as_float(r[cc * 8 + 1]));
                            out.y = pack2<T>(__uint_as_float(r[cc * 8 + 2]), __uint_as_float(r[cc * 8 + 3]));
                            out.z = pack2<T>(__uint_as_float(r[cc * 8 + 4]), __uint_as_float(r[cc * 8 + 5]));
                            out.w = pack2<T>(__uint_as_float(r[cc * 8 + 6]), __uint_as_float(r[cc * 8 + 7]));
                            *reinterpret_cast<uint4*>(orow + (c0 + cc) * 8) = out;
                        }
                    }
                }
            }
            fence_before();      // the O reads are ordered before this group's next P arrive (-> the next PV overwrites O)
        }
        if (TMA_STORE && store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    if (warp == 9) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

// 4-D view (d, rows, H, B) with explicit element strides between rows / heads / samples; boxes are 64 x box_rows x 1 x 1
static int make_map_strided(CUtensorMap* map, const void* base, int64_t row_stride, int64_t head_stride, int64_t batch_stride,
                            int B, int H, int rows, int d, int dtype, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail("%s: cuTensorMapEncodeTiled is unavailable", "dadd_cross_attn_fwd(tcgen05)");
    const cuuint64_t dims[4] = {(cuuint64_t)d, (cuuint64_t)rows, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t strides[3] = {(cuuint64_t)row_stride * 2, (cuuint64_t)head_stride * 2, (cuuint64_t)batch_stride * 2};
    const cuuint32_t box[4] = {64, (cuuint32_t)box_rows, 1, 1};
    const cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(map, dtype == DADD_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4,
                    const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return fail("%s: cuTensorMapEncodeTiled failed (CUresult %lld)", "dadd_cross_attn_fwd(tcgen05)", (long long)r);
    return 0;
}

template <typename T, int NP, int SEG, int NSEG, int STAGES>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, void* o, int64_t o_stride,
                  int B, int H, int N, int d, const float* gates, float scale, cudaStream_t s) {
    constexpr int L = SEG * NSEG;
    const size_t smem = (size_t)STAGES * NP * (128 * 128 + 2 * L * 128) + (NP == 1 ? 2 * 128 * 128 : 0) + sizeof(Bars<STAGES>) + 1024;
    const int items = ((N + BM - 1) / BM) * H * B;
    const int ctas = (NP == 1 ? 2 : 1) * num_sms();
    const int nq = (N + BM - 1) / BM, groups = H * B;
    // measured on B200 (profiles/r01_cross_attn_kv_reuse.txt): the group walk wins at N = 1024 with >= 2 groups per CTA
    const int group_walk = (nq >= 4 && groups >= 2 * ctas) ? 1 : 0;
    const int grid = group_walk ? ctas : (items < ctas ? items : ctas);
    auto kern = cross_attn_tc_kernel<T, NP, SEG, NSEG, STAGES>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cross_attn_tc smem")) return 2;
    kern<<<grid, NTHREADS, smem, s>>>(tq, tk, tv, to, (T*)o, o_stride, B, H, N, d, gates, scale * 1.4426950408889634f, group_walk);
    return launched("dadd_cross_attn_fwd(tcgen05)");
}

}  // namespace xtc

bool cross_attn_tc_supported(int N, int d, int seg_len, int n_seg) {
    return N >= 128 && d % 8 == 0 && d >= 8 && d <= 128 && ((seg_len == 16 && (n_seg == 2 || n_seg == 3)) || (seg_len == 32 && n_seg == 1));
}

int cross_attn_tc(const void* q, int64_t q_stride, const void* k_cat, const void* v_cat, void* o, int64_t o_stride, int B, int H,
                  int N, int d, int seg_len, int n_seg, const float* gates, float scale, int dtype, cudaStream_t s) {
    const int L = seg_len * n_seg;
    CUtensorMap tq, tk, tv, to;
    if (tc::make_map(&tq, q, q_stride, B, H, N, d, dtype, 128) || tc::make_map(&to, o, o_stride, B, H, N, d, dtype, 128) ||
        xtc::make_map_strided(&tk, k_cat, d, (int64_t)L * d, (int64_t)H * L * d, B, H, L, d, dtype, L) ||
        xtc::make_map_strided(&tv, v_cat, d, (int64_t)L * d, (int64_t)H * L * d, B, H, L, d, dtype, L))
        return 1;
    const int np = (d + 63) / 64;
#define DADD_XTC(NPV, SEGV, NSEGV, STG) \
    DADD_DISPATCH_16(dtype, T, return (xtc::launch<T, NPV, SEGV, NSEGV, STG>(tq, tk, tv, to, o, o_stride, B, H, N, d, gates, scale, s)))
    if (np == 1) {
        if (n_seg == 3) DADD_XTC(1, 16, 3, 2);
        if (n_seg == 2) DADD_XTC(1, 16, 2, 2);
        DADD_XTC(1, 32, 1, 2);
    }
    if (n_seg == 3) DADD_XTC(2, 16, 3, 3);
    if (n_seg == 2) DADD_XTC(2, 16, 2, 3);
    DADD_XTC(2, 32, 1, 3);
#undef DADD_XTC
    return 1;
}

}  // namespace daddk
