// K1 (second generation): self-attention core with two query tiles per CTA ping-ponged against the tensor core.
// Reference op: F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 (installed at
// src/models/attention_processor_routing_gates.py:284-286); o = softmax(q k^T / sqrt d) v per (sample, head).
//
// The first kernel (self_attn_tc.cu) ran QK^T -> softmax -> PV strictly in sequence per CTA and reached 13.5 % of the
// measured bf16 peak at (N = 1024, d = 40): the softmax warps waited for the MMAs and vice versa, the row-max / row-sum
// reductions were single dependent chains, S was read from TMEM twice and O was rescaled almost every tile
// (profiles/r01_ncu_self_attn_tc_v1.txt).  Here:
//   * a persistent CTA (one per SM) walks over work items = (sample, head, 256 query rows); the key/value steps of all its
//     items form one continuous stream, so the next item's loads and first QK^T overlap the current item's tail;
//   * warps 0-3 / 4-7 are two softmax groups, one 128-row query tile each (thread = row, TMEM lane = row);
//   * P has its own TMEM columns (it does not overwrite S), so S(j+1) = Q K(j+1)^T is issued as soon as the group has
//     pulled S(j) into registers - the whole MMA round trip hides behind the group's own exponentials, and the exp2 stream
//     on the MUFU pipe (the real bound at d = 40: 160 flop per exponential) never waits for the tensor core;
//   * S is read from TMEM once, max uses the 3-input max with independent accumulators, scale/sum use packed f32x2 ops;
//   * O is rescaled lazily: only when a row maximum grew by more than 2^8 since the offset in use was chosen;
//   * the two groups are kept half a step apart by a pair of named barriers around the exp2 phase, so that one group's TMEM
//     loads, row max, P stores and barrier traffic fall into the other's exponentials instead of both stalling together;
//     the exp2 loop is software-pipelined in batches of 8 (results of batch b-1 are retired while batch b is in flight);
//   * warp 8 = TMA producer of the K/V ring, warp 10 = TMA producer of the query tiles, warp 9 = tcgen05.mma issuer
//     (warp-uniform control flow, one elected lane issues); O leaves through a swizzled staging panel and a TMA store.
// TMEM (BN = keys per step: 128 for d <= 64, 64 for d <= 128):
//   S0 S1 [0, 2 BN)   P0 P1 [2 BN, 3 BN) (16-bit pairs)   O0 O1 [3 BN, 3 BN + 128 NP)     (512 resp. 448 columns)
// Layouts as in self_attn_tc.cu: Q/K/V are read in place from the fused (B, N, 3C) projection output through 4-D tensor
// maps, 64-element 128-byte-swizzled panels, out-of-range rows / columns zero-filled by TMA.
#include <cstdlib>

#include "tc_util.cuh"

namespace daddk {
namespace tc2 {

using namespace daddk::tc;

constexpr int NTHREADS = 352;      // 8 softmax warps + K/V producer + MMA issuer + Q producer
constexpr float RESCALE_LOG2 = 8.0f;

template <int STAGES>
struct Bars {
    uint64_t q_full[2], q_empty[2];
    uint64_t s_full[2], s_read[2], p_full[2], pv_done[2];
    uint64_t k_full[STAGES], v_full[STAGES], kv_empty[STAGES];
    uint32_t tmem_base;
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
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
// 2^a for a pair of exponents on the FMA / ALU pipes instead of the MUFU pipe (the bound of this kernel at d = 40 is the
// 16 exp2 per clock per SM of the MUFU unit; the FMA pipe is mostly idle): a is clamped to >= -125, rounded to the nearest
// integer n with the 1.5 * 2^23 trick, 2^(a - n) is a degree-3 minimax polynomial on [-0.5, 0.5] (max relative error
// 7.5e-5, far below the 2^-9 / 2^-11 rounding of the 16-bit P it feeds) and n is added into the exponent field with one
// integer shift-add.  Valid for a <= +126 (the lazy rescale keeps a <= 8).
__device__ __forceinline__ void exp2_poly2(uint64_t a, float& p0, float& p1) {
    float a0, a1;
    unpack_f2(a, a0, a1);
    a0 = fmaxf(a0, -125.0f);
    a1 = fmaxf(a1, -125.0f);
    const uint64_t ac = pack_f2(a0, a1);
    const uint64_t t = add_f2(ac, pack_f2(12582912.0f, 12582912.0f));
    const uint64_t n = add_f2(t, pack_f2(-12582912.0f, -12582912.0f));
    const uint64_t f = fma_f2(n, pack_f2(-1.0f, -1.0f), ac);
    uint64_t p = fma_f2(pack_f2(0.055171459913253784f, 0.055171459913253784f), f, pack_f2(0.2426108568906784f, 0.2426108568906784f));
    p = fma_f2(p, f, pack_f2(0.6932609677314758f, 0.6932609677314758f));
    p = fma_f2(p, f, pack_f2(0.9999281167984009f, 0.9999281167984009f));
    float t0, t1;
    unpack_f2(p, p0, p1);
    unpack_f2(t, t0, t1);
    p0 = __int_as_float(__float_as_int(p0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(p1) + (__float_as_int(t1) << 23));
}
// which of the 4 pairs of exp2 batch `bt` (8 exponentials) go to the polynomial: POLY = pairs per 2 batches (0..8)
template <int POLY>
__device__ __forceinline__ constexpr bool poly_pair(int bt, int k) {
    const int even = (POLY + 1) / 2, odd = POLY / 2;      // pairs in even / odd batches
    return k < ((bt & 1) ? odd : even);
}

// descriptor with the start address advanced by `bytes` (the address field holds addr >> 4 in bits [0,14): no carry out
// for shared-memory addresses below 256 KB)
__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }

// named barriers 1 / 2: the exp2 phases of the two softmax groups take turns on the MUFU pipe; 3 / 4: per-group epilogue
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void named_arrive(int id, int count) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(count) : "memory"); }

__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}

// TRACE (debug builds of the timeline only, DADD_ATTN_TRACE=<file>): CTA 0 records clock64() at the protocol events of its
// first 64 steps into trace[role][step][event], role 0/1 = softmax group leaders, 2 = MMA issuer.
#define DADD_TRACE_EVENT(role, step, ev)                                                           \
    do {                                                                                            \
        if constexpr (TRACE) {                                                                      \
            if (blockIdx.x == 0 && (step) < 64 && trace_on) trace[((role) * 64 + (step)) * 16 + (ev)] = clock64(); \
        }                                                                                           \
    } while (0)

template <typename T, int NP, int BN_, int STAGES, bool TRACE, int POLY>
__global__ void __launch_bounds__(NTHREADS, 1)
self_attn_tc2_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                     const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap to, int B, int H, int N, int d,
                     float scale_log2e, long long* __restrict__ trace) {
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t FMT = std::is_same_v<T, __nv_bfloat16> ? 1u : 0u;
    constexpr uint32_t IDESC_QK = instr_desc(FMT, BN_, 0);
    constexpr uint32_t IDESC_PV = instr_desc(FMT, 64, 1);
    constexpr uint32_t Q_PANEL = 128 * 128, KV_PANEL = BN_ * 128;
    constexpr uint32_t COL_S = 0, COL_P = 2 * BN_, COL_O = 3 * BN_;
    static_assert(COL_O + 128 * NP <= 512, "TMEM budget");

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sQ = smem;                                   // [2 tiles][NP panels]
    unsigned char* sO = sQ + 2 * NP * Q_PANEL;                  // [2 tiles][NP panels] output staging for the TMA store
    unsigned char* sK = sO + 2 * NP * Q_PANEL;                  // [STAGES][NP]
    unsigned char* sV = sK + STAGES * NP * KV_PANEL;            // [STAGES][NP]
    Bars<STAGES>* bars = reinterpret_cast<Bars<STAGES>*>(sV + STAGES * NP * KV_PANEL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkv = (N + BN_ - 1) / BN_;
    const int nq2 = (N + 2 * BM - 1) / (2 * BM);
    const int items = nq2 * H * B;
    const int my_items = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t steps = (uint32_t)my_items * (uint32_t)nkv;
    const int ksteps = (d + 15) >> 4;

    if (tid == 0) {
        for (int q = 0; q < 2; ++q) {
            mbar_init(&bars->q_full[q], 1);
            mbar_init(&bars->q_empty[q], 1);
            mbar_init(&bars->s_full[q], 1);
            mbar_init(&bars->s_read[q], 128);
            mbar_init(&bars->p_full[q], 128);
            mbar_init(&bars->pv_done[q], 1);
        }
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->k_full[s], 1);
            mbar_init(&bars->v_full[s], 1);
            mbar_init(&bars->kv_empty[s], 1);
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
            // ------------------------------------------------------------------ TMA producer: K/V ring over the step stream
            for (uint32_t it = 0; it < steps; ++it) {
                const int item = (int)blockIdx.x + (int)(it / (uint32_t)nkv) * (int)gridDim.x, j = (int)(it % (uint32_t)nkv);
                const int h = (item / nq2) % H, b = item / (nq2 * H);
                const uint32_t st = it % STAGES, use = it / STAGES;
                mbar_wait(&bars->kv_empty[st], (use & 1) ^ 1);
                mbar_expect_tx(&bars->k_full[st], NP * KV_PANEL);
                for (int p = 0; p < NP; ++p)
                    tma_load_4d(smem_u32(sK + (st * NP + p) * KV_PANEL), &tk, &bars->k_full[st], p * 64, j * BN_, h, b);
                mbar_expect_tx(&bars->v_full[st], NP * KV_PANEL);
                for (int p = 0; p < NP; ++p)
                    tma_load_4d(smem_u32(sV + (st * NP + p) * KV_PANEL), &tv, &bars->v_full[st], p * 64, j * BN_, h, b);
            }
        }
    } else if (warp == 10) {
        if (lane == 0) {
            // ------------------------------------------------------------------ TMA producer: the two query tiles of each item
            for (int ti = 0; ti < my_items; ++ti) {
                const int item = (int)blockIdx.x + ti * (int)gridDim.x;
                const int qb = item % nq2, h = (item / nq2) % H, b = item / (nq2 * H);
                for (int q = 0; q < 2; ++q) {
                    mbar_wait(&bars->q_empty[q], (ti & 1) ^ 1);             // the previous item's QK^T MMAs are done with this tile
                    mbar_expect_tx(&bars->q_full[q], NP * Q_PANEL);
                    for (int p = 0; p < NP; ++p)
                        tma_load_4d(smem_u32(sQ + (q * NP + p) * Q_PANEL), &tq, &bars->q_full[q], p * 64, (qb * 2 + q) * BM, h, b);
                }
            }
        }
    } else if (warp == 9) {
        // ---------------------------------------------------------------------- MMA issuer (whole warp walks the stream)
        const bool leader = elect_one();
        const bool trace_on = leader;
        const uint64_t dq0 = smem_desc(smem_u32(sQ), 16, 1024);
        const uint64_t dk0 = smem_desc(smem_u32(sK), 16, 1024);
        const uint64_t dv0 = smem_desc(smem_u32(sV), KV_PANEL, 1024);
        // S_q(step) = Q_q K(step)^T; waits for the operands, signals the group, releases the Q tile after the item's last one
        auto issue_qk = [&](int q, uint32_t step) {
            const uint32_t st = step % STAGES, j = step % (uint32_t)nkv;
            if (j == 0) mbar_wait(&bars->q_full[q], (step / (uint32_t)nkv) & 1);
            mbar_wait(&bars->k_full[st], (step / STAGES) & 1);
            fence_after();
            if (leader) {
                const uint64_t da = desc_add(dq0, q * NP * Q_PANEL), db = desc_add(dk0, st * NP * KV_PANEL);
                for (int ks = 0; ks < ksteps; ++ks) {
                    const uint32_t offq = (ks >> 2) * Q_PANEL + (ks & 3) * 32, offk = (ks >> 2) * KV_PANEL + (ks & 3) * 32;
                    mma_ss(tmem + COL_S + q * BN_, desc_add(da, offq), desc_add(db, offk), IDESC_QK, ks > 0);
                }
                mma_commit(&bars->s_full[q]);
                if (j + 1 == (uint32_t)nkv) mma_commit(&bars->q_empty[q]);
            }
            __syncwarp();
        };
        // O_q (+)= P_q(step) V(step), then release P_q / O_q (and, after group 1's, the K/V stage)
        auto issue_pv = [&](int q, uint32_t step) {
            const uint32_t st = step % STAGES, acc = step % (uint32_t)nkv > 0 ? 1u : 0u;
            mbar_wait(&bars->p_full[q], step & 1);
            mbar_wait(&bars->v_full[st], (step / STAGES) & 1);
            fence_after();
            if (leader) {
                const uint32_t tO = tmem + COL_O + q * (64 * NP), tP = tmem + COL_P + q * (BN_ / 2);
#pragma unroll
                for (int p = 0; p < NP; ++p) {
                    const uint64_t dv = desc_add(dv0, (st * NP + p) * KV_PANEL);
#pragma unroll
                    for (int kk = 0; kk < BN_ / 16; ++kk)
                        mma_ts(tO + p * 64, tP + kk * 8, desc_add(dv, kk * 2048), IDESC_PV, (acc || kk > 0) ? 1u : 0u);
                }
                mma_commit(&bars->pv_done[q]);
                if (q == 1) mma_commit(&bars->kv_empty[st]);
            }
            __syncwarp();
        };
        // Group 0 runs half a step ahead of group 1 (their exp2 phases alternate), so the stream is issued in the order the
        // events fire: S1(s+1) | P0(s) V | S0(s+2) | P1(s) V.
        if (steps > 0) {
            issue_qk(0, 0);
            issue_qk(1, 0);
            if (steps > 1) {
                mbar_wait(&bars->s_read[0], 0);
                issue_qk(0, 1);
            }
        }
        for (uint32_t s = 0; s < steps; ++s) {
            DADD_TRACE_EVENT(2, s, 0);
            if (s + 1 < steps) {
                mbar_wait(&bars->s_read[1], s & 1);
                DADD_TRACE_EVENT(2, s, 1);
                issue_qk(1, s + 1);
            }
            DADD_TRACE_EVENT(2, s, 2);
            issue_pv(0, s);
            DADD_TRACE_EVENT(2, s, 3);
            // at the end of an item group 0 is busy with its epilogue, so group 1's P arrives first: serve it first
            const bool item_end = (s + 1) % (uint32_t)nkv == 0;
            if (item_end) issue_pv(1, s);
            if (s + 2 < steps) {
                mbar_wait(&bars->s_read[0], (s + 1) & 1);
                DADD_TRACE_EVENT(2, s, 4);
                issue_qk(0, s + 2);
            }
            DADD_TRACE_EVENT(2, s, 5);
            if (!item_end) issue_pv(1, s);
            DADD_TRACE_EVENT(2, s, 6);
        }
    } else if (warp < 8) {
        // ---------------------------------------------------------------------- softmax / correction / epilogue
        const int q = warp >> 2;                                   // query tile of this warp's group
        const int wrow = (warp & 3) * 32 + lane;                   // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
        const uint32_t tS = tmem + COL_S + q * BN_ + lane_base;
        const uint32_t tP = tmem + COL_P + q * (BN_ / 2) + lane_base;
        const uint32_t tO = tmem + COL_O + q * (64 * NP) + lane_base;
        const uint64_t cc = pack_f2(scale_log2e, scale_log2e);
        const bool trace_on = (warp & 3) == 0 && lane == 0;
        const bool store_leader = (warp & 3) == 0 && lane == 0;
        if (q == 1 && steps > 0) named_arrive(1, 256);             // group 0 takes the first turn on the MUFU pipe
        uint32_t s = 0;
        for (int ti = 0; ti < my_items; ++ti) {
            const int item = (int)blockIdx.x + ti * (int)gridDim.x;
            const int qb = item % nq2, h = (item / nq2) % H, b = item / (nq2 * H);
            float m_used = -INFINITY, l = 0.0f;
            for (int j = 0; j < nkv; ++j, ++s) {
                DADD_TRACE_EVENT(q, s, 0);
                mbar_wait(&bars->s_full[q], s & 1);
                fence_after();
                DADD_TRACE_EVENT(q, s, 1);
                uint32_t sr[BN_];
                tmem_ld64(tS, sr);
                if constexpr (BN_ == 128) tmem_ld64(tS + 64, sr + 64);
                tmem_wait_ld();
                DADD_TRACE_EVENT(q, s, 2);
                fence_before();
                mbar_arrive(&bars->s_read[q]);                       // S_q may be overwritten by the next QK^T
                const bool ragged = (j + 1) * BN_ > N;
                if (ragged) {                                        // ragged last key tile
#pragma unroll
                    for (int i = 0; i < BN_; ++i)
                        if (j * BN_ + i >= N) sr[i] = 0xff800000u;   // -inf
                }
                float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]), mx2 = __uint_as_float(sr[2]),
                      mx3 = __uint_as_float(sr[3]);
#pragma unroll
                for (int i = 4; i < BN_; i += 8) {
                    mx0 = max3(mx0, __uint_as_float(sr[i]), __uint_as_float(sr[i + 1]));
                    mx1 = max3(mx1, __uint_as_float(sr[i + 2]), __uint_as_float(sr[i + 3]));
                    if (i + 4 < BN_) {
                        mx2 = max3(mx2, __uint_as_float(sr[i + 4]), __uint_as_float(sr[i + 5]));
                        mx3 = max3(mx3, __uint_as_float(sr[i + 6]), __uint_as_float(sr[i + 7]));
                    }
                }
                const float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
                DADD_TRACE_EVENT(q, s, 3);
                // PV_q(s-1) must be complete before O_q is rescaled or P_q is overwritten (it was issued half a step ago, while
                // the other group held the MUFU pipe, so this wait is normally already satisfied)
                if (j > 0) {
                    mbar_wait(&bars->pv_done[q], (s - 1) & 1);
                    fence_after();
                }
                // lazy offset update: keep exponentiating against m_used until a row maximum outgrows it by 2^8
                const bool grow = (tmax - m_used) * scale_log2e > RESCALE_LOG2;
                if (__any_sync(0xffffffffu, grow)) {
                    float alpha = 1.0f;
                    if (grow) {
                        alpha = ex2((m_used - tmax) * scale_log2e);   // 0 on the first tile (m_used = -inf)
                        m_used = tmax;
                        l *= alpha;
                    }
                    if (j > 0) {
#pragma unroll 1
                        for (int c = 0; c < 8 * NP; ++c) {            // rare path: 8 columns at a time keeps registers free
                            uint32_t v[8];
                            tmem_ld8(tO + c * 8, v);
                            tmem_wait_ld();
#pragma unroll
                            for (int i = 0; i < 8; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
                            tmem_st8(tO + c * 8, v);
                        }
                    }
                }
                const float nmb = -m_used * scale_log2e;
                const uint64_t nb = pack_f2(nmb, nmb);
                uint64_t acc[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) acc[i] = pack_f2(0.0f, 0.0f);
                // the exponents a = s * c - m * c are formed here, outside the exp2 phase: while a MUFU instruction dispatches
                // (8 clk for 32 lanes) its warp cannot issue, so inside the phase every other instruction adds to the MUFU time
                // one for one (per-phase timeline and what it implies: profiles/r01_attn_exp_phase.txt)
#pragma unroll
                for (int k = 0; k < BN_ / 2; ++k) {
                    float a0, a1;
                    unpack_f2(fma_f2(pack_f2(__uint_as_float(sr[2 * k]), __uint_as_float(sr[2 * k + 1])), cc, nb), a0, a1);
                    sr[2 * k] = __float_as_uint(a0);
                    sr[2 * k + 1] = __float_as_uint(a1);
                }
                DADD_TRACE_EVENT(q, s, 4);
                named_sync(1 + q, 256);                               // my turn on the MUFU pipe
                // (ptxas is free to start the exp2 stream before the barrier - measured best: a hard hand-over leaves one warp per
                // scheduler, whose in-order issue sustains only ~2/3 of the MUFU rate)
                DADD_TRACE_EVENT(q, s, 5);
                // Software-pipelined in batches of 8: the exp2 of batch b are issued back to back while the results of batch
                // b-1 (long in flight) are summed, converted and stored, so ONE warp per scheduler keeps the MUFU pipe busy.
                auto exp_phase = [&](auto poly_tag) {
                    constexpr int POLYX = decltype(poly_tag)::value;
                    constexpr int NB = BN_ / 8;
                    float pprev[8], pcur[8];
                    uint32_t pk[16];
                    auto scaled = [&](int base, uint64_t (&a)[4]) {
#pragma unroll
                        for (int k = 0; k < 4; ++k)
                            a[k] = pack_f2(__uint_as_float(sr[base + 2 * k]), __uint_as_float(sr[base + 2 * k + 1]));
                    };
                    // pair k of batch bt: MUFU, or the FMA-pipe polynomial for a fixed share of the pairs (a ragged tile runs
                    // the all-MUFU instance: its masked -inf scores must map to exactly 0)
                    auto expo = [&](int bt, int k, uint64_t a, float& p0, float& p1) {
                        if (poly_pair<POLYX>(bt, k)) {
                            exp2_poly2(a, p0, p1);
                        } else {
                            float a0, a1;
                            unpack_f2(a, a0, a1);
                            p0 = ex2(a0);
                            p1 = ex2(a1);
                        }
                    };
                    {
                        uint64_t a[4];
                        scaled(0, a);
#pragma unroll
                        for (int k = 0; k < 4; ++k) expo(0, k, a[k], pprev[2 * k], pprev[2 * k + 1]);
                    }
#pragma unroll
                    for (int bt = 1; bt <= NB; ++bt) {
                        uint64_t a[4];
                        if (bt < NB) scaled(bt * 8, a);
#pragma unroll
                        for (int k = 0; k < 4; ++k) {
                            if (bt < NB) expo(bt, k, a[k], pcur[2 * k], pcur[2 * k + 1]);
                            const int i = 2 * k + 1;                  // retire one pair of the previous batch per pair issued
                            acc[k] = add_f2(acc[k], pack_f2(pprev[i - 1], pprev[i]));
                            pk[(((bt - 1) * 8 + i) >> 1) & 15] = pack2<T>(pprev[i - 1], pprev[i]);
                        }
#pragma unroll
                        for (int i = 0; i < 8; ++i) pprev[i] = pcur[i];
                        if ((bt & 3) == 0) tmem_st16(tP + (bt / 4 - 1) * 16, pk);   // 32 probabilities = 16 packed columns done
                        if (bt == 4) DADD_TRACE_EVENT(q, s, 10);
                        if (bt == 8) DADD_TRACE_EVENT(q, s, 11);
                        if (bt == 12) DADD_TRACE_EVENT(q, s, 12);
                    }
                };
                if (POLY > 0 && !ragged) exp_phase(std::integral_constant<int, POLY>{});
                else exp_phase(std::integral_constant<int, 0>{});
                if (q == 0 || s + 1 < steps) named_arrive(2 - q, 256);   // the other group's turn (none left after the last step)
                DADD_TRACE_EVENT(q, s, 6);
                {
                    float r0, r1, r2, r3, r4, r5, r6, r7;
                    unpack_f2(acc[0], r0, r1);
                    unpack_f2(acc[1], r2, r3);
                    unpack_f2(acc[2], r4, r5);
                    unpack_f2(acc[3], r6, r7);
                    l += ((r0 + r1) + (r2 + r3)) + ((r4 + r5) + (r6 + r7));
                }
                tmem_wait_st();
                DADD_TRACE_EVENT(q, s, 7);
                fence_before();
                mbar_arrive(&bars->p_full[q]);
            }
            // epilogue of this item: O / l -> swizzled staging panel -> one TMA store per panel (rows / columns beyond N / d clipped)
            mbar_wait(&bars->pv_done[q], (s - 1) & 1);
            fence_after();
            DADD_TRACE_EVENT(q, s - 1, 8);
            if (store_leader) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");   // previous store has left the panel
            named_sync(3 + q, 128);
            const float inv = __fdividef(1.0f, l);
            unsigned char* stage = sO + q * NP * Q_PANEL + wrow * 128;
            const int chunks = (d + 7) >> 3;
#pragma unroll 1
            for (int c0 = 0; c0 < chunks; c0 += 4) {                 // 32 columns per round trip to TMEM
                uint32_t r[32];
                tmem_ld32(tO + c0 * 8, r);
                tmem_wait_ld();
#pragma unroll
                for (int cc4 = 0; cc4 < 4; ++cc4) {
                    const int c = c0 + cc4;
                    if (c < chunks) {
                        uint4 out;
                        out.x = pack2<T>(__uint_as_float(r[cc4 * 8 + 0]) * inv, __uint_as_float(r[cc4 * 8 + 1]) * inv);
                        out.y = pack2<T>(__uint_as_float(r[cc4 * 8 + 2]) * inv, __uint_as_float(r[cc4 * 8 + 3]) * inv);
                        out.z = pack2<T>(__uint_as_float(r[cc4 * 8 + 4]) * inv, __uint_as_float(r[cc4 * 8 + 5]) * inv);
                        out.w = pack2<T>(__uint_as_float(r[cc4 * 8 + 6]) * inv, __uint_as_float(r[cc4 * 8 + 7]) * inv);
                        *reinterpret_cast<uint4*>(stage + (c >> 3) * Q_PANEL + (((c & 7) ^ (wrow & 7)) << 4)) = out;   // 128-byte swizzle
                    }
                }
            }
            fence_before();      // the O reads above are ordered before the next item's first P arrive (-> PV overwrites O)
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            named_sync(3 + q, 128);
            if (store_leader) {
                for (int p = 0; p < NP; ++p)
                    tma_store_4d(&to, smem_u32(sO + (q * NP + p) * Q_PANEL), p * 64, (qb * 2 + q) * BM, h, b);
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
            DADD_TRACE_EVENT(q, s - 1, 9);
        }
        if (store_leader) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
    fence_before();
    __syncthreads();
    if (warp == 9) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

template <typename T, int NP, int BN_, int STAGES, int POLY>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, int B, int H, int N,
                  int d, float scale, cudaStream_t s) {
    const size_t smem = (size_t)4 * NP * 128 * 128 + (size_t)2 * STAGES * NP * BN_ * 128 + sizeof(Bars<STAGES>) + 1024;
    const int items = ((N + 2 * BM - 1) / (2 * BM)) * H * B;
    const int grid = items < num_sms() ? items : num_sms();
    const float sl2 = scale * 1.4426950408889634f;
    if constexpr (std::is_same_v<T, __nv_bfloat16> && NP == 1) {
        // debugging aid: DADD_ATTN_TRACE=<file> dumps CTA 0's event timeline of every launch (synchronises; never in a product run)
        static const char* trace_path = getenv("DADD_ATTN_TRACE");
        if (trace_path) {
            auto tk_ = self_attn_tc2_kernel<T, NP, BN_, STAGES, true, POLY>;
            if (cuda_ok(cudaFuncSetAttribute(tk_, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "self_attn_tc2 smem")) return 2;
            long long* buf = nullptr;
            const size_t n = 3 * 64 * 16;
            if (cuda_ok(cudaMalloc(&buf, n * sizeof(long long)), "trace alloc")) return 2;
            cudaMemsetAsync(buf, 0, n * sizeof(long long), s);
            tk_<<<grid, NTHREADS, smem, s>>>(tq, tk, tv, to, B, H, N, d, sl2, buf);
            cudaStreamSynchronize(s);
            long long* host = new long long[n];
            cudaMemcpy(host, buf, n * sizeof(long long), cudaMemcpyDeviceToHost);
            if (FILE* f = fopen(trace_path, "w")) {
                for (size_t i = 0; i < n; ++i) fprintf(f, "%lld%c", host[i], (i % 16 == 15) ? '\n' : ' ');
                fclose(f);
            }
            delete[] host;
            cudaFree(buf);
            return launched("dadd_self_attn_fwd(tcgen05 v2 trace)");
        }
    }
    auto kern = self_attn_tc2_kernel<T, NP, BN_, STAGES, false, POLY>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "self_attn_tc2 smem")) return 2;
    kern<<<grid, NTHREADS, smem, s>>>(tq, tk, tv, to, B, H, N, d, sl2, nullptr);
    return launched("dadd_self_attn_fwd(tcgen05 v2)");
}

}  // namespace tc2

bool self_attn_tc2_supported(int N, int d) { return N >= 128 && d % 8 == 0 && d >= 8 && d <= 128; }

int self_attn_tc2(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os, int B,
                  int H, int N, int d, float scale, int dtype, cudaStream_t s) {
    const int np = (d + 63) / 64;
    const int bn = np == 1 ? 128 : 64;
    CUtensorMap tq, tk, tv, to;
    if (tc::make_map(&tq, q, qs, B, H, N, d, dtype, 128) || tc::make_map(&tk, k, ks, B, H, N, d, dtype, bn) ||
        tc::make_map(&tv, v, vs, B, H, N, d, dtype, bn) || tc::make_map(&to, o, os, B, H, N, d, dtype, 128))
        return 1;
    // Share of the exponentials computed by the FMA-pipe polynomial, in pairs per 16.  Measured on B200 (N = 1024, d = 40,
    // B = 26): 0 -> 86.4 us, 2 -> 87.8, 3 -> 91.5, 4 -> 99.5 (profiles/r01_attn_poly_exp.txt): on sm_100a FFMA2/FADD2 issue at
    // half rate, so the polynomial costs the FMA pipe as many cycles per element as MUFU.EX2 costs the XU pipe and the
    // softmax warps (one per scheduler and turn) become issue-bound.  Default 0; DADD_ATTN_POLY=2 selects the mixed variant.
    static const int poly = getenv("DADD_ATTN_POLY") ? atoi(getenv("DADD_ATTN_POLY")) : 0;
#define DADD_TC2(NPV, BNV, STG, PL) DADD_DISPATCH_16(dtype, T, return (tc2::launch<T, NPV, BNV, STG, PL>(tq, tk, tv, to, B, H, N, d, scale, s)))
    if (np == 1) {
        if (poly == 2) DADD_TC2(1, 128, 4, 2);
        DADD_TC2(1, 128, 4, 0);
    }
    DADD_TC2(2, 64, 3, 0);
#undef DADD_TC2
    return 1;
}

}  // namespace daddk
