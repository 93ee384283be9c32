// K1: self-attention core, o = softmax(q k^T / sqrt d) v per (sample, head), on tcgen05 / TMEM / TMA.
// Reference op: F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 (installed at
// src/models/attention_processor_routing_gates.py:284-286; VAE mid block: src/models/vae/vae.py:90-112).
//
// What bounded the second generation (profiles/r01_attn_exp_phase.txt, r01_ncu_self_attn_tc2.txt): at d = 40 every score
// costs one exponential but only 160 MMA flop, so the kernel lives on the MUFU pipe (16 ex2 / clk / SM) - and that pipe
// was 70 % busy because (1) ONE warp per scheduler was ever inside the exp2 phase (in-order issue: ~45 other
// instructions per 32 exponentials could not overlap the 8-clk MUFU dispatch), (2) P was single-buffered in TMEM, so every
// step waited ~490 clk for PV(s-1), (3) with one warp per scheduler the FMA-pipe polynomial exp2 had nobody to overlap with,
// (4) Q was single-buffered: short sequences (N = 256) exposed a full TMA round trip per work item.
// This kernel:
//   * 16 softmax warps: a query row is shared by TWO threads (warps w and w+4 of a group own the same 32 TMEM lanes and
//     split the key tile's columns), so 4 warps per scheduler interleave exp2 / FMA / ALU work; the row maximum is
//     exchanged through shared memory behind a 64-thread named barrier, the row sum is combined once per item;
//   * S lives in NBUF rotating TMEM buffers and P (16-bit) is written IN PLACE over the S it came from, so nothing is
//     single-buffered: the MMA issuer runs one chain "wait P(i) -> PV(i) -> QK(i + NBUF)" (tcgen05.mma executes in issue
//     order, so QK(i + NBUF) cannot overtake the PV(i) that still reads the buffer) and a softmax group never waits for the
//     tensor core except on the rare lazy O rescale;
//   * row sums come from the tensor core (ONES): the key-tile of V gets a column of ones at column d (a spare column of
//     its last 64-column panel), so O[:, d] accumulates sum_k P[:, k] in fp32 - from the very 16-bit P the PV product uses -
//     and the exp2 phase is one scalar FFMA, one MUFU.EX2 and half an F2FP per score: with two warps per scheduler in the
//     phase that stream runs at 9.8 clk per exponential against 11.7 with packed f32x2 FMAs and FADD row sums
//     (scripts/micro/smsp_mix.cu, profiles/r02_attn_microbench.txt); a polynomial exp2 share on the FMA pipe measured
//     slower in every mix (the scheduler starves MUFU-bound warps when FMA-bound ones are ready) and is gone;
//   * Q is double-buffered across work items and O leaves through the dead Q buffer of its own item (no extra staging
//     memory): the Q producer thread stores it by TMA and refills the buffer with the Q of the item after next;
//   * O columns in TMEM and the PV MMAs cover round16(d) columns, not 64-column panels (d = 40: N = 48).
// Warps: 0-15 softmax (group g = warp / 8, column half h = (warp / 4) & 1, TMEM lane quarter = warp & 3), 16 = K/V TMA
// producer, 17 = MMA issuer, 18 = Q TMA producer + O store.  Work item = (sample, head, 256 query rows); a persistent
// CTA per SM walks its items as one continuous stream of key steps.
// TMEM: S/P buffers [0, NBUF * BN), O_g at NBUF * BN + g * round16(d).
#include <cstdlib>

#include "tc_util.cuh"

namespace daddk {
namespace tcsa {

using namespace daddk::tc;

constexpr int NSOFT_WARPS = 16;
constexpr int NTHREADS = 608;      // 19 warps: 65536 / 608 leaves 104 registers per thread (the softmax threads hold 64 scores)
constexpr float RESCALE_LOG2 = 8.0f;

template <int STAGES, int QSTAGES, int NBUF>
struct Bars {
    uint64_t q_full[2][QSTAGES], q_free[2][QSTAGES];      // q_free: only without O staging (QSTAGES == 1)
    uint64_t s_full[NBUF];
    uint64_t p_full[2][2], pv_done[2], o_ready[2];      // p_full[group][step parity]: a group may finish P(s + 1) before the issuer has looked at P(s)
    uint64_t k_full[STAGES], v_full[STAGES], v_ready[STAGES], kv_empty[STAGES];      // v_ready: V tile with its ones column (ONES)
    uint32_t tmem_base;
    uint32_t pad;
    float xch[2][2][2][128];      // [group][step parity][column half][row]: row maxima of the two halves of a row
    float xl[2][2][128];          // [group][column half][row]: row sums of the two halves (once per item)
};

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ float ex2_ordered(float x) {
    float y;
    asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ uint64_t desc_add(uint64_t desc, uint32_t bytes) { return desc + (uint64_t)(bytes >> 4); }
__device__ __forceinline__ void named_sync(int id, int count) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(count) : "memory"); }
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : DADD_R8(r, 0), DADD_R8(r, 8)
                 : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld32p(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : DADD_R8(r, 0), DADD_R8(r, 8), DADD_R8(r, 16), DADD_R8(r, 24)
        : "r"(taddr));
}

// Debugging aids (never on a product path): a protocol bug shows up as a bounded mbarrier wait running out.  The waiter then
// records {magic, CTA, warp, barrier id, parity, step} in a host-mapped buffer (readable after the context died) and traps.
// DADD_ATTN_TRACE=<file> runs the TRACE instance of the d <= 64 kernel, which stores clock64() at the protocol events of the
// first steps of CTA 0 (one row per (warp, step)); scripts/attn_trace.py prints it.
struct Dbg {
    unsigned int* rec;        // host-mapped, 16 words
    long long* trace;         // [20 warps][TRACE_STEPS][8 events] or nullptr
};
constexpr int TRACE_STEPS = 24;
static unsigned int* g_dbg_host = nullptr;
__device__ __forceinline__ void mbar_wait_dbg(uint64_t* bar, uint32_t parity, const Dbg& dbg, int id, uint32_t step) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
        if (done) return;
        if (++spins > SPIN_LIMIT) {
            if (dbg.rec && atomicCAS(dbg.rec, 0u, 0xdeadbeefu) == 0u) {
                dbg.rec[1] = blockIdx.x;
                dbg.rec[2] = threadIdx.x >> 5;
                dbg.rec[3] = (unsigned)id;
                dbg.rec[4] = parity;
                dbg.rec[5] = step;
                __threadfence_system();
            }
            __trap();
        }
    }
}
#define TC_EVENT(step, ev)                                                                                          \
    do {                                                                                                             \
        if constexpr (TRACE) {                                                                                       \
            if (blockIdx.x == 0 && lane == 0 && (step) < TRACE_STEPS)                                                \
                dbg.trace[((size_t)warp * TRACE_STEPS + (step)) * 8 + (ev)] = clock64();                             \
        }                                                                                                            \
    } while (0)

// T: element type; NP: 64-column panels per head (ceil(d / 64)); BN: keys per step; HALF = BN / 2 columns per thread.
template <typename T, int NP, int BN_, int NBUF, int STAGES, int QSTAGES, bool ONES, int KS, bool TRACE>
__global__ void __launch_bounds__(NTHREADS, 1)
self_attn_tc_kernel(const __grid_constant__ CUtensorMap tq, const __grid_constant__ CUtensorMap tk,
                     const __grid_constant__ CUtensorMap tv, const __grid_constant__ CUtensorMap to, T* __restrict__ o_ptr,
                     int64_t o_stride, int B, int H, int N, int d, float scale_log2e, const Dbg dbg) {
    constexpr uint32_t TMEM_COLS = 512;
    constexpr uint32_t FMT = std::is_same_v<T, __nv_bfloat16> ? 1u : 0u;
    constexpr uint32_t IDESC_QK = instr_desc(FMT, BN_, 0);
    constexpr uint32_t Q_PANEL = 128 * 128, KV_PANEL = BN_ * 128;
    constexpr int HALF = BN_ / 2;
    static_assert(BN_ == 64 || BN_ == 128, "key tile");
    // QSTAGES == 2: O leaves through the item's own dead Q buffer (store warp, TMA store): measured 10 % faster than 16-byte
    // stores from the softmax threads (rows 640 B apart: 32 half-used sectors per store instruction; same-box A/B in
    // profiles/r02_attn_v3.txt).  QSTAGES == 1 (large d: no room for a second Q stage): the next item's Q must not wait for this
    // item's epilogue - the MMA issuer would block on it before it has issued the PVs that epilogue needs - so the Q buffer is
    // released by the item's last QK^T and O is written straight from registers.
    constexpr bool O_VIA_Q = QSTAGES >= 2;
    static_assert(NBUF * BN_ + 2 * 16 <= 512, "TMEM budget");
    using BarsT = Bars<STAGES, QSTAGES, NBUF>;

    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sQ = smem;                                   // [2 groups][QSTAGES][NP panels]; also the O staging of its item
    unsigned char* sK = sQ + 2 * QSTAGES * NP * Q_PANEL;        // [STAGES][NP]
    unsigned char* sV = sK + STAGES * NP * KV_PANEL;            // [STAGES][NP]
    BarsT* bars = reinterpret_cast<BarsT*>(sV + STAGES * NP * KV_PANEL);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nkv = (N + BN_ - 1) / BN_;
    const int nq2 = (N + 2 * BM - 1) / (2 * BM);
    const int items = nq2 * H * B;
    const int my_items = (items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;
    const uint32_t steps = (uint32_t)my_items * (uint32_t)nkv;
    const uint32_t tiles = 2 * steps;                           // tile i = (step i / 2, group i & 1) uses S buffer i % NBUF
    const int ksteps = KS > 0 ? KS : (d + 15) >> 4;             // KS: the UNet's head sizes get a compile-time K loop (a lean issuer)
    const int ow = ((d + (ONES ? 16 : 15)) >> 4) << 4;          // O columns per group that the PV MMAs write (N of the MMA); ONES: + the row-sum column d
    const int og = (ow + 31) & ~31;                             // column stride between the groups' O tiles
    const uint32_t col_o = NBUF * BN_;

    if (tid == 0) {
        for (int g = 0; g < 2; ++g) {
            for (int s = 0; s < QSTAGES; ++s) {
                mbar_init(&bars->q_full[g][s], 1);
                mbar_init(&bars->q_free[g][s], 1);
            }
            mbar_init(&bars->p_full[g][0], 8);
            mbar_init(&bars->p_full[g][1], 8);
            mbar_init(&bars->pv_done[g], 1);
            mbar_init(&bars->o_ready[g], 8);
        }
        for (int b = 0; b < NBUF; ++b) mbar_init(&bars->s_full[b], 1);
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bars->k_full[s], 1);
            mbar_init(&bars->v_full[s], 1);
            mbar_init(&bars->v_ready[s], 1);
            mbar_init(&bars->kv_empty[s], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 17) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&bars->tmem_base)), "n"(TMEM_COLS));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_before();
    __syncthreads();
    fence_after();
    const uint32_t tmem = bars->tmem_base;

    if (warp >= NSOFT_WARPS) {
        if (warp == 16) {
            // ------------------------------------------------------------------ TMA producer: K/V ring over the step stream.
            // ONES: the warp also writes the column of ones into every V tile once it has landed - element d of each key row, the
            // first of a 16-byte chunk that TMA zero-filled (d % 8 == 0, d % 64 != 0) - LAG loads behind the one it issues, and
            // hands the tile to the issuer through v_ready.  Generic-proxy stores: a proxy fence orders them before the PV MMAs.
            constexpr uint32_t LAG = STAGES >= 3 ? 2 : 1;
            uint32_t st = 0, eph = 1, fst = 0, fph = 0;
            int j = 0, item = (int)blockIdx.x, h = (item / nq2) % H, b = item / (nq2 * H);
            auto ones_column = [&](uint32_t it) {
                mbar_wait_dbg(&bars->v_full[fst], fph, dbg, 11, it);
                unsigned char* vp = sV + (fst * NP + (d >> 6)) * KV_PANEL;
                const uint32_t chunk = (uint32_t)(d & 63) >> 3;
                const uint4 one = make_uint4(std::is_same_v<T, __nv_bfloat16> ? 0x3F80u : 0x3C00u, 0u, 0u, 0u);
#pragma unroll
                for (int r = lane; r < BN_; r += 32) *reinterpret_cast<uint4*>(vp + r * 128 + ((chunk ^ (r & 7)) << 4)) = one;
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->v_ready[fst]);
                if (++fst == STAGES) fst = 0, fph ^= 1;
            };
            for (uint32_t it = 0; it < steps; ++it) {
                if (lane == 0) {
                    mbar_wait_dbg(&bars->kv_empty[st], eph, dbg, 1, it);
                    mbar_expect_tx(&bars->k_full[st], NP * KV_PANEL);
                    for (int p = 0; p < NP; ++p)
                        tma_load_4d(smem_u32(sK + (st * NP + p) * KV_PANEL), &tk, &bars->k_full[st], p * 64, j * BN_, h, b);
                    mbar_expect_tx(&bars->v_full[st], NP * KV_PANEL);
                    for (int p = 0; p < NP; ++p)
                        tma_load_4d(smem_u32(sV + (st * NP + p) * KV_PANEL), &tv, &bars->v_full[st], p * 64, j * BN_, h, b);
                }
                if (++st == STAGES) st = 0, eph ^= 1;
                if (++j == nkv) {
                    j = 0;
                    item += (int)gridDim.x;
                    h = (item / nq2) % H, b = item / (nq2 * H);
                }
                if constexpr (ONES) {
                    __syncwarp();
                    if (it >= LAG) ones_column(it - LAG);
                }
            }
            if constexpr (ONES)
                for (uint32_t it = steps > LAG ? steps - LAG : 0; it < steps; ++it) ones_column(it);
        } else if (warp == 18) {
            if (lane == 0) {
                // -------------------------------------------------------------- TMA: the two query tiles of each item, QSTAGES items deep
                auto coords = [&](int ti, int& qb, int& h, int& b) {
                    const int item = (int)blockIdx.x + ti * (int)gridDim.x;
                    qb = item % nq2, h = (item / nq2) % H, b = item / (nq2 * H);
                };
                auto load_q = [&](int ti, int g) {
                    int qb, h, b;
                    coords(ti, qb, h, b);
                    const int qs = ti % QSTAGES;
                    mbar_expect_tx(&bars->q_full[g][qs], NP * Q_PANEL);
                    for (int p = 0; p < NP; ++p)
                        tma_load_4d(smem_u32(sQ + ((g * QSTAGES + qs) * NP + p) * Q_PANEL), &tq, &bars->q_full[g][qs], p * 64,
                                    (qb * 2 + g) * BM, h, b);
                };
                if constexpr (O_VIA_Q) {
                    // the buffer of item ti is refilled with the Q of item ti + QSTAGES as soon as its O store has been read out of
                    // shared memory: one thread runs loads and stores as a single chain
                    for (int ti = 0; ti < QSTAGES && ti < my_items; ++ti)
                        for (int g = 0; g < 2; ++g) load_q(ti, g);
                    for (int ti = 0; ti < my_items; ++ti) {
                        int qb, h, b;
                        coords(ti, qb, h, b);
                        const int qs = ti % QSTAGES;
                        for (int g = 0; g < 2; ++g) {
                            mbar_wait_dbg(&bars->o_ready[g], ti & 1, dbg, 2, ti);
                            for (int p = 0; p < NP; ++p)
                                tma_store_4d(&to, smem_u32(sQ + ((g * QSTAGES + qs) * NP + p) * Q_PANEL), p * 64, (qb * 2 + g) * BM, h, b);
                            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                            if (ti + QSTAGES < my_items) {
                                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
                                load_q(ti + QSTAGES, g);
                            }
                        }
                    }
                    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
                } else {
                    for (int ti = 0; ti < my_items; ++ti)
                        for (int g = 0; g < 2; ++g) {
                            mbar_wait_dbg(&bars->q_free[g][0], (ti & 1) ^ 1, dbg, 3, ti);     // the previous item's last QK^T has read the tile
                            load_q(ti, g);
                        }
                }
            }
        } else {
            // ------------------------------------------------------------------ MMA issuer (whole warp walks the tile stream)
            // The issuer shares its scheduler with four softmax warps and every instruction it executes queues behind theirs,
            // so the loop carries running counters (no division / modulo by run-time values) and ready-made descriptors.
            const bool leader = elect_one();
            const uint64_t dq0 = smem_desc(smem_u32(sQ), 16, 1024);
            const uint64_t dk0 = smem_desc(smem_u32(sK), 16, 1024);
            const uint64_t dv0 = smem_desc(smem_u32(sV), KV_PANEL, 1024);
            const uint32_t idesc_pv = instr_desc(FMT, (uint32_t)ow, 1);
            // position of a stream (QK^T or PV) inside the tile sequence
            struct Pos {
                uint32_t g = 0, j = 0, qs = 0, qph = 0, st = 0, kph = 0, buf = 0;
            };
            auto advance = [&](Pos& p) {
                if (++p.buf == NBUF) p.buf = 0;
                if (p.g == 0) {
                    p.g = 1;
                    return;
                }
                p.g = 0;
                if (++p.st == STAGES) p.st = 0, p.kph ^= 1;
                if (++p.j == (uint32_t)nkv) {
                    p.j = 0;
                    if (++p.qs == QSTAGES) p.qs = 0, p.qph ^= 1;
                }
            };
            Pos pq, pp;
            uint32_t ps0 = 0, ps1 = 0;                                      // steps consumed per group: barrier = step & 1, phase = (step >> 1) & 1
            // S(i) = Q_g K(step)^T into buffer i % NBUF
            auto issue_qk = [&](uint32_t i) {
                if (pq.j == 0) mbar_wait_dbg(&bars->q_full[pq.g][pq.qs], pq.qph, dbg, 4, i);
                if (pq.g == 0) mbar_wait_dbg(&bars->k_full[pq.st], pq.kph, dbg, 5, i);
                fence_after();
                TC_EVENT(i - NBUF, 4);
                if (leader) {
                    const uint64_t da = desc_add(dq0, (pq.g * QSTAGES + pq.qs) * NP * Q_PANEL), db = desc_add(dk0, pq.st * NP * KV_PANEL);
                    const uint32_t tS = tmem + pq.buf * BN_;
                    auto qk = [&](int ks) {
                        const uint32_t offq = (ks >> 2) * Q_PANEL + (ks & 3) * 32, offk = (ks >> 2) * KV_PANEL + (ks & 3) * 32;
                        mma_ss(tS, desc_add(da, offq), desc_add(db, offk), IDESC_QK, ks > 0);
                    };
                    if constexpr (KS > 0) {
#pragma unroll
                        for (int ks = 0; ks < KS; ++ks) qk(ks);
                    } else {
                        for (int ks = 0; ks < ksteps; ++ks) qk(ks);
                    }
                    mma_commit(&bars->s_full[pq.buf]);
                    if (!O_VIA_Q && pq.j + 1 == (uint32_t)nkv) mma_commit(&bars->q_free[pq.g][pq.qs]);
                }
                __syncwarp();
                advance(pq);
            };
            // O_g (+)= P(i) V(step); P(i) sits in the 16-bit view of S buffer i % NBUF
            auto issue_pv = [&](uint32_t i) {
                const uint32_t ps = pp.g ? ps1++ : ps0++;
                mbar_wait_dbg(&bars->p_full[pp.g][ps & 1], (ps >> 1) & 1, dbg, 6, i);
                if (pp.g == 0) mbar_wait_dbg(ONES ? &bars->v_ready[pp.st] : &bars->v_full[pp.st], pp.kph, dbg, 7, i);
                fence_after();
                TC_EVENT(i, 3);
                if (leader) {
                    const uint32_t tO = tmem + col_o + pp.g * og, tP = tmem + pp.buf * BN_;
                    // ONE MMA per 16 keys over all ow columns: V's 64-column panels are KV_PANEL bytes apart, which is the
                    // descriptor's leading byte offset, and an A-from-TMEM MMA costs ~62 clk whatever its N (scripts/micro/mma_rate.cu)
                    const uint64_t dv = desc_add(dv0, pp.st * NP * KV_PANEL);
#pragma unroll
                    for (int kk = 0; kk < BN_ / 16; ++kk)
                        mma_ts(tO, tP + kk * 8, desc_add(dv, kk * 2048), idesc_pv, (pp.j > 0 || kk > 0) ? 1u : 0u);
                    mma_commit(&bars->pv_done[pp.g]);
                    if (pp.g == 1) mma_commit(&bars->kv_empty[pp.st]);
                }
                __syncwarp();
                advance(pp);
            };
            for (uint32_t i = 0; i < (uint32_t)NBUF && i < tiles; ++i) issue_qk(i);
            for (uint32_t i = 0; i < tiles; ++i) {
                TC_EVENT(i, 0);
                issue_pv(i);
                TC_EVENT(i, 1);
                if (i + NBUF < tiles) issue_qk(i + NBUF);      // in issue order behind PV(i), which still reads that buffer
                TC_EVENT(i, 2);
            }
        }
    } else {
        // ---------------------------------------------------------------------- softmax / correction / epilogue
        const int g = warp >> 3;                                   // query tile of this warp's group
        const int half = (warp >> 2) & 1;                          // which half of the key tile's columns this thread owns
        const int quarter = warp & 3;
        const int wrow = quarter * 32 + lane;                      // row inside the tile == TMEM lane
        const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
        const uint32_t tO = tmem + col_o + g * og + lane_base;
        const int pair_bar = 1 + g * 4 + quarter;                  // named barrier shared with the partner warp
        // this thread's share of the O columns (8-column chunks) in the epilogue and - plus the row-sum chunk - in the rare rescale
        const int chunks = (d + 7) >> 3, c_lo = half == 0 ? 0 : (chunks + 1) / 2, c_hi = half == 0 ? (chunks + 1) / 2 : chunks;
        const int r_hi = (ONES && half == 1) ? chunks + 1 : c_hi;
        uint32_t s = 0, buf = (uint32_t)g % NBUF, sph = 0;         // tile i = 2 s + g lives in buffer i % NBUF, phase (i / NBUF) & 1
        // Epilogue of item te (its last step is s - 1): O / l -> the item's own (dead) Q buffer in the swizzled TMA layout -> store
        // warp, or straight to global memory.  (Deferring it into the first step of the next item - to hide the issuer's
        // P -> PV-complete latency - measured equal: same-box A/B in profiles/r02_attn_v3.txt.)
        auto epilogue = [&](int te, float l_own) {
            if constexpr (!ONES) bars->xl[g][half][wrow] = l_own;
            mbar_wait_dbg(&bars->pv_done[g], (s - 1) & 1, dbg, 10, s);
            fence_after();
            float inv;
            if constexpr (ONES) {                                    // the row sum sits in O's column d
                uint32_t r[8];
                tmem_ld8(tO + chunks * 8, r);
                tmem_wait_ld();
                inv = __fdividef(1.0f, __uint_as_float(r[0]));
            } else {
                named_sync(pair_bar, 64);
                inv = __fdividef(1.0f, l_own + bars->xl[g][half ^ 1][wrow]);
            }
            const int item = (int)blockIdx.x + te * (int)gridDim.x;
            const int row = ((item % nq2) * 2 + g) * BM + wrow;
            T* orow = o_ptr + ((int64_t)(item / (nq2 * H)) * N + row) * o_stride + (int64_t)((item / nq2) % H) * d;
            unsigned char* stage = sQ + (g * QSTAGES + te % QSTAGES) * NP * Q_PANEL + wrow * 128;
#pragma unroll 1
            for (int c = c_lo; c < c_hi; ++c) {
                uint32_t r[8];
                tmem_ld8(tO + c * 8, r);
                tmem_wait_ld();
                uint4 out;
                out.x = pack2<T>(__uint_as_float(r[0]) * inv, __uint_as_float(r[1]) * inv);
                out.y = pack2<T>(__uint_as_float(r[2]) * inv, __uint_as_float(r[3]) * inv);
                out.z = pack2<T>(__uint_as_float(r[4]) * inv, __uint_as_float(r[5]) * inv);
                out.w = pack2<T>(__uint_as_float(r[6]) * inv, __uint_as_float(r[7]) * inv);
                if constexpr (O_VIA_Q)
                    *reinterpret_cast<uint4*>(stage + (c >> 3) * Q_PANEL + (((c & 7) ^ (wrow & 7)) << 4)) = out;   // 128-byte swizzle
                else if (row < N)
                    *reinterpret_cast<uint4*>(orow + c * 8) = out;
            }
            fence_before();      // the O reads above are ordered before the p_full arrive that lets the next PV overwrite O
            if constexpr (O_VIA_Q) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&bars->o_ready[g]);
            }
        };
        float m_used = -INFINITY, l = 0.0f;
        int ti = 0, j = 0;
        auto step = [&]() {                                        // one key step of this group
            const uint32_t tS = tmem + buf * BN_ + lane_base;
            TC_EVENT(s, 0);
            mbar_wait_dbg(&bars->s_full[buf], sph, dbg, 8, s);
            buf += 2;
            if (buf >= NBUF) buf -= NBUF, sph ^= 1;
            fence_after();
            TC_EVENT(s, 1);
            uint32_t sr[HALF];
            if constexpr (HALF == 64) tmem_ld64(tS + half * HALF, sr);
            else tmem_ld32p(tS + half * HALF, sr);
            tmem_wait_ld();
            TC_EVENT(s, 2);
            const bool ragged = (j + 1) * BN_ > N;
            if (ragged) {                                        // ragged last key tile
#pragma unroll
                for (int e = 0; e < HALF; ++e)
                    if (j * BN_ + half * HALF + e >= N) sr[e] = 0xff800000u;   // -inf
            }
            float mx0 = __uint_as_float(sr[0]), mx1 = __uint_as_float(sr[1]), mx2 = __uint_as_float(sr[2]),
                  mx3 = __uint_as_float(sr[3]);
#pragma unroll
            for (int e = 4; e < HALF; e += 8) {
                mx0 = max3(mx0, __uint_as_float(sr[e]), __uint_as_float(sr[e + 1]));
                mx1 = max3(mx1, __uint_as_float(sr[e + 2]), __uint_as_float(sr[e + 3]));
                if (e + 4 < HALF) {
                    mx2 = max3(mx2, __uint_as_float(sr[e + 4]), __uint_as_float(sr[e + 5]));
                    mx3 = max3(mx3, __uint_as_float(sr[e + 6]), __uint_as_float(sr[e + 7]));
                }
            }
            float tmax = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
            // the other half of the row: exchange through shared memory (double-buffered by step parity; the pair barrier
            // also orders the partner's tcgen05.ld of its S columns before my in-place P store below)
            TC_EVENT(s, 3);
            bars->xch[g][s & 1][half][wrow] = tmax;
            fence_before();
            named_sync(pair_bar, 64);
            fence_after();
            tmax = fmaxf(tmax, bars->xch[g][s & 1][half ^ 1][wrow]);
            TC_EVENT(s, 4);
            // lazy offset update: keep exponentiating against m_used until a row maximum outgrows it by 2^8
            const bool grow = (tmax - m_used) * scale_log2e > RESCALE_LOG2;
            if (__any_sync(0xffffffffu, grow)) {                 // (same decision in the partner warp: same rows, same values)
                float alpha = 1.0f;
                if (grow) {
                    alpha = ex2((m_used - tmax) * scale_log2e);   // 0 on the first tile (m_used = -inf)
                    m_used = tmax;
                    if constexpr (!ONES) l *= alpha;
                }
                if (j > 0) {
                    mbar_wait_dbg(&bars->pv_done[g], (s - 1) & 1, dbg, 9, s);    // O_g must hold every PV up to step s-1 before it is scaled
                    fence_after();
#pragma unroll 1
                    for (int c = c_lo; c < r_hi; ++c) {
                        uint32_t v[8];
                        tmem_ld8(tO + c * 8, v);
                        tmem_wait_ld();
#pragma unroll
                        for (int e = 0; e < 8; ++e) v[e] = __float_as_uint(__uint_as_float(v[e]) * alpha);
                        tmem_st8(tO + c * 8, v);
                    }
                }
            }
            const float nmb = -m_used * scale_log2e;
            const uint32_t tP = tS + half * (HALF / 2);          // 16-bit pairs: my HALF probabilities = HALF / 2 columns
            TC_EVENT(s, 5);
            // exp2 phase: scalar FFMA + MUFU.EX2 + half an F2FP per score (a masked -inf score maps to exactly 0)
            float l0 = 0.0f, l1 = 0.0f, l2 = 0.0f, l3 = 0.0f;
#pragma unroll
            for (int c = 0; c < HALF / 16; ++c) {
                uint32_t pk[8];
#pragma unroll
                for (int k = 0; k < 8; ++k) {
                    // (volatile: keeps the MUFU stream in program order between the chunks' tcgen05.st, i.e. interleaved with the
                    // packs and stores of the previous chunk - ptxas otherwise hoists all 64 exponentials into one block)
                    const float p0 = ex2_ordered(fmaf(__uint_as_float(sr[c * 16 + 2 * k]), scale_log2e, nmb));
                    const float p1 = ex2_ordered(fmaf(__uint_as_float(sr[c * 16 + 2 * k + 1]), scale_log2e, nmb));
                    if constexpr (!ONES) {
                        if (k & 1) l2 += p0, l3 += p1;
                        else l0 += p0, l1 += p1;
                    }
                    pk[k] = pack2<T>(p0, p1);
                }
                tmem_st8(tP + c * 8, pk);
            }
            if constexpr (!ONES) l += (l0 + l1) + (l2 + l3);
            tmem_wait_st();
            fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bars->p_full[g][s & 1]);   // one arrival per warp (8 per group)
            TC_EVENT(s, 6);
        };
        for (ti = 0; ti < my_items; ++ti) {
            m_used = -INFINITY, l = 0.0f;
            for (j = 0; j < nkv; ++j, ++s) step();
            epilogue(ti, l);
        }
    }
    fence_before();
    __syncthreads();
    if (warp == 17) {
        fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(TMEM_COLS));
    }
}

static Dbg debug_buffers() {
    static Dbg d = [] {
        Dbg x{nullptr, nullptr};
        void* host = nullptr;
        if (cudaHostAlloc(&host, 64, cudaHostAllocMapped) == cudaSuccess) {
            memset(host, 0, 64);
            void* dev = nullptr;
            if (cudaHostGetDevicePointer(&dev, host, 0) == cudaSuccess) x.rec = (unsigned int*)dev;
            g_dbg_host = (unsigned int*)host;
        }
        return x;
    }();
    return d;
}

template <typename T, int NP, int BN_, int NBUF, int STAGES, int QSTAGES, bool ONES, int KS>
static int launch(const CUtensorMap& tq, const CUtensorMap& tk, const CUtensorMap& tv, const CUtensorMap& to, void* o, int64_t os,
                  int B, int H, int N, int d, float scale, cudaStream_t s) {
    const size_t smem = (size_t)2 * QSTAGES * NP * 128 * 128 + (size_t)2 * STAGES * NP * BN_ * 128 +
                        sizeof(Bars<STAGES, QSTAGES, NBUF>) + 1024;
    const int items = ((N + 2 * BM - 1) / (2 * BM)) * H * B;
    const int grid = items < num_sms() ? items : num_sms();
    const float sl2 = scale * 1.4426950408889634f;
    Dbg dbg = debug_buffers();
    static const char* trace_path = getenv("DADD_ATTN_TRACE");
    static const bool debug_sync = getenv("DADD_ATTN_DEBUG") != nullptr;
    if constexpr (NP == 1 && KS == 3 && std::is_same_v<T, __half>) {
        if (trace_path) {      // debugging aid: synchronises and writes CTA 0's event timeline (never in a product run)
            auto tkern = self_attn_tc_kernel<T, NP, BN_, NBUF, STAGES, QSTAGES, ONES, KS, true>;
            if (cuda_ok(cudaFuncSetAttribute(tkern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "self_attn_tc smem")) return 2;
            const size_t n = (size_t)20 * TRACE_STEPS * 8;
            if (cuda_ok(cudaMalloc(&dbg.trace, n * sizeof(long long)), "trace alloc")) return 2;
            cudaMemsetAsync(dbg.trace, 0, n * sizeof(long long), s);
            tkern<<<grid, NTHREADS, smem, s>>>(tq, tk, tv, to, (T*)o, os, B, H, N, d, sl2, dbg);
            cudaStreamSynchronize(s);
            long long* host = new long long[n];
            cudaMemcpy(host, dbg.trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
            if (FILE* f = fopen(trace_path, "w")) {
                for (size_t i = 0; i < n; ++i) fprintf(f, "%lld%c", host[i], (i % 8 == 7) ? '\n' : ' ');
                fclose(f);
            }
            delete[] host;
            cudaFree(dbg.trace);
            return launched("dadd_self_attn_fwd(tcgen05 trace)");
        }
    }
    auto kern = self_attn_tc_kernel<T, NP, BN_, NBUF, STAGES, QSTAGES, ONES, KS, false>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "self_attn_tc smem")) return 2;
    kern<<<grid, NTHREADS, smem, s>>>(tq, tk, tv, to, (T*)o, os, B, H, N, d, sl2, dbg);
    if (debug_sync) {          // DADD_ATTN_DEBUG=1: find out which wait ran out (see mbar_wait_dbg)
        const cudaError_t e = cudaStreamSynchronize(s);
        if (e != cudaSuccess && g_dbg_host)
            fprintf(stderr, "self_attn_tc<NP=%d BN=%d NBUF=%d STAGES=%d QSTAGES=%d>: %s; timeout record: magic=%x cta=%u warp=%u barrier=%u parity=%u step=%u\n",
                    NP, BN_, NBUF, STAGES, QSTAGES, cudaGetErrorString(e), g_dbg_host[0], g_dbg_host[1], g_dbg_host[2], g_dbg_host[3],
                    g_dbg_host[4], g_dbg_host[5]);
    }
    return launched("dadd_self_attn_fwd(tcgen05)");
}

}  // namespace tcsa

bool self_attn_tc_supported(int N, int d) { return N >= 128 && d % 8 == 0 && d >= 8 && d <= 160; }

int self_attn_tc(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os, int B,
                  int H, int N, int d, float scale, int dtype, cudaStream_t s) {
    const int np = (d + 63) / 64;
    const int bn = np <= 2 ? 128 : 64;
    CUtensorMap tq, tk, tv, to;
    if (tc::make_map(&tq, q, qs, B, H, N, d, dtype, 128) || tc::make_map(&tk, k, ks, B, H, N, d, dtype, bn) ||
        tc::make_map(&tv, v, vs, B, H, N, d, dtype, bn) || tc::make_map(&to, o, os, B, H, N, d, dtype, 128))
        return 1;
    // ONES (row sums from the tensor core) needs a spare column behind d in V's last panel and room for it in TMEM
    const bool ones = d % 64 != 0 && d <= 128;
#define DADD_TCSA(NPV, BNV, NBUFV, STG, QSTG, ON, KSV) \
    DADD_DISPATCH_16(dtype, T, return (tcsa::launch<T, NPV, BNV, NBUFV, STG, QSTG, ON, KSV>(tq, tk, tv, to, o, os, B, H, N, d, scale, s)))
    if (d == 40) DADD_TCSA(1, 128, 3, 4, 2, true, 3);              // the UNet's three head sizes: compile-time K loop
    if (d == 80) DADD_TCSA(2, 128, 2, 2, 1, true, 5);
    if (d == 160) DADD_TCSA(3, 64, 3, 2, 1, false, 10);
    if (np == 1) {
        if (ones) DADD_TCSA(1, 128, 3, 4, 2, true, 0);
        DADD_TCSA(1, 128, 3, 4, 2, false, 0);
    }
    if (np == 2) {
        if (ones) DADD_TCSA(2, 128, 2, 2, 1, true, 0);
        DADD_TCSA(2, 128, 2, 2, 1, false, 0);
    }
    DADD_TCSA(3, 64, 3, 2, 1, false, 0);
#undef DADD_TCSA
    return 1;
}

}  // namespace daddk
