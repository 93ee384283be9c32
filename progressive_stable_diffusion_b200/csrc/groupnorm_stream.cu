// K4s: NHWC GroupNorm (+ per-(sample, channel) additive term) (+ SiLU) as ONE streaming launch.
// Reference ops: diffusers ResnetBlock2D.norm1/norm2 + SiLU, conv_norm_out + conv_act, Transformer2DModel.norm (reached via
// src/models/unet/unet.py:140-146) and the up path's torch.cat in front of norm1.
//
// Why a third form.  The cluster kernel (groupnorm.cu) reads x once, but every CTA lives through load -> statistics -> cluster
// barrier -> normalise -> store, all 296 resident CTAs in phase: the memory system idles while the SMs compute and the other
// way round, and the two passes cost 21 instructions per element (profiles/r02_ncu_gn_cluster.txt: 42 % of the HBM peak,
// barrier stalls 3.3 per issue, 2.1 IPC).  Here:
//   * persistent CTAs walk work items (sample b, pixel slab r of S) in a fixed order; a producer warp keeps a ring of NST slabs
//     filled by TMA (3-D boxes {64 channels, PB panels, npix pixels}, 128-byte swizzle, PB odd: ldmatrix over 8 pixels of one
//     panel is conflict-free) together with the sample's chan_add row and its pixel-0 row, so HBM reads of the items ahead run
//     under the math of the current one and nothing on the consumers' path waits for global memory;
//   * statistics come from the tensor core: for a 16-pixel x 8-channel block X one mma.m16n8k16 with A = [X^T ; 1] and B = X
//     yields the Gram diagonal (sum x^2) in rows 0-7 and the column sums in rows 8-15: 3 instructions per 256 elements instead
//     of ~4 per element.  Values are shifted by the group's first element of the sample (one packed subtract per fragment) so
//     that E[d^2] - E[d]^2 never cancels;
//   * the S slabs of a sample exchange their per-group partial sums through global memory, off the consumers' path: a publisher
//     warp turns the per-channel sums of item k into group partials, stores them and a release flag; a combiner warp polls the S
//     flags of item k-1's sample, adds the partials in rank order (bit-reproducible) and hands (mean, rstd) to the consumers
//     through an mbarrier.  No cluster barrier, no clusters: all SMs work and S is not limited to 8.  Flags carry a launch
//     generation kept on the device, so a captured CUDA graph replays without a memset node;
//   * the normalise pass is thread = fixed 8-channel column (scale / shift in registers), LDS.128 -> FFMA -> tanh.approx -> FFMA
//     -> STG.128, 5.6 instructions per element.
// Forward progress: every CTA publishes item k before it normalises item k-1, whose wait is the only one that depends on other
// CTAs; a sample's items sit in at most two consecutive rounds (S <= gridDim.x) and all CTAs are resident - no wait cycle.
#include <cstdlib>
#include <mutex>

#include "tc_util.cuh"

namespace daddk {
namespace gns {

constexpr int MAX_G = 64;
constexpr int MAX_S = 128;
constexpr int MAX_ITEMS = 1 << 16;
constexpr int MAX_NST = 5;
constexpr uint32_t SPIN_LIMIT = 1u << 24;

struct Params {
    const void* x;
    const void* x2;
    const float* gamma;
    const float* beta;
    const float* chan_add;
    void* y;
    float2* part;           // [items][G]
    uint32_t* flags;        // [items]
    uint32_t* ctl;          // {generation, finished CTAs}
    int64_t add_stride;
    int C1, HW, C, G;
    int npix, S, items, V, PH, nst, PB1, PB2;
    uint32_t stage_bytes, aux_off;      // aux of a stage: [C halves pixel-0 row][C floats chan_add][G float2 (mean, rstd)]
    float eps;
    long long* trace;       // debugging aid (DADD_GN_TRACE): [4 roles][TRACE_ITEMS][8 events] clock64() of CTA 0, else nullptr
};
constexpr int TRACE_ITEMS = 16;
#define GN_EVENT(role, it, ev)                                                                                   \
    do {                                                                                                          \
        if (p.trace && blockIdx.x == 0 && lane == 0 && (it) < TRACE_ITEMS) p.trace[((role) * TRACE_ITEMS + (it)) * 8 + (ev)] = clock64(); \
    } while (0)

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* b, uint32_t n) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(s32(b)), "r"(n)); }
__device__ __forceinline__ void mbar_arrive(uint64_t* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(s32(b)) : "memory"); }
__device__ __forceinline__ void mbar_expect(uint64_t* b, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* b, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(s32(b)), "r"(parity) : "memory");
        if (done) return;
        if (++spins > SPIN_LIMIT) __trap();
    }
}
__device__ __forceinline__ void bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(s32(bar)) : "memory");
}
__device__ __forceinline__ void tma_box_3d(uint32_t dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(dst), "l"((uint64_t)map), "r"(s32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, uint32_t src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];"
                 ::"l"((uint64_t)map), "r"(src), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void named_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
__device__ __forceinline__ uint32_t ld_relaxed(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release(uint32_t* p, uint32_t v) { asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

template <typename T>
__device__ __forceinline__ uint32_t sub2(uint32_t a, uint32_t b) {
    uint32_t d;
    if constexpr (std::is_same_v<T, __nv_bfloat16>) asm("sub.rn.bf16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    else asm("sub.rn.f16x2 %0, %1, %2;" : "=r"(d) : "r"(a), "r"(b));
    return d;
}
template <typename T>
__device__ __forceinline__ void mma_stats(float (&d)[4], uint32_t lo, uint32_t hi) {
    constexpr uint32_t ONES = std::is_same_v<T, __nv_bfloat16> ? 0x3F803F80u : 0x3C003C00u;
    if constexpr (std::is_same_v<T, __nv_bfloat16>)
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(lo), "r"(ONES), "r"(hi), "r"(ONES), "r"(lo), "r"(hi));
    else
        asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                     : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(lo), "r"(ONES), "r"(hi), "r"(ONES), "r"(lo), "r"(hi));
}
__device__ __forceinline__ float tanh_fast(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// act: 0 identity, 1 SiLU as o / (1 + exp(-o)), 2 SiLU as h + h tanh(h) with scale / shift already halved (one MUFU op)
template <typename T, int ACT>
__device__ __forceinline__ void emit(const uint4 raw, const float (&sa)[8], const float (&sb)[8], unsigned char* dst) {
    Vec8<T> t;
    t.raw = raw;
    float f[8];
    t.unpack(f);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const float h = fmaf(f[i], sa[i], sb[i]);
        f[i] = ACT == 2 ? fmaf(h, tanh_fast(h), h) : (ACT == 1 ? silu(h) : h);
    }
    t.pack(f);
    *reinterpret_cast<uint4*>(dst) = t.raw;
}

// Where channel c (a multiple of 8) of the concatenated input lives inside a stage: byte offset of its 16-byte chunk at pixel 0,
// the 128-byte-row stride per pixel (PB rows) and the row index at pixel 0 (the swizzle XORs the chunk with row % 8).
struct ColAddr {
    uint32_t base;      // superpanel base (bytes from the stage start)
    uint32_t row0;      // 128-byte row of pixel 0: pb
    uint32_t rpp;       // rows per pixel: PB
    uint32_t chunk;     // logical 16-byte chunk inside the row
    __device__ __forceinline__ uint32_t at(uint32_t px) const {
        const uint32_t row = px * rpp + row0;
        return base + row * 128u + ((chunk ^ (row & 7u)) << 4);
    }
};
__device__ __forceinline__ ColAddr col_addr(int c, int C1, int npix, int PB1, int PB2) {
    const bool sec = c >= C1;
    const int cl = sec ? c - C1 : c, PB = sec ? PB2 : PB1;
    const int P = cl >> 6;
    ColAddr a;
    a.base = (sec ? (uint32_t)npix * (uint32_t)C1 * 2u : 0u) + (uint32_t)(P / PB) * (uint32_t)(npix * PB) * 128u;
    a.row0 = (uint32_t)(P % PB);
    a.rpp = (uint32_t)PB;
    a.chunk = (uint32_t)(cl & 63) >> 3;
    return a;
}

template <typename T, int ACT, int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB)
gn_stream_kernel(const __grid_constant__ CUtensorMap m1, const __grid_constant__ CUtensorMap m2, const __grid_constant__ CUtensorMap my1,
                 const __grid_constant__ CUtensorMap my2, const Params p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* ring = smem_raw + ((1024u - (s32(smem_raw) & 1023u)) & 1023u);      // (array arithmetic: accesses stay LDS / STS)
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int ncons = p.V * p.PH, ncw = ncons >> 5;
    const int C = p.C, C1 = p.C1, C2 = C - C1, G = p.G, cpg = C / G, S = p.S, npix = p.npix;
    float2* chs = reinterpret_cast<float2*>(ring + (size_t)p.nst * p.stage_bytes);      // [C] per-channel {sum d, sum d^2} of the slab
    uint64_t* full = reinterpret_cast<uint64_t*>(chs + C);
    uint64_t* empty = full + MAX_NST;
    uint64_t* ready = empty + MAX_NST;
    uint64_t* applied = ready + MAX_NST;
    uint64_t* chs_free = applied + MAX_NST;
    uint32_t* misc = reinterpret_cast<uint32_t*>(chs_free + 1);

    const int grid = (int)gridDim.x;
    const int n_my = (p.items - (int)blockIdx.x + grid - 1) / grid;
    const T* x1 = static_cast<const T*>(p.x);
    const T* x2 = static_cast<const T*>(p.x2);
    auto stage_ptr = [&](int st) { return ring + (size_t)st * p.stage_bytes; };
    auto aux_px0 = [&](int st) { return reinterpret_cast<const T*>(stage_ptr(st) + p.aux_off); };
    auto aux_add = [&](int st) { return reinterpret_cast<const float*>(stage_ptr(st) + p.aux_off + (size_t)C * sizeof(T)); };
    auto aux_grp = [&](int st) { return reinterpret_cast<float2*>(stage_ptr(st) + p.aux_off + (size_t)C * (sizeof(T) + sizeof(float))); };

    if (tid == 0) {
        for (int s = 0; s < p.nst; ++s) {
            mbar_init(&full[s], 1);
            mbar_init(&empty[s], 1);
            mbar_init(&ready[s], 1);
            mbar_init(&applied[s], (uint32_t)ncw);
        }
        mbar_init(chs_free, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        misc[0] = *reinterpret_cast<volatile uint32_t*>(p.ctl) + 1u;      // this launch's generation
    }
    __syncthreads();
    const uint32_t gen = misc[0];
    const float inv_n = __fdividef(1.0f, (float)cpg * (float)p.HW);
    // (mean, rstd) of a group from its shifted sums around M = kx + chan_add[first channel]
    auto finalise = [&](float t1, float t2, float shift) -> float2 {
        const float m = t1 * inv_n;
        const float var = fmaxf(t2 - t1 * m, 0.0f) * inv_n;
        return make_float2(shift + m, rsqrtf(var + p.eps));
    };
    // group g of the slab in stage st: shifted sums from the per-channel sums
    auto group_partial = [&](int st, int g) -> float2 {
        const float* add = aux_add(st);
        const int cg = g * cpg;
        const float a_first = p.chan_add ? add[cg] : 0.0f;
        float t1 = 0.0f, t2 = 0.0f;
        for (int i = 0; i < cpg; ++i) {
            const float2 sc = chs[cg + i];
            const float e = p.chan_add ? add[cg + i] - a_first : 0.0f;
            const float ne = (float)npix * e;
            t1 += sc.x + ne;
            t2 += sc.y + e * (2.0f * sc.x + ne);
        }
        return make_float2(t1, t2);
    };
    auto group_shift = [&](int st, int g) -> float {
        const int cg = g * cpg;
        return to_f(aux_px0(st)[cg]) + (p.chan_add ? aux_add(st)[cg] : 0.0f);
    };

    if (warp == ncw) {
        // ------------------------------------------------------------------ producer: slab ring
        if (lane == 0) {
            const uint32_t slab_bytes = (uint32_t)npix * (uint32_t)C * sizeof(T);
            const uint32_t aux_bytes = (uint32_t)C * sizeof(T) + (p.chan_add ? (uint32_t)C * sizeof(float) : 0u);
            for (int k = 0; k < n_my; ++k) {
                const int item = (int)blockIdx.x + k * grid, b = item / S, r = item - b * S;
                const int st = k % p.nst;
                GN_EVENT(3, k, 0);
                mbar_wait(&empty[st], (((uint32_t)k / (uint32_t)p.nst) & 1u) ^ 1u);
                GN_EVENT(3, k, 1);
                mbar_expect(&full[st], slab_bytes + aux_bytes);
                unsigned char* dst = stage_ptr(st);
                const int row = b * p.HW + r * npix;
                const uint32_t sp1 = (uint32_t)(npix * p.PB1) * 128u;
                for (int q = 0; q < (C1 >> 6) / p.PB1; ++q) tma_box_3d(s32(dst + (size_t)q * sp1), &m1, &full[st], 0, q * p.PB1, row);
                if (x2) {
                    unsigned char* dst2 = dst + (size_t)npix * C1 * sizeof(T);
                    const uint32_t sp2 = (uint32_t)(npix * p.PB2) * 128u;
                    for (int q = 0; q < (C2 >> 6) / p.PB2; ++q) tma_box_3d(s32(dst2 + (size_t)q * sp2), &m2, &full[st], 0, q * p.PB2, row);
                }
                unsigned char* aux = dst + p.aux_off;
                bulk_copy(s32(aux), x1 + (size_t)b * p.HW * C1, (uint32_t)C1 * sizeof(T), &full[st]);
                if (x2) bulk_copy(s32(aux + (size_t)C1 * sizeof(T)), x2 + (size_t)b * p.HW * C2, (uint32_t)C2 * sizeof(T), &full[st]);
                if (p.chan_add) bulk_copy(s32(aux + (size_t)C * sizeof(T)), p.chan_add + (size_t)b * p.add_stride, (uint32_t)C * sizeof(float), &full[st]);
            }
        }
    } else if (warp == ncw + 1) {
        // ------------------------------------------------------------------ publisher: group partials of item k -> global, release flag
        if (S > 1) {
            for (int k = 0; k < n_my; ++k) {
                const int item = (int)blockIdx.x + k * grid;
                const int st = k % p.nst;
                named_sync(1, ncons + 32);                                   // chs[] of item k is complete
                GN_EVENT(1, k, 0);
                for (int g = lane; g < G; g += 32) p.part[(size_t)item * G + g] = group_partial(st, g);
                __syncwarp();
                GN_EVENT(1, k, 1);
                if (lane == 0) {
                    mbar_arrive(chs_free);
                    GN_EVENT(1, k, 2);
                    st_release(p.flags + item, gen);      // (the warp barrier above ordered the other lanes' partials before it)
                    GN_EVENT(1, k, 3);
                }
            }
        }
    } else if (warp == ncw + 2) {
        // ------------------------------------------------------------------ combiner: partials of item j's sample -> (mean, rstd)
        if (S > 1) {
            for (int j = 0; j < n_my; ++j) {
                const int item = (int)blockIdx.x + j * grid, bj = item / S;
                const int st = j % p.nst;
                uint32_t spins = 0;
                GN_EVENT(2, j, 0);
                while (true) {
                    bool ok = true;
                    for (int r = lane; r < S; r += 32) ok = ok && (ld_relaxed(p.flags + (size_t)bj * S + r) == gen);
                    if (__all_sync(0xffffffffu, ok)) break;
                    if (++spins > SPIN_LIMIT) __trap();
                }
                __threadfence();                                              // acquire: the partials behind the flags
                GN_EVENT(2, j, 1);
                // lane g adds group g's partials in rank order (coalesced: a rank's G partials are contiguous), 16 loads in flight
                float t1[2] = {0.0f, 0.0f}, t2[2] = {0.0f, 0.0f};             // groups lane, lane + 32
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int g = lane + 32 * h;
                    if (g < G) {
                        const float2* src = p.part + (size_t)bj * S * G + g;
                        for (int r0 = 0; r0 < S; r0 += 16) {
                            float2 w[16];
#pragma unroll
                            for (int r = 0; r < 16; ++r) w[r] = (r0 + r < S) ? __ldcg(src + (size_t)(r0 + r) * G) : make_float2(0.0f, 0.0f);
#pragma unroll
                            for (int r = 0; r < 16; ++r) {
                                t1[h] += w[r].x;
                                t2[h] += w[r].y;
                            }
                        }
                    }
                }
                GN_EVENT(2, j, 2);
                mbar_wait(&full[st], ((uint32_t)j / (uint32_t)p.nst) & 1u);   // the stage's aux rows (long there: stats(j) ran before the flags)
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    const int g = lane + 32 * h;
                    if (g < G) aux_grp(st)[g] = finalise(t1[h], t2[h], group_shift(st, g));
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&ready[st]);
                GN_EVENT(2, j, 3);
            }
        }
    } else if (warp == ncw + 3) {
        // ------------------------------------------------------------------ store thread: normalised slabs leave by TMA, then the stage is free
        if (lane == 0) {
            for (int j = 0; j < n_my; ++j) {
                const int item = (int)blockIdx.x + j * grid, b = item / S, r = item - b * S;
                const int st = j % p.nst;
                mbar_wait(&applied[st], ((uint32_t)j / (uint32_t)p.nst) & 1u);
                const unsigned char* src = stage_ptr(st);
                const int row = b * p.HW + r * npix;
                const uint32_t sp1 = (uint32_t)(npix * p.PB1) * 128u;
                for (int q = 0; q < (C1 >> 6) / p.PB1; ++q) tma_store_3d(&my1, s32(src + (size_t)q * sp1), 0, q * p.PB1, row);
                if (x2) {
                    const unsigned char* src2 = src + (size_t)npix * C1 * sizeof(T);
                    const uint32_t sp2 = (uint32_t)(npix * p.PB2) * 128u;
                    for (int q = 0; q < (C2 >> 6) / p.PB2; ++q) tma_store_3d(&my2, s32(src2 + (size_t)q * sp2), 0, (C1 >> 6) + q * p.PB2, row);
                }
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");      // the slab has left shared memory
                mbar_arrive(&empty[st]);
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    } else if (warp < ncw) {
        // ------------------------------------------------------------------ consumers
        const int v = tid % p.V, ph = tid / p.V, c0 = v << 3;
        const ColAddr col = col_addr(c0, C1, npix, p.PB1, p.PB2);
        float gm[8], bt[8];
        {
            const float4 a0 = *reinterpret_cast<const float4*>(p.gamma + c0), a1 = *reinterpret_cast<const float4*>(p.gamma + c0 + 4);
            const float4 b0 = *reinterpret_cast<const float4*>(p.beta + c0), b1 = *reinterpret_cast<const float4*>(p.beta + c0 + 4);
            gm[0] = a0.x; gm[1] = a0.y; gm[2] = a0.z; gm[3] = a0.w; gm[4] = a1.x; gm[5] = a1.y; gm[6] = a1.z; gm[7] = a1.w;
            bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
        }
        const int g_first = c0 / cpg;
        const int split = min(8, (g_first + 1) * cpg - c0);      // channels [0, split) of my column are in group g_first

        // statistics of the slab in stage st -> chs[c].  Unit = 16 channels x all pixels of the slab; a warp's units are the same for
        // every item, so their addressing is worked out once (two units per warp cover C <= 32 * warps; wider inputs recompute).
        struct Unit {
            uint32_t off, step, ia, ib;      // lane's ldmatrix row offset in a stage, bytes per 16 pixels, pixel-0-row index of the two shifts
        };
        auto unit_of = [&](int u) -> Unit {
            const int ch0 = u << 4, gq = lane >> 2;
            const ColAddr a = col_addr(ch0 + ((lane >> 4) << 3), C1, npix, p.PB1, p.PB2);
            Unit un;
            // ldmatrix row of this lane: matrix lane / 8 = {pixels 0-7 | 8-15} x {channels 0-7 | 8-15}
            un.off = a.at((uint32_t)((lane & 7) + ((lane >> 3) & 1) * 8));
            un.step = 16u * a.rpp * 128u;                                    // 16 pixels on: row % 8 (the swizzle term) is unchanged
            un.ia = (uint32_t)(((ch0 + gq) / cpg) * cpg);
            un.ib = (uint32_t)(((ch0 + 8 + gq) / cpg) * cpg);
            return un;
        };
        const int nunits = C >> 4;
        const bool hoisted = nunits <= 2 * ncw;
        Unit u0 = unit_of(warp < nunits ? warp : 0), u1 = unit_of(warp + ncw < nunits ? warp + ncw : 0);
        auto stats_unit = [&](int st, int u, const Unit& un) {
            const int gq = lane >> 2, tq = lane & 3, ch0 = u << 4;
            const uint16_t* px0 = reinterpret_cast<const uint16_t*>(aux_px0(st));
            const uint32_t ka = (uint32_t)px0[un.ia] * 0x10001u, kb = (uint32_t)px0[un.ib] * 0x10001u;
            uint32_t addr = s32(stage_ptr(st)) + un.off;
            float d0[4] = {0.f, 0.f, 0.f, 0.f}, d1[4] = {0.f, 0.f, 0.f, 0.f}, e0[4] = {0.f, 0.f, 0.f, 0.f}, e1[4] = {0.f, 0.f, 0.f, 0.f};
            int pt = 0;
            for (; pt + 16 < npix; pt += 32, addr += 2u * un.step) {          // two independent accumulator chains per channel half
                uint32_t r0, r1, r2, r3, q0, q1, q2, q3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(q0), "=r"(q1), "=r"(q2), "=r"(q3) : "r"(addr + un.step) : "memory");
                mma_stats<T>(d0, sub2<T>(r0, ka), sub2<T>(r1, ka));
                mma_stats<T>(d1, sub2<T>(r2, kb), sub2<T>(r3, kb));
                mma_stats<T>(e0, sub2<T>(q0, ka), sub2<T>(q1, ka));
                mma_stats<T>(e1, sub2<T>(q2, kb), sub2<T>(q3, kb));
            }
            if (pt < npix) {
                uint32_t r0, r1, r2, r3;
                asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
                             : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(addr) : "memory");
                mma_stats<T>(d0, sub2<T>(r0, ka), sub2<T>(r1, ka));
                mma_stats<T>(d1, sub2<T>(r2, kb), sub2<T>(r3, kb));
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) { d0[i] += e0[i]; d1[i] += e1[i]; }
            // rows 8-15 of D: column sums (every row equal; lanes 0-3 write); rows 0-7: Gram, the diagonal is sum d^2
            if (gq == 0) {
                chs[ch0 + 2 * tq].x = d0[2];
                chs[ch0 + 2 * tq + 1].x = d0[3];
                chs[ch0 + 8 + 2 * tq].x = d1[2];
                chs[ch0 + 8 + 2 * tq + 1].x = d1[3];
            }
            if (tq == (gq >> 1)) {
                chs[ch0 + gq].y = (gq & 1) ? d0[1] : d0[0];
                chs[ch0 + 8 + gq].y = (gq & 1) ? d1[1] : d1[0];
            }
        };
        auto stats = [&](int st) {
            if (hoisted) {
                if (warp < nunits) stats_unit(st, warp, u0);
                if (warp + ncw < nunits) stats_unit(st, warp + ncw, u1);
            } else {
                for (int u = warp; u < nunits; u += ncw) stats_unit(st, u, unit_of(u));
            }
        };
        // normalise the slab of item j (ring stage st) with the statistics in the stage's aux
        auto apply = [&](int st) {
            float sa[8], sb[8];
            {
                const float* add = aux_add(st);
                const float2 sA = aux_grp(st)[g_first], sB = aux_grp(st)[min(g_first + 1, G - 1)];
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float2 stt = i < split ? sA : sB;
                    sa[i] = stt.y * gm[i];
                    sb[i] = bt[i] + ((p.chan_add ? add[c0 + i] : 0.0f) - stt.x) * sa[i];
                    if (ACT == 2) { sa[i] *= 0.5f; sb[i] *= 0.5f; }
                }
            }
            unsigned char* slab = stage_ptr(st);                              // normalised in place; the store thread sends the slab out
            int px = ph;
            const int PH = p.PH;
            for (; px + 3 * PH < npix; px += 4 * PH) {                        // four loads in flight per thread
                uint4 q[4];
                uint32_t o[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    o[i] = col.at((uint32_t)(px + i * PH));
                    q[i] = *reinterpret_cast<const uint4*>(slab + o[i]);
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) emit<T, ACT>(q[i], sa, sb, slab + o[i]);
            }
            {
                uint4 q[3];
                uint32_t o[3];
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (px + i * PH < npix) {
                        o[i] = col.at((uint32_t)(px + i * PH));
                        q[i] = *reinterpret_cast<const uint4*>(slab + o[i]);
                    }
#pragma unroll
                for (int i = 0; i < 3; ++i)
                    if (px + i * PH < npix) emit<T, ACT>(q[i], sa, sb, slab + o[i]);
            }
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // my generic-proxy writes, before the TMA store reads them
            __syncwarp();
            if (lane == 0) mbar_arrive(&applied[st]);
        };

        if (S == 1) {
            for (int k = 0; k < n_my; ++k) {
                const int st = k % p.nst;
                mbar_wait(&full[st], ((uint32_t)k / (uint32_t)p.nst) & 1u);
                stats(st);
                named_sync(1, ncons);
                if (tid < G) {
                    const float2 t = group_partial(st, tid);
                    aux_grp(st)[tid] = finalise(t.x, t.y, group_shift(st, tid));
                }
                named_sync(1, ncons);
                apply(st);
            }
        } else {
            // item k's statistics are published LA items before it is normalised (LA = 2 with five ring stages: loading, statistics,
            // two waiting for their sample's partials, storing), so the publish -> combine round trip through L2 is off the path
            const int LA = p.nst >= 5 ? 2 : 1;
            for (int k = 0; k < n_my; ++k) {
                const int st = k % p.nst;
                if (warp == 0) GN_EVENT(0, k, 0);
                mbar_wait(&full[st], ((uint32_t)k / (uint32_t)p.nst) & 1u);
                if (k > 0) mbar_wait(chs_free, (uint32_t)(k - 1) & 1u);      // the publisher has read chs[] of item k-1
                if (warp == 0) GN_EVENT(0, k, 1);
                stats(st);
                if (warp == 0) GN_EVENT(0, k, 2);
                named_sync(1, ncons + 32);
                if (warp == 0) GN_EVENT(0, k, 3);
                if (k >= LA) {
                    const int j = k - LA, sj = j % p.nst;
                    mbar_wait(&ready[sj], ((uint32_t)j / (uint32_t)p.nst) & 1u);
                    if (warp == 0) GN_EVENT(0, k, 4);
                    apply(sj);
                    if (warp == 0) GN_EVENT(0, k, 5);
                }
            }
            for (int j = max(0, n_my - LA); j < n_my; ++j) {
                const int sj = j % p.nst;
                mbar_wait(&ready[sj], ((uint32_t)j / (uint32_t)p.nst) & 1u);
                apply(sj);
            }
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        const uint32_t done = atomicAdd(p.ctl + 1, 1u);
        if (done == (uint32_t)grid - 1u) {       // last CTA out: the next launch sees a new generation
            p.ctl[1] = 0u;
            __threadfence();
            *reinterpret_cast<volatile uint32_t*>(p.ctl) = gen;
        }
    }
}

struct Plan {
    bool ok;
    int npix, S, items, V, PH, nst, grid, occ, PB1, PB2;
    uint32_t stage_bytes, aux_off;
    size_t smem;
};

static int odd_panel_block(int panels) {      // largest odd divisor of `panels` that is at most 15 (a TMA box holds <= 256 per dimension)
    int best = 1;
    for (int d = 1; d <= 15; d += 2)
        if (panels % d == 0) best = d;
    return best;
}

static Plan make_plan(int B, int C, int C1, int HW, int G, bool two, bool has_add) {
    Plan pl{};
    const int C2 = C - C1;
    if (C % 64 != 0 || HW % 16 != 0 || G > MAX_G || G % 2 != 0 || C % G != 0) return pl;
    if (two && (C1 % 64 != 0 || C2 % 64 != 0)) return pl;
    static const int occ_env = [] { const char* e = getenv("DADD_GN_OCC"); return e ? atoi(e) : 0; }();
    static const int slab_kb = [] { const char* e = getenv("DADD_GN_SLAB_KB"); return e ? atoi(e) : 0; }();
    pl.occ = occ_env == 2 ? 2 : 1;      // one CTA per SM: four ring stages (loading, statistics, normalise, storing)
    pl.V = C >> 3;
    if (pl.V > 512) return pl;
    if (pl.V > 320) pl.occ = 1;
    pl.PB1 = odd_panel_block(C1 >> 6);
    pl.PB2 = two ? odd_panel_block(C2 >> 6) : 1;
    const uint32_t row = (uint32_t)C * 2;
    (void)has_add;
    const uint32_t aux = (uint32_t)C * 6u + MAX_G * (uint32_t)sizeof(float2);      // pixel-0 row, chan_add row, (mean, rstd)
    for (;; pl.occ = 1) {
        const int cons_cap = pl.occ == 2 ? 320 : 512;
        const uint32_t ring_budget = pl.occ == 2 ? 100u * 1024u : 222u * 1024u - (uint32_t)C * 8u;
        const uint32_t slab_cap = slab_kb > 0 ? (uint32_t)slab_kb * 1024u : (pl.occ == 2 ? 42u * 1024u : 64u * 1024u);
        const int slots = num_sms() * pl.occ;
        int npix = 16;
        while (npix * 2 <= HW && HW % (npix * 2) == 0 && npix * 2 <= 256 && (uint32_t)(npix * 2) * row <= slab_cap) npix *= 2;
        auto eff = [&](int np) {
            const int64_t items = (int64_t)B * (HW / np);
            const int64_t rounds = (items + slots - 1) / slots;
            return (double)items / (double)(rounds * slots);
        };
        while (npix > 16 && eff(npix) < 0.85 && eff(npix / 2) > eff(npix)) npix /= 2;
        pl.npix = npix;
        pl.S = HW / npix;
        pl.PH = cons_cap / pl.V;
        if (pl.PH > npix) pl.PH = npix;
        while (pl.PH > 1 && (pl.V * pl.PH) % 32 != 0) --pl.PH;
        pl.aux_off = (uint32_t)npix * row;
        pl.stage_bytes = (pl.aux_off + aux + 1023u) & ~1023u;
        pl.nst = (int)(ring_budget / pl.stage_bytes);
        if (pl.nst > MAX_NST) pl.nst = MAX_NST;
        pl.grid = slots;
        // slabs too large for two CTAs per SM, or no whole number of warps: one CTA with the whole shared memory
        if ((pl.nst >= 2 && (pl.V * pl.PH) % 32 == 0) || pl.occ == 1) break;
    }
    if (pl.nst < 2 || (pl.V * pl.PH) % 32 != 0) return pl;
    if (pl.S > MAX_S || pl.S > num_sms()) return pl;
    const int64_t items = (int64_t)B * pl.S;
    if (items > MAX_ITEMS) return pl;
    pl.items = (int)items;
    if (pl.items < pl.grid) pl.grid = pl.items;
    pl.smem = 1024 + (size_t)pl.nst * pl.stage_bytes + (size_t)C * sizeof(float2) + (4 * MAX_NST + 1) * 8 + 64;
    if (pl.smem > (pl.occ == 2 ? 113u : 227u) * 1024u) return pl;
    pl.ok = true;
    return pl;
}

// Flags and the generation word are owned by the library (zeroed once, never reset).  Every stream handle gets its own control
// block, so launches on different streams never share a generation counter; replays of a captured graph use the block of the
// stream they were captured on.  (What must stay stream-ordered: launches - or graph replays - that share one block, i.e. one
// capture / launch stream.  The engines of this package run one denoising loop per device at a time.)  Blocks come from a pool
// that is allocated outside stream capture (the first eager call of a device; the engines run eager steps before they capture)
// and handed out without further allocation; a capture that finds the pool empty takes the older kernels for that call.
struct Ctl {
    uint32_t* flags;      // [MAX_ITEMS]
    uint32_t* ctl;        // [2]
};
static int get_ctl(cudaStream_t s, Ctl* out) {
    constexpr size_t SLOT_BYTES = (size_t)MAX_ITEMS * sizeof(uint32_t) + 256;
    constexpr int CHUNK = 16, MAX_SLOTS = 256;
    struct Dev {
        unsigned char* slot[MAX_SLOTS] = {};
        cudaStream_t owner[MAX_SLOTS] = {};
        int allocated = 0, used = 0;
    };
    static std::mutex mu;
    static Dev devs[64];
    int dev = 0;
    if (cuda_ok(cudaGetDevice(&dev), "dadd_groupnorm_fwd(stream) device")) return 2;
    if (dev < 0 || dev >= 64) return 3;
    std::lock_guard<std::mutex> lk(mu);
    Dev& d = devs[dev];
    int slot = -1;
    for (int i = 0; i < d.used && slot < 0; ++i)
        if (d.owner[i] == s) slot = i;
    if (slot < 0) {
        if (d.used == d.allocated) {
            cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
            if (cudaStreamIsCapturing(s, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone || d.allocated + CHUNK > MAX_SLOTS) {
                cudaGetLastError();
                return 3;
            }
            void* ptr = nullptr;
            if (cuda_ok(cudaMalloc(&ptr, SLOT_BYTES * CHUNK), "dadd_groupnorm_fwd(stream) control blocks")) return 2;
            if (cuda_ok(cudaMemset(ptr, 0, SLOT_BYTES * CHUNK), "dadd_groupnorm_fwd(stream) control blocks")) return 2;
            for (int i = 0; i < CHUNK; ++i) d.slot[d.allocated + i] = static_cast<unsigned char*>(ptr) + (size_t)i * SLOT_BYTES;
            d.allocated += CHUNK;
        }
        slot = d.used++;
        d.owner[slot] = s;
    }
    out->ctl = reinterpret_cast<uint32_t*>(d.slot[slot]);
    out->flags = reinterpret_cast<uint32_t*>(d.slot[slot] + 256);
    return 0;
}

// 3-D view {64 channels, C / 64 panels, rows} of an NHWC tensor with C channels; boxes {64, PB, npix}, 128-byte swizzle
static int make_map(CUtensorMap* map, const void* base, int64_t rows, int C, int PB, int npix, bool bf16) {
    tc::EncodeTiledFn fn = tc::encode_fn();
    if (!fn) return 1;
    const cuuint64_t dims[3] = {64, (cuuint64_t)(C >> 6), (cuuint64_t)rows};
    const cuuint64_t strides[2] = {128, (cuuint64_t)C * 2};
    const cuuint32_t box[3] = {64, (cuuint32_t)PB, (cuuint32_t)npix};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = fn(map, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 3, const_cast<void*>(base), dims,
                          strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? 0 : 1;
}

}  // namespace gns

int64_t gn_stream_workspace_bytes(int B, int C, int HW, int G) {
    // partial sums for the smallest slab the planner may pick (16 pixels)
    return (int64_t)B * ((HW + 15) / 16) * G * (int64_t)sizeof(float2);
}

// Returns -1 when the shape (or the control-block pool) is not served by this kernel - the caller then takes another path -
// 0 on success, > 0 on error.
template <typename T>
int gn_stream_launch(const T* x, const T* x2, int C1, const float* gamma, const float* beta, const float* chan_add, int64_t add_stride,
                     T* y, int B, int C, int HW, int G, float eps, int act, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    using namespace gns;
    const Plan pl = make_plan(B, C, x2 ? C1 : C, HW, G, x2 != nullptr, chan_add != nullptr);
    if (!pl.ok) return -1;
    if ((int64_t)pl.items * G * (int64_t)sizeof(float2) > workspace_bytes || !workspace) return -1;
    if (((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(y) | reinterpret_cast<uintptr_t>(workspace)) & 15) != 0) return -1;
    if (chan_add && ((reinterpret_cast<uintptr_t>(chan_add) & 15) != 0 || (add_stride & 3) != 0)) return -1;
    if ((int64_t)B * HW > 0x7fffffffll) return -1;
    constexpr bool bf16 = std::is_same_v<T, __nv_bfloat16>;
    const int Ca = x2 ? C1 : C;
    CUtensorMap m1, m2;
    if (make_map(&m1, x, (int64_t)B * HW, Ca, pl.PB1, pl.npix, bf16)) return -1;
    if (x2) {
        if (make_map(&m2, x2, (int64_t)B * HW, C - C1, pl.PB2, pl.npix, bf16)) return -1;
    } else {
        m2 = m1;
    }
    Ctl ctl;
    const int rc = get_ctl(s, &ctl);
    if (rc == 3) return -1;
    if (rc) return rc;
    Params p{};
    p.x = x; p.x2 = x2; p.gamma = gamma; p.beta = beta; p.chan_add = chan_add; p.y = y;
    p.part = static_cast<float2*>(workspace); p.flags = ctl.flags; p.ctl = ctl.ctl;
    p.add_stride = add_stride;
    p.C1 = Ca; p.HW = HW; p.C = C; p.G = G;
    p.npix = pl.npix; p.S = pl.S; p.items = pl.items; p.V = pl.V; p.PH = pl.PH; p.nst = pl.nst; p.PB1 = pl.PB1; p.PB2 = pl.PB2;
    p.stage_bytes = pl.stage_bytes; p.aux_off = pl.aux_off; p.eps = eps;
    CUtensorMap my1, my2;
    if (make_map(&my1, y, (int64_t)B * HW, C, pl.PB1, pl.npix, bf16)) return -1;
    if (make_map(&my2, y, (int64_t)B * HW, C, pl.PB2, pl.npix, bf16)) return -1;
    const int threads = pl.V * pl.PH + 128;
    static const char* trace_path = getenv("DADD_GN_TRACE");      // debugging aid: synchronises, never on a product path
    auto go = [&](auto kern) -> int {
        if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem), "gn stream smem")) return 2;
        if (trace_path) {
            const size_t n = (size_t)4 * TRACE_ITEMS * 8;
            if (cuda_ok(cudaMalloc(&p.trace, n * sizeof(long long)), "gn trace")) return 2;
            cudaMemsetAsync(p.trace, 0, n * sizeof(long long), s);
            kern<<<pl.grid, threads, pl.smem, s>>>(m1, m2, my1, my2, p);
            cudaStreamSynchronize(s);
            long long* host = new long long[n];
            cudaMemcpy(host, p.trace, n * sizeof(long long), cudaMemcpyDeviceToHost);
            if (FILE* f = fopen(trace_path, "w")) {
                fprintf(f, "# B=%d C=%d HW=%d npix=%d S=%d items=%d grid=%d threads=%d nst=%d occ=%d stage=%u smem=%zu\n", B, C, HW, pl.npix, pl.S,
                        pl.items, pl.grid, threads, pl.nst, pl.occ, pl.stage_bytes, pl.smem);
                for (size_t i = 0; i < n; ++i) fprintf(f, "%lld%c", host[i], (i % 8 == 7) ? '\n' : ' ');
                fclose(f);
            }
            delete[] host;
            cudaFree(p.trace);
            return launched("dadd_groupnorm_fwd(NHWC stream trace)");
        }
        kern<<<pl.grid, threads, pl.smem, s>>>(m1, m2, my1, my2, p);
        return launched("dadd_groupnorm_fwd(NHWC stream)");
    };
    if (pl.occ == 2) {
        if (act == 2) return go(gn_stream_kernel<T, 2, 448, 2>);
        if (act == 1) return go(gn_stream_kernel<T, 1, 448, 2>);
        return go(gn_stream_kernel<T, 0, 448, 2>);
    }
    if (act == 2) return go(gn_stream_kernel<T, 2, 640, 1>);
    if (act == 1) return go(gn_stream_kernel<T, 1, 640, 1>);
    return go(gn_stream_kernel<T, 0, 640, 1>);
}

template int gn_stream_launch<__half>(const __half*, const __half*, int, const float*, const float*, const float*, int64_t, __half*, int, int,
                                      int, int, float, int, void*, int64_t, cudaStream_t);
template int gn_stream_launch<__nv_bfloat16>(const __nv_bfloat16*, const __nv_bfloat16*, int, const float*, const float*, const float*, int64_t,
                                             __nv_bfloat16*, int, int, int, int, float, int, void*, int64_t, cudaStream_t);

}  // namespace daddk
