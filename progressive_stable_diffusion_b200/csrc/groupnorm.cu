// K4/K4b: GroupNorm (+ per-(sample,channel) additive term) (+ SiLU), fp32 statistics, one HBM read + one write.
// Reference ops: diffusers ResnetBlock2D.norm1/norm2 + SiLU, conv_norm_out + conv_act, Transformer2DModel.norm
// (reached via src/models/unet/unet.py:140-146).
//
// NHWC (channels-last, the layout the B200 UNet runs in): flat, fully parallel passes (partial statistics, a tiny
// finalise, then normalise) whose grids fill all 148 SMs; the second pass re-reads the activation from the 126 MB L2, so
// HBM sees one read and one write.  Every thread keeps a fixed 8-channel column (coalesced 16 B accesses, per-channel scale/shift in registers).
// NCHW (the reference's layout): a group is contiguous, one CTA per (sample, group).
#include <cstdlib>

#include "common.cuh"

namespace daddk {

constexpr int GN_MAX_G = 64;
constexpr int GN_CACHE = 12;  // cached 8-element vectors per thread

struct Moments {
    float n, mean, m2;
};

__device__ __forceinline__ Moments chan_combine(Moments a, Moments b) {
    if (b.n == 0.0f) return a;
    if (a.n == 0.0f) return b;
    const float n = a.n + b.n;
    const float d = b.mean - a.mean;
    Moments r;
    r.n = n;
    r.mean = a.mean + d * (b.n / n);
    r.m2 = a.m2 + b.m2 + d * d * (a.n * b.n / n);
    return r;
}

// ------------------------------------------------------------------------------------------------ NHWC
// Three lean launches over x[b][pixel][channel], each with ONE global-memory round trip on its critical path:
//   gn_stats_nhwc:  CTA = (pixel chunk, sample); thread = fixed 8-channel column x pixel phase.  Shifted sums
//                   sum(v - K_g), sum((v - K_g)^2) with one shift per (sample, group) -- K_g = the group's first channel
//                   at pixel 0 -- so that the partials of different chunks simply add.  Out: part[b][chunk][g] (float2).
//   gn_final:       one CTA per sample: partials -> (mean, rstd) per group -> per-channel scale / shift
//                   coef[b][0][c] = rstd * gamma[c],  coef[b][1][c] = beta[c] + (chan_add[b][c] - mean) * rstd * gamma[c].
//   gn_apply_nhwc:  flat pass  y = act(x * scale + shift); x is re-read from L2 (a whole activation is <= 50 MB).
// No atomics anywhere: results are bit-reproducible run to run.
struct GnPlan {
    int V, PH, threads, npx, chunks;
};

static GnPlan gn_plan(int B, int C, int HW) {
    GnPlan p;
    p.V = C >> 3;
    p.PH = 512 / p.V;
    if (p.PH < 1) p.PH = 1;
    if (p.PH > HW) p.PH = HW;
    p.threads = p.V * p.PH;
    p.npx = p.PH * 16;                                  // up to 16 vectors (256 B) per thread ...
    if (p.npx > HW) p.npx = HW;
    p.chunks = (HW + p.npx - 1) / p.npx;
    const int want = num_sms() + num_sms() / 2;         // ... unless that leaves SMs idle
    while ((int64_t)p.chunks * B < want && p.npx > p.PH) {
        p.npx = (p.npx / 2 < p.PH) ? p.PH : p.npx / 2;
        p.chunks = (HW + p.npx - 1) / p.npx;
    }
    return p;
}

template <typename T>
__global__ void __launch_bounds__(512) gn_stats_nhwc_kernel(const T* __restrict__ x, const T* __restrict__ x2, int C1,
                                                            const float* __restrict__ chan_add,
                                                            int64_t add_stride, float2* __restrict__ part, int HW, int C, int G,
                                                            int npx) {
    extern __shared__ float smem[];      // [PH][C] x 2
    const int V = C >> 3, PH = blockDim.x / V;
    const int tid = threadIdx.x, v = tid % V, ph = tid / V, c0 = v << 3;
    const int b = blockIdx.y, chunk = blockIdx.x, cpg = C / G;
    // two-source form (x2 != nullptr): the input is the channel concatenation [x | x2] with C1 and C - C1 channels
    const int C2 = C - C1;
    const T* xs1 = x + (size_t)b * HW * C1;
    const T* xs2 = x2 ? x2 + (size_t)b * HW * C2 : nullptr;
    const float* addb = chan_add ? chan_add + (size_t)b * add_stride : nullptr;
    const int p0 = chunk * npx, p1 = min(HW, p0 + npx);
    const int sC = c0 >= C1 ? C2 : C1;                  // this thread's source and its row stride
    const T* xp = c0 >= C1 ? xs2 + (c0 - C1) : xs1 + c0;
    int p = p0 + ph;
    // first batch of loads goes out before anything else
    constexpr int U = 4;
    Vec8<T> t[U];
    const bool full0 = p + (U - 1) * PH < p1;
    if (full0) {
#pragma unroll
        for (int u = 0; u < U; ++u) t[u].load(xp + (size_t)(p + u * PH) * sC);
    }
    float sh[8];                                        // chan_add[c] - K_g(c),  K_g = x[b][0][g*cpg] + chan_add[g*cpg]
    {
        int g = c0 / cpg, r = c0 - g * cpg;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int cg = g * cpg;
            const float k = to_f(cg < C1 ? xs1[cg] : xs2[cg - C1]) + (addb ? addb[cg] : 0.0f);
            sh[i] = (addb ? addb[c0 + i] : 0.0f) - k;
            if (++r == cpg) { r = 0; ++g; }
        }
    }
    float a1[8], a2[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a1[i] = 0.0f; a2[i] = 0.0f; }
    auto acc = [&](const Vec8<T>& tv) {
        float f[8];
        tv.unpack(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = f[i] + sh[i]; a1[i] += d; a2[i] = fmaf(d, d, a2[i]); }
    };
    if (full0) {
#pragma unroll
        for (int u = 0; u < U; ++u) acc(t[u]);
        p += U * PH;
    }
    for (; p + (U - 1) * PH < p1; p += U * PH) {
#pragma unroll
        for (int u = 0; u < U; ++u) t[u].load(xp + (size_t)(p + u * PH) * sC);
#pragma unroll
        for (int u = 0; u < U; ++u) acc(t[u]);
    }
    for (; p < p1; p += PH) {
        t[0].load(xp + (size_t)p * sC);
        acc(t[0]);
    }
    float* s1 = smem;
    float* s2 = smem + (size_t)PH * C;
    *reinterpret_cast<float4*>(s1 + (size_t)ph * C + c0) = make_float4(a1[0], a1[1], a1[2], a1[3]);
    *reinterpret_cast<float4*>(s1 + (size_t)ph * C + c0 + 4) = make_float4(a1[4], a1[5], a1[6], a1[7]);
    *reinterpret_cast<float4*>(s2 + (size_t)ph * C + c0) = make_float4(a2[0], a2[1], a2[2], a2[3]);
    *reinterpret_cast<float4*>(s2 + (size_t)ph * C + c0 + 4) = make_float4(a2[4], a2[5], a2[6], a2[7]);
    __syncthreads();
    for (int c = tid; c < C; c += blockDim.x) {         // pixel phases -> one (sum, sumsq) per channel
        float t1 = 0.0f, t2 = 0.0f;
        for (int q = 0; q < PH; ++q) { t1 += s1[(size_t)q * C + c]; t2 += s2[(size_t)q * C + c]; }
        s1[c] = t1;
        s2[c] = t2;
    }
    __syncthreads();
    if (tid < G) {
        float t1 = 0.0f, t2 = 0.0f;
        for (int i = 0; i < cpg; ++i) { t1 += s1[tid * cpg + i]; t2 += s2[tid * cpg + i]; }
        part[((size_t)b * gridDim.x + chunk) * G + tid] = make_float2(t1, t2);
    }
}

// (mean, rstd) of group g of sample b from the chunk partials (shifted sums around K_g)
template <typename T>
__device__ __forceinline__ float2 gn_group_stats(const float2* __restrict__ part, const T* __restrict__ x, const T* __restrict__ x2, int C1,
                                                 const float* addb, int b, int g, int chunks, int HW, int C, int G, float eps) {
    const int cpg = C / G;
    float t1 = 0.0f, t2 = 0.0f;
    for (int k = 0; k < chunks; ++k) {
        const float2 w = part[((size_t)b * chunks + k) * G + g];
        t1 += w.x;
        t2 += w.y;
    }
    const float inv_n = __fdividef(1.0f, (float)cpg * (float)HW);
    const float m = t1 * inv_n;
    const float var = fmaxf(t2 - t1 * m, 0.0f) * inv_n;
    const int cg = g * cpg;
    const float first = to_f(cg < C1 ? x[(size_t)b * HW * C1 + cg] : x2[(size_t)b * HW * (C - C1) + cg - C1]);
    return make_float2(first + (addb ? addb[cg] : 0.0f) + m, rsqrtf(var + eps));
}

template <typename T>
__global__ void __launch_bounds__(256) gn_final_kernel(const float2* __restrict__ part, const T* __restrict__ x, const T* __restrict__ x2,
                                                       int C1, const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       const float* __restrict__ chan_add, int64_t add_stride,
                                                       float* __restrict__ coef, int chunks, int HW, int C, int G, float eps) {
    __shared__ float2 grp[GN_MAX_G];
    const int b = blockIdx.x, tid = threadIdx.x, cpg = C / G;
    const float* addb = chan_add ? chan_add + (size_t)b * add_stride : nullptr;
    if (tid < G) grp[tid] = gn_group_stats(part, x, x2, C1, addb, b, tid, chunks, HW, C, G, eps);
    __syncthreads();
    float* sc = coef + (size_t)b * 2 * C;
    for (int c = tid; c < C; c += blockDim.x) {
        const float2 st = grp[c / cpg];
        const float a = st.y * gamma[c];
        sc[c] = a;
        sc[C + c] = beta[c] + ((addb ? addb[c] : 0.0f) - st.x) * a;
    }
}

__device__ __forceinline__ float tanh_approx(float x) {
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// act: 0 = identity, 1 = SiLU as o / (1 + exp(-o)) (two MUFU ops per element), 2 = SiLU as h + h * tanh(h) with h = o / 2
// (one MUFU op; the caller passes scale / shift already halved, so the kernel computes h directly).
template <typename T>
__device__ __forceinline__ void gn_emit(Vec8<T>& t, const float (&sa)[8], const float (&sb)[8], int act, T* dst) {
    float f[8];
    t.unpack(f);
    if (act == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float h = fmaf(f[i], sa[i], sb[i]);
            f[i] = fmaf(h, tanh_approx(h), h);
        }
    } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const float o = fmaf(f[i], sa[i], sb[i]);
            f[i] = act ? silu(o) : o;
        }
    }
    t.pack(f);
    t.store(dst);
}

// coef != nullptr: per-channel scale / shift precomputed by gn_final_kernel (many chunks per sample: the VAE decoder);
// coef == nullptr: every CTA finalises its sample's group statistics itself from the (few) chunk partials - two launches.
template <typename T>
__global__ void __launch_bounds__(512) gn_apply_nhwc_kernel(const T* __restrict__ x, const T* __restrict__ x2, int C1,
                                                            const float* __restrict__ coef, const float2* __restrict__ part,
                                                            const float* __restrict__ gamma, const float* __restrict__ beta,
                                                            const float* __restrict__ chan_add, int64_t add_stride,
                                                            T* __restrict__ y, int HW, int C, int G, float eps, int apply_silu, int npx) {
    __shared__ float2 grp[GN_MAX_G];
    const int V = C >> 3, PH = blockDim.x / V;
    const int tid = threadIdx.x, v = tid % V, ph = tid / V, c0 = v << 3;
    const int b = blockIdx.y, chunk = blockIdx.x;
    const int p0 = chunk * npx, p1 = min(HW, p0 + npx);
    const int C2 = C - C1;
    const int sC = c0 >= C1 ? C2 : C1;
    const T* xp = c0 >= C1 ? x2 + (size_t)b * HW * C2 + (c0 - C1) : x + (size_t)b * HW * C1 + c0;
    T* yp = y + (size_t)b * HW * C + c0;
    int p = p0 + ph;
    constexpr int U = 4;
    Vec8<T> t[U];
    const bool full0 = p + (U - 1) * PH < p1;
    if (full0) {
#pragma unroll
        for (int u = 0; u < U; ++u) t[u].load(xp + (size_t)(p + u * PH) * sC);
    }
    float sa[8], sb[8];
    if (coef) {
        const float* sc = coef + (size_t)b * 2 * C + c0;
        const float4 a0 = *reinterpret_cast<const float4*>(sc), a1 = *reinterpret_cast<const float4*>(sc + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(sc + C), b1 = *reinterpret_cast<const float4*>(sc + C + 4);
        sa[0] = a0.x; sa[1] = a0.y; sa[2] = a0.z; sa[3] = a0.w; sa[4] = a1.x; sa[5] = a1.y; sa[6] = a1.z; sa[7] = a1.w;
        sb[0] = b0.x; sb[1] = b0.y; sb[2] = b0.z; sb[3] = b0.w; sb[4] = b1.x; sb[5] = b1.y; sb[6] = b1.z; sb[7] = b1.w;
    } else {
        const float* addb = chan_add ? chan_add + (size_t)b * add_stride : nullptr;
        if (tid < G) grp[tid] = gn_group_stats(part, x, x2, C1, addb, b, tid, gridDim.x, HW, C, G, eps);
        __syncthreads();
        const int cpg = C / G;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = c0 + i;
            const float2 st = grp[c / cpg];
            const float a = st.y * gamma[c];
            sa[i] = a;
            sb[i] = beta[c] + ((addb ? addb[c] : 0.0f) - st.x) * a;
        }
    }
    if (apply_silu == 2) {
#pragma unroll
        for (int i = 0; i < 8; ++i) { sa[i] *= 0.5f; sb[i] *= 0.5f; }
    }
    if (full0) {
#pragma unroll
        for (int u = 0; u < U; ++u) gn_emit(t[u], sa, sb, apply_silu, yp + (size_t)(p + u * PH) * C);
        p += U * PH;
    }
    for (; p + (U - 1) * PH < p1; p += U * PH) {
#pragma unroll
        for (int u = 0; u < U; ++u) t[u].load(xp + (size_t)(p + u * PH) * sC);
#pragma unroll
        for (int u = 0; u < U; ++u) gn_emit(t[u], sa, sb, apply_silu, yp + (size_t)(p + u * PH) * C);
    }
    for (; p < p1; p += PH) {
        t[0].load(xp + (size_t)p * sC);
        gn_emit(t[0], sa, sb, apply_silu, yp + (size_t)p * C);
    }
}

// ------------------------------------------------------------------------------------------------ NHWC, one launch
// Samples up to ~1.6 MB (every GroupNorm of the UNet): a thread-block cluster of S <= 8 CTAs owns one sample, CTA r the
// contiguous slab of pixels [r*npix, (r+1)*npix).  The slab is pulled into shared memory ONCE with bulk async copies
// (cp.async.bulk + mbarrier: no registers spent on loads in flight), statistics are taken from shared memory, the
// per-group partials of the S CTAs are combined through distributed shared memory (two cluster barriers, no global round
// trip), and the normalise pass reads the slab back from shared memory: HBM/L2 sees one read and one write, one launch.
// Pixels beyond the shared-memory budget (only the 960-channel 32x32 site) are streamed from L2 in both passes.
// Needs channels-per-group >= 8 so that a thread's 8 channels touch at most two groups.
constexpr int GNC_THREADS = 512;
constexpr int GNC_SLAB_BUDGET = 96 * 1024;              // two CTAs per SM

__device__ __forceinline__ uint32_t gn_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <typename T>
__global__ void __launch_bounds__(GNC_THREADS, 2) gn_cluster_kernel(const T* __restrict__ x, const T* __restrict__ x2, int C1,
                                                                    const float* __restrict__ gamma,
                                                                    const float* __restrict__ beta,
                                                                    const float* __restrict__ chan_add, int64_t add_stride,
                                                                    T* __restrict__ y, int HW, int C, int G, float eps,
                                                                    int apply_silu, int npix, int spix, int S) {
    extern __shared__ __align__(128) unsigned char gn_smem[];
    const int V = C >> 3, PH = blockDim.x / V;
    const int tid = threadIdx.x, v = tid % V, ph = tid / V, c0 = v << 3;
    const int b = blockIdx.y, rank = blockIdx.x, cpg = C / G;
    const size_t slab_bytes = (size_t)spix * C * sizeof(T);
    T* slab = reinterpret_cast<T*>(gn_smem);
    float4* red = reinterpret_cast<float4*>(gn_smem + ((slab_bytes + 127) & ~(size_t)127));   // [PH][V]: {A1, A2, B1, B2}
    float2* part = reinterpret_cast<float2*>(red + (size_t)PH * V);                            // [G] this CTA's partials
    float2* grp = part + GN_MAX_G;                                                             // [G] (mean, rstd)
    float* Kg = reinterpret_cast<float*>(grp + GN_MAX_G);                                      // [G] shifts
    uint64_t* mbar = reinterpret_cast<uint64_t*>(Kg + GN_MAX_G);

    // two-source form (x2 != nullptr): the input is the channel concatenation [x | x2] of two NHWC tensors with C1 and
    // C - C1 channels (the skip connections of the up path: torch.cat is never materialised); each source has its own slab
    const int C2 = C - C1;
    const float* addb = chan_add ? chan_add + (size_t)b * add_stride : nullptr;
    const int p0 = rank * npix;
    const int my = max(0, min(npix, HW - p0));          // pixels of this CTA
    const int ms = min(my, spix);                       // ... of which staged in shared memory
    const T* xs1 = x + (size_t)b * HW * C1;
    const T* xs2 = x2 ? x2 + (size_t)b * HW * C2 : nullptr;
    const T* xg1 = xs1 + (size_t)p0 * C1;
    const T* xg2 = x2 ? xs2 + (size_t)p0 * C2 : nullptr;
    T* yg = y + ((size_t)b * HW + p0) * C;
    T* slab2 = slab + (size_t)spix * C1;
    // this thread's column: source, row stride and base pointers
    const bool second = c0 >= C1;
    const int sC = second ? C2 : C1;
    const T* sl = second ? slab2 + (c0 - C1) : slab + c0;
    const T* xg = second ? xg2 + (c0 - C1) : xg1 + c0;

    if (tid == 0) {      // the slab copy goes out first: everything else overlaps with it
        asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(gn_smem_u32(mbar)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t total = (uint32_t)((size_t)ms * C * sizeof(T));
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(gn_smem_u32(mbar)), "r"(total) : "memory");
        for (int src = 0; src < (x2 ? 2 : 1); ++src) {
            const uint32_t bytes = (uint32_t)((size_t)ms * (src ? C2 : C1) * sizeof(T));
            const unsigned char* g = reinterpret_cast<const unsigned char*>(src ? xg2 : xg1);
            unsigned char* d = reinterpret_cast<unsigned char*>(src ? slab2 : slab);
            for (uint32_t off = 0; off < bytes; off += 32768u) {
                const uint32_t n = min(32768u, bytes - off);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(gn_smem_u32(d + off)), "l"(g + off), "r"(n), "r"(gn_smem_u32(mbar))
                             : "memory");
            }
        }
    }
    if (tid < G) {
        const int cg = tid * cpg;
        Kg[tid] = to_f(cg < C1 ? xs1[cg] : xs2[cg - C1]) + (addb ? addb[cg] : 0.0f);
    }
    __syncthreads();
    // per-thread constants while the copies fly
    const int g0 = c0 / cpg;
    const int split = min(8, (g0 + 1) * cpg - c0);      // channels [0, split) of this thread are in group g0, the rest in g0+1
    float gm[8], bt[8], sh[8];
    {
        const float4 a0 = *reinterpret_cast<const float4*>(gamma + c0), a1 = *reinterpret_cast<const float4*>(gamma + c0 + 4);
        const float4 b0 = *reinterpret_cast<const float4*>(beta + c0), b1 = *reinterpret_cast<const float4*>(beta + c0 + 4);
        gm[0] = a0.x; gm[1] = a0.y; gm[2] = a0.z; gm[3] = a0.w; gm[4] = a1.x; gm[5] = a1.y; gm[6] = a1.z; gm[7] = a1.w;
        bt[0] = b0.x; bt[1] = b0.y; bt[2] = b0.z; bt[3] = b0.w; bt[4] = b1.x; bt[5] = b1.y; bt[6] = b1.z; bt[7] = b1.w;
#pragma unroll
        for (int i = 0; i < 8; ++i) sh[i] = (addb ? addb[c0 + i] : 0.0f) - Kg[i < split ? g0 : g0 + 1];
    }
    {   // wait for the slab
        uint32_t done = 0;
        while (!done) {
            asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                         : "=r"(done) : "r"(gn_smem_u32(mbar)) : "memory");
        }
    }
    // ---- pass 1: shifted sums per channel -> two group partials per thread
    float a1s[8], a2s[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) { a1s[i] = 0.0f; a2s[i] = 0.0f; }
    {
        auto acc = [&](const Vec8<T>& tv) {
            float f[8];
            tv.unpack(f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = f[i] + sh[i]; a1s[i] += d; a2s[i] = fmaf(d, d, a2s[i]); }
        };
        int p = ph;
        for (; p + PH < ms; p += 2 * PH) {
            Vec8<T> t0, t1;
            t0.load(sl + (size_t)p * sC);
            t1.load(sl + (size_t)(p + PH) * sC);
            acc(t0);
            acc(t1);
        }
        for (; p < ms; p += PH) {
            Vec8<T> t0;
            t0.load(sl + (size_t)p * sC);
            acc(t0);
        }
        for (; p + PH < my; p += 2 * PH) {               // streamed remainder (p >= spix)
            Vec8<T> t0, t1;
            t0.load(xg + (size_t)p * sC);
            t1.load(xg + (size_t)(p + PH) * sC);
            acc(t0);
            acc(t1);
        }
        for (; p < my; p += PH) {
            Vec8<T> t0;
            t0.load(xg + (size_t)p * sC);
            acc(t0);
        }
    }
    {
        float A1 = 0.0f, A2 = 0.0f, B1 = 0.0f, B2 = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (i < split) { A1 += a1s[i]; A2 += a2s[i]; }
            else { B1 += a1s[i]; B2 += a2s[i]; }
        }
        red[ph * V + v] = make_float4(A1, A2, B1, B2);
    }
    __syncthreads();
    {   // warp w reduces groups w, w + nwarps, ...: entries of the columns that overlap the group, all pixel phases
        const int warp = tid >> 5, lane = tid & 31, nwarps = blockDim.x >> 5;
        for (int g = warp; g < G; g += nwarps) {
            const int vlo = (g * cpg) >> 3, vhi = ((g + 1) * cpg - 1) >> 3, nv = vhi - vlo + 1;
            float t1 = 0.0f, t2 = 0.0f;
            for (int e = lane; e < nv * PH; e += 32) {
                const int vv = vlo + e % nv, pp = e / nv;
                const float4 r = red[pp * V + vv];
                const bool first = ((vv << 3) / cpg) == g;      // this group is the column's first (A) or second (B) group
                t1 += first ? r.x : r.z;
                t2 += first ? r.y : r.w;
            }
            t1 = warp_sum(t1);
            t2 = warp_sum(t2);
            if (lane == 0) part[g] = make_float2(t1, t2);
        }
    }
    // ---- combine the S CTAs of the sample through distributed shared memory
    if (S > 1) {
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    if (tid < G) {
        float t1 = 0.0f, t2 = 0.0f;
        if (S > 1) {
            const uint32_t local = gn_smem_u32(&part[tid]);
            for (int r = 0; r < S; ++r) {
                uint32_t remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(r));
                float2 w;
                asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(w.x), "=f"(w.y) : "r"(remote) : "memory");
                t1 += w.x;
                t2 += w.y;
            }
        } else {
            t1 = part[tid].x;
            t2 = part[tid].y;
        }
        const float inv_n = __fdividef(1.0f, (float)cpg * (float)HW);
        const float m = t1 * inv_n;
        const float var = fmaxf(t2 - t1 * m, 0.0f) * inv_n;
        grp[tid] = make_float2(Kg[tid] + m, rsqrtf(var + eps));
    }
    if (S > 1) {     // also keeps every CTA's shared memory alive until its peers have read it
        asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
    } else {
        __syncthreads();
    }
    // ---- pass 2: normalise (+ SiLU) from shared memory
    float sa[8], sb[8];
    {
        const int g1 = min(g0 + 1, G - 1);
        const float2 sA = grp[g0], sB = grp[g1];
        const float dA = Kg[g0] - sA.x, dB = Kg[g1] - sB.x;      // chan_add - mean = sh + (K_g - mean)
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sa[i] = (i < split ? sA.y : sB.y) * gm[i];
            sb[i] = bt[i] + (sh[i] + (i < split ? dA : dB)) * sa[i];
            if (apply_silu == 2) { sa[i] *= 0.5f; sb[i] *= 0.5f; }
        }
    }
    {
        int p = ph;
        for (; p + PH < ms; p += 2 * PH) {
            Vec8<T> t0, t1;
            t0.load(sl + (size_t)p * sC);
            t1.load(sl + (size_t)(p + PH) * sC);
            gn_emit(t0, sa, sb, apply_silu, yg + (size_t)p * C + c0);
            gn_emit(t1, sa, sb, apply_silu, yg + (size_t)(p + PH) * C + c0);
        }
        for (; p < ms; p += PH) {
            Vec8<T> t0;
            t0.load(sl + (size_t)p * sC);
            gn_emit(t0, sa, sb, apply_silu, yg + (size_t)p * C + c0);
        }
        for (; p + PH < my; p += 2 * PH) {
            Vec8<T> t0, t1;
            t0.load(xg + (size_t)p * sC);
            t1.load(xg + (size_t)(p + PH) * sC);
            gn_emit(t0, sa, sb, apply_silu, yg + (size_t)p * C + c0);
            gn_emit(t1, sa, sb, apply_silu, yg + (size_t)(p + PH) * C + c0);
        }
        for (; p < my; p += PH) {
            Vec8<T> t0;
            t0.load(xg + (size_t)p * sC);
            gn_emit(t0, sa, sb, apply_silu, yg + (size_t)p * C + c0);
        }
    }
}

struct GncPlan {
    bool ok;
    int V, PH, threads, S, npix, spix;
    size_t smem;
};

static GncPlan gnc_plan(int B, int C, int HW, int G, size_t esize) {
    GncPlan p{};
    p.V = C >> 3;
    const int cpg = C / G;
    if (cpg < 8 || p.V > GNC_THREADS || C % 8 != 0) return p;
    p.PH = GNC_THREADS / p.V;
    if (p.PH > HW) p.PH = HW;
    p.threads = p.V * p.PH;
    const size_t pix_bytes = (size_t)C * esize;
    const size_t sample = pix_bytes * HW;
    if (sample > (size_t)24 * GNC_SLAB_BUDGET) return p;                    // at most 2/3 of a slab streamed from L2
    int S = 1;
    while (S < 8 && sample > (size_t)S * GNC_SLAB_BUDGET) S *= 2;           // fit the slabs in shared memory ...
    while (S < 8 && (int64_t)B * S < num_sms() && HW / (S * 2) >= p.PH) S *= 2;   // ... and use the SMs
    while (S > 1 && HW < S) S /= 2;
    p.S = S;
    p.npix = (HW + S - 1) / S;
    p.spix = (int)(GNC_SLAB_BUDGET / pix_bytes);
    if (p.spix > p.npix) p.spix = p.npix;
    const size_t slab = (((size_t)p.spix * pix_bytes) + 127) & ~(size_t)127;
    p.smem = slab + (size_t)p.PH * p.V * sizeof(float4) + 2 * GN_MAX_G * sizeof(float2) + GN_MAX_G * sizeof(float) + 16;
    p.ok = true;
    return p;
}

template <typename T>
static int launch_cluster(const GncPlan& p, const T* x, const T* x2, int C1, const float* gamma, const float* beta, const float* chan_add,
                          int64_t add_stride, T* y, int B, int C, int HW, int G, float eps, int silu, cudaStream_t s) {
    DADD_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(y) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(x2) & 15) == 0, "dadd_groupnorm_fwd(NHWC)");
    auto kern = gn_cluster_kernel<T>;
    static bool configured = false;       // per template instance
    if (!configured) {
        if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024), "gn smem")) return 2;
        configured = true;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(p.S, B, 1);
    cfg.blockDim = dim3(p.threads, 1, 1);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = p.S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, x, x2, C1, gamma, beta, chan_add, add_stride, y, HW, C, G, eps, silu, p.npix, p.spix, p.S);
    if (e != cudaSuccess) return cuda_ok(e, "dadd_groupnorm_fwd(NHWC cluster) launch");
    return launched("dadd_groupnorm_fwd(NHWC cluster)");
}

// ------------------------------------------------------------------------------------------------ NCHW
template <typename T, bool CACHED>
__global__ void __launch_bounds__(512) gn_nchw_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta,
                                                       const float* __restrict__ chan_add, int64_t add_stride, T* __restrict__ y, int HW,
                                                       int C, int G, float eps, int apply_silu) {
    __shared__ float red[32 * 3];
    __shared__ float stat[2];
    const int cpg = C / G;
    const int b = blockIdx.x / G, g = blockIdx.x % G;
    const int64_t len = (int64_t)cpg * HW;          // contiguous run of this (sample, group)
    const int nvec = (int)(len >> 3);
    const T* xg = x + ((int64_t)b * C + (int64_t)g * cpg) * HW;
    T* yg = y + ((int64_t)b * C + (int64_t)g * cpg) * HW;
    const float* addg = chan_add ? chan_add + (int64_t)b * add_stride + g * cpg : nullptr;
    const int tid = threadIdx.x, nt = blockDim.x;

    const float K = to_f(xg[0]) + (addg ? addg[0] : 0.0f);
    float t1 = 0.0f, t2 = 0.0f;
    Vec8<T> cache[CACHED ? GN_CACHE : 1];
    auto accum = [&](const Vec8<T>& t, int iv) {
        float f[8];
        t.unpack(f);
        const float ad = addg ? addg[(iv << 3) / HW] : 0.0f;   // HW % 8 == 0: a vector never straddles channels
#pragma unroll
        for (int i = 0; i < 8; ++i) { const float d = f[i] + ad - K; t1 += d; t2 += d * d; }
    };
    if (CACHED) {
#pragma unroll
        for (int j = 0; j < GN_CACHE; ++j) {
            const int iv = tid + j * nt;
            if (iv < nvec) cache[j].load(xg + ((int64_t)iv << 3));
        }
#pragma unroll
        for (int j = 0; j < GN_CACHE; ++j) {
            const int iv = tid + j * nt;
            if (iv < nvec) accum(cache[j], iv);
        }
    } else {
        for (int iv = tid; iv < nvec; iv += nt) {
            Vec8<T> t;
            t.load(xg + ((int64_t)iv << 3));
            accum(t, iv);
        }
    }
    // thread-local (n, mean, M2) -> warp -> block via Chan
    const int cnt = nvec > tid ? ((nvec - tid + nt - 1) / nt) * 8 : 0;
    Moments m{(float)cnt, 0.0f, 0.0f};
    if (cnt > 0) { m.mean = t1 / (float)cnt; m.m2 = fmaxf(t2 - t1 * m.mean, 0.0f); }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Moments other{__shfl_xor_sync(0xffffffffu, m.n, o), __shfl_xor_sync(0xffffffffu, m.mean, o),
                      __shfl_xor_sync(0xffffffffu, m.m2, o)};
        m = chan_combine(m, other);
    }
    const int w = tid >> 5, l = tid & 31;
    if (l == 0) { red[w * 3] = m.n; red[w * 3 + 1] = m.mean; red[w * 3 + 2] = m.m2; }
    __syncthreads();
    if (w == 0) {
        const int nw = nt >> 5;
        Moments r = l < nw ? Moments{red[l * 3], red[l * 3 + 1], red[l * 3 + 2]} : Moments{0.0f, 0.0f, 0.0f};
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            Moments other{__shfl_xor_sync(0xffffffffu, r.n, o), __shfl_xor_sync(0xffffffffu, r.mean, o),
                          __shfl_xor_sync(0xffffffffu, r.m2, o)};
            r = chan_combine(r, other);
        }
        if (l == 0) { stat[0] = K + r.mean; stat[1] = rsqrtf(r.m2 / r.n + eps); }
    }
    __syncthreads();
    const float mean = stat[0], rstd = stat[1];
    auto emit = [&](Vec8<T>& t, int iv) {
        const int cl = (iv << 3) / HW;
        const int c = g * cpg + cl;
        const float a = rstd * gamma[c];
        const float bb = beta[c] + ((addg ? addg[cl] : 0.0f) - mean) * a;
        float f[8];
        t.unpack(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float o = fmaf(f[i], a, bb);
            f[i] = apply_silu ? silu(o) : o;
        }
        t.pack(f);
        t.store(yg + ((int64_t)iv << 3));
    };
    if (CACHED) {
#pragma unroll
        for (int j = 0; j < GN_CACHE; ++j) {
            const int iv = tid + j * nt;
            if (iv < nvec) emit(cache[j], iv);
        }
    } else {
        for (int iv = tid; iv < nvec; iv += nt) {
            Vec8<T> t;
            t.load(xg + ((int64_t)iv << 3));
            emit(t, iv);
        }
    }
}

// groupnorm_stream.cu: the one-launch streaming kernel (16-bit NHWC); -1 = shape not served
int64_t gn_stream_workspace_bytes(int B, int C, int HW, int G);
template <typename T>
int gn_stream_launch(const T* x, const T* x2, int C1, const float* gamma, const float* beta, const float* chan_add, int64_t add_stride,
                     T* y, int B, int C, int HW, int G, float eps, int act, void* workspace, int64_t workspace_bytes, cudaStream_t s);

static std::atomic<int> g_gn_impl{[] { const char* e = getenv("DADD_GN_IMPL"); return (e && !strcmp(e, "stream")) ? 1 : 0; }()};

static int64_t gn_workspace_bytes(int B, int C, int HW, int G) {
    const GnPlan p = gn_plan(B, C, HW);
    const int64_t flat = (int64_t)B * p.chunks * G * sizeof(float2) + (int64_t)B * 2 * C * sizeof(float);
    const int64_t stream = gn_stream_workspace_bytes(B, C, HW, G);
    return flat > stream ? flat : stream;
}

// The one-launch cluster kernel serves every shape it can hold; the flat passes (stats [-> final] -> apply) take the rest (samples
// above ~2.3 MB: 512x512 latents, the VAE decoder).  On B200 the two paths are within run-to-run noise of each other for the
// large-batch 32x32 sites (profiles/r01_groupnorm_paths.txt); DADD_GN_FLAT=1 forces the flat passes, DADD_GN_FLAT_MB=<n> sends
// activations of at least n MB to them.
static bool gn_use_flat(const GncPlan& cp, int B, int C, int HW, size_t esize) {
    static const int force_flat = [] { const char* e = getenv("DADD_GN_FLAT"); return e ? atoi(e) : -1; }();
    static const long long flat_mb = [] { const char* e = getenv("DADD_GN_FLAT_MB"); return e ? atoll(e) : -1ll; }();
    if (!cp.ok || force_flat > 0) return true;
    if (force_flat == 0 || flat_mb < 0) return false;
    return (long long)B * C * HW * (long long)esize >= (flat_mb << 20);
}

template <typename T>
static int launch_nhwc(const T* x, const T* x2, int C1, const float* gamma, const float* beta, const float* chan_add, int64_t add_stride,
                       T* y, int B, int C, int HW, int G, float eps, int silu, void* workspace, int64_t workspace_bytes, cudaStream_t s) {
    DADD_REQUIRE(C / 8 <= 512, "dadd_groupnorm_fwd(NHWC)");
    if constexpr (sizeof(T) == 2) {
        if (g_gn_impl.load(std::memory_order_relaxed) == 1) {      // dadd_groupnorm_select(1) / DADD_GN_IMPL=stream
            const int rc = gn_stream_launch<T>(x, x2, C1, gamma, beta, chan_add, add_stride, y, B, C, HW, G, eps, silu, workspace, workspace_bytes, s);
            if (rc >= 0) return rc;
        }
    }
    const GncPlan cp = gnc_plan(B, C, HW, G, sizeof(T));
    if (!gn_use_flat(cp, B, C, HW, sizeof(T))) return launch_cluster(cp, x, x2, C1, gamma, beta, chan_add, add_stride, y, B, C, HW, G, eps, silu, s);
    DADD_REQUIRE(workspace != nullptr && workspace_bytes >= gn_workspace_bytes(B, C, HW, G), "dadd_groupnorm_fwd(NHWC)");
    DADD_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 15) == 0, "dadd_groupnorm_fwd(NHWC)");
    DADD_REQUIRE(((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(x2) | reinterpret_cast<uintptr_t>(y)) & 15) == 0,
                 "dadd_groupnorm_fwd(NHWC)");
    const GnPlan p = gn_plan(B, C, HW);
    float* coef = static_cast<float*>(workspace);                                   // [B][2][C], 16-byte aligned rows
    float2* part = reinterpret_cast<float2*>(coef + (size_t)B * 2 * C);             // [B][chunks][G]
    const size_t smem = (size_t)2 * p.PH * C * sizeof(float);
    auto stats = gn_stats_nhwc_kernel<T>;
    if (smem > 48 * 1024) {
        if (cuda_ok(cudaFuncSetAttribute(stats, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "gn smem")) return 2;
    }
    const dim3 grid(p.chunks, B, 1);
    stats<<<grid, p.threads, smem, s>>>(x, x2, C1, chan_add, add_stride, part, HW, C, G, p.npx);
    if (int rc = launched("dadd_groupnorm_fwd(NHWC stats)")) return rc;
    if (p.chunks > 16) {      // many chunks per sample (the VAE decoder): finalise once per sample instead of once per CTA
        gn_final_kernel<T><<<B, 256, 0, s>>>(part, x, x2, C1, gamma, beta, chan_add, add_stride, coef, p.chunks, HW, C, G, eps);
        if (int rc = launched("dadd_groupnorm_fwd(NHWC final)")) return rc;
    } else {
        coef = nullptr;
    }
    gn_apply_nhwc_kernel<T><<<grid, p.threads, 0, s>>>(x, x2, C1, coef, part, gamma, beta, chan_add, add_stride, y, HW, C, G, eps, silu, p.npx);
    return launched("dadd_groupnorm_fwd(NHWC apply)");
}

template <typename T>
static int launch_nchw(const T* x, const float* gamma, const float* beta, const float* chan_add, int64_t add_stride, T* y, int B, int C,
                       int HW, int G, float eps, int silu, cudaStream_t s) {
    DADD_REQUIRE(HW % 8 == 0, "dadd_groupnorm_fwd(NCHW)");
    const int64_t nvec = (int64_t)(C / G) * HW / 8;
    int threads = 256;
    while (threads < 512 && nvec > (int64_t)threads * GN_CACHE) threads *= 2;
    if (nvec < 256) threads = (int)((nvec + 31) / 32 * 32);
    const bool cached = sizeof(T) == 2 && nvec <= (int64_t)threads * GN_CACHE;
    auto kern = gn_nchw_kernel<T, false>;
    if constexpr (sizeof(T) == 2) {
        if (cached) kern = gn_nchw_kernel<T, true>;
    }
    kern<<<B * G, threads, 0, s>>>(x, gamma, beta, chan_add, add_stride, y, HW, C, G, eps, silu);
    return launched("dadd_groupnorm_fwd(NCHW)");
}

}  // namespace daddkk

using namespace daddk;

extern "C" int64_t dadd_groupnorm_workspace_bytes(int B, int C, int HW, int G, int layout) {
    if (layout != DADD_LAYOUT_NHWC || B <= 0 || C <= 0 || HW <= 0 || G <= 0 || C % 8 != 0) return 0;
    return gn_workspace_bytes(B, C, HW, G);
}

extern "C" int dadd_groupnorm_fwd(const void* x, const float* gamma, const float* beta, const float* chan_add,
                                  int64_t chan_add_stride, void* y,
                                  int B, int C, int HW, int G, float eps, int apply_silu, int layout, int dtype,
                                  void* workspace, int64_t workspace_bytes, void* stream) {
    DADD_REQUIRE(x && y && gamma && beta, "dadd_groupnorm_fwd");
    DADD_REQUIRE(B >= 0 && C > 0 && HW > 0 && G > 0 && G <= GN_MAX_G, "dadd_groupnorm_fwd");
    DADD_REQUIRE(C % G == 0 && C % 8 == 0, "dadd_groupnorm_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_groupnorm_fwd");
    DADD_REQUIRE(layout == DADD_LAYOUT_NCHW || layout == DADD_LAYOUT_NHWC, "dadd_groupnorm_fwd");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    // 16-bit outputs take SiLU as h + h * tanh(h) (one MUFU op per element; |error| <= 2^-11 |h|, below the output's own
    // rounding for all but strongly negative pre-activations); DADD_SILU_EXACT=1 forces o / (1 + exp(-o)) everywhere.
    static const bool silu_exact = [] { const char* e = getenv("DADD_SILU_EXACT"); return e && atoi(e) != 0; }();
    if (apply_silu) apply_silu = (dtype != DADD_F32 && layout == DADD_LAYOUT_NHWC && !silu_exact) ? 2 : 1;
    if (layout == DADD_LAYOUT_NHWC)
        DADD_DISPATCH_ANY(dtype, T, return launch_nhwc((const T*)x, (const T*)nullptr, C, gamma, beta, chan_add, chan_add_stride, (T*)y, B, C, HW, G, eps, apply_silu, workspace, workspace_bytes, s));
    DADD_DISPATCH_ANY(dtype, T, return launch_nchw((const T*)x, gamma, beta, chan_add, chan_add_stride, (T*)y, B, C, HW, G, eps, apply_silu, s));
    return 1;
}

extern "C" int dadd_groupnorm_select(int impl) { return g_gn_impl.exchange(impl == 1 ? 1 : 0); }

extern "C" int dadd_groupnorm_cat_supported(int B, int C1, int C2, int HW, int G, int dtype) {
    return (B > 0 && C1 > 0 && C2 > 0 && HW > 0 && G > 0 && G <= GN_MAX_G && C1 % 8 == 0 && C2 % 8 == 0 && (C1 + C2) % G == 0 &&
            (C1 + C2) / 8 <= 512 && dtype16_ok(dtype)) ? 1 : 0;
}

extern "C" int dadd_groupnorm_cat_fwd(const void* x1, int C1, const void* x2, int C2, const float* gamma, const float* beta,
                                      const float* chan_add, int64_t chan_add_stride, void* y, int B, int HW, int G, float eps,
                                      int apply_silu, int dtype, void* workspace, int64_t workspace_bytes, void* stream) {
    DADD_REQUIRE(x1 && x2 && y && gamma && beta, "dadd_groupnorm_cat_fwd");
    DADD_REQUIRE(B >= 0, "dadd_groupnorm_cat_fwd");
    if (B == 0) return 0;
    if (!dadd_groupnorm_cat_supported(B, C1, C2, HW, G, dtype))
        return fail("%s: needs 16-bit NHWC operands with C1 %% 8 == 0, C2 %% 8 == 0, (C1 + C2) %% G == 0 (C1 + C2 = %lld, HW = %lld)",
                    "dadd_groupnorm_cat_fwd", (long long)C1 + C2, (long long)HW);
    cudaStream_t s = (cudaStream_t)stream;
    static const bool silu_exact = [] { const char* e = getenv("DADD_SILU_EXACT"); return e && atoi(e) != 0; }();
    if (apply_silu) apply_silu = silu_exact ? 1 : 2;
    DADD_DISPATCH_16(dtype, T, return launch_nhwc((const T*)x1, (const T*)x2, C1, gamma, beta, chan_add, chan_add_stride, (T*)y, B, C1 + C2, HW,
                                                  G, eps, apply_silu, workspace, workspace_bytes, s));
    return 1;
}
