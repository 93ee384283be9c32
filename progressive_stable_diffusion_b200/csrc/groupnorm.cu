// K4/K4b: GroupNorm (+ per-(sample,channel) additive term) (+ SiLU), fp32 statistics, one HBM read + one write.
// Reference ops: diffusers ResnetBlock2D.norm1/norm2 + SiLU, conv_norm_out + conv_act, Transformer2DModel.norm
// (reached via src/models/unet/unet.py:140-146).
//
// NHWC (channels-last, the layout the B200 UNet runs in): a thread-block cluster of S CTAs covers one sample, each CTA
// owns HW/S pixels x all C channels (fully coalesced 16 B accesses).  Every thread keeps a fixed 8-channel column, so
// per-channel shifted sums live in registers; per-channel -> per-group -> per-cluster combination uses Chan's
// formula (no E[x^2]-E[x]^2 cancellation), the cross-CTA hop goes through distributed shared memory.
// The CTA's slab stays in registers between the statistics pass and the normalise pass when it fits.
// NCHW (the reference's layout): a group is contiguous, one CTA per (sample, group).
#include <cooperative_groups.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

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
template <typename T, bool CACHED>
__global__ void __launch_bounds__(512) gn_nhwc_kernel(const T* __restrict__ x, const float* __restrict__ gamma,
                                                      const float* __restrict__ beta,
                                                      const float* __restrict__ chan_add, int64_t add_stride, T* __restrict__ y, int HW,
                                                      int C, int G, float eps, int apply_silu, int cluster_size) {
    extern __shared__ float smem[];
    __shared__ float cta_stats[GN_MAX_G * 3];
    __shared__ float grp[GN_MAX_G * 2];

    const int V = C >> 3;                 // 8-channel vectors per pixel
    const int PH = blockDim.x / V;        // pixel phases per CTA
    const int tid = threadIdx.x;
    const int v = tid % V, ph = tid / V;
    const int c0 = v << 3;
    const int b = blockIdx.y;
    const int rank = blockIdx.x;          // == rank in cluster (cluster spans gridDim.x)
    const int npix = HW / cluster_size;
    const int p0 = rank * npix;
    const int cpg = C / G;

    float* s1 = smem;                     // [PH][C]
    float* s2 = smem + (size_t)PH * C;    // [PH][C]
    float* ksh = s2 + (size_t)PH * C;     // [C]

    const T* xb = x + ((size_t)b * HW + p0) * C + c0;
    T* yb = y + ((size_t)b * HW + p0) * C + c0;

    float add[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) add[i] = chan_add ? chan_add[(size_t)b * add_stride + c0 + i] : 0.0f;

    float K[8], a1[8], a2[8];
    {
        Vec8<T> k;
        k.load(xb);
        k.unpack(K);
#pragma unroll
        for (int i = 0; i < 8; ++i) { K[i] += add[i]; a1[i] = 0.0f; a2[i] = 0.0f; }
    }

    Vec8<T> cache[CACHED ? GN_CACHE : 1];
    if (CACHED) {
#pragma unroll
        for (int j = 0; j < GN_CACHE; ++j) {
            const int p = ph + j * PH;
            if (p < npix) cache[j].load(xb + (size_t)p * C);
        }
#pragma unroll
        for (int j = 0; j < GN_CACHE; ++j) {
            const int p = ph + j * PH;
            if (p < npix) {
                float f[8];
                cache[j].unpack(f);
#pragma unroll
                for (int i = 0; i < 8; ++i) { const float d = f[i] + add[i] - K[i]; a1[i] += d; a2[i] += d * d; }
            }
        }
    } else {
        for (int p = ph; p < npix; p += PH) {
            Vec8<T> t;
            t.load(xb + (size_t)p * C);
            float f[8];
            t.unpack(f);
#pragma unroll
            for (int i = 0; i < 8; ++i) { const float d = f[i] + add[i] - K[i]; a1[i] += d; a2[i] += d * d; }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        s1[(size_t)ph * C + c0 + i] = a1[i];
        s2[(size_t)ph * C + c0 + i] = a2[i];
        if (ph == 0) ksh[c0 + i] = K[i];
    }
    __syncthreads();

    // per-channel moments of this CTA's slab (overwrite row 0 of s1/s2 with mean / M2)
    for (int c = tid; c < C; c += blockDim.x) {
        float t1 = 0.0f, t2 = 0.0f;
        for (int q = 0; q < PH; ++q) { t1 += s1[(size_t)q * C + c]; t2 += s2[(size_t)q * C + c]; }
        const float n = (float)npix;
        const float m = t1 / n;
        s1[c] = ksh[c] + m;            // in place: column c is touched by this thread only
        s2[c] = fmaxf(t2 - t1 * m, 0.0f);
    }
    __syncthreads();

    if (tid < G) {
        float mg = 0.0f;
        for (int i = 0; i < cpg; ++i) mg += s1[tid * cpg + i];
        mg /= (float)cpg;
        float m2 = 0.0f;
        for (int i = 0; i < cpg; ++i) {
            const float d = s1[tid * cpg + i] - mg;
            m2 += s2[tid * cpg + i] + (float)npix * d * d;
        }
        cta_stats[tid * 3 + 0] = (float)npix * (float)cpg;
        cta_stats[tid * 3 + 1] = mg;
        cta_stats[tid * 3 + 2] = m2;
    }
    if (cluster_size > 1) {
        cg::cluster_group cluster = cg::this_cluster();
        cluster.sync();
        if (tid < G) {
            Moments acc{0.0f, 0.0f, 0.0f};
            for (int r = 0; r < cluster_size; ++r) {
                const float* rs = cluster.map_shared_rank(cta_stats, r);
                acc = chan_combine(acc, Moments{rs[tid * 3], rs[tid * 3 + 1], rs[tid * 3 + 2]});
            }
            grp[tid * 2] = acc.mean;
            grp[tid * 2 + 1] = rsqrtf(acc.m2 / acc.n + eps);
        }
        cluster.sync();
    } else {
        __syncthreads();
        if (tid < G) {
            grp[tid * 2] = cta_stats[tid * 3 + 1];
            grp[tid * 2 + 1] = rsqrtf(cta_stats[tid * 3 + 2] / cta_stats[tid * 3] + eps);
        }
        __syncthreads();
    }

    float sa[8], sb[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int c = c0 + i, g = c / cpg;
        const float a = grp[g * 2 + 1] * gamma[c];
        sa[i] = a;
        sb[i] = beta[c] + (add[i] - grp[g * 2]) * a;
    }
    auto emit = [&](Vec8<T>& t, int p) {
        float f[8];
        t.unpack(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            float o = fmaf(f[i], sa[i], sb[i]);
            f[i] = apply_silu ? silu(o) : o;
        }
        t.pack(f);
        t.store(yb + (size_t)p * C);
    };
    if (CACHED) {
#pragma unroll
        for (int j = 0; j < GN_CACHE; ++j) {
            const int p = ph + j * PH;
            if (p < npix) emit(cache[j], p);
        }
    } else {
        for (int p = ph; p < npix; p += PH) {
            Vec8<T> t;
            t.load(xb + (size_t)p * C);
            emit(t, p);
        }
    }
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

template <typename T>
static int launch_nhwc(const T* x, const float* gamma, const float* beta, const float* chan_add, int64_t add_stride, T* y, int B, int C,
                       int HW, int G, float eps, int silu, cudaStream_t s) {
    const int V = C / 8;
    DADD_REQUIRE(V <= 512, "dadd_groupnorm_fwd(NHWC)");
    const int PH = 512 / V;
    const int threads = V * PH;
    // cluster size: largest power of two <= 8 dividing HW that keeps >= 8 pixels per CTA
    static const int max_cluster = [] {
        const char* e = getenv("DADD_GN_MAX_CLUSTER");
        int v = e ? atoi(e) : 8;
        return v >= 16 ? 16 : (v >= 8 ? 8 : (v >= 4 ? 4 : (v >= 2 ? 2 : 1)));
    }();
    int S = 1;
    while (S < max_cluster && HW % (S * 2) == 0 && HW / (S * 2) >= 8) S *= 2;
    const int npix = HW / S;
    const bool cached = sizeof(T) == 2 && (npix + PH - 1) / PH <= GN_CACHE;
    const size_t smem = ((size_t)2 * PH * C + C) * sizeof(float);
    auto kern = gn_nhwc_kernel<T, false>;
    if constexpr (sizeof(T) == 2) {
        if (cached) kern = gn_nhwc_kernel<T, true>;
    }
    if (smem > 48 * 1024) {
        if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "gn smem")) return 2;
    }
    if (S > 8) {
        if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1), "gn cluster16")) return 2;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(S, B, 1);
    cfg.blockDim = dim3(threads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = S;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    cudaError_t e = cudaLaunchKernelEx(&cfg, kern, x, gamma, beta, chan_add, add_stride, y, HW, C, G, eps, silu, S);
    if (e != cudaSuccess) return cuda_ok(e, "dadd_groupnorm_fwd(NHWC) launch");
    return launched("dadd_groupnorm_fwd(NHWC)");
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

extern "C" int dadd_groupnorm_fwd(const void* x, const float* gamma, const float* beta, const float* chan_add,
                                  int64_t chan_add_stride, void* y,
                                  int B, int C, int HW, int G, float eps, int apply_silu, int layout, int dtype,
                                  void* stream) {
    DADD_REQUIRE(x && y && gamma && beta, "dadd_groupnorm_fwd");
    DADD_REQUIRE(B >= 0 && C > 0 && HW > 0 && G > 0 && G <= GN_MAX_G, "dadd_groupnorm_fwd");
    DADD_REQUIRE(C % G == 0 && C % 8 == 0, "dadd_groupnorm_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_groupnorm_fwd");
    DADD_REQUIRE(layout == DADD_LAYOUT_NCHW || layout == DADD_LAYOUT_NHWC, "dadd_groupnorm_fwd");
    if (B == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    if (layout == DADD_LAYOUT_NHWC)
        DADD_DISPATCH_ANY(dtype, T, return launch_nhwc((const T*)x, gamma, beta, chan_add, chan_add_stride, (T*)y, B, C, HW, G, eps, apply_silu, s));
    DADD_DISPATCH_ANY(dtype, T, return launch_nchw((const T*)x, gamma, beta, chan_add, chan_add_stride, (T*)y, B, C, HW, G, eps, apply_silu, s));
    return 1;
}
