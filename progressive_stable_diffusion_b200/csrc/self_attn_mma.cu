// K1 bring-up path: flash-style self-attention on warp-level mma.sync (64 query rows x 64-key tiles, online softmax).
// Reference op: F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0 (installed at
// src/models/attention_processor_routing_gates.py:284-286).  Serves the short sequences (N <= 64: the d = 160 sites,
// one KV tile, latency-bound) and is the correctness anchor for the tcgen05/TMEM kernel in self_attn_tc.cu, which
// takes the long-sequence sites.
#include "mma_util.cuh"

namespace daddk {

// 16-byte global -> shared copy that does not pass through registers; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void sa_cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}
// B fragment (16 keys x 8 columns) of a row-major [key][column] tile: ldmatrix with transpose
__device__ __forceinline__ void sa_load_b_frag_trans(uint32_t& b0, uint32_t& b1, const void* row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                 : "=r"(b0), "=r"(b1)
                 : "r"((uint32_t)__cvta_generic_to_shared(row_ptr)));
}

template <typename T, int DK>
__global__ void __launch_bounds__(128) self_attn_mma_kernel(const T* __restrict__ q,
                                                            const T* __restrict__ k,
                                                            const T* __restrict__ v, int64_t q_stride,
                                                            int64_t k_stride, int64_t v_stride,
                                                            T* __restrict__ o, int64_t o_stride, int N,
                                                            int d, float scale_log2e) {
    constexpr int QS = DK + 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Qs = reinterpret_cast<T*>(smem_raw);   // [64][QS]
    T* Ks = Qs + 64 * QS;                                  // [64][QS]
    T* Vs = Ks + 64 * QS;                                  // [64][QS]  (row-major; PV reads it through ldmatrix.trans)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y;
    const int row0 = blockIdx.x * 64;
    const int dv = d >> 3;
    constexpr int DKV = DK >> 3;

    const T* qb = q + ((int64_t)b * N + row0) * q_stride + (int64_t)h * d;
    const T* kb = k + ((int64_t)b * N) * k_stride + (int64_t)h * d;
    const T* vb = v + ((int64_t)b * N) * v_stride + (int64_t)h * d;

    for (int i = tid; i < 64 * DKV; i += 128) {
        const int r = i / DKV, c = i % DKV;
        const bool ok = c < dv && row0 + r < N;
        sa_cp_async16(Qs + r * QS + c * 8, ok ? qb + (int64_t)r * q_stride + c * 8 : q, ok ? 16 : 0);
    }

    float acc[DK / 8][4];
#pragma unroll
    for (int nd = 0; nd < DK / 8; ++nd) { acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.0f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;

    const T* qw = Qs + (warp * 16) * QS;
    for (int kv0 = 0; kv0 < N; kv0 += 64) {
        __syncthreads();   // previous tile fully consumed (also orders the Q fill before first use)
        for (int i = tid; i < 64 * DKV; i += 128) {       // one global round trip for the whole K / V tile (and Q the first time)
            const int r = i / DKV, c = i % DKV;
            const bool ok = c < dv && kv0 + r < N;
            sa_cp_async16(Ks + r * QS + c * 8, ok ? kb + (int64_t)(kv0 + r) * k_stride + c * 8 : k, ok ? 16 : 0);
            sa_cp_async16(Vs + r * QS + c * 8, ok ? vb + (int64_t)(kv0 + r) * v_stride + c * 8 : v, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();

        float s[8][4];
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f; }
#pragma unroll
        for (int kk = 0; kk < DK / 16; ++kk) {
            uint32_t a[4];
            load_a_frag(a, qw, QS, kk * 16, g, t);
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                uint32_t b0, b1;
                load_b_frag(b0, b1, Ks + (nt * 8) * QS, QS, kk * 16, g, t);
                mma_16816<T>(s[nt], a, b0, b1);
            }
        }
        if (kv0 + 64 > N) {   // mask the keys past the end of the sequence
#pragma unroll
            for (int nt = 0; nt < 8; ++nt) {
                const int key = kv0 + nt * 8 + 2 * t;
                if (key >= N) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                if (key + 1 >= N) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            }
        }
        float t0 = m0, t1 = m1;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            t0 = fmaxf(t0, fmaxf(s[nt][0], s[nt][1]));
            t1 = fmaxf(t1, fmaxf(s[nt][2], s[nt][3]));
        }
        t0 = quad_max(t0);
        t1 = quad_max(t1);
        const float c0 = fast_exp2((m0 - t0) * scale_log2e), c1 = fast_exp2((m1 - t1) * scale_log2e);
        m0 = t0;
        m1 = t1;
        float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 8; ++nt) {
            s[nt][0] = fast_exp2((s[nt][0] - m0) * scale_log2e);
            s[nt][1] = fast_exp2((s[nt][1] - m0) * scale_log2e);
            s[nt][2] = fast_exp2((s[nt][2] - m1) * scale_log2e);
            s[nt][3] = fast_exp2((s[nt][3] - m1) * scale_log2e);
            r0 += s[nt][0] + s[nt][1];
            r1 += s[nt][2] + s[nt][3];
        }
        l0 = l0 * c0 + r0;     // per-thread partial row sums; reduced across the quad at the end
        l1 = l1 * c1 + r1;
#pragma unroll
        for (int nd = 0; nd < DK / 8; ++nd) { acc[nd][0] *= c0; acc[nd][1] *= c0; acc[nd][2] *= c1; acc[nd][3] *= c1; }
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
            uint32_t a[4];
            a[0] = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
            a[1] = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
            a[2] = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            a[3] = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int nd = 0; nd < DK / 8; ++nd) {
                if (nd < dv) {
                    uint32_t b0, b1;      // lanes 0-15 address the 16 key rows of this k-step, 8 columns wide
                    sa_load_b_frag_trans(b0, b1, Vs + (kk * 16 + (lane & 15)) * QS + nd * 8);
                    mma_16816<T>(acc[nd], a, b0, b1);
                }
            }
        }
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __syncwarp();
    T* ow = Qs + (warp * 16) * QS;   // this warp's Q rows are no longer needed by anyone else
#pragma unroll
    for (int nd = 0; nd < DK / 8; ++nd) {
        if (nd < dv) {
            *reinterpret_cast<uint32_t*>(ow + g * QS + nd * 8 + 2 * t) = pack2<T>(acc[nd][0] * i0, acc[nd][1] * i0);
            *reinterpret_cast<uint32_t*>(ow + (g + 8) * QS + nd * 8 + 2 * t) = pack2<T>(acc[nd][2] * i1, acc[nd][3] * i1);
        }
    }
    __syncwarp();
    T* ob = o + ((int64_t)b * N + row0 + warp * 16) * o_stride + (int64_t)h * d;
    for (int i = lane; i < 16 * dv; i += 32) {
        const int r = i / dv, c = i % dv;
        if (row0 + warp * 16 + r < N)
            *reinterpret_cast<uint4*>(ob + (int64_t)r * o_stride + c * 8) = *reinterpret_cast<const uint4*>(ow + r * QS + c * 8);
    }
}


// ---------------------------------------------------------------------------------------------------------- wide heads
// d = 256 / 512 (the VAE mid block's single 512-wide head, src/models/vae/vae.py:90-112 -> diffusers Attention with heads = 1): the O
// accumulator of a 16-row block is split over FOUR warps by columns (d / 4 each: 64 fp32 registers per thread at d = 512), every one
// of which recomputes the block's 16 x 32 score tile - 2.5x the MMAs of a shared score tile, but no cross-warp exchange inside
// the online softmax; the site runs once per image (0.6 TFLOP for 104 images), not once per denoising step.
// CTA = 32 query rows (2 row blocks x 4 column quarters = 8 warps), 32-key tiles, Q / K / V tiles in padded shared memory.
template <typename T, int DK>
__global__ void __launch_bounds__(256, 2) self_attn_wide_kernel(const T* __restrict__ q, const T* __restrict__ k, const T* __restrict__ v,
                                                                int64_t q_stride, int64_t k_stride, int64_t v_stride, T* __restrict__ o,
                                                                int64_t o_stride, int N, float scale_log2e) {
    constexpr int QS = DK + 8, DKV = DK >> 3, KT = 32, CW = DK / 4, NDW = CW / 8;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Qs = reinterpret_cast<T*>(smem_raw);   // [32][QS]
    T* Ks = Qs + 32 * QS;                     // [KT][QS]
    T* Vs = Ks + KT * QS;                     // [KT][QS]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int rb = warp >> 2, cq = warp & 3;
    const int b = blockIdx.z, h = blockIdx.y;
    const int row0 = blockIdx.x * 32;
    const T* qb = q + ((int64_t)b * N + row0) * q_stride + (int64_t)h * DK;
    const T* kb = k + ((int64_t)b * N) * k_stride + (int64_t)h * DK;
    const T* vb = v + ((int64_t)b * N) * v_stride + (int64_t)h * DK;
    for (int i = tid; i < 32 * DKV; i += 256) {
        const int r = i / DKV, c = i % DKV;
        const bool ok = row0 + r < N;
        sa_cp_async16(Qs + r * QS + c * 8, ok ? qb + (int64_t)r * q_stride + c * 8 : q, ok ? 16 : 0);
    }
    float acc[NDW][4];
#pragma unroll
    for (int nd = 0; nd < NDW; ++nd) { acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.0f; }
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.0f, l1 = 0.0f;
    const T* qw = Qs + (rb * 16) * QS;
    for (int kv0 = 0; kv0 < N; kv0 += KT) {
        __syncthreads();
        for (int i = tid; i < KT * DKV; i += 256) {
            const int r = i / DKV, c = i % DKV;
            const bool ok = kv0 + r < N;
            sa_cp_async16(Ks + r * QS + c * 8, ok ? kb + (int64_t)(kv0 + r) * k_stride + c * 8 : k, ok ? 16 : 0);
            sa_cp_async16(Vs + r * QS + c * 8, ok ? vb + (int64_t)(kv0 + r) * v_stride + c * 8 : v, ok ? 16 : 0);
        }
        asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        float s[4][4];
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f; }
#pragma unroll 4
        for (int kk = 0; kk < DK / 16; ++kk) {
            uint32_t a[4];
            load_a_frag(a, qw, QS, kk * 16, g, t);
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                uint32_t b0, b1;
                load_b_frag(b0, b1, Ks + (nt * 8) * QS, QS, kk * 16, g, t);
                mma_16816<T>(s[nt], a, b0, b1);
            }
        }
        if (kv0 + KT > N) {
#pragma unroll
            for (int nt = 0; nt < 4; ++nt) {
                const int key = kv0 + nt * 8 + 2 * t;
                if (key >= N) { s[nt][0] = -INFINITY; s[nt][2] = -INFINITY; }
                if (key + 1 >= N) { s[nt][1] = -INFINITY; s[nt][3] = -INFINITY; }
            }
        }
        float t0 = m0, t1 = m1;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            t0 = fmaxf(t0, fmaxf(s[nt][0], s[nt][1]));
            t1 = fmaxf(t1, fmaxf(s[nt][2], s[nt][3]));
        }
        t0 = quad_max(t0);
        t1 = quad_max(t1);
        const float c0 = fast_exp2((m0 - t0) * scale_log2e), c1 = fast_exp2((m1 - t1) * scale_log2e);
        m0 = t0;
        m1 = t1;
        float r0 = 0.0f, r1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < 4; ++nt) {
            s[nt][0] = fast_exp2((s[nt][0] - m0) * scale_log2e);
            s[nt][1] = fast_exp2((s[nt][1] - m0) * scale_log2e);
            s[nt][2] = fast_exp2((s[nt][2] - m1) * scale_log2e);
            s[nt][3] = fast_exp2((s[nt][3] - m1) * scale_log2e);
            r0 += s[nt][0] + s[nt][1];
            r1 += s[nt][2] + s[nt][3];
        }
        l0 = l0 * c0 + r0;
        l1 = l1 * c1 + r1;
#pragma unroll
        for (int nd = 0; nd < NDW; ++nd) { acc[nd][0] *= c0; acc[nd][1] *= c0; acc[nd][2] *= c1; acc[nd][3] *= c1; }
#pragma unroll
        for (int kk = 0; kk < KT / 16; ++kk) {
            uint32_t a[4];
            a[0] = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
            a[1] = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
            a[2] = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            a[3] = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int nd = 0; nd < NDW; ++nd) {
                uint32_t b0, b1;
                sa_load_b_frag_trans(b0, b1, Vs + (kk * 16 + (lane & 15)) * QS + cq * CW + nd * 8);
                mma_16816<T>(acc[nd], a, b0, b1);
            }
        }
    }
    l0 = quad_sum(l0);
    l1 = quad_sum(l1);
    const float i0 = 1.0f / l0, i1 = 1.0f / l1;
    __syncthreads();                          // every warp is done with Q: its rows become the output staging
    T* ow = Qs + (rb * 16) * QS + cq * CW;
#pragma unroll
    for (int nd = 0; nd < NDW; ++nd) {
        *reinterpret_cast<uint32_t*>(ow + g * QS + nd * 8 + 2 * t) = pack2<T>(acc[nd][0] * i0, acc[nd][1] * i0);
        *reinterpret_cast<uint32_t*>(ow + (g + 8) * QS + nd * 8 + 2 * t) = pack2<T>(acc[nd][2] * i1, acc[nd][3] * i1);
    }
    __syncthreads();
    T* ob = o + ((int64_t)b * N + row0) * o_stride + (int64_t)h * DK;
    for (int i = tid; i < 32 * DKV; i += 256) {
        const int r = i / DKV, c = i % DKV;
        if (row0 + r < N) *reinterpret_cast<uint4*>(ob + (int64_t)r * o_stride + c * 8) = *reinterpret_cast<const uint4*>(Qs + r * QS + c * 8);
    }
}

template <typename T, int DK>
static int launch_self_wide(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os, int B, int H,
                            int N, float scale, cudaStream_t s) {
    const size_t smem = (size_t)(32 + 2 * 32) * (DK + 8) * sizeof(T);
    auto kern = self_attn_wide_kernel<T, DK>;
    if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "self_attn(wide) smem")) return 2;
    dim3 grid((N + 31) / 32, H, B);
    kern<<<grid, 256, smem, s>>>((const T*)q, (const T*)k, (const T*)v, qs, ks, vs, (T*)o, os, N, scale * 1.4426950408889634f);
    return launched("dadd_self_attn_fwd(wide head)");
}

bool self_attn_wide_supported(int d) { return d == 256 || d == 512; }

int self_attn_wide(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os, int B, int H, int N,
                   int d, float scale, int dtype, cudaStream_t s) {
    if (d == 256) DADD_DISPATCH_16(dtype, T, return (launch_self_wide<T, 256>(q, k, v, qs, ks, vs, o, os, B, H, N, scale, s)));
    if (d == 512) DADD_DISPATCH_16(dtype, T, return (launch_self_wide<T, 512>(q, k, v, qs, ks, vs, o, os, B, H, N, scale, s)));
    return fail("%s: wide heads are d = 256 or 512", "dadd_self_attn_fwd");
}

template <typename T, int DK>
static int launch_self_mma(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o,
                           int64_t os, int B, int H, int N, int d, float scale, cudaStream_t s) {
    const size_t smem = (size_t)3 * 64 * (DK + 8) * sizeof(T);
    auto kern = self_attn_mma_kernel<T, DK>;
    if (smem > 48 * 1024) {
        if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "self_attn smem"))
            return 2;
    }
    dim3 grid((N + 63) / 64, H, B);
    kern<<<grid, 128, smem, s>>>((const T*)q, (const T*)k, (const T*)v, qs, ks, vs,
                                 (T*)o, os, N, d, scale * 1.4426950408889634f);
    return launched("dadd_self_attn_fwd(mma)");
}

int self_attn_mma(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os,
                  int B, int H, int N, int d, float scale, int dtype, cudaStream_t s) {
#define DADD_X(DKV) DADD_DISPATCH_16(dtype, T, return (launch_self_mma<T, DKV>(q, k, v, qs, ks, vs, o, os, B, H, N, d, scale, s)))
    if (d <= 48) DADD_X(48);
    if (d <= 64) DADD_X(64);
    if (d <= 80) DADD_X(80);
    if (d <= 128) DADD_X(128);
    DADD_X(160);
#undef DADD_X
    return 1;
}

}  // namespace daddk
