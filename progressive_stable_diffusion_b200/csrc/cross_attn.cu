// K2/K2b: fused multi-pathway cross-attention core (48 or 32 condition tokens).
// Reference: SplitInjectionAttentionProcessor.__call__ (src/models/attention_processor_routing_gates.py:148-178):
//   z = g_a softmax(Q Ka^T/sqrt d) Va + g_d softmax(Q Kd^T/sqrt d) Vd (+ lambda softmax(Q Kx^T/sqrt d) Vx)
// and OrdinalIPAttnProcessor2_0.__call__ (src/models/attention_processor_base.py:96-118) as one 32-token segment.
//
// The whole K/V of one (sample, head) is <= 64 x 160 16-bit elements and lives in shared memory (pulled in with cp.async:
// one global round trip for Q, K and V; V stays row-major and is read through ldmatrix.trans); each warp owns 16
// query rows: S = Q K_cat^T (one pass, all segments), an independent softmax per 16-token segment in registers, the
// gate of the segment folded into the normalisation (invariant I11: sum_s g_s P_s V_s = [g_s P_s]_s V_cat), O = P V_cat.
// No score tensor is materialised; HBM traffic is Q in + O out (+ the tiny K/V), which is the roofline of this op
// (48 FLOP/B).  The contraction sizes (K = d <= 160, N = 48) are far below one tcgen05 tile, so the warp-level
// mma.sync path is used here; the op is bound by HBM/L2 bytes, not by the tensor pipe.
#include "mma_util.cuh"

namespace daddk {

// 16-byte global -> shared copy that does not pass through registers; src_bytes = 0 zero-fills the destination
__device__ __forceinline__ void cp_async16(void* dst, const void* src, int src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src), "r"(src_bytes)
                 : "memory");
}

// B fragment (16 keys x 8 columns) of a row-major [key][column] tile: ldmatrix with transpose
__device__ __forceinline__ void load_b_frag_trans(uint32_t& b0, uint32_t& b1, const void* row_ptr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0, %1}, [%2];"
                 : "=r"(b0), "=r"(b1)
                 : "r"((uint32_t)__cvta_generic_to_shared(row_ptr)));
}

template <typename T, int DK>
__global__ void __launch_bounds__(128) cross_attn_kernel(const T* __restrict__ q, int64_t q_stride,
                                                         const T* __restrict__ k_cat, const T* __restrict__ v_cat,
                                                         T* __restrict__ o, int64_t o_stride, int H, int N, int d,
                                                         int seg_len, int n_seg, const float* __restrict__ gates,
                                                         float scale_log2e) {
    constexpr int QS = DK + 8;          // smem row stride (elements): conflict-free 32-bit fragment loads and ldmatrix rows
    constexpr int LMAX = 64;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* Qs = reinterpret_cast<T*>(smem_raw);   // [64][QS]
    T* Ks = Qs + 64 * QS;                      // [LMAX][QS]
    T* Vs = Ks + LMAX * QS;                    // [LMAX][QS]  (row-major; PV reads it through ldmatrix.trans)

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;
    const int b = blockIdx.z, h = blockIdx.y;
    const int row0 = blockIdx.x * 64;
    const int L = seg_len * n_seg;
    const int dv = d >> 3;              // 16-byte chunks per row

    const T* qb = q + ((int64_t)b * N + row0) * q_stride + (int64_t)h * d;
    const T* kb = k_cat + ((int64_t)(b * H + h) * L) * d;
    const T* vb = v_cat + ((int64_t)(b * H + h) * L) * d;

    // all copies are issued back to back (cp.async), one wait: a single global round trip for Q, K and V
    constexpr int DKV = DK >> 3;
    for (int i = tid; i < 64 * DKV; i += 128) {
        const int r = i / DKV, c = i % DKV;
        const bool ok = c < dv && row0 + r < N;
        cp_async16(Qs + r * QS + c * 8, ok ? qb + (int64_t)r * q_stride + c * 8 : q, ok ? 16 : 0);
    }
    for (int i = tid; i < L * DKV; i += 128) {
        const int r = i / DKV, c = i % DKV;
        const bool ok = c < dv;
        cp_async16(Ks + r * QS + c * 8, ok ? kb + (int64_t)r * d + c * 8 : k_cat, ok ? 16 : 0);
        cp_async16(Vs + r * QS + c * 8, ok ? vb + (int64_t)r * d + c * 8 : v_cat, ok ? 16 : 0);
    }
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
    __syncthreads();

    // ---- S = Q K^T for this warp's 16 rows, all L tokens -----------------------------------------------------
    float s[LMAX / 8][4];
#pragma unroll
    for (int nt = 0; nt < LMAX / 8; ++nt) { s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.0f; }
    const T* qw = Qs + (warp * 16) * QS;
#pragma unroll
    for (int kk = 0; kk < DK / 16; ++kk) {
        uint32_t a[4];
        load_a_frag(a, qw, QS, kk * 16, g, t);
#pragma unroll
        for (int nt = 0; nt < LMAX / 8; ++nt) {
            if (nt * 8 < L) {
                uint32_t b0, b1;
                load_b_frag(b0, b1, Ks + (nt * 8) * QS, QS, kk * 16, g, t);
                mma_16816<T>(s[nt], a, b0, b1);
            }
        }
    }
    // ---- per-segment softmax, gate folded into the normaliser ------------------------------------------------
    const int tiles_per_seg = seg_len >> 3;
    for (int sg = 0; sg < n_seg; ++sg) {
        const float gate = gates[sg];
        float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
        for (int nt = 0; nt < LMAX / 8; ++nt) {
            if (nt / tiles_per_seg == sg) {
                m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
                m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
            }
        }
        m0 = quad_max(m0);
        m1 = quad_max(m1);
        float l0 = 0.0f, l1 = 0.0f;
#pragma unroll
        for (int nt = 0; nt < LMAX / 8; ++nt) {
            if (nt / tiles_per_seg == sg) {
                s[nt][0] = fast_exp2((s[nt][0] - m0) * scale_log2e);
                s[nt][1] = fast_exp2((s[nt][1] - m0) * scale_log2e);
                s[nt][2] = fast_exp2((s[nt][2] - m1) * scale_log2e);
                s[nt][3] = fast_exp2((s[nt][3] - m1) * scale_log2e);
                l0 += s[nt][0] + s[nt][1];
                l1 += s[nt][2] + s[nt][3];
            }
        }
        l0 = quad_sum(l0);
        l1 = quad_sum(l1);
        const float r0 = gate / l0, r1 = gate / l1;
#pragma unroll
        for (int nt = 0; nt < LMAX / 8; ++nt) {
            if (nt / tiles_per_seg == sg) {
                s[nt][0] *= r0; s[nt][1] *= r0;
                s[nt][2] *= r1; s[nt][3] *= r1;
            }
        }
    }
    // ---- O = P V ---------------------------------------------------------------------------------------------
    float acc[DK / 8][4];
#pragma unroll
    for (int nd = 0; nd < DK / 8; ++nd) { acc[nd][0] = acc[nd][1] = acc[nd][2] = acc[nd][3] = 0.0f; }
#pragma unroll
    for (int kk = 0; kk < LMAX / 16; ++kk) {
        if (kk * 16 < L) {
            uint32_t a[4];
            a[0] = pack2<T>(s[2 * kk][0], s[2 * kk][1]);
            a[1] = pack2<T>(s[2 * kk][2], s[2 * kk][3]);
            a[2] = pack2<T>(s[2 * kk + 1][0], s[2 * kk + 1][1]);
            a[3] = pack2<T>(s[2 * kk + 1][2], s[2 * kk + 1][3]);
#pragma unroll
            for (int nd = 0; nd < DK / 8; ++nd) {
                if (nd < dv) {
                    uint32_t b0, b1;      // lanes 0-15 address the 16 key rows of this k-step, 8 columns wide
                    load_b_frag_trans(b0, b1, Vs + (kk * 16 + (lane & 15)) * QS + nd * 8);
                    mma_16816<T>(acc[nd], a, b0, b1);
                }
            }
        }
    }
    // ---- stage this warp's 16 x d tile in its own Q rows, then 16-byte coalesced stores -----------------------
    __syncwarp();
    T* ow = Qs + (warp * 16) * QS;
#pragma unroll
    for (int nd = 0; nd < DK / 8; ++nd) {
        if (nd < dv) {
            *reinterpret_cast<uint32_t*>(ow + g * QS + nd * 8 + 2 * t) = pack2<T>(acc[nd][0], acc[nd][1]);
            *reinterpret_cast<uint32_t*>(ow + (g + 8) * QS + nd * 8 + 2 * t) = pack2<T>(acc[nd][2], acc[nd][3]);
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

template <typename T, int DK>
static int launch_cross(const void* q, int64_t q_stride, const void* k_cat, const void* v_cat, void* o,
                        int64_t o_stride, int B, int H, int N, int d, int seg_len, int n_seg, const float* gates,
                        float scale, cudaStream_t s) {
    const size_t smem = (size_t)3 * 64 * (DK + 8) * sizeof(T);
    auto kern = cross_attn_kernel<T, DK>;
    if (smem > 48 * 1024) {
        if (cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "cross_attn smem"))
            return 2;
    }
    dim3 grid((N + 63) / 64, H, B);
    kern<<<grid, 128, smem, s>>>((const T*)q, q_stride, (const T*)k_cat, (const T*)v_cat, (T*)o, o_stride, H, N, d, seg_len,
                                 n_seg, gates, scale * 1.4426950408889634f);
    return launched("dadd_cross_attn_fwd");
}

// cross_attn_tc.cu
bool cross_attn_tc_supported(int N, int d, int seg_len, int n_seg);
int cross_attn_tc(const void* q, int64_t q_stride, const void* k_cat, const void* v_cat, void* o, int64_t o_stride, int B, int H,
                  int N, int d, int seg_len, int n_seg, const float* gates, float scale, int dtype, cudaStream_t s);

}  // namespace daddk

using namespace daddk;

extern "C" int dadd_cross_attn_fwd(const void* q, int64_t q_stride, const void* k_cat, const void* v_cat, void* o,
                                   int64_t o_stride, int B, int H, int N, int d, int seg_len, int n_seg,
                                   const float* gates, float scale, int dtype, int impl, void* stream) {
    DADD_REQUIRE(q && k_cat && v_cat && o && gates, "dadd_cross_attn_fwd");
    DADD_REQUIRE(dtype16_ok(dtype), "dadd_cross_attn_fwd");
    DADD_REQUIRE(B >= 0 && H > 0 && N >= 0 && B <= 65535 && H <= 65535, "dadd_cross_attn_fwd");
    DADD_REQUIRE(d > 0 && d % 8 == 0 && d <= 160, "dadd_cross_attn_fwd");
    DADD_REQUIRE(seg_len > 0 && seg_len % 16 == 0 && n_seg > 0 && seg_len * n_seg <= 64, "dadd_cross_attn_fwd");
    DADD_REQUIRE(q_stride % 8 == 0 && o_stride % 8 == 0 && q_stride >= (int64_t)H * d && o_stride >= (int64_t)H * d,
                 "dadd_cross_attn_fwd");
    DADD_REQUIRE(((uintptr_t)q | (uintptr_t)k_cat | (uintptr_t)v_cat | (uintptr_t)o) % 16 == 0, "dadd_cross_attn_fwd");
    DADD_REQUIRE(impl >= 0 && impl <= 2, "dadd_cross_attn_fwd");
    if (B == 0 || N == 0) return 0;
    cudaStream_t s = (cudaStream_t)stream;
    const bool tc_ok = cross_attn_tc_supported(N, d, seg_len, n_seg);
    if (impl == 2 && !tc_ok)
        return fail("%s: impl = 2 (tcgen05) needs N >= 128, d <= 128 and (seg_len, n_seg) in {(16,2), (16,3), (32,1)}", "dadd_cross_attn_fwd");
    if (tc_ok && impl != 1) return cross_attn_tc(q, q_stride, k_cat, v_cat, o, o_stride, B, H, N, d, seg_len, n_seg, gates, scale, dtype, s);
#define DADD_X(DKV) \
    DADD_DISPATCH_16(dtype, T, return (launch_cross<T, DKV>(q, q_stride, k_cat, v_cat, o, o_stride, B, H, N, d, seg_len, n_seg, gates, scale, s)))
    if (d <= 48) DADD_X(48);
    if (d <= 64) DADD_X(64);
    if (d <= 80) DADD_X(80);
    if (d <= 128) DADD_X(128);
    DADD_X(160);
#undef DADD_X
    return 1;
}
