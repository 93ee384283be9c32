// Row-wise helpers of the transformer blocks and the Feature Purifier:
//   LayerNorm (48 per UNet step + purifier), GEGLU gate, purifier multi-head attention core (16x16 per head),
//   purifier gating epilogue (sigmoid gate * disease subtracted from the image tokens, fused with the final LayerNorm).
// Reference: diffusers BasicTransformerBlock; src/models/feature_purifier.py:81-95.
#include "common.cuh"

namespace daddk {

constexpr int LN_MAX_VEC = 8;  // 8 vectors x 8 elements x 32 lanes = C <= 2048

// (optional residual add +) LayerNorm.  A row is owned by LPR lanes (32, 16 or 8: C = 320 is 40 16-byte vectors, which 8 lanes
// x 5 vectors cover exactly, where a full warp would leave 24 of its 64 slots empty), so a warp works on 32 / LPR rows at a time,
// R such passes in flight, the rows packed in registers.
//   s = x (+ y) rounded to T (written to sum_out when given, plus sum_bias[c] when given);  out = LayerNorm(s) * gamma + beta.
// Exact two-pass mean / variance in fp32.  NV = ceil(C / (8 LPR)) 16-byte vectors per lane.  All loads are issued before the
// first reduction: R * (32 / LPR) * C * sizeof(T) bytes in flight per warp (the kernel is latency-bound otherwise).
template <int LPR>
__device__ __forceinline__ float group_sum(float v) {
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

template <typename T, int NV, int R, bool ADD, int LPR>
__global__ void __launch_bounds__(256) layernorm_kernel(const T* __restrict__ x, const T* __restrict__ yadd, T* __restrict__ sum_out,
                                                        const float* __restrict__ sum_bias,
                                                        const float* __restrict__ gamma, const float* __restrict__ beta,
                                                        T* __restrict__ out, int64_t rows, int C, float eps) {
    constexpr int G = 32 / LPR;                                  // rows per warp pass
    const int lane = threadIdx.x & 31, sub = lane % LPR, grp = lane / LPR;
    const int64_t row0 = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * (R * G);
    if (row0 >= rows) return;
    const int nvec = C >> 3;
    Vec8<T> vx[R][NV], vy[ADD ? R : 1][ADD ? NV : 1];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const int64_t row = row0 + r * G + grp;
        if (row < rows) {
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const int iv = sub + j * LPR;
                if (iv < nvec) {
                    vx[r][j].load(x + row * C + (iv << 3));
                    if constexpr (ADD) vy[r][j].load(yadd + row * C + (iv << 3));
                }
            }
        }
    }
    const float inv_c = 1.0f / (float)C;
#pragma unroll
    for (int r = 0; r < R; ++r) {
        if (row0 + r * G >= rows) break;                         // warp-uniform: no row of this pass exists
        const int64_t row = row0 + r * G + grp;
        const bool live = row < rows;                            // (dead lanes carry zeros through the shuffles)
        float f[NV][8];
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int iv = sub + j * LPR;
#pragma unroll
            for (int i = 0; i < 8; ++i) f[j][i] = 0.0f;
            if (live && iv < nvec) {
                vx[r][j].unpack(f[j]);
                if constexpr (ADD) {
                    float g[8];
                    vy[r][j].unpack(g);
#pragma unroll
                    for (int i = 0; i < 8; ++i) f[j][i] += g[i];
                    Vec8<T> t;
                    t.pack(f[j]);
                    if (sum_out) {
                        if (sum_bias) {                   // the stored sum (a later residual) carries an extra channel bias
                            const float4 c0 = *reinterpret_cast<const float4*>(sum_bias + (iv << 3));
                            const float4 c1 = *reinterpret_cast<const float4*>(sum_bias + (iv << 3) + 4);
                            const float cb[8] = {c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w};
                            float fb[8];
#pragma unroll
                            for (int i = 0; i < 8; ++i) fb[i] = f[j][i] + cb[i];
                            Vec8<T> tb;
                            tb.pack(fb);
                            tb.store(sum_out + row * C + (iv << 3));
                        } else {
                            t.store(sum_out + row * C + (iv << 3));
                        }
                    }
                    t.unpack(f[j]);                       // LayerNorm sees the rounded sum, like the unfused graph
                }
#pragma unroll
                for (int i = 0; i < 8; i += 2) { s0 += f[j][i]; s1 += f[j][i + 1]; }
            }
        }
        const float mean = group_sum<LPR>(s0 + s1) * inv_c;
        float q0 = 0.0f, q1 = 0.0f;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int iv = sub + j * LPR;
            if (live && iv < nvec) {
#pragma unroll
                for (int i = 0; i < 8; i += 2) {
                    const float d0 = f[j][i] - mean, d1 = f[j][i + 1] - mean;
                    q0 = fmaf(d0, d0, q0);
                    q1 = fmaf(d1, d1, q1);
                }
            }
        }
        const float rstd = rsqrtf(group_sum<LPR>(q0 + q1) * inv_c + eps);
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const int iv = sub + j * LPR;
            if (live && iv < nvec) {
                const float4 g0 = *reinterpret_cast<const float4*>(gamma + (iv << 3));
                const float4 g1 = *reinterpret_cast<const float4*>(gamma + (iv << 3) + 4);
                const float4 b0 = *reinterpret_cast<const float4*>(beta + (iv << 3));
                const float4 b1 = *reinterpret_cast<const float4*>(beta + (iv << 3) + 4);
                const float g[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
                const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
                float o[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] = (f[j][i] - mean) * rstd * g[i] + bb[i];
                Vec8<T> t;
                t.pack(o);
                t.store(out + row * C + (iv << 3));
            }
        }
    }
}

template <typename T, bool ADD>
static int launch_layernorm(const T* x, const T* yadd, T* sum_out, const float* sum_bias, const float* gamma, const float* beta, T* out,
                            int64_t rows, int C, float eps, cudaStream_t s) {
    const int nvec = C / 8;
    const int nv = (nvec + 31) / 32;
    const int wpb = 8;
#define DADD_LN(NVV, RR, LL)                                                                                              \
    do {                                                                                                                  \
        const int64_t per_block = (int64_t)wpb * RR * (32 / LL);                                                          \
        layernorm_kernel<T, NVV, RR, ADD, LL><<<(unsigned)((rows + per_block - 1) / per_block), wpb * 32, 0, s>>>(        \
            x, yadd, sum_out, sum_bias, gamma, beta, out, rows, C, eps);                                                  \
        return launched("dadd_layernorm_fwd");                                                                            \
    } while (0)
    if constexpr (sizeof(T) == 2) {
        if (nvec == 40) DADD_LN(5, 1, 8);               // C = 320: 8 lanes x 5 vectors per row, 4 rows per warp
        if (nvec == 80) DADD_LN(5, 1, 16);              // C = 640: 16 lanes x 5 vectors, 2 rows per warp
        if (nv <= 1) DADD_LN(1, 4, 32);
        if (nv == 2) DADD_LN(2, 4, 32);
        if (nv == 3) DADD_LN(3, 2, 32);
        if (nv <= 5) DADD_LN(5, 1, 32);
        DADD_LN(8, 1, 32);
    } else {
        if (nv <= 1) DADD_LN(1, 2, 32);
        if (nv <= 3) DADD_LN(3, 1, 32);
        DADD_LN(8, 1, 32);
    }
#undef DADD_LN
}

__device__ __forceinline__ float gelu_erf(float g) { return 0.5f * g * (1.0f + erff(g * 0.70710678118654752440f)); }

template <typename T>
__global__ void __launch_bounds__(256) geglu_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t rows, int inner) {
    const int vec_per_row = inner >> 3;
    const int64_t total = rows * vec_per_row;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t r = i / vec_per_row;
        const int j = (int)(i - r * vec_per_row) << 3;
        Vec8<T> a, g;
        a.load(x + r * 2 * inner + j);
        g.load(x + r * 2 * inner + inner + j);
        float fa[8], fg[8];
        a.unpack(fa);
        g.unpack(fg);
#pragma unroll
        for (int k = 0; k < 8; ++k) fa[k] *= gelu_erf(fg[k]);
        a.pack(fa);
        a.store(y + r * inner + j);
    }
}

// nn.MultiheadAttention core for the purifier (16 x 16 tokens) and the Perceiver resampler of the conditioning front end
// (16 queries x 257 patches): one CTA per (sample, head); everything in shared memory, fp32.
__global__ void __launch_bounds__(256) purifier_attn_kernel(const float* __restrict__ q, const float* __restrict__ k,
                                                            const float* __restrict__ v, float* __restrict__ o, int Lq,
                                                            int Lk, int D, int heads) {
    extern __shared__ float sm[];
    const int hd = D / heads;
    const int b = blockIdx.x / heads, h = blockIdx.x % heads;
    float* sq = sm;                 // [Lq][hd+1]
    float* sk = sq + Lq * (hd + 1); // [Lk][hd+1]
    float* sv = sk + Lk * (hd + 1); // [Lk][hd+1]
    float* sp = sv + Lk * (hd + 1); // [Lq][Lk]
    const int tid = threadIdx.x;
    for (int i = tid; i < Lq * hd; i += blockDim.x) {
        const int r = i / hd, c = i % hd;
        sq[r * (hd + 1) + c] = q[((int64_t)b * Lq + r) * D + h * hd + c];
    }
    for (int i = tid; i < Lk * hd; i += blockDim.x) {
        const int r = i / hd, c = i % hd;
        sk[r * (hd + 1) + c] = k[((int64_t)b * Lk + r) * D + h * hd + c];
        sv[r * (hd + 1) + c] = v[((int64_t)b * Lk + r) * D + h * hd + c];
    }
    __syncthreads();
    const float scale = rsqrtf((float)hd);
    for (int i = tid; i < Lq * Lk; i += blockDim.x) {
        const int r = i / Lk, c = i % Lk;
        float acc = 0.0f;
        for (int e = 0; e < hd; ++e) acc = fmaf(sq[r * (hd + 1) + e], sk[c * (hd + 1) + e], acc);
        sp[i] = acc * scale;
    }
    __syncthreads();
    if (tid < Lq) {
        float m = -INFINITY;
        for (int c = 0; c < Lk; ++c) m = fmaxf(m, sp[tid * Lk + c]);
        float s = 0.0f;
        for (int c = 0; c < Lk; ++c) { const float e = expf(sp[tid * Lk + c] - m); sp[tid * Lk + c] = e; s += e; }
        const float inv = 1.0f / s;
        for (int c = 0; c < Lk; ++c) sp[tid * Lk + c] *= inv;
    }
    __syncthreads();
    for (int i = tid; i < Lq * hd; i += blockDim.x) {
        const int r = i / hd, c = i % hd;
        float acc = 0.0f;
        for (int j = 0; j < Lk; ++j) acc = fmaf(sp[r * Lk + j], sv[j * (hd + 1) + c], acc);
        o[((int64_t)b * Lq + r) * D + h * hd + c] = acc;
    }
}

// e_clean = img - sigmoid(logit) * disease ; y = LayerNorm(e_clean).  One warp per row, D <= 2048.
__global__ void __launch_bounds__(256) purifier_gate_ln_kernel(const float* __restrict__ img,
                                                               const float* __restrict__ logits,
                                                               const float* __restrict__ disease,
                                                               const float* __restrict__ gamma,
                                                               const float* __restrict__ beta, float* __restrict__ y,
                                                               int64_t rows, int D, float eps) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= rows) return;
    const int per = (D + 31) / 32;      // <= 64
    float f[64];
    float sum = 0.0f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        const int c = lane + j * 32;
        if (j < per && c < D) {
            const int64_t i = row * D + c;
            const float gate = 1.0f / (1.0f + expf(-logits[i]));
            f[j] = img[i] - gate * disease[i];
            sum += f[j];
        }
    }
    const float mean = warp_sum(sum) / (float)D;
    float sq = 0.0f;
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        const int c = lane + j * 32;
        if (j < per && c < D) { const float d = f[j] - mean; sq += d * d; }
    }
    const float rstd = rsqrtf(warp_sum(sq) / (float)D + eps);
#pragma unroll
    for (int j = 0; j < 64; ++j) {
        const int c = lane + j * 32;
        if (j < per && c < D) y[row * D + c] = (f[j] - mean) * rstd * gamma[c] + beta[c];
    }
}

}  // namespace daddk

using namespace daddk;

extern "C" {

int dadd_layernorm_fwd(const void* x, const float* gamma, const float* beta, void* y, int64_t rows, int C, float eps,
                       int dtype, void* stream) {
    DADD_REQUIRE(x && y && gamma && beta && rows >= 0, "dadd_layernorm_fwd");
    DADD_REQUIRE(C > 0 && C % 8 == 0 && C <= LN_MAX_VEC * 256, "dadd_layernorm_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_layernorm_fwd");
    if (rows == 0) return 0;
    DADD_DISPATCH_ANY(dtype, T, return (launch_layernorm<T, false>((const T*)x, nullptr, nullptr, nullptr, gamma, beta, (T*)y, rows, C, eps,
                                                                   (cudaStream_t)stream)));
    return 1;
}

int dadd_add_layernorm_fwd(const void* x, const void* r, void* sum_out, const float* sum_bias, const float* gamma, const float* beta,
                           void* y, int64_t rows, int C, float eps, int dtype, void* stream) {
    DADD_REQUIRE(x && r && y && gamma && beta && rows >= 0, "dadd_add_layernorm_fwd");
    DADD_REQUIRE(sum_bias == nullptr || sum_out != nullptr, "dadd_add_layernorm_fwd");
    DADD_REQUIRE(C > 0 && C % 8 == 0 && C <= LN_MAX_VEC * 256, "dadd_add_layernorm_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_add_layernorm_fwd");
    if (rows == 0) return 0;
    DADD_DISPATCH_ANY(dtype, T, return (launch_layernorm<T, true>((const T*)x, (const T*)r, (T*)sum_out, sum_bias, gamma, beta, (T*)y, rows, C,
                                                                  eps, (cudaStream_t)stream)));
    return 1;
}

int dadd_geglu_fwd(const void* x, void* y, int64_t rows, int inner, int dtype, void* stream) {
    DADD_REQUIRE(x && y && rows >= 0 && inner > 0 && inner % 8 == 0, "dadd_geglu_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_geglu_fwd");
    if (rows == 0) return 0;
    const int64_t total = rows * (inner / 8);
    int64_t g = (total + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (g > cap) g = cap;
    DADD_DISPATCH_ANY(dtype, T, (geglu_kernel<T><<<(unsigned)g, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, rows, inner)));
    return launched("dadd_geglu_fwd");
}

int dadd_purifier_attn_fwd(const float* q, const float* k, const float* v, float* o, int B, int Lq, int Lk, int D,
                           int heads, void* stream) {
    DADD_REQUIRE(q && k && v && o, "dadd_purifier_attn_fwd");
    DADD_REQUIRE(B >= 0 && Lq > 0 && Lk > 0 && Lq <= 256, "dadd_purifier_attn_fwd");
    DADD_REQUIRE(heads > 0 && D % heads == 0 && D / heads <= 128, "dadd_purifier_attn_fwd");
    if (B == 0) return 0;
    const int hd = D / heads;
    const size_t smem = ((size_t)(Lq + 2 * Lk) * (hd + 1) + (size_t)Lq * Lk) * sizeof(float);
    if (smem > 227 * 1024)
        return fail("%s: (Lq + 2 Lk) (D / heads + 1) + Lq Lk floats must fit 227 KB of shared memory (Lq = %lld, Lk = %lld)",
                    "dadd_purifier_attn_fwd", (long long)Lq, (long long)Lk);
    if (smem > 48 * 1024) {
        if (cuda_ok(cudaFuncSetAttribute(purifier_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem),
                    "dadd_purifier_attn_fwd"))
            return 2;
    }
    purifier_attn_kernel<<<B * heads, 256, smem, (cudaStream_t)stream>>>(q, k, v, o, Lq, Lk, D, heads);
    return launched("dadd_purifier_attn_fwd");
}

int dadd_purifier_gate_ln_fwd(const float* img, const float* gate_logits, const float* disease, const float* gamma,
                              const float* beta, float* y, int64_t rows, int D, float eps, void* stream) {
    DADD_REQUIRE(img && gate_logits && disease && gamma && beta && y, "dadd_purifier_gate_ln_fwd");
    DADD_REQUIRE(rows >= 0 && D > 0 && D <= 2048, "dadd_purifier_gate_ln_fwd");
    if (rows == 0) return 0;
    const int wpb = 8;
    purifier_gate_ln_kernel<<<(unsigned)((rows + wpb - 1) / wpb), wpb * 32, 0, (cudaStream_t)stream>>>(
        img, gate_logits, disease, gamma, beta, y, rows, D, eps);
    return launched("dadd_purifier_gate_ln_fwd");
}

}  // extern "C"
