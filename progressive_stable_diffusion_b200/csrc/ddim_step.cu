// K5: fused CFG combine + x0 prediction + clamp + DDIM update (fp32 state), plus the small per-step
// bookkeeping kernel that lets one captured graph serve every step of the schedule.
// Reference: src/pipelines/inference/inference_pipeline_ip.py:427-430, 434-468.
#include "common.cuh"

namespace daddk {

thread_local char g_last_error[512] = "";
std::atomic<int64_t> g_launches{0};

struct DdimCoef {
    float sa, so, sap, ec, sigma;
    int is_last;
};

template <typename T>
__device__ __forceinline__ float ldf(const T* p, int64_t i) { return to_f(p[i]); }

// Each op rounds separately (the __f*_rn intrinsics are never contracted into FMAs) so the result is
// bit-identical to the reference's chain of eager fp32 tensor ops.
__device__ __forceinline__ float ddim_one(float x, float ec_, float eu_, bool has_u, float g, const DdimCoef& c,
                                          float noise, bool has_noise, float clampv) {
    float e = has_u ? __fadd_rn(eu_, __fmul_rn(g, __fsub_rn(ec_, eu_))) : ec_;
    float x0 = __fdiv_rn(__fsub_rn(x, __fmul_rn(c.so, e)), c.sa);
    if (x0 == x0) x0 = fminf(fmaxf(x0, -clampv), clampv);   // torch.clamp propagates NaN
    if (c.is_last) return x0;
    float r = __fadd_rn(__fmul_rn(c.sap, x0), __fmul_rn(c.ec, e));
    if (has_noise) r = __fadd_rn(r, __fmul_rn(c.sigma, noise));
    return r;
}

template <typename T, bool TABLE>
__global__ void __launch_bounds__(256) ddim_kernel(float* __restrict__ x, const T* __restrict__ ec,
                                                   const T* __restrict__ eu, float g, DdimCoef c,
                                                   const float* __restrict__ table, const int32_t* __restrict__ state,
                                                   int n_steps, const float* __restrict__ noise, float clampv, int64_t n) {
    if (TABLE) {
        // a replay beyond the schedule must not index past the tables: it re-applies the last row (and is reported by
        // the host wrapper, which knows how many replays it queued)
        const int s = min(max(state[0], 0), n_steps - 1);
        const float* r = table + (int64_t)s * 8;
        c.sa = r[0]; c.so = r[1]; c.sap = r[2]; c.ec = r[3]; c.sigma = r[4];
        c.is_last = r[5] != 0.0f;
        if (noise) noise += (int64_t)s * n;
    }
    const bool has_u = eu != nullptr, has_n = noise != nullptr && c.sigma != 0.0f;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        x[i] = ddim_one(x[i], ldf(ec, i), has_u ? ldf(eu, i) : 0.0f, has_u, g, c, has_n ? noise[i] : 0.0f, has_n, clampv);
    }
}

__global__ void step_begin_kernel(int32_t* state, const uint4* __restrict__ table, uint4* __restrict__ row_out,
                                  int64_t row_vecs, int n_steps) {
    const int cur = state[1];
    const int row = min(max(cur, 0), n_steps - 1);       // never read past the table (see ddim_kernel)
    __syncthreads();
    for (int64_t i = threadIdx.x; i < row_vecs; i += blockDim.x) row_out[i] = table[(int64_t)row * row_vecs + i];
    if (threadIdx.x == 0) {
        state[0] = cur;
        state[1] = cur + 1;
    }
}

template <typename T>
__global__ void image_post_kernel(const T* __restrict__ x, float* __restrict__ y, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        float v = ldf(x, i);
        v = fminf(fmaxf(v, -1.0f), 1.0f);
        v = __fdiv_rn(__fadd_rn(v, 1.0f), 2.0f);
        y[i] = fminf(fmaxf(v, 0.0f), 1.0f);
    }
}

// One block per label; exact restatement of ordinal_embedder.py:107-127,155-171,15-40 in fp32 (cumsum order kept).
__global__ void aoe_interp_kernel(const float* __restrict__ base, const float* __restrict__ deltas,
                                  const float* __restrict__ labels, float* __restrict__ out, int K, int D) {
    const int b = blockIdx.x;
    float y = labels[b];
    y = fminf(fmaxf(y, 0.0f), (float)(K - 1));
    const float lo_f = floorf(y);
    const int lo = (int)lo_f;
    const int hi = min(lo + 1, K - 1);
    const float a = __fsub_rn(y, lo_f);
    const float oma = __fsub_rn(1.0f, a);
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        float cum = 0.0f, e_lo = 0.0f, e_hi = 0.0f;
        const float bj = base[j];
        for (int k = 0; k < K; ++k) {
            if (k > 0) cum = __fadd_rn(cum, deltas[(int64_t)(k - 1) * D + j]);
            const float e = __fadd_rn(bj, cum);
            if (k == lo) e_lo = e;
            if (k == hi) e_hi = e;
        }
        out[(int64_t)b * D + j] = __fadd_rn(__fmul_rn(e_lo, oma), __fmul_rn(e_hi, a));
    }
}

static inline int grid_for(int64_t n, int threads) {
    int64_t g = (n + threads - 1) / threads;
    int64_t cap = (int64_t)num_sms() * 8;
    return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace daddkk

using namespace daddk;

extern "C" {

int dadd_abi_version(void) { return DADD_ABI_VERSION; }
const char* dadd_last_error(void) { return g_last_error; }
int64_t dadd_launch_count(void) { return g_launches.load(); }
void dadd_reset_launch_count(void) { g_launches.store(0); }

int dadd_ddim_step(float* x, const void* eps_cond, const void* eps_uncond, int eps_dtype, float guidance,
                   float sqrt_ab_t, float sqrt_1mab_t, float sqrt_ab_prev, float eps_coef, float sigma,
                   const float* noise, float clampv, int is_last, int64_t n, void* stream) {
    DADD_REQUIRE(x && eps_cond && n >= 0, "dadd_ddim_step");
    DADD_REQUIRE(dtype_ok(eps_dtype), "dadd_ddim_step");
    DADD_REQUIRE(sqrt_ab_t != 0.0f, "dadd_ddim_step");
    if (n == 0) return 0;
    DdimCoef c{sqrt_ab_t, sqrt_1mab_t, sqrt_ab_prev, eps_coef, sigma, is_last};
    cudaStream_t s = (cudaStream_t)stream;
    DADD_DISPATCH_ANY(eps_dtype, T, (ddim_kernel<T, false><<<grid_for(n, 256), 256, 0, s>>>(
                                        x, (const T*)eps_cond, (const T*)eps_uncond, guidance, c, nullptr, nullptr, 1, noise, clampv, n)));
    return launched("dadd_ddim_step");
}

int dadd_ddim_step_table(float* x, const void* eps_cond, const void* eps_uncond, int eps_dtype, float guidance,
                         const float* coef_table, const int32_t* step_state, int n_steps, const float* noise,
                         float clampv, int64_t n, void* stream) {
    DADD_REQUIRE(x && eps_cond && coef_table && step_state && n >= 0 && n_steps >= 1, "dadd_ddim_step_table");
    DADD_REQUIRE(dtype_ok(eps_dtype), "dadd_ddim_step_table");
    if (n == 0) return 0;
    DdimCoef c{};
    cudaStream_t s = (cudaStream_t)stream;
    DADD_DISPATCH_ANY(eps_dtype, T, (ddim_kernel<T, true><<<grid_for(n, 256), 256, 0, s>>>(
                                        x, (const T*)eps_cond, (const T*)eps_uncond, guidance, c, coef_table, step_state, n_steps, noise, clampv, n)));
    return launched("dadd_ddim_step_table");
}

int dadd_step_begin(int32_t* step_state, const void* table, void* row_out, int64_t row_bytes, int n_steps, void* stream) {
    DADD_REQUIRE(step_state != nullptr && n_steps >= 1, "dadd_step_begin");
    DADD_REQUIRE(row_bytes % 16 == 0, "dadd_step_begin");
    DADD_REQUIRE(row_bytes == 0 || (table && row_out), "dadd_step_begin");
    step_begin_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(step_state, (const uint4*)table, (uint4*)row_out,
                                                            row_bytes / 16, n_steps);
    return launched("dadd_step_begin");
}

int dadd_image_post_fwd(const void* x, float* y, int64_t n, int dtype, void* stream) {
    DADD_REQUIRE(x && y && n >= 0, "dadd_image_post_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_image_post_fwd");
    if (n == 0) return 0;
    DADD_DISPATCH_ANY(dtype, T, (image_post_kernel<T><<<grid_for(n, 256), 256, 0, (cudaStream_t)stream>>>((const T*)x, y, n)));
    return launched("dadd_image_post_fwd");
}

int dadd_aoe_interp_fwd(const float* base, const float* deltas, const float* labels, float* out, int B, int K, int D,
                        void* stream) {
    DADD_REQUIRE(base && deltas && labels && out, "dadd_aoe_interp_fwd");
    DADD_REQUIRE(B >= 0 && K >= 2 && D > 0, "dadd_aoe_interp_fwd");
    if (B == 0) return 0;
    aoe_interp_kernel<<<B, 256, 0, (cudaStream_t)stream>>>(base, deltas, labels, out, K, D);
    return launched("dadd_aoe_interp_fwd");
}

}  // extern "C"
