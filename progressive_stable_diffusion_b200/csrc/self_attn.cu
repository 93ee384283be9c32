// K1 entry point: validation + shape dispatch of the self-attention core.
// Reference op: F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0
// (src/models/attention_processor_routing_gates.py:284-286).
//   N >= 128            -> tcgen05 / TMEM / TMA flash kernel, two query tiles per persistent CTA (self_attn_tc.cu)
//   N <  128 (64, 16)   -> warp-level mma.sync kernel (self_attn_mma.cu): one KV tile, latency-bound sites
//   d = 256 / 512       -> column-split mma.sync kernel (self_attn_mma.cu): the VAE mid block's single wide head
#include <cstdlib>

#include "common.cuh"

namespace daddk {
int self_attn_mma(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os,
                  int B, int H, int N, int d, float scale, int dtype, cudaStream_t s);
int self_attn_tc(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os,
                 int B, int H, int N, int d, float scale, int dtype, cudaStream_t s);
bool self_attn_tc_supported(int N, int d);
bool self_attn_wide_supported(int d);
int self_attn_wide(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os, int B, int H, int N,
                   int d, float scale, int dtype, cudaStream_t s);
}  // namespace daddk

using namespace daddk;

// DADD_SELF_ATTN=mma|tc forces one implementation (tests and A/B timing); default is the shape dispatch above.
static int forced_impl() {
    static const int v = [] {
        const char* e = getenv("DADD_SELF_ATTN");
        if (!e) return 0;
        return e[0] == 'm' ? 1 : (e[0] == 't' ? 2 : 0);
    }();
    return v;
}

extern "C" int dadd_self_attn_fwd(const void* q, const void* k, const void* v, int64_t q_stride, int64_t k_stride,
                                  int64_t v_stride, void* o, int64_t o_stride, int B, int H, int N, int d, float scale,
                                  int dtype, int impl, void* stream) {
    DADD_REQUIRE(q && k && v && o, "dadd_self_attn_fwd");
    DADD_REQUIRE(dtype16_ok(dtype), "dadd_self_attn_fwd");
    DADD_REQUIRE(impl >= 0 && impl <= 2, "dadd_self_attn_fwd");
    DADD_REQUIRE(B >= 0 && H > 0 && N >= 0 && B <= 65535 && H <= 65535, "dadd_self_attn_fwd");
    DADD_REQUIRE(d > 0 && d % 8 == 0 && (d <= 160 || self_attn_wide_supported(d)), "dadd_self_attn_fwd");
    DADD_REQUIRE(q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && o_stride % 8 == 0, "dadd_self_attn_fwd");
    DADD_REQUIRE(q_stride >= (int64_t)H * d && k_stride >= (int64_t)H * d && v_stride >= (int64_t)H * d &&
                     o_stride >= (int64_t)H * d, "dadd_self_attn_fwd");
    DADD_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) % 16 == 0, "dadd_self_attn_fwd");
    if (B == 0 || N == 0) return 0;
    if (d > 160)      // wide single heads (the VAE mid block): column-split mma.sync kernel whatever `impl` says
        return self_attn_wide(q, k, v, q_stride, k_stride, v_stride, o, o_stride, B, H, N, d, scale, dtype, (cudaStream_t)stream);
    if (impl == 0) impl = forced_impl();
    if (impl == 0) impl = (N >= 128 && self_attn_tc_supported(N, d)) ? 2 : 1;
    if (impl == 2) {
        DADD_REQUIRE(self_attn_tc_supported(N, d), "dadd_self_attn_fwd(tcgen05)");
        return self_attn_tc(q, k, v, q_stride, k_stride, v_stride, o, o_stride, B, H, N, d, scale, dtype, (cudaStream_t)stream);
    }
    return self_attn_mma(q, k, v, q_stride, k_stride, v_stride, o, o_stride, B, H, N, d, scale, dtype, (cudaStream_t)stream);
}
