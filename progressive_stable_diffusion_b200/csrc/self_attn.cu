// K1 entry point: validation + shape dispatch of the self-attention core.
// Reference op: F.scaled_dot_product_attention inside diffusers' AttnProcessor2_0
// (src/models/attention_processor_routing_gates.py:284-286).
#include <cstdlib>

#include "common.cuh"

namespace daddk {
int self_attn_mma(const void* q, const void* k, const void* v, int64_t qs, int64_t ks, int64_t vs, void* o, int64_t os,
                  int B, int H, int N, int d, float scale, cudaStream_t s);
}  // namespace daddk

using namespace daddk;

extern "C" int dadd_self_attn_fwd(const void* q, const void* k, const void* v, int64_t q_stride, int64_t k_stride,
                                  int64_t v_stride, void* o, int64_t o_stride, int B, int H, int N, int d, float scale,
                                  void* stream) {
    DADD_REQUIRE(q && k && v && o, "dadd_self_attn_fwd");
    DADD_REQUIRE(B >= 0 && H > 0 && N >= 0 && B <= 65535 && H <= 65535, "dadd_self_attn_fwd");
    DADD_REQUIRE(d > 0 && d % 8 == 0 && d <= 160, "dadd_self_attn_fwd");
    DADD_REQUIRE(q_stride % 8 == 0 && k_stride % 8 == 0 && v_stride % 8 == 0 && o_stride % 8 == 0, "dadd_self_attn_fwd");
    DADD_REQUIRE(q_stride >= (int64_t)H * d && k_stride >= (int64_t)H * d && v_stride >= (int64_t)H * d &&
                     o_stride >= (int64_t)H * d, "dadd_self_attn_fwd");
    DADD_REQUIRE(((uintptr_t)q | (uintptr_t)k | (uintptr_t)v | (uintptr_t)o) % 16 == 0, "dadd_self_attn_fwd");
    if (B == 0 || N == 0) return 0;
    return self_attn_mma(q, k, v, q_stride, k_stride, v_stride, o, o_stride, B, H, N, d, scale, (cudaStream_t)stream);
}
