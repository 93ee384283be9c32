// Channel-bias / residual adds around the library convolutions and GEMMs, as one vectorised pass:
//   y[r][c] = a[r][c] (+ res[r][c]) (+ bias[c])      a, res, y: [rows][C] (NHWC activations or token matrices)
// Replaces the separate cuDNN-bias broadcast add and the residual adds of diffusers' ResnetBlock2D.forward
// (output = shortcut(x) + conv2(...)), Transformer2DModel.forward (proj_out(...) + residual) and
// BasicTransformerBlock.forward (ff(...) + hidden_states), reached through src/models/unet/unet.py:140-146.
#include "common.cuh"

namespace daddk {

template <typename T, bool RES, bool BIAS>
__global__ void __launch_bounds__(256) bias_residual_kernel(const T* __restrict__ a, const T* __restrict__ res,
                                                            const float* __restrict__ bias, T* __restrict__ y, int64_t nvec, int V) {
    constexpr int U = 4;
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto one = [&](const Vec8<T>& va, const Vec8<T>& vr, int64_t idx) {
        float f[8];
        va.unpack(f);
        if constexpr (RES) {
            float g[8];
            vr.unpack(g);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[k] += g[k];
        }
        if constexpr (BIAS) {
            const int c0 = (int)(idx % V) << 3;
            const float4 b0 = *reinterpret_cast<const float4*>(bias + c0), b1 = *reinterpret_cast<const float4*>(bias + c0 + 4);
            f[0] += b0.x; f[1] += b0.y; f[2] += b0.z; f[3] += b0.w;
            f[4] += b1.x; f[5] += b1.y; f[6] += b1.z; f[7] += b1.w;
        }
        Vec8<T> o;
        o.pack(f);
        o.store(y + (idx << 3));
    };
    for (; i + (U - 1) * stride < nvec; i += U * stride) {
        Vec8<T> va[U], vr[RES ? U : 1];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            va[u].load(a + ((i + u * stride) << 3));
            if constexpr (RES) vr[u].load(res + ((i + u * stride) << 3));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) one(va[u], vr[RES ? u : 0], i + u * stride);
    }
    for (; i < nvec; i += stride) {
        Vec8<T> va, vr;
        va.load(a + (i << 3));
        if constexpr (RES) vr.load(res + (i << 3));
        one(va, vr, i);
    }
}

template <typename T>
static int launch_bias_residual(const T* a, const T* res, const float* bias, T* y, int64_t rows, int C, cudaStream_t s) {
    const int V = C >> 3;
    const int64_t nvec = rows * V;
    int64_t grid = (nvec + 256 * 4 - 1) / (256 * 4);
    const int64_t cap = (int64_t)num_sms() * 8;
    if (grid > cap) grid = cap;
    if (grid < 1) grid = 1;
#define DADD_BR(R, B) bias_residual_kernel<T, R, B><<<(unsigned)grid, 256, 0, s>>>(a, res, bias, y, nvec, V)
    if (res && bias) DADD_BR(true, true);
    else if (res) DADD_BR(true, false);
    else DADD_BR(false, true);
#undef DADD_BR
    return launched("dadd_bias_residual_fwd");
}

// Nearest-neighbour 2x upsampling of an NHWC activation (diffusers Upsample2D: F.interpolate(scale_factor=2, mode="nearest")
// ahead of its 3x3 convolution): each 16-byte input vector is read once and stored to its four output pixels.
__global__ void __launch_bounds__(256) upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int64_t nvec, int W, int V) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        const int v = (int)(i % V);
        const int64_t pix = i / V;
        const int w = (int)(pix % W);
        const int64_t bh = pix / W;                     // b * H + h
        const uint4 t = x[i];
        uint4* o = y + (((bh * 2) * (int64_t)(2 * W)) + 2 * w) * V + v;       // output pixel (2h, 2w)
        const int64_t row = (int64_t)2 * W * V;
        o[0] = t;
        o[V] = t;
        o[row] = t;
        o[row + V] = t;
    }
}

// CLIP's activation: y = x * sigmoid(1.702 x) (transformers `quick_gelu`, the MLP of the ViT-L/14 tower the reference loads at
// src/models/image_encoder.py:34-38), vectorised, in place allowed.
template <typename T>
__global__ void __launch_bounds__(256) quick_gelu_kernel(const T* __restrict__ x, T* __restrict__ y, int64_t nvec) {
    const int64_t stride = (int64_t)gridDim.x * blockDim.x;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += stride) {
        Vec8<T> t;
        t.load(x + (i << 3));
        float f[8];
        t.unpack(f);
#pragma unroll
        for (int k = 0; k < 8; ++k) f[k] = __fdividef(f[k], 1.0f + __expf(-1.702f * f[k]));
        t.pack(f);
        t.store(y + (i << 3));
    }
}

}  // namespace daddk

using namespace daddk;

extern "C" int dadd_quick_gelu_fwd(const void* x, void* y, int64_t n, int dtype, void* stream) {
    DADD_REQUIRE(x && y && n >= 0 && n % 8 == 0, "dadd_quick_gelu_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_quick_gelu_fwd");
    if (n == 0) return 0;
    const int64_t nvec = n >> 3;
    int64_t grid = (nvec + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    DADD_DISPATCH_ANY(dtype, T, quick_gelu_kernel<T><<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const T*)x, (T*)y, nvec));
    return launched("dadd_quick_gelu_fwd");
}

extern "C" int dadd_upsample_nearest2x_fwd(const void* x, void* y, int B, int H, int W, int C, int dtype, void* stream) {
    DADD_REQUIRE(x && y && B >= 0 && H > 0 && W > 0, "dadd_upsample_nearest2x_fwd");
    DADD_REQUIRE(C > 0 && C % 8 == 0, "dadd_upsample_nearest2x_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_upsample_nearest2x_fwd");
    if (B == 0) return 0;
    const int esz = dtype == DADD_F32 ? 4 : 2;
    const int V = C * esz / 16;                          // 16-byte vectors per pixel (fp32: two per 8 channels)
    const int64_t nvec = (int64_t)B * H * W * V;
    int64_t grid = (nvec + 255) / 256;
    const int64_t cap = (int64_t)num_sms() * 16;
    if (grid > cap) grid = cap;
    upsample2x_kernel<<<(unsigned)grid, 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, nvec, W, V);
    return launched("dadd_upsample_nearest2x_fwd");
}

extern "C" int dadd_bias_residual_fwd(const void* a, const void* res, const float* bias, void* y, int64_t rows, int C, int dtype,
                                      void* stream) {
    DADD_REQUIRE(a && y && (res || bias) && rows >= 0, "dadd_bias_residual_fwd");
    DADD_REQUIRE(C > 0 && C % 8 == 0, "dadd_bias_residual_fwd");
    DADD_REQUIRE(dtype_ok(dtype), "dadd_bias_residual_fwd");
    if (rows == 0) return 0;
    DADD_DISPATCH_ANY(dtype, T, return launch_bias_residual((const T*)a, (const T*)res, bias, (T*)y, rows, C, (cudaStream_t)stream));
    return 1;
}
