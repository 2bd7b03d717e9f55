// Weight packing and small glue kernels used by the model runtime (run once at load, or once per step
// on tiny tensors): row gather + fp32->bf16 cast, SiLU+cast of the time embedding, the VAE's latent
// pre-scale + post_quant_conv, and the VAE output map to [0,1] / uint8.
#include "../common.cuh"
#include "../host_common.h"
#include "../launch_count.h"
#include "glue.h"

namespace sdod {

__global__ void pack_rows_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int N, int K, int Kpad, const int* __restrict__ rowmap) {
    const size_t total = static_cast<size_t>(N) * Kpad;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = o % Kpad;
        const int n = o / Kpad;
        const int r = rowmap ? rowmap[n] : n;
        dst[o] = __float2bfloat16(k < K ? src[static_cast<size_t>(r) * K + k] : 0.f);
    }
}
// dst[n*ld + col0 + k] = bf16(src[n*K + k]): a [N,K] fp32 matrix into a column window of a wider bf16 matrix (K-concatenated weights)
__global__ void pack_rows_window_kernel(const float* __restrict__ src, bf16* __restrict__ dst, int N, int K, long long ld, int col0) {
    const size_t total = static_cast<size_t>(N) * K;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = o % K;
        const size_t n = o / K;
        dst[n * ld + col0 + k] = __float2bfloat16(src[o]);
    }
}
__global__ void add_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ dst, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = a[i] + b[i];
}
__global__ void gather_f32_kernel(const float* __restrict__ src, float* __restrict__ dst, int n, const int* __restrict__ map) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[map ? map[i] : i];
}
__global__ void silu_f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n, int apply_silu) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float v = x[i];
        y[i] = __float2bfloat16(apply_silu ? silu_f(v) : v);
    }
}
// z [rows,4] fp32 -> (z * inv_scale) * Wpq^T + bpq -> bf16 [rows,4]   (AutoencoderKL.decode: post_quant_conv(z / 0.18215))
__global__ void latent_prequant_kernel(const float* __restrict__ z, bf16* __restrict__ y, size_t rows, const float* __restrict__ w,
                                       const float* __restrict__ b, float inv_scale) {
    for (size_t r = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; r < rows; r += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const float4 v = *reinterpret_cast<const float4*>(z + r * 4);
        const float in[4] = {v.x * inv_scale, v.y * inv_scale, v.z * inv_scale, v.w * inv_scale};
#pragma unroll
        for (int o = 0; o < 4; ++o) {
            float acc = b ? b[o] : 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) acc += w[o * 4 + i] * in[i];
            y[r * 4 + o] = __float2bfloat16(acc);
        }
    }
}
// conv_out fp32 [n] -> img = clamp((x+1)/2, 0, 1); u8 = uint8(clamp(255*img, 0, 255))  (truncation; context.cpp:392-395)
__global__ void vae_post_kernel(const float* __restrict__ x, uint8_t* __restrict__ u8, float* __restrict__ img, size_t n) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        float f = __fdiv_rn(__fadd_rn(x[i], 1.0f), 2.0f);
        f = fminf(fmaxf(f, 0.0f), 1.0f);
        if (img) img[i] = f;
        if (u8) {
            float v = __fmul_rn(255.0f, f);
            v = fminf(fmaxf(v, 0.0f), 255.0f);
            u8[i] = static_cast<uint8_t>(v);
        }
    }
}
__global__ void fill_f32_kernel(float* __restrict__ x, size_t n, float v) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) x[i] = v;
}
__global__ void scale_f32_kernel(float* __restrict__ x, size_t n, float s) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x) x[i] *= s;
}

__global__ void broadcast_rows_kernel(float* __restrict__ dst, const float* __restrict__ src, int rows, int width) {
    const size_t total = static_cast<size_t>(rows) * width;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        dst[i] = src[i % width];
}

__global__ void clip_embed_kernel(const int* __restrict__ tokens, const float* __restrict__ tok_emb, const float* __restrict__ pos_emb,
                                  float* __restrict__ out, int rows, int T, int D, int vocab) {
    const int d4 = D / 4;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < static_cast<size_t>(rows) * d4; i += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int r = static_cast<int>(i / d4), c = static_cast<int>(i - static_cast<size_t>(r) * d4);
        int id = tokens[r];
        id = id < 0 ? 0 : (id >= vocab ? vocab - 1 : id);
        const float4 a = reinterpret_cast<const float4*>(tok_emb + static_cast<size_t>(id) * D)[c];
        const float4 b = reinterpret_cast<const float4*>(pos_emb + static_cast<size_t>(r % T) * D)[c];
        reinterpret_cast<float4*>(out + static_cast<size_t>(r) * D)[c] = make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
    }
}

__global__ void cast_bf16_to_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, size_t n) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        y[i] = __bfloat162float(x[i]);
}

static inline int g1(size_t n) {
    size_t g = (n + 255) / 256;
    if (g > 148 * 32) g = 148 * 32;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

int pack_rows(cudaStream_t s, const float* src, void* dst, int N, int K, int Kpad, const int* rowmap) {
    pack_rows_kernel<<<g1(static_cast<size_t>(N) * Kpad), 256, 0, s>>>(src, static_cast<bf16*>(dst), N, K, Kpad, rowmap);
    count_launch();
    return check_launch("pack_rows_kernel");
}
int pack_rows_window(cudaStream_t s, const float* src, void* dst, int N, int K, long long ld, int col0) {
    pack_rows_window_kernel<<<g1(static_cast<size_t>(N) * K), 256, 0, s>>>(src, static_cast<bf16*>(dst), N, K, ld, col0);
    count_launch();
    return check_launch("pack_rows_window_kernel");
}
int add_f32(cudaStream_t s, const float* a, const float* b, float* dst, int n) {
    add_f32_kernel<<<(n + 255) / 256, 256, 0, s>>>(a, b, dst, n);
    count_launch();
    return check_launch("add_f32_kernel");
}
int gather_f32(cudaStream_t s, const float* src, float* dst, int n, const int* map) {
    gather_f32_kernel<<<(n + 255) / 256, 256, 0, s>>>(src, dst, n, map);
    count_launch();
    return check_launch("gather_f32_kernel");
}
int silu_f32_to_bf16(cudaStream_t s, const float* x, void* y, size_t n, int apply_silu) {
    silu_f32_to_bf16_kernel<<<g1(n), 256, 0, s>>>(x, static_cast<bf16*>(y), n, apply_silu);
    count_launch();
    return check_launch("silu_f32_to_bf16_kernel");
}
int latent_prequant(cudaStream_t s, const float* z, void* y, size_t rows, const float* w, const float* b, float inv_scale) {
    latent_prequant_kernel<<<g1(rows), 256, 0, s>>>(z, static_cast<bf16*>(y), rows, w, b, inv_scale);
    count_launch();
    return check_launch("latent_prequant_kernel");
}
int vae_post(cudaStream_t s, const float* x, uint8_t* u8, float* img, size_t n) {
    vae_post_kernel<<<g1(n), 256, 0, s>>>(x, u8, img, n);
    count_launch();
    return check_launch("vae_post_kernel");
}
int broadcast_rows(cudaStream_t s, float* dst, const float* src, int rows, int width) {
    broadcast_rows_kernel<<<g1(static_cast<size_t>(rows) * width), 256, 0, s>>>(dst, src, rows, width);
    count_launch();
    return check_launch("broadcast_rows_kernel");
}
int clip_embed(cudaStream_t s, const int* tokens, const float* tok_emb, const float* pos_emb, float* out, int rows, int T, int D, int vocab) {
    if (D % 4 != 0) return fail(kInvalidArgument, "clip_embed: width must be a multiple of 4");
    clip_embed_kernel<<<g1(static_cast<size_t>(rows) * (D / 4)), 256, 0, s>>>(tokens, tok_emb, pos_emb, out, rows, T, D, vocab);
    count_launch();
    return check_launch("clip_embed_kernel");
}
int cast_bf16_to_f32(cudaStream_t s, const void* x, float* y, size_t n) {
    cast_bf16_to_f32_kernel<<<g1(n), 256, 0, s>>>(static_cast<const bf16*>(x), y, n);
    count_launch();
    return check_launch("cast_bf16_to_f32_kernel");
}
int fill_f32(cudaStream_t s, float* x, size_t n, float v) {
    fill_f32_kernel<<<g1(n), 256, 0, s>>>(x, n, v);
    count_launch();
    return check_launch("fill_f32_kernel");
}
int scale_f32(cudaStream_t s, float* x, size_t n, float v) {
    scale_f32_kernel<<<g1(n), 256, 0, s>>>(x, n, v);
    count_launch();
    return check_launch("scale_f32_kernel");
}

}  // namespace sdod
