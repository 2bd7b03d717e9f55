// Memory-bound helpers of the hot path: LayerNorm, row softmax, layout conversion, nearest-2x
// upsample, channel concat, im2col (conv_in / stride-2 downsample), casts, weight packing.
// All bf16 tensors are channels-last; every kernel moves 16-byte vectors where alignment allows.
#include <cstdlib>
#include "../common.cuh"
#include "../host_common.h"
#include "../launch_count.h"
#include "sdod_kernels.h"

namespace sdod {

static inline int ew_grid(size_t work_items, int block) {
    size_t g = (work_items + block - 1) / block;
    const size_t cap = 148 * 32;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

SDOD_DEVICE void unpack8(const uint4& u, float* f) {
    float2 t;
    t = unpack_bf16x2(u.x); f[0] = t.x; f[1] = t.y;
    t = unpack_bf16x2(u.y); f[2] = t.x; f[3] = t.y;
    t = unpack_bf16x2(u.z); f[4] = t.x; f[5] = t.y;
    t = unpack_bf16x2(u.w); f[6] = t.x; f[7] = t.y;
}
SDOD_DEVICE uint4 pack8(const float* f) {
    uint4 u;
    u.x = pack_bf16x2(f[0], f[1]); u.y = pack_bf16x2(f[2], f[3]);
    u.z = pack_bf16x2(f[4], f[5]); u.w = pack_bf16x2(f[6], f[7]);
    return u;
}

// ---------------------------------------------------------------- LayerNorm: one warp per row
constexpr int kLnMaxVecPerLane = 8;   // width <= 32*8*8 = 2048 (instantiated per width class to keep registers low)
template <typename T> SDOD_DEVICE void ld8(const T* p, float* f);
template <> SDOD_DEVICE void ld8<bf16>(const bf16* p, float* f) { unpack8(*reinterpret_cast<const uint4*>(p), f); }
template <> SDOD_DEVICE void ld8<float>(const float* p, float* f) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// One warp normalises R consecutive rows at once: all R x NV 32-byte loads are issued before the first reduction, so a warp
// keeps R rows in flight (with one row per warp the kernel ran at ~2.3 TB/s, bound by load latency and CTA turnover).
// (Round 2: a <= 64-register form with half the rows per warp and four CTAs per SM was slower at batch 32 — 77.5 vs 75.2 us at 131,072 x 320,
// 57.7 vs 45.0 us at 32,768 x 640 — rows in flight per warp matter more than resident warps; dropped.)
template <typename T, int NV, int R>
__global__ void __launch_bounds__(256) layer_norm_kernel(const T* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ w,
                                                         const float* __restrict__ b, int rows, int width, float eps) {
    griddep_wait();
    griddep_launch();
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    const int row0 = warp * R;
    if (row0 >= rows) return;
    const int nvec = width >> 3;
    float v[R][NV][8];
    float s[R];
#pragma unroll
    for (int r = 0; r < R; ++r) {
        const T* xr = x + static_cast<size_t>(row0 + r) * width;
        const bool row_ok = row0 + r < rows;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            const int i = lane + k * 32;
            if (row_ok && i < nvec) {
                ld8<T>(xr + i * 8, v[r][k]);
            } else {
#pragma unroll
                for (int j = 0; j < 8; ++j) v[r][k][j] = 0.f;
            }
        }
    }
#pragma unroll
    for (int r = 0; r < R; ++r) {
        s[r] = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
#pragma unroll
            for (int j = 0; j < 8; ++j) s[r] += v[r][k][j];
        }
    }
    float mean[R], rstd[R];
#pragma unroll
    for (int r = 0; r < R; ++r) mean[r] = warp_sum(s[r]) / static_cast<float>(width);
#pragma unroll
    for (int r = 0; r < R; ++r) {
        float q = 0.f;
#pragma unroll
        for (int k = 0; k < NV; ++k) {
            if (lane + k * 32 < nvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float d = v[r][k][j] - mean[r]; q += d * d; }
            }
        }
        s[r] = q;
    }
#pragma unroll
    for (int r = 0; r < R; ++r) rstd[r] = rsqrtf(warp_sum(s[r]) / static_cast<float>(width) + eps);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
        const int i = lane + k * 32;
        if (i < nvec) {
            float wv[8], bv[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) { wv[j] = w ? w[i * 8 + j] : 1.f; bv[j] = b ? b[i * 8 + j] : 0.f; }
#pragma unroll
            for (int r = 0; r < R; ++r) {
                if (row0 + r < rows) {
                    float o[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) o[j] = (v[r][k][j] - mean[r]) * rstd[r] * wv[j] + bv[j];
                    reinterpret_cast<uint4*>(y + static_cast<size_t>(row0 + r) * width)[i] = pack8(o);
                }
            }
        }
    }
}

// ---------------------------------------------------------------- row softmax (bf16 in/out, fp32 math)
__global__ void __launch_bounds__(256) softmax_rows_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, long long rows, int cols,
                                                           long long ld, float scale) {
    __shared__ float red[8];
    __shared__ float bc;
    for (long long r = blockIdx.x; r < rows; r += gridDim.x) {
        const bf16* xr = x + r * ld;
        bf16* yr = y + r * ld;
        float m = -INFINITY;
        for (int c = threadIdx.x; c < cols; c += 256) m = fmaxf(m, __bfloat162float(xr[c]) * scale);
        m = warp_max(m);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = m;
        __syncthreads();
        if (threadIdx.x == 0) { float t = red[0]; for (int i = 1; i < 8; ++i) t = fmaxf(t, red[i]); bc = t; }
        __syncthreads();
        m = bc;
        float s = 0.f;
        for (int c = threadIdx.x; c < cols; c += 256) s += __expf(__bfloat162float(xr[c]) * scale - m);
        s = warp_sum(s);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) { float t = 0.f; for (int i = 0; i < 8; ++i) t += red[i]; bc = t; }
        __syncthreads();
        const float inv = 1.0f / bc;
        for (int c = threadIdx.x; c < cols; c += 256) yr[c] = __float2bfloat16(__expf(__bfloat162float(xr[c]) * scale - m) * inv);
    }
}

// ---------------------------------------------------------------- layout conversion (latents / images)
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int N, int C, int HW) {
    const size_t total = static_cast<size_t>(N) * C * HW;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int c = o % C;
        const size_t p = o / C;
        const int hw = p % HW;
        const int n = p / HW;
        y[o] = __float2bfloat16(x[(static_cast<size_t>(n) * C + c) * HW + hw]);
    }
}
template <typename T>
__global__ void nhwc_to_nchw_f32_kernel(const T* __restrict__ x, float* __restrict__ y, int N, int C, int HW) {
    const size_t total = static_cast<size_t>(N) * C * HW;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int hw = o % HW;
        const size_t p = o / HW;
        const int c = p % C;
        const int n = p / C;
        const T v = x[(static_cast<size_t>(n) * HW + hw) * C + c];
        if constexpr (sizeof(T) == 4) y[o] = v; else y[o] = __bfloat162float(v);
    }
}

// ---------------------------------------------------------------- nearest 2x upsample, NHWC, 16-B vectors
template <typename T>
__global__ void upsample2x_kernel(const T* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int cvec) {
    const size_t total = static_cast<size_t>(N) * (2 * H) * (2 * W) * cvec;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cv = o % cvec;
        size_t p = o / cvec;
        const int ox = p % (2 * W); p /= (2 * W);
        const int oy = p % (2 * H);
        const int n = p / (2 * H);
        float f[8];
        ld8<T>(x + (((static_cast<size_t>(n) * H + (oy >> 1)) * W + (ox >> 1)) * cvec + cv) * 8, f);
        y[o] = pack8(f);
    }
}

// ---------------------------------------------------------------- channel concat [rows,Ca] ++ [rows,Cb]
__global__ void concat_kernel(const uint4* __restrict__ a, int va, const uint4* __restrict__ b, int vb, uint4* __restrict__ y, long long rows) {
    griddep_wait();
    griddep_launch();
    const int vt = va + vb;
    const size_t total = static_cast<size_t>(rows) * vt;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int v = o % vt;
        const size_t r = o / vt;
        y[o] = v < va ? a[r * va + v] : b[r * vb + (v - va)];
    }
}

// ---------------------------------------------------------------- im2col 3x3 pad 1 (stride 1|2), k = (ky*3+kx)*C + c
template <typename T>
__global__ void im2col3x3_kernel(const T* __restrict__ x, bf16* __restrict__ y, int N, int H, int W, int C, int stride, int Kpad) {
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const size_t total = static_cast<size_t>(N) * Ho * Wo * Kpad;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = o % Kpad;
        size_t p = o / Kpad;
        const int ox = p % Wo; p /= Wo;
        const int oy = p % Ho;
        const int n = p / Ho;
        bf16 v = __float2bfloat16(0.f);
        if (k < 9 * C) {
            const int tap = k / C, c = k - tap * C;
            const int iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
            if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
                const T e = x[((static_cast<size_t>(n) * H + iy) * W + ix) * C + c];
                if constexpr (sizeof(T) == 4) v = __float2bfloat16(e); else v = e;
            }
        }
        y[o] = v;
    }
}
// vectorised variant (C % 8 == 0, Kpad == 9*C)
template <typename T>
__global__ void im2col3x3_vec_kernel(const T* __restrict__ x, uint4* __restrict__ y, int N, int H, int W, int cvec, int stride) {
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    const size_t total = static_cast<size_t>(N) * Ho * Wo * 9 * cvec;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int cv = o % cvec;
        size_t p = o / cvec;
        const int tap = p % 9; p /= 9;
        const int ox = p % Wo; p /= Wo;
        const int oy = p % Ho;
        const int n = p / Ho;
        const int iy = oy * stride + tap / 3 - 1, ix = ox * stride + tap % 3 - 1;
        uint4 v = make_uint4(0, 0, 0, 0);
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) {
            float f[8];
            ld8<T>(x + (((static_cast<size_t>(n) * H + iy) * W + ix) * cvec + cv) * 8, f);
            v = pack8(f);
        }
        y[o] = v;
    }
}

__global__ void cast_f32_to_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, size_t n) {
    griddep_wait();
    griddep_launch();
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x, nth = static_cast<size_t>(gridDim.x) * blockDim.x;
    const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
    const size_t n8 = vec ? n / 8 : 0;                      // 8 elements per thread: two 16-B loads, one 16-B store
    for (size_t i = tid; i < n8; i += nth) {
        const float4 lo = reinterpret_cast<const float4*>(x)[2 * i], hi = reinterpret_cast<const float4*>(x)[2 * i + 1];
        float f[8] = {lo.x, lo.y, lo.z, lo.w, hi.x, hi.y, hi.z, hi.w};
        reinterpret_cast<uint4*>(y)[i] = pack8(f);
    }
    for (size_t i = n8 * 8 + tid; i < n; i += nth) y[i] = __float2bfloat16(x[i]);
}
__global__ void silu_bf16_kernel(const bf16* __restrict__ x, bf16* __restrict__ y, size_t n) {
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        y[i] = __float2bfloat16(silu_f(__bfloat162float(x[i])));
}

// OIHW fp32 -> [O][ky][kx][I] bf16, row length Kpad (zero padded)
__global__ void pack_conv3x3_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin, int Kpad) {
    const size_t total = static_cast<size_t>(Cout) * Kpad;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int k = o % Kpad;
        const int co = o / Kpad;
        float v = 0.f;
        if (k < 9 * Cin) {
            const int tap = k / Cin, ci = k - tap * Cin;
            v = w[(static_cast<size_t>(co) * Cin + ci) * 9 + tap];
        }
        out[o] = __float2bfloat16(v);
    }
}

// OIHW fp32 -> [parity][O][2a+b][I] bf16: the sub-pixel form of conv3x3 after a nearest 2x upsample.  Output row 2i+py reads source rows
// i-1+py (a=0) and i+py (a=1); the 3x3 rows ky that land on them are {0} | {1,2} for py=0 and {0,1} | {2} for py=1 (same for columns).
__global__ void pack_conv3x3_up2_kernel(const float* __restrict__ w, bf16* __restrict__ out, int Cout, int Cin) {
    const size_t per = static_cast<size_t>(Cout) * 4 * Cin, total = 4 * per;
    for (size_t o = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; o < total; o += static_cast<size_t>(gridDim.x) * blockDim.x) {
        const int par = static_cast<int>(o / per);
        const size_t r = o - par * per;
        const int co = static_cast<int>(r / (4 * Cin)), k = static_cast<int>(r - static_cast<size_t>(co) * 4 * Cin);
        const int tap = k / Cin, ci = k - tap * Cin;
        const int py = par >> 1, px = par & 1, a = tap >> 1, b = tap & 1;
        const int ky0 = py == 0 ? (a == 0 ? 0 : 1) : (a == 0 ? 0 : 2), ky1 = py == 0 ? (a == 0 ? 0 : 2) : (a == 0 ? 1 : 2);
        const int kx0 = px == 0 ? (b == 0 ? 0 : 1) : (b == 0 ? 0 : 2), kx1 = px == 0 ? (b == 0 ? 0 : 2) : (b == 0 ? 1 : 2);
        float v = 0.f;
        for (int ky = ky0; ky <= ky1; ++ky)
            for (int kx = kx0; kx <= kx1; ++kx) v += w[(static_cast<size_t>(co) * Cin + ci) * 9 + ky * 3 + kx];
        out[o] = __float2bfloat16(v);
    }
}

}  // namespace sdod

using namespace sdod;
#define ST(s) static_cast<cudaStream_t>(s)

extern "C" {

SDOD_API int sdod_layer_norm(sdod_stream_t stream, const void* x, int in_dtype, void* y, const float* weight, const float* bias, int rows, int width, float eps) {
    if (!x || !y || rows <= 0 || width <= 0) return fail(kInvalidArgument, "layer_norm: bad arguments");
    if (width % 8 != 0 || width > 32 * 8 * kLnMaxVecPerLane) return fail(kUnsupported, "layer_norm: width must be a multiple of 8 and <= 2048");
    const int warps_per_block = 8;
    const int grid = (rows + warps_per_block - 1) / warps_per_block;
    const int nv = (width / 8 + 31) / 32;
#define SDOD_LN(NV, R)                                                                                                                           \
    do {                                                                                                                                         \
        const unsigned grid = static_cast<unsigned>((rows + 8 * R - 1) / (8 * R));                                                               \
        if (in_dtype == SDOD_F32)                                                                                                                \
            launch_pdl(layer_norm_kernel<float, NV, R>, dim3(grid), dim3(256), 0, ST(stream), static_cast<const float*>(x), static_cast<bf16*>(y), weight, bias, rows, width, eps); \
        else                                                                                                                                     \
            launch_pdl(layer_norm_kernel<bf16, NV, R>, dim3(grid), dim3(256), 0, ST(stream), static_cast<const bf16*>(x), static_cast<bf16*>(y), weight, bias, rows, width, eps);  \
    } while (0)
    if (nv <= 2) SDOD_LN(2, 4);
    else if (nv <= 3) SDOD_LN(3, 2);
    else if (nv <= 5) SDOD_LN(5, 2);
    else SDOD_LN(8, 1);
#undef SDOD_LN
    count_launch();
    return check_launch("layer_norm_kernel");
}

SDOD_API int sdod_softmax_rows(sdod_stream_t stream, const void* x, void* y, long long rows, int cols, long long ld, float scale) {
    if (!x || !y || rows <= 0 || cols <= 0) return fail(kInvalidArgument, "softmax_rows: bad arguments");
    const long long grid = rows < 148 * 16 ? rows : 148 * 16;
    softmax_rows_kernel<<<static_cast<unsigned>(grid), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), rows, cols, ld, scale);
    count_launch();
    return check_launch("softmax_rows_kernel");
}

SDOD_API int sdod_nchw_f32_to_nhwc_bf16(sdod_stream_t stream, const float* x, void* y, int N, int C, int HW) {
    if (!x || !y) return fail(kInvalidArgument, "nchw_f32_to_nhwc_bf16: NULL tensor");
    const size_t n = static_cast<size_t>(N) * C * HW;
    nchw_f32_to_nhwc_bf16_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(x, static_cast<bf16*>(y), N, C, HW);
    count_launch();
    return check_launch("nchw_f32_to_nhwc_bf16_kernel");
}

SDOD_API int sdod_nhwc_to_nchw_f32(sdod_stream_t stream, const void* x, int dtype, float* y, int N, int C, int HW) {
    if (!x || !y) return fail(kInvalidArgument, "nhwc_to_nchw_f32: NULL tensor");
    const size_t n = static_cast<size_t>(N) * C * HW;
    if (dtype == SDOD_F32) nhwc_to_nchw_f32_kernel<float><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const float*>(x), y, N, C, HW);
    else nhwc_to_nchw_f32_kernel<bf16><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), y, N, C, HW);
    count_launch();
    return check_launch("nhwc_to_nchw_f32_kernel");
}

SDOD_API int sdod_upsample2x_nhwc(sdod_stream_t stream, const void* x, int in_dtype, void* y, int N, int H, int W, int C) {
    if (!x || !y || C % 8 != 0) return fail(kInvalidArgument, "upsample2x: NULL tensor or C % 8 != 0");
    const size_t n = static_cast<size_t>(N) * 4 * H * W * (C / 8);
    if (in_dtype == SDOD_F32)
        upsample2x_kernel<float><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const float*>(x), static_cast<uint4*>(y), N, H, W, C / 8);
    else
        upsample2x_kernel<bf16><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<uint4*>(y), N, H, W, C / 8);
    count_launch();
    return check_launch("upsample2x_kernel");
}

SDOD_API int sdod_concat_channels(sdod_stream_t stream, const void* a, int Ca, const void* b, int Cb, void* y, long long rows, int dtype) {
    if (!a || !b || !y || Ca % 8 != 0 || Cb % 8 != 0) return fail(kInvalidArgument, "concat_channels: NULL tensor or C % 8 != 0");
    const int per = dtype == SDOD_F32 ? 4 : 8;     // elements per 16-byte unit
    const size_t n = static_cast<size_t>(rows) * ((Ca + Cb) / per);
    if (launch_pdl(concat_kernel, dim3(ew_grid(n, 256)), dim3(256), 0, ST(stream), static_cast<const uint4*>(a), Ca / per, static_cast<const uint4*>(b),
                   Cb / per, static_cast<uint4*>(y), rows) != cudaSuccess)
        return check_launch("concat_kernel");
    count_launch();
    return check_launch("concat_kernel");
}

SDOD_API int sdod_im2col3x3(sdod_stream_t stream, const void* x, int in_dtype, void* y, int N, int H, int W, int C, int stride, int Kpad) {
    if (!x || !y || (stride != 1 && stride != 2) || Kpad < 9 * C) return fail(kInvalidArgument, "im2col3x3: bad arguments");
    const int Ho = (H + stride - 1) / stride, Wo = (W + stride - 1) / stride;
    if (C % 8 == 0 && Kpad == 9 * C) {
        const size_t n = static_cast<size_t>(N) * Ho * Wo * 9 * (C / 8);
        if (in_dtype == SDOD_F32)
            im2col3x3_vec_kernel<float><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const float*>(x), static_cast<uint4*>(y), N, H, W, C / 8, stride);
        else
            im2col3x3_vec_kernel<bf16><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<uint4*>(y), N, H, W, C / 8, stride);
    } else {
        const size_t n = static_cast<size_t>(N) * Ho * Wo * Kpad;
        if (in_dtype == SDOD_F32)
            im2col3x3_kernel<float><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const float*>(x), static_cast<bf16*>(y), N, H, W, C, stride, Kpad);
        else
            im2col3x3_kernel<bf16><<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), N, H, W, C, stride, Kpad);
    }
    count_launch();
    return check_launch("im2col3x3_kernel");
}

SDOD_API int sdod_cast_f32_to_bf16(sdod_stream_t stream, const float* x, void* y, size_t n) {
    if (!x || !y) return fail(kInvalidArgument, "cast_f32_to_bf16: NULL tensor");
    if (n == 0) return kOk;
    if (launch_pdl(cast_f32_to_bf16_kernel, dim3(ew_grid((n + 7) / 8, 256)), dim3(256), 0, ST(stream), x, static_cast<bf16*>(y), n) != cudaSuccess)
        return check_launch("cast_f32_to_bf16_kernel");
    count_launch();
    return check_launch("cast_f32_to_bf16_kernel");
}

SDOD_API int sdod_silu_bf16(sdod_stream_t stream, const void* x, void* y, size_t n) {
    if (!x || !y) return fail(kInvalidArgument, "silu_bf16: NULL tensor");
    if (n == 0) return kOk;
    silu_bf16_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(static_cast<const bf16*>(x), static_cast<bf16*>(y), n);
    count_launch();
    return check_launch("silu_bf16_kernel");
}

SDOD_API int sdod_pack_conv3x3_weight(sdod_stream_t stream, const float* w_oihw, void* out, int Cout, int Cin, int Kpad) {
    if (!w_oihw || !out || Kpad < 9 * Cin) return fail(kInvalidArgument, "pack_conv3x3_weight: bad arguments");
    const size_t n = static_cast<size_t>(Cout) * Kpad;
    pack_conv3x3_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(w_oihw, static_cast<bf16*>(out), Cout, Cin, Kpad);
    count_launch();
    return check_launch("pack_conv3x3_kernel");
}

SDOD_API int sdod_pack_conv3x3_up2_weight(sdod_stream_t stream, const float* w_oihw, void* out, int Cout, int Cin) {
    if (!w_oihw || !out || Cout <= 0 || Cin <= 0) return fail(kInvalidArgument, "pack_conv3x3_up2_weight: bad arguments");
    const size_t n = static_cast<size_t>(4) * Cout * 4 * Cin;
    pack_conv3x3_up2_kernel<<<ew_grid(n, 256), 256, 0, ST(stream)>>>(w_oihw, static_cast<bf16*>(out), Cout, Cin);
    count_launch();
    return check_launch("pack_conv3x3_up2_kernel");
}

}  // extern "C"
