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

// LayerNorm for large row counts (UNet batch >= 8): persistent CTAs stream row blocks through shared memory with 1-D TMA bulk copies.
// Warp 0 keeps a 3-deep ring of row blocks in flight (mbarrier complete_tx); each of the 8 consumer warps owns a contiguous slice of the
// block's rows, normalises them out of shared memory (16-byte conflict-free reads, two-pass variance in registers), writes the bf16 rows into
// its own double-buffered output slice and hands that slice to a bulk store itself — no block-wide barrier anywhere in the loop.  The
// register-staged kernel above spends its time in CTA turnover at these sizes (39 % of DRAM bandwidth at 131,072 x 320, ncu).
constexpr int kLnTmaStages = 3;
constexpr int kLnTmaThreads = 288;          // warp 0: producer; warps 1..8: consumers
template <typename T, int NJ, int RI = (NJ <= 3 ? 4 : (NJ <= 5 ? 2 : 1))>
__global__ void __launch_bounds__(kLnTmaThreads, 1) layer_norm_tma_kernel(const T* __restrict__ x, bf16* __restrict__ y, const float* __restrict__ w,
                                                                          const float* __restrict__ b, int rows, int width, float eps, int rb, int n_blocks) {
    extern __shared__ __align__(128) uint8_t ln_smem[];
    __shared__ __align__(8) uint64_t full_bar[kLnTmaStages], empty_bar[kLnTmaStages];
    const uint32_t in_bytes = static_cast<uint32_t>(rb) * width * sizeof(T), out_bytes = static_cast<uint32_t>(rb) * width * 2;
    uint8_t* s_in = ln_smem;
    uint8_t* s_out = ln_smem + kLnTmaStages * in_bytes;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < kLnTmaStages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 8); }
        fence_mbar_init();
    }
    __syncthreads();
    griddep_wait();
    griddep_launch();
    if (warp == 0) {
        if (lane == 0) {
            int g = 0;
            for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++g) {
                const int s = g % kLnTmaStages;
                mbar_wait(&empty_bar[s], ((g / kLnTmaStages) & 1) ^ 1);
                const int r0 = blk * rb, nr = min(rb, rows - r0);
                const uint32_t bytes = static_cast<uint32_t>(nr) * width * sizeof(T);
                mbar_arrive_expect_tx(&full_bar[s], bytes);
                bulk_load_1d(s_in + s * in_bytes, x + static_cast<size_t>(r0) * width, bytes, &full_bar[s]);
            }
        }
        return;
    }
    const int cw = warp - 1;                          // consumer warp 0..7
    const int rpw = rb / 8;                           // rows per warp and block (rb % 8 == 0)
    const int nq = width >> 2;                        // 4-element units per row
    float wv[NJ][4], bv[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
        const int q = lane + 32 * j;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            wv[j][i] = (q < nq && w) ? w[4 * q + i] : 1.f;
            bv[j][i] = (q < nq && b) ? b[4 * q + i] : 0.f;
        }
    }
    const float inv_n = 1.0f / static_cast<float>(width);
    int g = 0;
    for (int blk = blockIdx.x; blk < n_blocks; blk += gridDim.x, ++g) {
        const int s = g % kLnTmaStages, o = g & 1;
        const int r0 = blk * rb, nr = min(rb, rows - r0);
        mbar_wait(&full_bar[s], (g / kLnTmaStages) & 1);
        const T* in = reinterpret_cast<const T*>(s_in + s * in_bytes);
        bf16* out = reinterpret_cast<bf16*>(s_out + o * out_bytes);
        const int rw0 = cw * rpw, rw1 = min(nr, rw0 + rpw);
        // RI rows at a time per warp: the per-row chain (load -> reduce -> mean -> reduce -> rstd -> store) is latency-bound, two rows in flight hide it
        for (int r = rw0; r < rw1; r += RI) {
            float v[RI][NJ][4];
            float sum[RI];
#pragma unroll
            for (int u = 0; u < RI; ++u) {
                sum[u] = 0.f;
                const bool row_ok = r + u < rw1;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int q = lane + 32 * j;
                    if (row_ok && q < nq) {
                        if (sizeof(T) == 4) {
                            const float4 t = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(in) + static_cast<size_t>(r + u) * width + 4 * q);
                            v[u][j][0] = t.x; v[u][j][1] = t.y; v[u][j][2] = t.z; v[u][j][3] = t.w;
                        } else {
                            const uint2 t = *reinterpret_cast<const uint2*>(reinterpret_cast<const bf16*>(in) + static_cast<size_t>(r + u) * width + 4 * q);
                            const float2 a = unpack_bf16x2(t.x), c = unpack_bf16x2(t.y);
                            v[u][j][0] = a.x; v[u][j][1] = a.y; v[u][j][2] = c.x; v[u][j][3] = c.y;
                        }
                        sum[u] += (v[u][j][0] + v[u][j][1]) + (v[u][j][2] + v[u][j][3]);
                    } else {
                        v[u][j][0] = v[u][j][1] = v[u][j][2] = v[u][j][3] = 0.f;
                    }
                }
            }
            float mean[RI], rstd[RI];
#pragma unroll
            for (int u = 0; u < RI; ++u) mean[u] = warp_sum(sum[u]) * inv_n;
#pragma unroll
            for (int u = 0; u < RI; ++u) {
                float q2 = 0.f;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    if (lane + 32 * j < nq) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) { const float d = v[u][j][i] - mean[u]; q2 += d * d; }
                    }
                }
                sum[u] = q2;
            }
#pragma unroll
            for (int u = 0; u < RI; ++u) rstd[u] = rsqrtf(warp_sum(sum[u]) * inv_n + eps);
#pragma unroll
            for (int u = 0; u < RI; ++u) {
                if (r + u >= rw1) break;
#pragma unroll
                for (int j = 0; j < NJ; ++j) {
                    const int q = lane + 32 * j;
                    if (q < nq) {
                        uint2 pk;
                        pk.x = pack_bf16x2((v[u][j][0] - mean[u]) * rstd[u] * wv[j][0] + bv[j][0], (v[u][j][1] - mean[u]) * rstd[u] * wv[j][1] + bv[j][1]);
                        pk.y = pack_bf16x2((v[u][j][2] - mean[u]) * rstd[u] * wv[j][2] + bv[j][2], (v[u][j][3] - mean[u]) * rstd[u] * wv[j][3] + bv[j][3]);
                        *reinterpret_cast<uint2*>(out + static_cast<size_t>(r + u) * width + 4 * q) = pk;
                    }
                }
            }
        }
        fence_proxy_async_smem();                       // this lane's output writes -> visible to the bulk-copy engine
        __syncwarp();
        if (lane == 0) {
            mbar_arrive(&empty_bar[s]);                 // the warp has read its rows of the input stage
            if (rw1 > rw0) {
                bulk_store_1d(y + (static_cast<size_t>(r0) + rw0) * width, out + static_cast<size_t>(rw0) * width, static_cast<uint32_t>(rw1 - rw0) * width * 2);
                bulk_commit();
            }
            if (rw1 > rw0) bulk_wait_read_but_last();   // the slice of the OTHER output buffer (stored one block ago) may be rewritten
            else bulk_wait_read_all();                  // (no store this time: nothing may stay pending on either buffer)
        }
        __syncwarp();
    }
    if (lane == 0) bulk_wait_read_all();
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
    {
        // large launches: the TMA-pipelined persistent kernel (row blocks of ~40 KB through a 3-stage shared-memory ring)
        static const int tma_env = [] { const char* e = std::getenv("SDOD_LN_TMA"); return e ? std::atoi(e) : 1; }();     // 0: register-staged kernel only (A/B)
        const size_t es = in_dtype == SDOD_F32 ? 4 : 2;
        const bool aligned = ((reinterpret_cast<uintptr_t>(x) | reinterpret_cast<uintptr_t>(y)) & 15) == 0 && (!weight || (reinterpret_cast<uintptr_t>(weight) & 3) == 0);
        if (tma_env && rows >= 16384 && aligned && width <= 2048) {
            int rb = static_cast<int>((40 * 1024) / (width * es)) & ~7;
            if (rb >= 8) {
                if (rb > 64) rb = 64;
                const int n_blocks = (rows + rb - 1) / rb;
                const size_t smem = kLnTmaStages * static_cast<size_t>(rb) * width * es + 2 * static_cast<size_t>(rb) * width * 2;
                const int pgrid = n_blocks < device_sm_count() ? n_blocks : device_sm_count();
                const int nj = (width / 4 + 31) / 32;
#define SDOD_LNT(T, NJ)                                                                                                                              \
    do {                                                                                                                                             \
        static bool cfgd = false;                                                                                                                    \
        if (!cfgd) { SDOD_TRY(check_cuda(cudaFuncSetAttribute(layer_norm_tma_kernel<T, NJ>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024), "cudaFuncSetAttribute(ln tma)")); cfgd = true; } \
        launch_pdl(layer_norm_tma_kernel<T, NJ>, dim3(pgrid), dim3(kLnTmaThreads), smem, ST(stream), static_cast<const T*>(x), static_cast<bf16*>(y), weight, bias, rows, width, eps, rb, n_blocks); \
        count_launch();                                                                                                                              \
        return check_launch("layer_norm_tma_kernel");                                                                                                \
    } while (0)
                if (smem <= 200 * 1024) {
                    if (in_dtype == SDOD_F32) {
                        if (nj <= 3) SDOD_LNT(float, 3); else if (nj <= 5) SDOD_LNT(float, 5); else if (nj <= 10) SDOD_LNT(float, 10); else SDOD_LNT(float, 16);
                    } else {
                        if (nj <= 3) SDOD_LNT(bf16, 3); else if (nj <= 5) SDOD_LNT(bf16, 5); else if (nj <= 10) SDOD_LNT(bf16, 10); else SDOD_LNT(bf16, 16);
                    }
                }
#undef SDOD_LNT
            }
        }
    }
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
