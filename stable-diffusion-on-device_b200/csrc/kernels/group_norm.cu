// GroupNorm (+SiLU, + per-(n,c) timestep-embedding add) for sm_100a.
//
// Contract restated from the reference operator surface: sdod/efficient_gn.py:9-12 (forward ==
// F.group_norm), :29-30, :61-70 and the op schema csrc/sdod_ops/config/group_norm.xml:16-107
// (NCHW) / group_norm.json:6-21 (NHWC):  biased variance over (C/G)*HW per (n,g), affine per channel.
//
// Three kernels:
//  * gn_nchw_cluster_kernel  — single pass over HBM.  A thread-block cluster owns one (n,g) group
//    (contiguous in NCHW); each CTA stages its slab in shared memory with TMA bulk copies
//    (cp.async.bulk + mbarrier, 4 chunks in flight), accumulates pivot-shifted moments with 128-bit
//    smem reads + warp shuffles, converts them to (count, mean, M2) and merges across the cluster
//    with Chan/Welford updates over distributed shared memory, then normalises out of smem.
//    Algorithmic traffic = read x once + write y once.
//  * gn_nhwc_stats_kernel + gn_nhwc_apply_kernel — channels-last path used inside the UNet/VAE.
//    Stats are pivot-shifted sums (pivot = first element of the group, shared by all CTAs, so partial
//    sums add exactly like Welford partials with a common origin); the last CTA of each image folds
//    the partials in fixed order (deterministic) and publishes mean / rstd.
//  * gn_nchw_generic_kernel  — any shape/alignment (two reads), used when the TMA path's
//    alignment or capacity preconditions do not hold.
#include <cstdlib>
#include "../common.cuh"
#include "../host_common.h"
#include "../launch_count.h"
#include "sdod_kernels.h"

#include <algorithm>

namespace sdod {

template <typename T> struct VecOf;
template <> struct VecOf<float> { static constexpr int N = 4; };
template <> struct VecOf<bf16> { static constexpr int N = 8; };

template <typename T> SDOD_DEVICE void load_vec(const T* p, float* out);
template <> SDOD_DEVICE void load_vec<float>(const float* p, float* out) {
    float4 v = *reinterpret_cast<const float4*>(p);
    out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
}
template <> SDOD_DEVICE void load_vec<bf16>(const bf16* p, float* out) {
    uint4 u = *reinterpret_cast<const uint4*>(p);
    float2 f;
    f = unpack_bf16x2(u.x); out[0] = f.x; out[1] = f.y;
    f = unpack_bf16x2(u.y); out[2] = f.x; out[3] = f.y;
    f = unpack_bf16x2(u.z); out[4] = f.x; out[5] = f.y;
    f = unpack_bf16x2(u.w); out[6] = f.x; out[7] = f.y;
}
template <typename T> SDOD_DEVICE void store_vec(T* p, const float* v);
template <> SDOD_DEVICE void store_vec<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
}
template <> SDOD_DEVICE void store_vec<bf16>(bf16* p, const float* v) {
    uint4 u;
    u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
    u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
    *reinterpret_cast<uint4*>(p) = u;
}
template <typename T> SDOD_DEVICE float to_f(T v);
template <> SDOD_DEVICE float to_f<float>(float v) { return v; }
template <> SDOD_DEVICE float to_f<bf16>(bf16 v) { return __bfloat162float(v); }
template <typename T> SDOD_DEVICE T from_f(float v);
template <> SDOD_DEVICE float from_f<float>(float v) { return v; }
template <> SDOD_DEVICE bf16 from_f<bf16>(float v) { return __float2bfloat16(v); }


// Chan et al. parallel Welford merge of (n, mean, M2)
SDOD_DEVICE void welford_merge(float& n, float& mean, float& m2, float nb, float meanb, float m2b) {
    if (nb == 0.f) return;
    float nt = n + nb;
    float delta = meanb - mean;
    float f = nb / nt;
    mean += delta * f;
    m2 += m2b + delta * delta * n * f;
    n = nt;
}

constexpr int kGnThreads = 256;
constexpr int kGnMaxChunks = 4;

constexpr int kGnClusterThreads = 512;
constexpr int kGnMaxCpgSmem = 256;      // per-channel scale / shift staged in shared memory up to this many channels per group

// Round 2: the round-1 form of this kernel spent its time on address arithmetic, not on memory — a 64-bit division per vector to find the
// channel, per-element weight / bias loads, 16-byte global stores from 256 threads (13.4 us warm for the 21 MB of config C1).  Now: 512
// threads, the channel walked incrementally, per-channel scale / shift folded with mean / rstd once into shared memory, the normalised slab
// written back IN PLACE and handed to TMA bulk stores (cp.async.bulk shared -> global, one per chunk), the second cluster barrier moved to the
// end of the kernel (it only protects the peers' DSMEM reads), and programmatic dependent launch.
template <typename T>
__global__ void __launch_bounds__(kGnClusterThreads) gn_nchw_cluster_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                             const float* __restrict__ weight, const float* __restrict__ bias,
                                                                             const float* __restrict__ add_nc, int C, int HW, int G, float eps,
                                                                             int fuse_silu, int slab_elems) {
    constexpr int VEC = VecOf<T>::N;
    constexpr int NT = kGnClusterThreads;
    extern __shared__ __align__(128) uint8_t gn_smem[];
    T* slab = reinterpret_cast<T*>(gn_smem);
    __shared__ __align__(8) uint64_t bars[kGnMaxChunks];
    __shared__ float red_s[NT / 32], red_ss[NT / 32];
    __shared__ float cta_stats[4];   // n, mean, M2
    __shared__ __align__(16) float peer_stats[16][4];   // [rank]{n, mean, M2, -}: pushed by every CTA of the cluster (st.async)
    __shared__ __align__(8) uint64_t xbar;
    __shared__ float pivot_sh;
    __shared__ float sc_sh[kGnMaxCpgSmem], sh_sh[kGnMaxCpgSmem];

    const uint32_t cs = cluster_nctarank(), rank = cluster_ctarank();
    const int group = blockIdx.x / cs;
    const int n = group / G, g = group - n * G;
    const int cpg = C / G;
    const long long L = static_cast<long long>(cpg) * HW;
    const long long gbase = (static_cast<long long>(n) * C + static_cast<long long>(g) * cpg) * HW;
    const uint32_t my_off = rank * static_cast<uint32_t>(slab_elems);      // < cpg * HW, which the launcher keeps below 2^31
    const T* src = x + gbase + my_off;
    const uint32_t slab_bytes = static_cast<uint32_t>(slab_elems) * sizeof(T);
    const int nch = slab_bytes >= 16384 ? kGnMaxChunks : 1;
    const uint32_t chunk_bytes = ((slab_bytes / nch) + 15) & ~15u;
    const float* addp = add_nc ? add_nc + static_cast<long long>(n) * C + g * cpg : nullptr;

    if (threadIdx.x == 0) {
        for (int c = 0; c < nch; ++c) mbar_init(&bars[c], 1);
        mbar_init(&xbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    cluster_arrive_relaxed();             // "my mbarriers exist": waited for just before the statistics push
    griddep_wait();                       // PDL: x may be the previous kernel's output
    if (threadIdx.x == 0) {
        for (int c = 0; c < nch; ++c) {
            uint32_t off = c * chunk_bytes;
            uint32_t bytes = (c == nch - 1) ? slab_bytes - off : chunk_bytes;
            mbar_arrive_expect_tx(&bars[c], bytes);
            bulk_load_1d(reinterpret_cast<uint8_t*>(slab) + off, reinterpret_cast<const uint8_t*>(src) + off, bytes, &bars[c]);
        }
    }
    griddep_launch();
    // ---- pass A (smem): pivot-shifted moments
    mbar_wait(&bars[0], 0);
    if (threadIdx.x == 0) pivot_sh = to_f<T>(slab[0]) + (addp ? addp[my_off / static_cast<uint32_t>(HW)] : 0.f);
    __syncthreads();
    const float K = pivot_sh;
    float s = 0.f, ss = 0.f;
    const int nvec = slab_elems / VEC;
    const int chunk_vecs = chunk_bytes / 16;
    int ready = 1;
    for (int v = threadIdx.x; v < nvec; v += NT) {
        while (ready < nch && v >= ready * chunk_vecs) { mbar_wait(&bars[ready], 0); ++ready; }
        float e[VEC];
        load_vec<T>(slab + v * VEC, e);
        if (addp) {
            const uint32_t gi = my_off + static_cast<uint32_t>(v) * VEC;
#pragma unroll
            for (int i = 0; i < VEC; ++i) e[i] += addp[(gi + i) / static_cast<uint32_t>(HW)];
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) { float d = e[i] - K; s += d; ss += d * d; }
    }
    while (ready < nch) { mbar_wait(&bars[ready], 0); ++ready; }
    s = warp_sum(s); ss = warp_sum(ss);
    if ((threadIdx.x & 31) == 0) { red_s[threadIdx.x >> 5] = s; red_ss[threadIdx.x >> 5] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float S = 0.f, SS = 0.f;
        for (int w = 0; w < NT / 32; ++w) { S += red_s[w]; SS += red_ss[w]; }
        const float cnt = static_cast<float>(slab_elems);
        cta_stats[0] = cnt;
        cta_stats[1] = K + S / cnt;
        cta_stats[2] = fmaxf(SS - S * S / cnt, 0.f);
    }
    // ---- cluster merge: every CTA pushes (n, mean, M2) into every CTA's peer_stats[rank] (st.async + mbarrier transaction bytes: no
    // cluster-scope fence, no exit barrier), then merges in fixed rank order => identical result in every CTA
    cluster_wait();
    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(&xbar, cs * 16u);
        for (uint32_t r = 0; r < cs; ++r) st_async_f32x4(&peer_stats[rank][0], &xbar, r, cta_stats[0], cta_stats[1], cta_stats[2], 0.f);
    }
    mbar_wait(&xbar, 0);
    float cn = 0.f, cmean = 0.f, cm2 = 0.f;
    for (uint32_t r = 0; r < cs; ++r) {
        float nb = peer_stats[r][0], mb = peer_stats[r][1], qb = peer_stats[r][2];
        if (r == 0) { cn = nb; cmean = mb; cm2 = qb; }
        else welford_merge(cn, cmean, cm2, nb, mb, qb);
    }
    const float mean = cmean;
    const float rstd = rsqrtf(cm2 / static_cast<float>(L) + eps);

    // ---- pass B (in place in smem): y = x * scale[ch] + shift[ch] with scale = rstd * w, shift = b + (add - mean) * scale
    const bool staged = cpg <= kGnMaxCpgSmem;
    if (staged) {
        for (int ch = threadIdx.x; ch < cpg; ch += NT) {
            const int c = g * cpg + ch;
            const float sc = rstd * (weight ? weight[c] : 1.f);
            sc_sh[ch] = sc;
            sh_sh[ch] = (bias ? bias[c] : 0.f) + ((addp ? addp[ch] : 0.f) - mean) * sc;
        }
        __syncthreads();
    }
    auto scale_of = [&](int ch) { return staged ? sc_sh[ch] : rstd * (weight ? weight[g * cpg + ch] : 1.f); };
    auto shift_of = [&](int ch, float sc) { return staged ? sh_sh[ch] : (bias ? bias[g * cpg + ch] : 0.f) + ((addp ? addp[ch] : 0.f) - mean) * sc; };
    const uint32_t uHW = static_cast<uint32_t>(HW);
    if (uHW % VEC == 0) {
        // a vector never straddles channels: walk (channel, vector-in-channel) incrementally, no divisions in the loop
        const uint32_t vpc = uHW / VEC;
        const uint32_t v0 = my_off / VEC + threadIdx.x;
        uint32_t ch = v0 / vpc, rem = v0 - ch * vpc;
        const uint32_t step_ch = NT / vpc, step_rem = NT - step_ch * vpc;
        for (int v = threadIdx.x; v < nvec; v += NT) {
            float e[VEC];
            load_vec<T>(slab + v * VEC, e);
            const float sc = scale_of(static_cast<int>(ch));
            const float sh = shift_of(static_cast<int>(ch), sc);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const float o = fmaf(e[i], sc, sh);
                e[i] = fuse_silu ? silu_f(o) : o;
            }
            store_vec<T>(slab + v * VEC, e);
            ch += step_ch; rem += step_rem;
            if (rem >= vpc) { rem -= vpc; ++ch; }
        }
    } else {
        for (int v = threadIdx.x; v < nvec; v += NT) {
            float e[VEC];
            load_vec<T>(slab + v * VEC, e);
            const uint32_t gi = my_off + static_cast<uint32_t>(v) * VEC;
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                const int ch = static_cast<int>((gi + i) / uHW);
                const float sc = scale_of(ch);
                const float o = fmaf(e[i], sc, shift_of(ch, sc));
                e[i] = fuse_silu ? silu_f(o) : o;
            }
            store_vec<T>(slab + v * VEC, e);
        }
    }
    // ---- smem -> HBM: TMA bulk stores, one per chunk, issued together (handing each chunk over as soon as it is normalised — a barrier and a
    // bulk group per chunk — measured slower: 11.2 vs 9.6 us on config C1)
    fence_proxy_async_smem();
    __syncthreads();
    if (threadIdx.x == 0) {
        uint8_t* dst = reinterpret_cast<uint8_t*>(y + gbase + my_off);
        for (int c = 0; c < nch; ++c) {
            const uint32_t off = c * chunk_bytes;
            const uint32_t bytes = (c == nch - 1) ? slab_bytes - off : chunk_bytes;
            bulk_store_1d(dst + off, reinterpret_cast<const uint8_t*>(slab) + off, bytes);
        }
        bulk_commit();
        bulk_wait_read_all();             // shared memory may go away once the stores have read it
    }
}

// Any shape: one CTA per (n,g), exact two-pass statistics straight from global memory.
template <typename T>
__global__ void __launch_bounds__(kGnThreads) gn_nchw_generic_kernel(const T* __restrict__ x, T* __restrict__ y,
                                                                      const float* __restrict__ weight, const float* __restrict__ bias,
                                                                      const float* __restrict__ add_nc, int C, int HW, int G, float eps,
                                                                      int fuse_silu) {
    __shared__ float red[kGnThreads / 32];
    __shared__ float bc;
    const int n = blockIdx.x / G, g = blockIdx.x - n * G;
    const int cpg = C / G;
    const long long L = static_cast<long long>(cpg) * HW;
    const long long gbase = (static_cast<long long>(n) * C + static_cast<long long>(g) * cpg) * HW;
    const float* addp = add_nc ? add_nc + static_cast<long long>(n) * C + g * cpg : nullptr;
    auto block_sum = [&](float v) -> float {
        v = warp_sum(v);
        __syncthreads();
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
        __syncthreads();
        if (threadIdx.x == 0) { float t = 0.f; for (int w = 0; w < kGnThreads / 32; ++w) t += red[w]; bc = t; }
        __syncthreads();
        return bc;
    };
    float s = 0.f;
    for (long long i = threadIdx.x; i < L; i += kGnThreads) s += to_f<T>(x[gbase + i]) + (addp ? addp[static_cast<int>(i / HW)] : 0.f);
    const float mean = block_sum(s) / static_cast<float>(L);
    float q = 0.f;
    for (long long i = threadIdx.x; i < L; i += kGnThreads) {
        float d = to_f<T>(x[gbase + i]) + (addp ? addp[static_cast<int>(i / HW)] : 0.f) - mean;
        q += d * d;
    }
    const float rstd = rsqrtf(block_sum(q) / static_cast<float>(L) + eps);
    for (long long i = threadIdx.x; i < L; i += kGnThreads) {
        const int ch = static_cast<int>(i / HW), c = g * cpg + ch;
        float o = (to_f<T>(x[gbase + i]) + (addp ? addp[ch] : 0.f) - mean) * rstd;
        if (weight) o *= weight[c];
        if (bias) o += bias[c];
        y[gbase + i] = from_f<T>(fuse_silu ? silu_f(o) : o);
    }
}

// ------------------------------------------------------------------------------------ NHWC
// workspace layout: [counters: 256 ints][stats: N*G*2 floats][partials: N*S*G*2 floats]
constexpr int kGnCounterInts = 256;
constexpr int kGnMaxSlabs = 512;
constexpr int kNhwcVec = 8;      // channels per thread (16 B of bf16, 32 B of fp32)

template <typename T> SDOD_DEVICE void load8(const T* p, float* out);
template <> SDOD_DEVICE void load8<float>(const float* p, float* out) {
    const float4 a = *reinterpret_cast<const float4*>(p), b = *reinterpret_cast<const float4*>(p + 4);
    out[0] = a.x; out[1] = a.y; out[2] = a.z; out[3] = a.w; out[4] = b.x; out[5] = b.y; out[6] = b.z; out[7] = b.w;
}
template <> SDOD_DEVICE void load8<bf16>(const bf16* p, float* out) { load_vec<bf16>(p, out); }
template <typename T> SDOD_DEVICE void store8v(T* p, const float* v);
template <> SDOD_DEVICE void store8v<float>(float* p, const float* v) {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
}
template <> SDOD_DEVICE void store8v<bf16>(bf16* p, const float* v) { store_vec<bf16>(p, v); }

template <typename T>
__global__ void gn_nhwc_stats_kernel(const T* __restrict__ x, const float* __restrict__ add_nc, float* __restrict__ partials,
                                     unsigned int* __restrict__ counters, float* __restrict__ stats, int C, int HW, int G,
                                     int rows_per_slab, float eps) {
    constexpr int VEC = kNhwcVec;
    griddep_wait();
    griddep_launch();
    extern __shared__ float gsm[];          // [r][C][2] then reused
    __shared__ float pivots[64];
    __shared__ int is_last;
    const int n = blockIdx.y, slab = blockIdx.x, S = gridDim.x;
    const int cvec = C / VEC;
    const int r = blockDim.x / cvec;
    const int col = threadIdx.x % cvec, rr = threadIdx.x / cvec;
    const bool active = rr < r;
    const int cpg = C / G;
    const T* xn = x + static_cast<long long>(n) * HW * C;
    const float* addp = add_nc ? add_nc + static_cast<long long>(n) * C : nullptr;
    for (int gg = threadIdx.x; gg < G; gg += blockDim.x) {
        const int c = gg * cpg;
        pivots[gg] = to_f<T>(xn[c]) + (addp ? addp[c] : 0.f);
    }
    __syncthreads();
    float K[VEC], ad[VEC], s[VEC], ss[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = col * VEC + i;
        K[i] = pivots[c / cpg];
        ad[i] = addp ? addp[c] : 0.f;
        s[i] = 0.f; ss[i] = 0.f;
    }
    const int p0 = slab * rows_per_slab;
    const int p1 = min(HW, p0 + rows_per_slab);
    if (active) {
        // four rows per trip: the loads are issued back to back, so each thread keeps 4 x 32 B (fp32) in flight — with one
        // row per trip the kernel ran at ~45 % of HBM bandwidth, latency-bound (profiles/r01_unet_step_b8_per_op_hot_d.txt)
        int p = p0 + rr;
        for (; p + 3 * r < p1; p += 4 * r) {
            float e[4][VEC];
#pragma unroll
            for (int u = 0; u < 4; ++u) load8<T>(xn + static_cast<long long>(p + u * r) * C + col * VEC, e[u]);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
#pragma unroll
                for (int i = 0; i < VEC; ++i) { float d = e[u][i] + ad[i] - K[i]; s[i] += d; ss[i] += d * d; }
            }
        }
        for (; p < p1; p += r) {
            float e[VEC];
            load8<T>(xn + static_cast<long long>(p) * C + col * VEC, e);
#pragma unroll
            for (int i = 0; i < VEC; ++i) { float d = e[i] + ad[i] - K[i]; s[i] += d; ss[i] += d * d; }
        }
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            gsm[(rr * C + col * VEC + i) * 2 + 0] = s[i];
            gsm[(rr * C + col * VEC + i) * 2 + 1] = ss[i];
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int q = 0; q < r; ++q) { a += gsm[(q * C + c) * 2]; b += gsm[(q * C + c) * 2 + 1]; }
        gsm[c * 2] = a; gsm[c * 2 + 1] = b;     // row 0 doubles as the per-channel totals (q==0 read first)
    }
    __syncthreads();
    for (int gg = threadIdx.x; gg < G; gg += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = gg * cpg; c < (gg + 1) * cpg; ++c) { a += gsm[c * 2]; b += gsm[c * 2 + 1]; }
        float* dst = partials + ((static_cast<long long>(n) * S + slab) * G + gg) * 2;
        dst[0] = a; dst[1] = b;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = (atomicAdd(&counters[n], 1u) == static_cast<unsigned>(S - 1));
    __syncthreads();
    if (is_last) {
        __threadfence();
        // fold the S slab partials: `lanes` threads per group each take a strided subset, then a fixed-order smem fold
        const int lanes = max(1, min(static_cast<int>(blockDim.x) / G, 16));
        float* fold = gsm;                                    // [G][lanes][2], gsm is free now
        if (static_cast<int>(threadIdx.x) < G * lanes) {
            const int gg = threadIdx.x / lanes, l = threadIdx.x - gg * lanes;
            float a = 0.f, b = 0.f;
            const float2* src = reinterpret_cast<const float2*>(partials + (static_cast<long long>(n) * S * G + gg) * 2);
            for (int q = l; q < S; q += lanes) { const float2 v = __ldcg(src + static_cast<long long>(q) * G); a += v.x; b += v.y; }
            fold[(gg * lanes + l) * 2] = a; fold[(gg * lanes + l) * 2 + 1] = b;
        }
        __syncthreads();
        for (int gg = threadIdx.x; gg < G; gg += blockDim.x) {
            float a = 0.f, b = 0.f;
            for (int l = 0; l < lanes; ++l) { a += fold[(gg * lanes + l) * 2]; b += fold[(gg * lanes + l) * 2 + 1]; }
            const float cnt = static_cast<float>(cpg) * static_cast<float>(HW);
            const float m = a / cnt;
            const float var = fmaxf(b / cnt - m * m, 0.f);
            stats[(n * G + gg) * 2 + 0] = pivots[gg] + m;
            stats[(n * G + gg) * 2 + 1] = rsqrtf(var + eps);
        }
        if (threadIdx.x == 0) counters[n] = 0;   // self-cleaning: workspace stays reusable
    }
}

template <typename TI, typename TO>
__global__ void gn_nhwc_apply_kernel(const TI* __restrict__ x, TO* __restrict__ y, const float* __restrict__ weight,
                                     const float* __restrict__ bias, const float* __restrict__ add_nc,
                                     const float* __restrict__ stats, int C, int HW, int G, int rows_per_slab, int fuse_silu) {
    constexpr int VEC = kNhwcVec;
    griddep_wait();
    griddep_launch();
    const int n = blockIdx.y, slab = blockIdx.x;
    const int cvec = C / VEC;
    const int r = blockDim.x / cvec;
    const int col = threadIdx.x % cvec, rr = threadIdx.x / cvec;
    if (rr >= r) return;
    const int cpg = C / G;
    float A[VEC], B[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = col * VEC + i;
        const int g = c / cpg;
        const float mean = stats[(n * G + g) * 2], rstd = stats[(n * G + g) * 2 + 1];
        const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
        const float ad = add_nc ? add_nc[static_cast<long long>(n) * C + c] : 0.f;
        A[i] = rstd * w;
        B[i] = (ad - mean) * rstd * w + b;
    }
    const int p0 = slab * rows_per_slab;
    const int p1 = min(HW, p0 + rows_per_slab);
    const TI* xn = x + static_cast<long long>(n) * HW * C;
    TO* yn = y + static_cast<long long>(n) * HW * C;
    int p = p0 + rr;
    for (; p + 3 * r < p1; p += 4 * r) {          // four rows in flight per thread (see the stats kernel)
        float e[4][VEC];
#pragma unroll
        for (int u = 0; u < 4; ++u) load8<TI>(xn + static_cast<long long>(p + u * r) * C + col * VEC, e[u]);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float o = fmaf(e[u][i], A[i], B[i]);
                e[u][i] = fuse_silu ? silu_f(o) : o;
            }
            store8v<TO>(yn + static_cast<long long>(p + u * r) * C + col * VEC, e[u]);
        }
    }
    for (; p < p1; p += r) {
        float e[VEC];
        load8<TI>(xn + static_cast<long long>(p) * C + col * VEC, e);
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            float o = fmaf(e[i], A[i], B[i]);
            e[i] = fuse_silu ? silu_f(o) : o;
        }
        store8v<TO>(yn + static_cast<long long>(p) * C + col * VEC, e);
    }
}

// Group-owned NHWC GroupNorm: one CTA (or a cluster of `cs` CTAs splitting the rows) owns one (sample, group) — its rows x (C/G) channels
// slab — stages it in shared memory as fp32 while accumulating pivot-shifted moments, reduces within the block (+ across the cluster over
// distributed shared memory, fixed rank order), and normalises out of shared memory.  x is read once, y written once, there is no grid-wide
// synchronisation and no workspace: at the batch-2 UNet step a GroupNorm costs one short dependent kernel instead of a cooperative two-pass one
// (the grid barrier + partial-sum round trips of gn_nhwc_fused_kernel put ~9 us of pure L2 latency on the step's critical path, x61).
// x is the channel concatenation [xa | xb] (xb may be NULL); `raw` optionally receives the bf16 copy of the un-normalised concatenation.
// Thread layout: a thread owns one 2-channel unit (column) of the slab and walks the rows with a fixed stride, so its affine
// coefficients are loop constants; 8-byte (fp32) / 4-byte (bf16) accesses, adjacent groups (adjacent CTAs) complete each 32-byte sector.
template <typename TI, typename TO>
__global__ void __launch_bounds__(512) gn_nhwc_group_kernel(const TI* __restrict__ xa, int Ca, const TI* __restrict__ xb, int Cb,
                                                            TO* __restrict__ y, bf16* __restrict__ raw, const float* __restrict__ weight,
                                                            const float* __restrict__ bias, int HW, int G, int cs, int rows_per_cta, float eps,
                                                            int fuse_silu) {
    extern __shared__ float2 slab[];            // [rows_per_cta][upr]
    __shared__ float red[2][16];
    __shared__ float cta_tot[2];
    __shared__ __align__(16) float peer_tot[8][2];          // [rank]{sum, sum of squares}: every CTA of the cluster pushes its totals here (st.async)
    __shared__ __align__(8) uint64_t xbar;                  // ... and credits their bytes to this mbarrier
    const int C = Ca + Cb, cpg = C / G, upr = cpg / 2;
    const int g = blockIdx.x / cs, rank = blockIdx.x - g * cs, n = blockIdx.y;
    const int rpb = blockDim.x / upr;                       // rows per pass
    const int u = threadIdx.x % upr, r0 = threadIdx.x / upr;
    const bool active = r0 < rpb;
    const int c = g * cpg + 2 * u;                          // this thread's first channel (of the concatenation)
    const bool from_b = c >= Ca;
    const int ldx = from_b ? Cb : Ca;
    const int p0 = rank * rows_per_cta, p1 = min(HW, p0 + rows_per_cta);
    if (cs > 1) {
        if (threadIdx.x == 0) { mbar_init(&xbar, 1); fence_mbar_init(); }
        cluster_arrive_relaxed();                           // "my mbarrier exists": waited for just before the push, after the whole load phase
    }
    griddep_wait();
    griddep_launch();
    const TI* xcol = from_b ? xb + static_cast<long long>(n) * HW * Cb + (c - Ca) : xa + static_cast<long long>(n) * HW * Ca + c;
    const int cg0 = g * cpg;                                // pivot: the group's first element of row 0 (identical in every CTA of the cluster)
    const float K = cg0 < Ca ? to_f<TI>(xa[static_cast<long long>(n) * HW * Ca + cg0]) : to_f<TI>(xb[static_cast<long long>(n) * HW * Cb + (cg0 - Ca)]);
    float s = 0.f, ss = 0.f;
    if (active) {
        // eight rows per trip, all loads issued before the first use: the slab is L2-resident and each access is only 8 (4) bytes, so the
        // kernel lives on memory-level parallelism (one load in flight per thread ran this phase at ~0.4 us per row)
        constexpr int UN = 8;
        for (int pb = p0 + r0; pb < p1; pb += UN * rpb) {
            float2 v[UN];
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                const int p = pb + k * rpb;
                if (p < p1) {
                    if (sizeof(TI) == 4) v[k] = *reinterpret_cast<const float2*>(xcol + static_cast<long long>(p) * ldx);
                    else v[k] = unpack_bf16x2(*reinterpret_cast<const uint32_t*>(xcol + static_cast<long long>(p) * ldx));
                }
            }
#pragma unroll
            for (int k = 0; k < UN; ++k) {
                const int p = pb + k * rpb;
                if (p < p1) {
                    slab[(p - p0) * upr + u] = v[k];
                    const float d0 = v[k].x - K, d1 = v[k].y - K;
                    s += d0 + d1;
                    ss += d0 * d0 + d1 * d1;
                }
            }
        }
    }
    s = warp_sum(s);
    ss = warp_sum(ss);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = (blockDim.x + 31) >> 5;
    if (lane == 0) { red[0][warp] = s; red[1][warp] = ss; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int w = 0; w < nwarps; ++w) { a += red[0][w]; b += red[1][w]; }
        cta_tot[0] = a; cta_tot[1] = b;
    }
    float ts, tss;
    if (cs > 1) {
        // all-gather of the cs CTA totals by push: no cluster-scope fence (the release form of the cluster barrier is a MEMBAR.ALL.GPU per warp,
        // ~9 % of this kernel's stall samples at batch 32 together with the exit barrier it needed; profiles/r02_gn_group_b32_source_top.txt)
        cluster_wait();                                     // every peer's mbarrier is initialised
        if (threadIdx.x == 0) {
            mbar_arrive_expect_tx(&xbar, static_cast<uint32_t>(cs) * 8u);
            for (int r = 0; r < cs; ++r) st_async_f32x2(&peer_tot[rank][0], &xbar, r, cta_tot[0], cta_tot[1]);
        }
        mbar_wait(&xbar, 0);
        ts = 0.f; tss = 0.f;
        for (int r = 0; r < cs; ++r) { ts += peer_tot[r][0]; tss += peer_tot[r][1]; }      // fixed rank order: identical in every CTA
    } else {
        __syncthreads();
        ts = cta_tot[0]; tss = cta_tot[1];
    }
    const float cnt = static_cast<float>(cpg) * static_cast<float>(HW);
    const float m = ts / cnt;
    const float var = fmaxf(tss / cnt - m * m, 0.f);
    const float mean = K + m, rstd = rsqrtf(var + eps);
    if (active) {
        const float w0 = weight ? weight[c] : 1.f, w1 = weight ? weight[c + 1] : 1.f;
        const float b0 = bias ? bias[c] : 0.f, b1 = bias ? bias[c + 1] : 0.f;
        const float A0 = rstd * w0, A1 = rstd * w1, B0 = -mean * A0 + b0, B1 = -mean * A1 + b1;
        // pointers advance by a fixed step (no 64-bit multiply per row); two rows per trip for instruction-level parallelism
        const long long step = static_cast<long long>(rpb) * C;
        TO* yp = y + (static_cast<long long>(n) * HW + p0 + r0) * C + c;
        bf16* rp = raw ? raw + (static_cast<long long>(n) * HW + p0 + r0) * C + c : nullptr;
        const float2* sp = slab + r0 * upr + u;
        const int sstep = rpb * upr;
#pragma unroll 2
        for (int p = p0 + r0; p < p1; p += rpb, yp += step, sp += sstep) {
            const float2 v = *sp;
            if (rp) { *reinterpret_cast<uint32_t*>(rp) = pack_bf16x2(v.x, v.y); rp += step; }
            float o0 = fmaf(v.x, A0, B0), o1 = fmaf(v.y, A1, B1);
            if (fuse_silu) { o0 = silu_f(o0); o1 = silu_f(o1); }
            if (sizeof(TO) == 4) *reinterpret_cast<float2*>(yp) = make_float2(o0, o1);
            else *reinterpret_cast<uint32_t*>(yp) = pack_bf16x2(o0, o1);
        }
    }
    // (no exit barrier: nobody reads this CTA's shared memory remotely, and every push INTO it had landed before its mbarrier wait returned)
}

// Single-launch cooperative NHWC GroupNorm for L2-resident tensors (the batch-2 UNet step), register-resident:
// one grid of <= #SMs CTAs, all co-resident.  Each CTA (sample n, row slab) loads its rows ONCE — every thread keeps its <= MAXR rows x 8
// channels in registers — accumulates moments about a pivot taken from its own slab, converts them to a Welford partial (mean, M2) per group,
// publishes it and joins the sample's arrive counter; once the sample's S slabs have arrived EVERY CTA merges the S partials with Chan's
// update in the same fixed order (deterministic, bit-identical statistics everywhere) and normalises straight out of its registers.
// Critical path: one row-load round trip, one publish + arrive, the spin, one fold round trip, the stores (the first version re-read the rows
// and used a shared pivot from global memory: ~10 dependent L2 round trips, 13 us per GroupNorm in the step's timeline).
// x is the channel concatenation [xa | xb] (xb may be NULL): the UNet's skip concat never exists in memory, and `raw` (optional) receives
// the bf16 copy of the un-normalised concatenation that the ResBlock's 1x1 skip convolution consumes.
// Co-residency: the grid is sized <= #SMs at one CTA per SM, the predecessor kernel has completed when griddepcontrol.wait returns and a PDL
// successor can only be launched once every CTA of this grid runs, so every CTA a spinning CTA waits for is resident or about to be; the
// spin is bounded and traps instead of hanging the GPU.
template <typename TI, typename TO, int MAXR>
__global__ void __launch_bounds__(512, 1) gn_nhwc_fused_kernel(const TI* __restrict__ xa, int Ca, const TI* __restrict__ xb, int Cb,
                                                               TO* __restrict__ y, bf16* __restrict__ raw, const float* __restrict__ weight,
                                                               const float* __restrict__ bias, float* __restrict__ partials,
                                                               unsigned int* __restrict__ counters, int HW, int G, int rows_per_slab, float eps,
                                                               int fuse_silu) {
    constexpr int VEC = kNhwcVec;
    extern __shared__ float gsm[];          // [r][C][2], then [G][lanes][3]
    __shared__ float pivots[64], s_mean[64], s_rstd[64];
    const int C = Ca + Cb;
    const int n = blockIdx.y, slab = blockIdx.x, S = gridDim.x;
    const int cvec = C / VEC;
    const int r = blockDim.x / cvec;
    const int col = threadIdx.x % cvec, rr = threadIdx.x / cvec;
    const bool active = rr < r;
    const int cpg = C / G;
    const int c0 = col * VEC;                                  // this thread's 8 channels live in one of the two sources (Ca % 8 == 0)
    const bool from_b = c0 >= Ca;
    const int ldx = from_b ? Cb : Ca;
    const int p0 = slab * rows_per_slab;
    const int p1 = min(HW, p0 + rows_per_slab);
    griddep_wait();
    griddep_launch();
    const TI* xcol = (from_b ? xb + static_cast<long long>(n) * HW * Cb + (c0 - Ca) : xa + static_cast<long long>(n) * HW * Ca + c0);
    float e[MAXR][VEC];
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
        const int p = p0 + rr + k * r;
        if (active && p < p1) {
            load8<TI>(xcol + static_cast<long long>(p) * ldx, e[k]);
        } else {
#pragma unroll
            for (int i = 0; i < VEC; ++i) e[k][i] = 0.f;
        }
    }
    // pivot of group g = its first channel in this slab's first row (held by a thread of row 0)
    if (active && rr == 0) {
#pragma unroll
        for (int i = 0; i < VEC; ++i)
            if ((c0 + i) % cpg == 0) pivots[(c0 + i) / cpg] = e[0][i];
    }
    __syncthreads();
    if (active) {
#pragma unroll
        for (int i = 0; i < VEC; ++i) {
            const float K = pivots[(c0 + i) / cpg];
            float s = 0.f, ss = 0.f;
#pragma unroll
            for (int k = 0; k < MAXR; ++k) {
                if (p0 + rr + k * r < p1) { const float d = e[k][i] - K; s += d; ss += d * d; }
            }
            gsm[(rr * C + c0 + i) * 2 + 0] = s;
            gsm[(rr * C + c0 + i) * 2 + 1] = ss;
        }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int q = 0; q < r; ++q) { a += gsm[(q * C + c) * 2]; b += gsm[(q * C + c) * 2 + 1]; }
        gsm[c * 2] = a; gsm[c * 2 + 1] = b;
    }
    __syncthreads();
    for (int gg = threadIdx.x; gg < G; gg += blockDim.x) {
        float a = 0.f, b = 0.f;
        for (int c = gg * cpg; c < (gg + 1) * cpg; ++c) { a += gsm[c * 2]; b += gsm[c * 2 + 1]; }
        const float cnt = static_cast<float>(cpg) * static_cast<float>(p1 - p0);
        float* dst = partials + ((static_cast<long long>(n) * S + slab) * G + gg) * 2;
        dst[0] = pivots[gg] + a / cnt;                         // slab mean
        dst[1] = fmaxf(b - a * a / cnt, 0.f);                  // slab M2
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(&counters[n], 1u);
        unsigned int seen = 0, spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(counters + n) : "memory");
            if (seen < static_cast<unsigned>(S) && ++spins > (1u << 22)) __trap();   // seconds, not a hot loop: a scheduling bug traps instead of hanging
        } while (seen < static_cast<unsigned>(S));
    }
    __syncthreads();
    {
        // fold: `lanes` threads per group take a strided subset of the S slab partials (all loads in flight), Chan-merge them in order,
        // then one thread per group merges the lane results in lane order
        const int lanes = max(1, min(static_cast<int>(blockDim.x) / G, 16));
        float* fold = gsm;                                    // [G][lanes][3] = (count, mean, M2)
        if (static_cast<int>(threadIdx.x) < G * lanes) {
            const int gg = threadIdx.x / lanes, l = threadIdx.x - gg * lanes;
            const float2* src = reinterpret_cast<const float2*>(partials + (static_cast<long long>(n) * S * G + gg) * 2);
            constexpr int kMaxPer = 16;                        // S <= 148 and lanes >= 10 at the block sizes used (checked by the host)
            float2 v[kMaxPer];
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int q = l + j * lanes;
                v[j] = q < S ? __ldcg(src + static_cast<long long>(q) * G) : make_float2(0.f, 0.f);
            }
            float cn = 0.f, cm = 0.f, c2 = 0.f;
#pragma unroll
            for (int j = 0; j < kMaxPer; ++j) {
                const int q = l + j * lanes;
                if (q < S) {
                    const float nb = static_cast<float>(cpg) * static_cast<float>(min(HW, (q + 1) * rows_per_slab) - q * rows_per_slab);
                    welford_merge(cn, cm, c2, nb, v[j].x, v[j].y);
                }
            }
            fold[(gg * lanes + l) * 3] = cn; fold[(gg * lanes + l) * 3 + 1] = cm; fold[(gg * lanes + l) * 3 + 2] = c2;
        }
        __syncthreads();
        for (int gg = threadIdx.x; gg < G; gg += blockDim.x) {
            float cn = 0.f, cm = 0.f, c2 = 0.f;
            for (int l = 0; l < lanes; ++l) welford_merge(cn, cm, c2, fold[(gg * lanes + l) * 3], fold[(gg * lanes + l) * 3 + 1], fold[(gg * lanes + l) * 3 + 2]);
            s_mean[gg] = cm;
            s_rstd[gg] = rsqrtf(fmaxf(c2 / cn, 0.f) + eps);
        }
    }
    __syncthreads();
    // every thread of this CTA is done with the partials: depart; the last CTA of the sample to depart re-arms both counters
    if (threadIdx.x == 0) {
        if (atomicAdd(&counters[kGnCounterInts / 2 + n], 1u) == static_cast<unsigned>(S - 1)) {
            counters[n] = 0;
            counters[kGnCounterInts / 2 + n] = 0;
        }
    }
    if (!active) return;
    float A[VEC], B[VEC];
#pragma unroll
    for (int i = 0; i < VEC; ++i) {
        const int c = c0 + i;
        const int g = c / cpg;
        const float w = weight ? weight[c] : 1.f, b = bias ? bias[c] : 0.f;
        A[i] = s_rstd[g] * w;
        B[i] = -s_mean[g] * s_rstd[g] * w + b;
    }
    TO* ycol = y + static_cast<long long>(n) * HW * C + c0;
    bf16* rcol = raw ? raw + static_cast<long long>(n) * HW * C + c0 : nullptr;
#pragma unroll
    for (int k = 0; k < MAXR; ++k) {
        const int p = p0 + rr + k * r;
        if (p < p1) {
            if (rcol) store8v<bf16>(rcol + static_cast<long long>(p) * C, e[k]);
#pragma unroll
            for (int i = 0; i < VEC; ++i) {
                float o = fmaf(e[k][i], A[i], B[i]);
                e[k][i] = fuse_silu ? silu_f(o) : o;
            }
            store8v<TO>(ycol + static_cast<long long>(p) * C, e[k]);
        }
    }
}

static void nhwc_geometry(int N, int C, int HW, int vec, int* threads, int* slabs, int* rows_per_slab) {
    const int cvec = C / vec;
    int r = cvec >= 256 ? 1 : 256 / cvec;
    if (r > HW) r = HW;
    if (r < 1) r = 1;
    *threads = cvec * r;
    int S = (2 * 148 + N - 1) / N;
    int max_s = HW / (r * 4);
    if (max_s < 1) max_s = 1;
    S = std::max(1, std::min(std::min(S, max_s), kGnMaxSlabs));
    int rps = (HW + S - 1) / S;
    *rows_per_slab = rps;
    *slabs = (HW + rps - 1) / rps;
}

template <typename TI, typename TO>
static int group_norm_nhwc_typed(cudaStream_t stream, const TI* x, TO* y, const float* weight, const float* bias, const float* add_nc,
                                 int N, int C, int HW, int G, float eps, int fuse_silu, void* ws, size_t ws_bytes) {
    constexpr int VEC = kNhwcVec;
    if (C % VEC != 0) return fail(kUnsupported, "group_norm NHWC: C must be a multiple of 8");
    if (C / VEC > 1024) return fail(kUnsupported, "group_norm NHWC: C too large");
    if (G > 64) return fail(kUnsupported, "group_norm NHWC: num_groups > 64");
    if (N > kGnCounterInts) return fail(kUnsupported, "group_norm NHWC: batch > 256");
    int threads, S, rps;
    nhwc_geometry(N, C, HW, VEC, &threads, &S, &rps);
    const size_t need = sdod_group_norm_workspace(N, C, HW, G, SDOD_NHWC);
    if (!ws || ws_bytes < need) return fail(kInvalidArgument, "group_norm NHWC: workspace too small (need " + std::to_string(need) + " bytes)");
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    float* stats = reinterpret_cast<float*>(counters + kGnCounterInts);
    float* partials = stats + static_cast<size_t>(N) * G * 2;
    const int r = threads / (C / VEC);
    const size_t smem = static_cast<size_t>(r) * C * 2 * sizeof(float);
    if (smem > 48 * 1024) {
        SDOD_TRY(check_cuda(cudaFuncSetAttribute(gn_nhwc_stats_kernel<TI>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                            "cudaFuncSetAttribute(gn stats)"));
    }
    SDOD_TRY(check_cuda(launch_pdl(gn_nhwc_stats_kernel<TI>, dim3(S, N), dim3(threads), smem, stream, x, add_nc, partials, counters, stats, C, HW, G,
                                   rps, eps), "launch gn_nhwc_stats_kernel"));
    SDOD_TRY(check_launch("gn_nhwc_stats_kernel"));
    SDOD_TRY(check_cuda(launch_pdl(gn_nhwc_apply_kernel<TI, TO>, dim3(S, N), dim3(threads), 0, stream, x, y, weight, bias, add_nc,
                                   static_cast<const float*>(stats), C, HW, G, rps, fuse_silu), "launch gn_nhwc_apply_kernel"));
    count_launch(2);
    return check_launch("gn_nhwc_apply_kernel");
}

// Single-launch path (gn_nhwc_fused_kernel): the whole tensor must sit in L2 (it is read twice) and the grid must be co-resident.
static bool fused_env() {
    static const int env = [] { const char* e = std::getenv("SDOD_GN_FUSED"); return e ? std::atoi(e) : 1; }();
    return env != 0;
}
static bool fused_size_ok(int N, int Ca, int Cb, int HW, int G, int in_dtype) {
    const int C = Ca + Cb;
    if (!fused_env() || N < 1 || N > kGnCounterInts / 2 || N > device_sm_count()) return false;
    if (C % kNhwcVec != 0 || Ca % kNhwcVec != 0 || C / kNhwcVec > 512 || G > 64 || C % G != 0) return false;
    const size_t bytes = static_cast<size_t>(N) * HW * C * (in_dtype == SDOD_F32 ? 4 : 2);
    return bytes <= (static_cast<size_t>(48) << 20);
}

constexpr int kGnCoopMaxRows = 8;
// Geometry of the cooperative kernel: threads per CTA, slabs per sample, rows per slab; false when the rows do not fit the registers.
static bool coop_geometry(int N, int C, int HW, int G, int* threads, int* slabs, int* rows_per_slab) {
    static const int env = [] { const char* e = std::getenv("SDOD_GN_COOP"); return e ? std::atoi(e) : 1; }();
    if (!env) return false;
    const int cvec = C / kNhwcVec;
    if (cvec < 1 || cvec > 512) return false;
    int r = std::max(1, 512 / cvec);
    if (r > HW) r = HW;
    int S = std::max(1, std::min(device_sm_count() / N, (HW + r - 1) / r));
    const int rps = (HW + S - 1) / S;
    S = (HW + rps - 1) / rps;
    if ((rps + r - 1) / r > kGnCoopMaxRows) return false;
    const int lanes = std::max(1, std::min((cvec * r) / G, 16));        // fold lanes per group, as the kernel computes them
    if (S > 16 * lanes) return false;
    *threads = cvec * r; *slabs = S; *rows_per_slab = rps;
    return true;
}

template <typename TI, typename TO>
static int group_norm_fused_typed(cudaStream_t stream, const TI* xa, int Ca, const TI* xb, int Cb, TO* y, bf16* raw, const float* weight,
                                  const float* bias, int N, int HW, int G, float eps, int fuse_silu, void* ws, size_t ws_bytes) {
    const int C = Ca + Cb;
    int threads = 0, S = 0, rps = 0;
    if (!coop_geometry(N, C, HW, G, &threads, &S, &rps)) return fail(kUnsupported, "group_norm (cooperative): rows per thread exceed the register budget");
    const int r = threads / (C / kNhwcVec);
    const size_t need = kGnCounterInts * sizeof(unsigned int) + static_cast<size_t>(N) * G * 2 * sizeof(float) * (1 + S);
    if (!ws || ws_bytes < need) return fail(kInvalidArgument, "group_norm (fused): workspace too small (need " + std::to_string(need) + " bytes)");
    unsigned int* counters = reinterpret_cast<unsigned int*>(ws);
    float* partials = reinterpret_cast<float*>(counters + kGnCounterInts) + static_cast<size_t>(N) * G * 2;
    const size_t smem = std::max(static_cast<size_t>(r) * C * 2 * sizeof(float), static_cast<size_t>(G) * 16 * 3 * sizeof(float));
    if (smem > 48 * 1024)
        SDOD_TRY(check_cuda(cudaFuncSetAttribute(gn_nhwc_fused_kernel<TI, TO, kGnCoopMaxRows>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)),
                            "cudaFuncSetAttribute(gn fused)"));
    SDOD_TRY(check_cuda(launch_pdl(gn_nhwc_fused_kernel<TI, TO, kGnCoopMaxRows>, dim3(S, N), dim3(threads), smem, stream, xa, Ca, xb, Cb, y, raw, weight, bias,
                                   partials, counters, HW, G, rps, eps, fuse_silu), "launch gn_nhwc_fused_kernel"));
    count_launch();
    return check_launch("gn_nhwc_fused_kernel");
}

// Geometry of the group-owned kernel: cluster size (rows split) and rows per CTA; cs == 0 -> not eligible.  It needs neither workspace nor
// co-residency, so large batches qualify too.  Measured at UNet batch 32 (B200 r2, profiles/r02_unet_step_b32_per_op_*.txt): serving every
// in-network GroupNorm with it (no size bound) instead of the stats + apply pair above 48 MB removes the materialised concat / cast launches
// (374 -> 328 per pass) and takes the pass from 42.40 to 42.08 ms; the CTAs of one image's 32 groups run together, so the 40-byte-per-row
// slices they read share their DRAM sectors in L2.  SDOD_GN_GROUP_MAXMB bounds the input size for A/B runs (default: unbounded).
static int group_kernel_geometry(int N, int Ca, int Cb, int HW, int G, int in_dtype, int* rows_per_cta) {
    static const int env = [] { const char* e = std::getenv("SDOD_GN_GROUP"); return e ? std::atoi(e) : 1; }();
    static const long long max_mb = [] { const char* e = std::getenv("SDOD_GN_GROUP_MAXMB"); return e ? std::atoll(e) : (1LL << 40); }();
    const int C = Ca + Cb;
    if (!env || !fused_env() || G <= 0 || C % G != 0 || N < 1 || HW < 1) return 0;
    if (static_cast<long long>(N) * HW * C * (in_dtype == SDOD_F32 ? 4 : 2) > (max_mb << 20)) return 0;
    const int cpg = C / G;
    if (cpg % 2 != 0 || Ca % 2 != 0 || cpg / 2 > 512 || N > 65535) return 0;
    const size_t slab = static_cast<size_t>(HW) * cpg * sizeof(float);
    int cs = 1;
    static const long long slab_kb = [] { const char* e = std::getenv("SDOD_GN_GROUP_SLABKB"); return e ? std::atoll(e) : 100LL; }();     // A/B knob
    while (cs < 8 && (slab / cs > (static_cast<size_t>(slab_kb) << 10) || HW % cs != 0)) cs *= 2;      // <= 100 KB: two CTAs per SM
    // small batch: spread each group over more CTAs (two 512-thread CTAs per SM) while a thread still walks >= 8 rows
    const int rpb = std::max(1, 512 / (cpg / 2));
    static const long long spread = [] { const char* e = std::getenv("SDOD_GN_GROUP_SPREAD"); return e ? std::atoll(e) : 256LL; }();   // A/B knob
    while (cs < 8 && static_cast<long long>(N) * G * cs < spread && HW / (cs * 2) >= 8 * rpb && HW % (cs * 2) == 0) cs *= 2;
    if (HW % cs != 0 || slab / cs > (static_cast<size_t>(200) << 10)) return 0;
    // Tensors beyond L2 whose groups only fit as one 100-200 KB CTA per SM (the VAE's 128x128 x 512-channel level at batch 8) stay on the
    // stats + apply pair: measured 30.7 vs 31.8 ms for the batch-8 decoder (B200 r2).
    const long long in_bytes = static_cast<long long>(N) * HW * C * (in_dtype == SDOD_F32 ? 4 : 2);
    if (in_bytes > (48LL << 20) && slab / cs > (static_cast<size_t>(100) << 10)) return 0;
    *rows_per_cta = HW / cs;
    return cs;
}

template <typename TI, typename TO>
static int group_norm_group_typed(cudaStream_t stream, const TI* xa, int Ca, const TI* xb, int Cb, TO* y, bf16* raw, const float* weight,
                                  const float* bias, int N, int HW, int G, float eps, int fuse_silu, int cs, int rows_per_cta) {
    const int C = Ca + Cb, upr = (C / G) / 2;
    int rpb = std::max(1, std::min(512 / upr, rows_per_cta));
    const int threads = std::min(512, ((upr * rpb + 31) / 32) * 32);      // whole warps; threads beyond upr*rpb idle (r0 >= rpb)
    const size_t smem = static_cast<size_t>(rows_per_cta) * upr * sizeof(float2);
    static size_t configured = 0;
    if (smem > 48 * 1024 && smem > configured) {
        SDOD_TRY(check_cuda(cudaFuncSetAttribute(gn_nhwc_group_kernel<TI, TO>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                            "cudaFuncSetAttribute(gn group)"));
        configured = 200 * 1024;
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(static_cast<unsigned>(G) * cs, N); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    int na = 0;
    if (cs > 1) {
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = cs; attr[na].val.clusterDim.y = 1; attr[na].val.clusterDim.z = 1;
        ++na;
    }
    if (pdl_enabled() && static_cast<long long>(G) * cs * N <= 296) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    cfg.attrs = attr; cfg.numAttrs = na;
    SDOD_TRY(check_cuda(cudaLaunchKernelEx(&cfg, gn_nhwc_group_kernel<TI, TO>, xa, Ca, xb, Cb, y, raw, weight, bias, HW, G, cs, rows_per_cta, eps, fuse_silu),
                        "launch gn_nhwc_group_kernel"));
    count_launch();
    return check_launch("gn_nhwc_group_kernel");
}

// Can the single-launch forms (cooperative register-resident kernel, else the group-owned kernel) take this shape?
bool group_norm_fused_eligible(int N, int Ca, int Cb, int HW, int G, int in_dtype) {
    if (N <= 0 || HW <= 0 || G <= 0) return false;
    int a = 0, b = 0, c = 0;
    return group_kernel_geometry(N, Ca, Cb, HW, G, in_dtype, &a) != 0 || (fused_size_ok(N, Ca, Cb, HW, G, in_dtype) && coop_geometry(N, Ca + Cb, HW, G, &a, &b, &c));
}

int group_norm_nhwc2(cudaStream_t stream, const void* xa, int Ca, const void* xb, int Cb, int in_dtype, void* y, int out_dtype, void* raw_bf16,
                     const float* weight, const float* bias, int N, int HW, int G, float eps, int fuse_silu, void* ws, size_t ws_bytes) {
    int ct = 0, cS = 0, crps = 0;
    const bool coop_ok = xa && y && N > 0 && HW > 0 && G > 0 && fused_size_ok(N, Ca, xb ? Cb : 0, HW, G, in_dtype) &&
                         coop_geometry(N, Ca + (xb ? Cb : 0), HW, G, &ct, &cS, &crps);
    {   // the group-owned kernel first (measured B200 r2: batch-2 step 5.42 ms vs 5.67 ms with the cooperative kernel, whose 128 registers x
        // 512 threads own the whole SM and so forfeit the PDL overlap with its neighbours); the cooperative kernel serves odd channels-per-group
        int rpc = 0;
        const int cs = (xa && y && (Cb == 0 || xb)) ? group_kernel_geometry(N, Ca, xb ? Cb : 0, HW, G, in_dtype, &rpc) : 0;
        if (cs && ((weight == nullptr) == (bias == nullptr))) {
            if (!xb) Cb = 0;
#define SDOD_GNG_CASE(TI, TO) \
    return group_norm_group_typed<TI, TO>(stream, static_cast<const TI*>(xa), Ca, static_cast<const TI*>(xb), Cb, static_cast<TO*>(y), \
                                          static_cast<bf16*>(raw_bf16), weight, bias, N, HW, G, eps, fuse_silu, cs, rpc)
            if (in_dtype == SDOD_F32 && out_dtype == SDOD_F32) SDOD_GNG_CASE(float, float);
            if (in_dtype == SDOD_F32 && out_dtype == SDOD_BF16) SDOD_GNG_CASE(float, bf16);
            if (in_dtype == SDOD_BF16 && out_dtype == SDOD_BF16) SDOD_GNG_CASE(bf16, bf16);
            if (in_dtype == SDOD_BF16 && out_dtype == SDOD_F32) SDOD_GNG_CASE(bf16, float);
#undef SDOD_GNG_CASE
        }
    }
    if (!xa || !y || (Cb > 0 && !xb)) return fail(kInvalidArgument, "group_norm: NULL tensor");
    if (!xb) Cb = 0;
    if (N <= 0 || Ca <= 0 || Cb < 0 || HW <= 0 || G <= 0) return fail(kInvalidArgument, "group_norm: non-positive extent");
    if ((Ca + Cb) % G != 0) return fail(kInvalidArgument, "num_channels must be divisible by num_groups");
    if ((weight == nullptr) != (bias == nullptr)) return fail(kInvalidArgument, "group_norm: weight and bias must both be given or both be NULL");
    if (!coop_ok)
        return fail(kUnsupported, "group_norm (two-source / single-launch form): tensor too large for the L2-resident kernel or unsupported shape");
#define SDOD_GN2_CASE(TI, TO) \
    return group_norm_fused_typed<TI, TO>(stream, static_cast<const TI*>(xa), Ca, static_cast<const TI*>(xb), Cb, static_cast<TO*>(y), \
                                          static_cast<bf16*>(raw_bf16), weight, bias, N, HW, G, eps, fuse_silu, ws, ws_bytes)
    if (in_dtype == SDOD_F32 && out_dtype == SDOD_F32) SDOD_GN2_CASE(float, float);
    if (in_dtype == SDOD_F32 && out_dtype == SDOD_BF16) SDOD_GN2_CASE(float, bf16);
    if (in_dtype == SDOD_BF16 && out_dtype == SDOD_BF16) SDOD_GN2_CASE(bf16, bf16);
    if (in_dtype == SDOD_BF16 && out_dtype == SDOD_F32) SDOD_GN2_CASE(bf16, float);
#undef SDOD_GN2_CASE
    return fail(kInvalidArgument, "group_norm: unknown dtype");
}

int group_norm_nhwc(cudaStream_t stream, const void* x, int in_dtype, void* y, int out_dtype, const float* weight, const float* bias,
                    const float* add_nc, int N, int C, int HW, int G, float eps, int fuse_silu, void* ws, size_t ws_bytes) {
    if (!x || !y) return fail(kInvalidArgument, "group_norm: NULL tensor");
    if (N <= 0 || C <= 0 || HW <= 0 || G <= 0) return fail(kInvalidArgument, "group_norm: non-positive extent");
    if (C % G != 0) return fail(kInvalidArgument, "num_channels must be divisible by num_groups");
    if ((weight == nullptr) != (bias == nullptr)) return fail(kInvalidArgument, "group_norm: weight and bias must both be given or both be NULL");
    if (!add_nc && group_norm_fused_eligible(N, C, 0, HW, G, in_dtype))
        return group_norm_nhwc2(stream, x, C, nullptr, 0, in_dtype, y, out_dtype, nullptr, weight, bias, N, HW, G, eps, fuse_silu, ws, ws_bytes);
#define SDOD_GN_CASE(TI, TO) \
    return group_norm_nhwc_typed<TI, TO>(stream, static_cast<const TI*>(x), static_cast<TO*>(y), weight, bias, add_nc, N, C, HW, G, eps, fuse_silu, ws, ws_bytes)
    if (in_dtype == SDOD_F32 && out_dtype == SDOD_F32) SDOD_GN_CASE(float, float);
    if (in_dtype == SDOD_F32 && out_dtype == SDOD_BF16) SDOD_GN_CASE(float, bf16);
    if (in_dtype == SDOD_BF16 && out_dtype == SDOD_BF16) SDOD_GN_CASE(bf16, bf16);
    if (in_dtype == SDOD_BF16 && out_dtype == SDOD_F32) SDOD_GN_CASE(bf16, float);
#undef SDOD_GN_CASE
    return fail(kInvalidArgument, "group_norm: unknown dtype");
}

template <typename T>
static int group_norm_typed(cudaStream_t stream, const T* x, T* y, const float* weight, const float* bias, const float* add_nc,
                            int N, int C, int HW, int G, float eps, int layout, int fuse_silu, void* ws, size_t ws_bytes) {
    constexpr int VEC = VecOf<T>::N;
    const int cpg = C / G;
    if (layout == SDOD_NCHW) {
        const long long L = static_cast<long long>(cpg) * HW;
        const long long Lb = L * sizeof(T);
        const bool aligned = (Lb % 16 == 0) && ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((reinterpret_cast<uintptr_t>(y) & 15) == 0);
        int cs = 1;
        while (Lb / cs > 49152 && cs < 8) cs *= 2;
        while (static_cast<long long>(N) * G * cs < 148 && cs < 8 && Lb / (cs * 2) >= 8192) cs *= 2;     // (C1: 4 beats 8, 9.6 vs 14.2 us, B200 r2)
        {
            static const int env_cs = [] { const char* e = std::getenv("SDOD_GN_CS"); return e ? std::atoi(e) : 0; }();   // experiment knob
            if (env_cs >= 1 && env_cs <= 8 && (env_cs & (env_cs - 1)) == 0 && Lb / env_cs <= 200 * 1024) cs = env_cs;
        }
        while (cs > 1 && ((L % cs) != 0 || ((Lb / cs) % 16) != 0)) cs /= 2;
        const long long slab_bytes = Lb / cs;
        if (aligned && slab_bytes <= 200 * 1024 && slab_bytes >= 16 && (L / cs) % VEC == 0 && L < (1LL << 31)) {
            static bool configured = false;
            if (!configured) {
                SDOD_TRY(check_cuda(cudaFuncSetAttribute(gn_nchw_cluster_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
                                    "cudaFuncSetAttribute(gn)"));
                configured = true;
            }
            cudaLaunchConfig_t cfg{};
            cfg.gridDim = dim3(static_cast<unsigned>(N) * G * cs);
            cfg.blockDim = dim3(kGnClusterThreads);
            cfg.dynamicSmemBytes = static_cast<size_t>(slab_bytes);
            cfg.stream = stream;
            cudaLaunchAttribute attr[2];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
            attr[1].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
            int slab_elems = static_cast<int>(L / cs);
            SDOD_TRY(check_cuda(cudaLaunchKernelEx(&cfg, gn_nchw_cluster_kernel<T>, x, y, weight, bias, add_nc, C, HW, G, eps, fuse_silu, slab_elems),
                                "gn_nchw_cluster_kernel"));
            count_launch();
            return kOk;
        }
        gn_nchw_generic_kernel<T><<<N * G, kGnThreads, 0, stream>>>(x, y, weight, bias, add_nc, C, HW, G, eps, fuse_silu);
        count_launch();
        return check_launch("gn_nchw_generic_kernel");
    }
    return fail(kInvalidArgument, "group_norm: unknown layout");
}

int group_norm(cudaStream_t stream, const void* x, void* y, const float* weight, const float* bias, const float* add_nc, int N, int C,
               int HW, int G, float eps, int dtype, int layout, int fuse_silu, void* ws, size_t ws_bytes) {
    if (!x || !y) return fail(kInvalidArgument, "group_norm: NULL tensor");
    if (N <= 0 || C <= 0 || HW <= 0 || G <= 0) return fail(kInvalidArgument, "group_norm: non-positive extent");
    if (C % G != 0) return fail(kInvalidArgument, "num_channels must be divisible by num_groups");   // efficient_gn.py:37-38
    if ((weight == nullptr) != (bias == nullptr)) return fail(kInvalidArgument, "group_norm: weight and bias must both be given or both be NULL");  // efficient_gn.py:17-22
    if (layout == SDOD_NHWC) return group_norm_nhwc(stream, x, dtype, y, dtype, weight, bias, add_nc, N, C, HW, G, eps, fuse_silu, ws, ws_bytes);
    if (dtype == SDOD_F32)
        return group_norm_typed<float>(stream, static_cast<const float*>(x), static_cast<float*>(y), weight, bias, add_nc, N, C, HW, G, eps, layout, fuse_silu, ws, ws_bytes);
    if (dtype == SDOD_BF16)
        return group_norm_typed<bf16>(stream, static_cast<const bf16*>(x), static_cast<bf16*>(y), weight, bias, add_nc, N, C, HW, G, eps, layout, fuse_silu, ws, ws_bytes);
    return fail(kInvalidArgument, "group_norm: unknown dtype");
}

}  // namespace sdod

extern "C" {
SDOD_API size_t sdod_group_norm_workspace(int N, int C, int HW, int num_groups, int layout) {
    (void)C; (void)HW;
    if (layout != SDOD_NHWC) return 0;
    return sdod::kGnCounterInts * sizeof(unsigned int) + static_cast<size_t>(N) * num_groups * 2 * sizeof(float) * (1 + sdod::kGnMaxSlabs);
}
SDOD_API int sdod_group_norm_nhwc(sdod_stream_t stream, const void* x, int in_dtype, void* y, int out_dtype, const float* weight,
                                  const float* bias, const float* add_nc, int N, int C, int HW, int num_groups, float eps, int fuse_silu,
                                  void* workspace, size_t workspace_bytes) {
    return sdod::group_norm_nhwc(static_cast<cudaStream_t>(stream), x, in_dtype, y, out_dtype, weight, bias, add_nc, N, C, HW, num_groups, eps,
                                 fuse_silu, workspace, workspace_bytes);
}
SDOD_API int sdod_group_norm_nhwc2(sdod_stream_t stream, const void* xa, int Ca, const void* xb, int Cb, int in_dtype, void* y, int out_dtype,
                                   void* raw_bf16, const float* weight, const float* bias, int N, int HW, int num_groups, float eps,
                                   int fuse_silu, void* workspace, size_t workspace_bytes) {
    return sdod::group_norm_nhwc2(static_cast<cudaStream_t>(stream), xa, Ca, xb, Cb, in_dtype, y, out_dtype, raw_bf16, weight, bias, N, HW,
                                  num_groups, eps, fuse_silu, workspace, workspace_bytes);
}
SDOD_API int sdod_group_norm_nhwc2_supported(int N, int Ca, int Cb, int HW, int num_groups, int in_dtype) {
    return sdod::group_norm_fused_eligible(N, Ca, Cb, HW, num_groups, in_dtype) ? 1 : 0;
}
SDOD_API int sdod_group_norm(sdod_stream_t stream, const void* x, void* y, const float* weight, const float* bias, const float* add_nc,
                             int N, int C, int HW, int num_groups, float eps, int dtype, int layout, int fuse_silu, void* workspace,
                             size_t workspace_bytes) {
    return sdod::group_norm(static_cast<cudaStream_t>(stream), x, y, weight, bias, add_nc, N, C, HW, num_groups, eps, dtype, layout, fuse_silu,
                            workspace, workspace_bytes);
}
}
