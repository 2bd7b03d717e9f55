// tcgen05/TMEM GEMM and implicit-GEMM conv3x3 for sm_100a.
//
//   C[M,N] = epilogue( alpha * A[M,K] * W[N,K]^T ),  bf16 operands, fp32 accumulation in TMEM.
//
// One CTA computes 128 x BN output tiles.  Warp roles (320 threads; 576 in the persistent variant):
//   warp 0      : TMA producer  (cp.async.bulk.tensor, 128B-swizzled K-major tiles, mbarrier ring)
//   warp 1      : MMA issuer    (one thread issues tcgen05.mma 128xBNx16 — or 256xBNx16 cta_group::2 for CTA pairs — and commits to mbarriers)
//   warps 2..   : epilogue      (tcgen05.ld 32x32b -> alpha / bias / temb row-bias / SiLU / GELU / GEGLU / residual -> staged in shared
//                                memory -> TMA stores; attention head layouts as TMA boxes; split-K partial dump; a coalesced LSU fallback).
//                                8 warps (16 when persistent): NPART warps share each TMEM lane quarter and split its columns.
// Variants (template flags, chosen per layer shape by the host code at the bottom of this file): shallow ring with 2 CTAs/SM, DEEP ring for
// single-wave grids, PAIR = 2-CTA clusters on cta_group::2, PERSIST = one CTA per SM walking the tile list with a double-buffered TMEM accumulator.
// Conv mode feeds the same mainloop: the A tile for tap (ky,kx) and channel block c is one 4-D TMA
// box {64 ch, bw, bh, bb} of the NHWC activation at (c, x0+kx-1, y0+ky-1, b0); TMA's out-of-bounds
// zero fill supplies the padding, so no im2col buffer exists in HBM.
//
// Layer list served (reference evidence: analyze_results.py:25-87; executed by the opaque
// unet/decoder graphs at csrc/libsdod/src/context.cpp:352,366,387).
#include <cstdlib>
#include "../common.cuh"
#include "../host_common.h"
#include "../launch_count.h"
#include "launchers.h"
#include "sdod_kernels.h"

#include <cstring>
#include <type_traits>

namespace sdod {

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;                    // 64 bf16 = one 128-byte swizzle row
constexpr int kABytes = kBlockM * kBlockK * 2;  // 16 KiB
constexpr int kGemmThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 epilogue (two per TMEM lane quarter)
constexpr int kGemmThreadsPersist = 576;   // persistent variant: 16 epilogue warps (four per lane quarter) on the one CTA of an SM

// DEEP = false: shallow ring so that 2 CTAs co-reside per SM (one CTA's epilogue overlaps the other's mainloop) — used when
//                the grid is larger than one wave.
// DEEP = true : single-wave grids (<= 148 CTAs, the batch-2 UNet case): one CTA owns the SM, so the ring takes all of
//                shared memory; by Little's law the per-SM load bandwidth is bytes-in-flight / L2 latency, and at 3 stages
//                the K loop ran latency-bound at ~1.3 us per K block (profiles/r01_*).
// PAIR = true : two CTAs on the SMs of one TPC run each K block as ONE tcgen05.mma.cta_group::2 (M = 256): every CTA loads
//                its own 128 A rows but only half of the B tile, so the weight traffic L2 -> SM halves.  Long-K layers are
//                bound by exactly that traffic (~6300 B/clk chip-wide from L2; profiles/r01_gemm_l2_bound.txt).
// PERSIST = true: one CTA per SM walks a strided list of output tiles.  The TMA ring keeps running across tile boundaries,
//                the accumulator is double-buffered in TMEM (2 x BN columns) and the epilogue stages its output in its
//                own shared-memory region, so tile i's epilogue overlaps tile i+1's loads and MMAs.  Short-K layers
//                (K = 320 / 640: 5-10 K blocks) were bound by the per-CTA fill + drain chain (~6 us per 128x128 tile whose
//                MMAs take 0.7 us, profiles/r01_hot_kernels_ncu.txt); persistence hides that chain behind the next tile.
template <int BN, bool DEEP, bool PAIR = false, bool PERSIST = false>
struct GemmCfg {
    static constexpr int kBBytes = (PAIR ? BN / 2 : BN) * kBlockK * 2;   // B rows held by this CTA
    static constexpr int kStageBytes = kABytes + kBBytes;
    static constexpr int kShallow = PAIR ? (BN >= 256 ? 5 : 4) : ((BN >= 256) ? 4 : (BN >= 128 ? 3 : 4));
    static constexpr int kDeepRaw = (220 * 1024 - 2048) / kStageBytes;
    static constexpr int kEpiBytes = PERSIST ? BN * 512 : 0;          // dedicated epilogue staging (4 quarters x BN/32 boxes x 4 KiB)
    static constexpr int kBiasBytes = 2 * BN * 4 < 1024 ? 1024 : 2 * BN * 4;      // two bias tiles (double-buffered when persistent)
    static constexpr int kPersistRaw = (232448 - 1024 - 256 - kBiasBytes - kEpiBytes) / kStageBytes;
    static constexpr int kStages = PERSIST ? (kPersistRaw > 8 ? 8 : kPersistRaw) : (DEEP ? (kDeepRaw > 8 ? 8 : kDeepRaw) : kShallow);
    // persistent: as many accumulators as the 512 TMEM columns hold (4 x 128 or 3 x 160) — the MMA thread runs up to kAccs - 1 tiles ahead of the
    // epilogue groups (with two buffers it waited 30 % of its time for a drained accumulator, profiles/r02_geglu_source_top.txt)
    static constexpr int kAccs = PERSIST ? (512 / BN > 4 ? 4 : 512 / BN) : 1;
    static constexpr int kTmemCols = PERSIST ? 512 : (BN <= 32 ? 32 : (BN <= 64 ? 64 : (BN <= 128 ? 128 : 256)));
    static constexpr int kSmemBytes = kStages * kStageBytes + kEpiBytes + 1024 /*align slack*/ + 256 /*barriers*/ + kBiasBytes;
    static_assert(PERSIST || kStages * kStageBytes >= 512 * (BN + 4), "the idle ring doubles as the epilogue staging area");
    static_assert(kSmemBytes <= 232448, "exceeds 227 KiB of shared memory");
};

// Profiling hook: one thread stamps phase `slot` of this CTA with the global nanosecond timer (results are unaffected).
SDOD_DEVICE void tstamp(const MainloopParams& mp, int slot) {
    if (mp.tlog) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        const unsigned long long cta = (static_cast<unsigned long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        mp.tlog[cta * 16 + slot] = t;
    }
}

SDOD_DEVICE float apply_act(float v, int act) {
    if (act == SDOD_ACT_SILU) return silu_f(v);
    if (act == SDOD_ACT_GELU) return gelu_f(v);
    if (act == SDOD_ACT_QUICK_GELU) return quick_gelu_f(v);
    return v;
}

// Store 8 consecutive output columns [n, n+8) of row m (values already activated).
SDOD_DEVICE void store8(const sdod_epilogue& ep, long long zoff_c, long long zoff_r, int m, int n, int N, float (&v)[8]) {
    if (ep.residual && ep.residual_f32) {
        const float* r = reinterpret_cast<const float*>(ep.residual) + zoff_r + static_cast<long long>(m) * ep.ldr + n;
        if (n + 8 <= N && ((reinterpret_cast<uintptr_t>(r) & 15) == 0)) {
            const float4 a = *reinterpret_cast<const float4*>(r), b = *reinterpret_cast<const float4*>(r + 4);
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w; v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (n + i < N) v[i] += r[i];
        }
    } else if (ep.residual) {
        const bf16* r = reinterpret_cast<const bf16*>(ep.residual) + zoff_r + static_cast<long long>(m) * ep.ldr + n;
        if (n + 8 <= N && ((reinterpret_cast<uintptr_t>(r) & 15) == 0)) {
            uint4 u = *reinterpret_cast<const uint4*>(r);
            float2 f;
            f = unpack_bf16x2(u.x); v[0] += f.x; v[1] += f.y;
            f = unpack_bf16x2(u.y); v[2] += f.x; v[3] += f.y;
            f = unpack_bf16x2(u.z); v[4] += f.x; v[5] += f.y;
            f = unpack_bf16x2(u.w); v[6] += f.x; v[7] += f.y;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (n + i < N) v[i] += __bfloat162float(r[i]);
        }
    }
    int mode = ep.out_mode;
    void* base = ep.C;
    int nn = n;
    if (mode == SDOD_OUT_QKV) {
        int Cw = ep.heads * ep.head_dim;
        int which = n / Cw;
        nn = n - which * Cw;
        base = which == 0 ? ep.C : (which == 1 ? ep.C2 : ep.C3);
        mode = which == 2 ? SDOD_OUT_HEADS_T : SDOD_OUT_HEADS;
    }
    if (mode == SDOD_OUT_BF16) {
        bf16* c = reinterpret_cast<bf16*>(base) + zoff_c + static_cast<long long>(m) * ep.ldc + nn;
        if (n + 8 <= N && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
            uint4 u;
            u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
            u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(c) = u;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (n + i < N) c[i] = __float2bfloat16(v[i]);
        }
    } else if (mode == SDOD_OUT_F32) {
        float* c = reinterpret_cast<float*>(base) + zoff_c + static_cast<long long>(m) * ep.ldc + nn;
        if (n + 8 <= N && ((reinterpret_cast<uintptr_t>(c) & 15) == 0)) {
            *reinterpret_cast<float4*>(c) = make_float4(v[0], v[1], v[2], v[3]);
            *reinterpret_cast<float4*>(c + 4) = make_float4(v[4], v[5], v[6], v[7]);
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (n + i < N) c[i] = v[i];
        }
    } else if (mode == SDOD_OUT_HEADS) {
        int b = m / ep.tokens, t = m - b * ep.tokens;
        int h = nn / ep.head_dim, d = nn - h * ep.head_dim;
        bf16* c = reinterpret_cast<bf16*>(base) + ((static_cast<long long>(b) * ep.heads + h) * ep.tokens + t) * ep.dpad + d;
        if (n + 8 <= N && d + 8 <= ep.head_dim) {
            uint4 u;
            u.x = pack_bf16x2(v[0], v[1]); u.y = pack_bf16x2(v[2], v[3]);
            u.z = pack_bf16x2(v[4], v[5]); u.w = pack_bf16x2(v[6], v[7]);
            *reinterpret_cast<uint4*>(c) = u;
        } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (n + i < N) {
                    int hh = (nn + i) / ep.head_dim, dd = (nn + i) - hh * ep.head_dim;
                    reinterpret_cast<bf16*>(base)[((static_cast<long long>(b) * ep.heads + hh) * ep.tokens + t) * ep.dpad + dd] = __float2bfloat16(v[i]);
                }
            }
        }
    } else {  // SDOD_OUT_HEADS_T: [B*heads, dpad, tok_pad], token contiguous (lanes = consecutive tokens)
        int b = m / ep.tokens, t = m - b * ep.tokens;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (n + i < N) {
                int hh = (nn + i) / ep.head_dim, dd = (nn + i) - hh * ep.head_dim;
                reinterpret_cast<bf16*>(base)[((static_cast<long long>(b) * ep.heads + hh) * ep.vt_rows + dd) * ep.tok_pad + t] = __float2bfloat16(v[i]);
            }
        }
    }
}

// v[0..NV) = act(v * alpha + bias[n..] + rb[n..]) with every runtime branch hoisted out of the element loop and 16-B
// parameter loads (the per-element branchy form cost ~70 SASS instructions per element and made the epilogue
// instruction-bound: 15 us of a 23 us K=320 GEMM, profiles/r01_gemm_phase_timing.txt).
template <int NV>
SDOD_DEVICE void bias_act(float (&v)[NV], const sdod_epilogue& ep, const float* rb, int n, int N) {
    static_assert(NV % 4 == 0, "NV must be a multiple of 4");
    if (ep.alpha != 1.0f) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] *= ep.alpha;
    }
    const bool full = (n + NV <= N);
    if (ep.bias) {
        const float* b = ep.bias + n;
        if (full && ((reinterpret_cast<uintptr_t>(b) & 15) == 0)) {
#pragma unroll
            for (int q4 = 0; q4 < NV / 4; ++q4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(b) + q4);
                v[4 * q4] += t.x; v[4 * q4 + 1] += t.y; v[4 * q4 + 2] += t.z; v[4 * q4 + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (n + i < N) v[i] += b[i];
        }
    }
    if (rb) {
        const float* b = rb + n;
        if (full && ((reinterpret_cast<uintptr_t>(b) & 15) == 0)) {
#pragma unroll
            for (int q4 = 0; q4 < NV / 4; ++q4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(b) + q4);
                v[4 * q4] += t.x; v[4 * q4 + 1] += t.y; v[4 * q4 + 2] += t.z; v[4 * q4 + 3] += t.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < NV; ++i)
                if (n + i < N) v[i] += b[i];
        }
    }
    if (ep.act == SDOD_ACT_SILU) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = silu_f(v[i]);
    } else if (ep.act == SDOD_ACT_GELU) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = gelu_f(v[i]);
    } else if (ep.act == SDOD_ACT_QUICK_GELU) {
#pragma unroll
        for (int i = 0; i < NV; ++i) v[i] = quick_gelu_f(v[i]);
    }
}

// Epilogue of 8 accumulator columns [n, n+8) of output row m: alpha, bias, timestep row-bias, activation, residual, store.
SDOD_DEVICE void epilogue_plain8(const sdod_epilogue& ep, const MainloopParams& mp, int bz, int m, int n, const float (&acc)[8]) {
    if (m >= mp.M || n >= mp.N) return;
    const long long zc = static_cast<long long>(bz) * ep.strideC, zr = static_cast<long long>(bz) * ep.strideR;
    const float* rb = ep.row_bias ? ep.row_bias + static_cast<long long>(m / ep.rows_per_group) * (ep.ld_row_bias ? ep.ld_row_bias : mp.N) : nullptr;
    float v[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = acc[i];
    bias_act<8>(v, ep, rb, n, mp.N);
    store8(ep, zc, zr, m, n, mp.N, v);
}
SDOD_DEVICE void epilogue_plain16(const sdod_epilogue& ep, const MainloopParams& mp, int bz, int m, int n, const uint32_t (&acc)[16]) {
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
        float a8[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) a8[i] = __uint_as_float(acc[h8 * 8 + i]);
        epilogue_plain8(ep, mp, bz, m, n + h8 * 8, a8);
    }
}

// GEGLU: tile columns [0,BN/2) hold the value half, [BN/2,BN) the gate half; j indexes the value half.
template <int BN>
SDOD_DEVICE void epilogue_geglu16(const sdod_epilogue& ep, const MainloopParams& mp, int bz, int m, int n_tile, int j, const uint32_t (&a)[16],
                                  const uint32_t (&g)[16]) {
    if (m >= mp.M) return;
    constexpr int HALF = BN / 2;
    const long long zc = static_cast<long long>(bz) * ep.strideC, zr = static_cast<long long>(bz) * ep.strideR;
    const int n_out_total = mp.N / 2;
    const int n0 = n_tile * BN;
    float av[16], gv[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) { av[i] = __uint_as_float(a[i]) * ep.alpha; gv[i] = __uint_as_float(g[i]) * ep.alpha; }
    if (ep.bias) {
        const int na = n0 + j, ng = na + HALF;
        if (ng + 16 <= mp.N && ((reinterpret_cast<uintptr_t>(ep.bias + na) & 15) == 0) && ((reinterpret_cast<uintptr_t>(ep.bias + ng) & 15) == 0)) {
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 ta = __ldg(reinterpret_cast<const float4*>(ep.bias + na) + q4), tg = __ldg(reinterpret_cast<const float4*>(ep.bias + ng) + q4);
                av[4 * q4] += ta.x; av[4 * q4 + 1] += ta.y; av[4 * q4 + 2] += ta.z; av[4 * q4 + 3] += ta.w;
                gv[4 * q4] += tg.x; gv[4 * q4 + 1] += tg.y; gv[4 * q4 + 2] += tg.z; gv[4 * q4 + 3] += tg.w;
            }
        } else {
#pragma unroll
            for (int i = 0; i < 16; ++i)
                if (ng + i < mp.N) { av[i] += ep.bias[na + i]; gv[i] += ep.bias[ng + i]; }
        }
    }
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
        float v[8];
        const int no = n_tile * HALF + j + h8 * 8;   // output column
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = av[h8 * 8 + i] * gelu_f(gv[h8 * 8 + i]);
        if (no < n_out_total) store8(ep, zc, zr, m, no, n_out_total, v);
    }
}

// Fold the `split` fp32 partials of 8 accumulator columns (two float4 per split, `zstride4` float4 apart) in fixed z order.
SDOD_DEVICE void fold_partials8(const float4* src, int split, long long zstride4, float (&acc)[8]) {
    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
    int z = 0;
    for (; z + 4 <= split; z += 4) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 4; ++u) { v[2 * u] = __ldcg(src + (z + u) * zstride4); v[2 * u + 1] = __ldcg(src + (z + u) * zstride4 + 1); }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            s0.x += v[2 * u].x; s0.y += v[2 * u].y; s0.z += v[2 * u].z; s0.w += v[2 * u].w;
            s1.x += v[2 * u + 1].x; s1.y += v[2 * u + 1].y; s1.z += v[2 * u + 1].z; s1.w += v[2 * u + 1].w;
        }
    }
    for (; z < split; ++z) {
        const float4 a = __ldcg(src + z * zstride4), b = __ldcg(src + z * zstride4 + 1);
        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w; s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
    }
    acc[0] = s0.x; acc[1] = s0.y; acc[2] = s0.z; acc[3] = s0.w; acc[4] = s1.x; acc[5] = s1.y; acc[6] = s1.z; acc[7] = s1.w;
}

// Split-K second phase (fallback when the split CTAs cannot form a cluster, and for GEGLU): folds the `split` partial tiles in
// fixed order (deterministic) and runs the shared epilogue.
// Plain epilogues: 256 threads = 128 rows x 2 column octets, so a warp reads 16 rows x 64 B contiguous partials and writes
// 16 row segments of 32 B (fp32) — coalesced both ways.  grid = (BN/16 chunks, tiles).
template <int BN>
__global__ void __launch_bounds__(256) splitk_reduce_kernel(const MainloopParams mp, const sdod_epilogue ep, int n_tiles) {
    griddep_wait();
    griddep_launch();
    const int tile = blockIdx.y;
    const int m_tile = tile / n_tiles, n_tile = tile - m_tile * n_tiles;
    const float* base = mp.ws + static_cast<long long>(tile) * mp.split * (BN * kBlockM);
    const long long zstride4 = BN * kBlockM / 4;
    if (ep.act != SDOD_ACT_GEGLU) {
        const int row = threadIdx.x >> 1, half = threadIdx.x & 1;
        const int j = blockIdx.x * 16;
        const float4* src = reinterpret_cast<const float4*>(base + ((j >> 4) * kBlockM + row) * 16 + half * 8);
        float acc[8];
        fold_partials8(src, mp.split, zstride4, acc);
        epilogue_plain8(ep, mp, 0, m_tile * kBlockM + row, n_tile * BN + j + half * 8, acc);
        return;
    }
    // GEGLU: value and gate halves are BN/2 apart; one thread per row
    if (threadIdx.x >= 128) return;
    const int row = threadIdx.x;
    constexpr int HALF = BN / 2;
    const int j = blockIdx.x * 16;
    if (j >= HALF) return;
    auto fold = [&](int jj, uint32_t (&acc)[16]) {
        float sacc[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) sacc[i] = 0.f;
        const float4* src = reinterpret_cast<const float4*>(base + ((jj >> 4) * kBlockM + row) * 16);
        for (int z = 0; z < mp.split; ++z)
#pragma unroll
            for (int q4 = 0; q4 < 4; ++q4) {
                const float4 v = __ldcg(src + z * zstride4 + q4);
                sacc[4 * q4] += v.x; sacc[4 * q4 + 1] += v.y; sacc[4 * q4 + 2] += v.z; sacc[4 * q4 + 3] += v.w;
            }
#pragma unroll
        for (int i = 0; i < 16; ++i) acc[i] = __float_as_uint(sacc[i]);
    };
    uint32_t a[16], g[16];
    fold(j, a);
    fold(HALF + j, g);
    epilogue_geglu16<BN>(ep, mp, 0, m_tile * kBlockM + row, n_tile, j, a, g);
}

// persistent CTAs take every gridDim.x-th tile (N fastest).  (Contiguous chunks per CTA, which would let the producer keep the A rows
// resident across the N tiles of one M tile, measured 10 % slower on the 64x64 GEGLU projection: the CTAs of a wave then touch
// 148 different A tiles at once instead of ~8, and the layer is epilogue-bound, not L2-bound — profiles/r01_gemm_load_vs_mma.txt.)
// Stream-K (persistent CTAs only, mp.streamk): instead of whole tiles each CTA takes a contiguous, equal share [u, u_end) of the linearised
// (tile, K block) space and walks it as per-tile segments [seg_lo, seg_hi).  Layers with few tiles and long K (every conv and most linear
// layers of the batch-2 step) otherwise leave SMs idle or run a second, mostly empty wave; here all CTAs finish together.  A segment that does
// not start at K block 0 publishes its fp32 partial tile (each CTA publishes at most one: the head of its share) and raises a flag; the CTA
// whose segment starts the tile — always the LAST segment of its own share, so the partials of the later K ranges were published at the very
// start of their owners' shares — folds them into its accumulator reads, in fixed order (deterministic), and runs the epilogue.
struct WorkWalk {
    int t, iter, seg_lo, seg_hi, t_step, t_end, kblocks;
    long long u, u_end;
    bool live, sk;
    __device__ WorkWalk(const MainloopParams& mp, bool persist, int kb0, int kb1) {
        iter = 0; kblocks = mp.k_blocks; sk = persist && mp.streamk; u = u_end = 0;
        if (!persist) { t = 0; t_step = 1; t_end = 1; seg_lo = kb0; seg_hi = kb1; live = true; }
        else if (!sk) { t = blockIdx.x; t_step = gridDim.x; t_end = mp.tiles_total; seg_lo = 0; seg_hi = kblocks; live = t < t_end; }
        else {
            const long long total = static_cast<long long>(mp.tiles_total) * kblocks;
            t_step = 0; t_end = mp.tiles_total;
            u = total * blockIdx.x / gridDim.x;
            u_end = total * (blockIdx.x + 1) / gridDim.x;
            live = u < u_end;
            if (live) cut();
        }
    }
    __device__ void cut() {
        t = static_cast<int>(u / kblocks);
        seg_lo = static_cast<int>(u - static_cast<long long>(t) * kblocks);
        const long long hi = seg_lo + (u_end - u);
        seg_hi = hi < kblocks ? static_cast<int>(hi) : kblocks;
    }
    __device__ void next() {
        ++iter;
        if (!sk) { t += t_step; live = t < t_end; }
        else { u += seg_hi - seg_lo; live = u < u_end; if (live) cut(); }
    }
    // first unit of CTA c's share (stream-K)
    static __device__ long long share_begin(const MainloopParams& mp, int c) {
        return static_cast<long long>(mp.tiles_total) * mp.k_blocks * c / gridDim.x;
    }
};
#define SDOD_TILE_LOOP                                                                                                               \
    for (WorkWalk wk(mp, PERSIST, kb0, kb1); wk.live; wk.next())
#define SDOD_TILE_COORDS                                                                                               \
    const int t = wk.t, iter = wk.iter; (void)t; (void)iter;                                                           \
    const int n_tile = PERSIST ? t % mp.n_tiles : static_cast<int>(PAIR ? blockIdx.y : blockIdx.x);                    \
    const int m_tile = PERSIST ? (t / mp.n_tiles) % mp.m_tiles : static_cast<int>(PAIR ? blockIdx.x : blockIdx.y);     \
    const int bz = PERSIST ? t / (mp.n_tiles * mp.m_tiles) : (mp.split > 1 ? 0 : static_cast<int>(blockIdx.z));        \
    const int m0 = m_tile * kBlockM, n0 = n_tile * BN;                                                                 \
    (void)bz; (void)m0; (void)n0;

template <int BN, bool DEEP, bool PAIR, bool PERSIST>
__global__ void __launch_bounds__(PERSIST ? kGemmThreadsPersist : kGemmThreads, (DEEP || PERSIST) ? 1 : 2) gemm_tcgen05_kernel(const __grid_constant__ CUtensorMap tmA,
                                                                      const __grid_constant__ CUtensorMap tmW,
                                                                      const __grid_constant__ CUtensorMap tmC,
                                                                      const __grid_constant__ CUtensorMap tmR,
                                                                      const __grid_constant__ CUtensorMap tmC2,
                                                                      const __grid_constant__ CUtensorMap tmA2,
                                                                      const MainloopParams mp, const sdod_epilogue ep) {
    using Cfg = GemmCfg<BN, DEEP, PAIR, PERSIST>;
    static_assert(!(PAIR && PERSIST), "persistent CTA pairs are not implemented");
    constexpr int STAGES = Cfg::kStages;
    constexpr int EW = PERSIST ? 16 : 8;          // epilogue warps: NPART per TMEM lane quarter, each owning a column range
    constexpr int NPART = EW / 4;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned, still a shared-space pointer
    const uint32_t rank = PAIR ? cluster_ctarank() : 0u;      // 0 = leader of the CTA pair (issues the MMAs)
    uint8_t* sA = smem;
    uint8_t* sB = smem + STAGES * kABytes;
    uint8_t* stage = PERSIST ? sB + STAGES * Cfg::kBBytes : smem;     // epilogue staging: own region, or the idle ring
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sB + STAGES * Cfg::kBBytes + Cfg::kEpiBytes);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;             // [4] accumulator ready (one per TMEM buffer)
    uint64_t* tmem_empty_bar = tmem_full_bar + 4;             // [4] accumulator drained by the epilogue (persistent only)
    uint64_t* res_bar = tmem_empty_bar + 4;                   // [4] residual tiles landed (TMA epilogue), one per lane quarter
    static_assert((2 * STAGES + 14) * 8 + 4 <= 256, "barrier block");
    const int naccs = PERSIST ? min(Cfg::kAccs, mp.tmem_accs > 0 ? mp.tmem_accs : Cfg::kAccs) : 1;
    uint64_t* ln_bar = res_bar + 4;                           // [2] LayerNorm epilogue: the cluster's row sums / squared deviations have landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(ln_bar + 2);
    float* s_bias_base = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(full_bar) + 256);   // [2][BN] bias tiles

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;
    const bool pair_relaxed = PAIR && mp.pair_relaxed;
    if (threadIdx.x == 0) tstamp(mp, 0);                                   // CTA started
    // Tile coordinates come from SDOD_TILE_COORDS inside each role's tile loop (one trip unless PERSIST).  Pair grids put M on
    // x: the two CTAs of a cluster (dims 2x1x1, as cta_group::2 kernels must be launched) are consecutive M tiles.
    const int zs = (!PERSIST && mp.split > 1) ? blockIdx.z : 0;            // split index (split-K only when batch == 1)
    const int kb0 = zs * mp.kb_per_split;
    const int kb1 = (!PERSIST && mp.split > 1) ? min(mp.k_blocks, kb0 + mp.kb_per_split) : mp.k_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA);
        tma_prefetch_desc(&tmW);
        if (mp.kb_a2 < mp.k_blocks) tma_prefetch_desc(&tmA2);
    }
    if (warp == 0 && lane >= 1 && mp.pf_bytes > 0) {
        // Pull the NEXT layer's weights towards L2 while this kernel runs (weights are constants: no dependency on the predecessor, so
        // this is issued before griddepcontrol.wait).  At batch 2 the step streams 1.7 GB of weights through ~200 short kernels; without
        // the prefetch every kernel starts with an exposed HBM round trip and the deep levels run at ~1/3 of HBM bandwidth.
        const long long ncta = static_cast<long long>(gridDim.x) * gridDim.y * gridDim.z;
        const long long cta = (static_cast<long long>(blockIdx.z) * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
        const long long chunk = (((mp.pf_bytes + ncta - 1) / ncta) + 4095) & ~4095LL;
        const long long end = min(mp.pf_bytes, (cta + 1) * chunk);
        for (long long off = cta * chunk + (lane - 1) * 4096LL; off < end; off += 31 * 4096LL) {
            const uint32_t nb = static_cast<uint32_t>(min(4096LL, end - off)) & ~15u;
            if (nb) asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(mp.pf_ptr + off), "r"(nb) : "memory");
        }
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        for (int i = 0; i < 4; ++i) { mbar_init(&tmem_full_bar[i], 1); mbar_init(&tmem_empty_bar[i], (PERSIST && mp.epi_groups == 2) ? 16 * EW : 32 * EW); }
        for (int i = 0; i < 4; ++i) mbar_init(&res_bar[i], 1);
        for (int i = 0; i < 2; ++i) mbar_init(&ln_bar[i], 1);
        fence_mbar_init();
    }
    if (warp == 2) {
        if (PAIR) tmem_alloc_pair<Cfg::kTmemCols>(tmem_slot);
        else tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    }
    tc_fence_before();
    __syncthreads();
    // the peer's barriers exist before any TMA / commit of ours can signal them.  Execution barrier only (fence.mbarrier_init above publishes the
    // barriers): the release/acquire form costs a MEMBAR.ALL.GPU per warp
    if (PAIR) { if (pair_relaxed) { cluster_arrive_relaxed(); cluster_wait(); } else cluster_sync_all(); }
    if (!PAIR && !PERSIST && mp.ln_fuse) cluster_arrive_relaxed();   // "my ln_bar exists": waited for right before the first row-sum push
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    // PDL: everything above overlapped the previous kernel's tail; its outputs are visible from here.  CTA-pair kernels stay out of
    // it by default (ordinary launch, no early trigger): a cta_group::2 TMEM allocation racing the neighbour kernel's allocation on the
    // same SM pair deadlocked the step in round 1.  With the allocation now ahead of the trigger that no longer reproduces
    // (SDOD_PAIR_PDL=1: a batch-2 step with every eligible layer forced onto pairs replays cleanly), but it gains nothing measurable
    // where pairs are used (batch-32 pass 39.85 vs 39.96 ms), so the switch stays off.
    if (threadIdx.x == 0) tstamp(mp, 1);                                   // barriers + TMEM ready
    if (!PAIR || mp.pair_pdl) {
        griddep_wait();
        griddep_launch();
    }
    if (threadIdx.x == 0) tstamp(mp, 2);                                   // predecessor complete

    if (warp == 0) {
        if (lane == 0) {
            // ---------------------------------------------------------------- TMA producer
            const uint32_t leader_full = PAIR ? mapa_shared(&full_bar[0], 0) : 0u;   // pair: all bytes are credited to the leader
            int g = 0;                                         // ring position, keeps counting across tiles
            SDOD_TILE_LOOP {
            SDOD_TILE_COORDS
            int b0 = 0, y0 = 0, x0 = 0;
            if (mp.conv) {
                const int hw = mp.H * mp.W;
                if (mp.bb > 1) {
                    b0 = m_tile * mp.bb;
                } else {
                    b0 = m0 / hw;
                    int rem = m0 - b0 * hw;
                    y0 = rem / mp.W;
                    x0 = rem - y0 * mp.W;
                }
            }
            // K rotation: M tile i starts its K loop i*rot blocks in, so the CTAs of a wave do not all pull the same weight
            // tile out of the same few L2 slices at the same moment (the sum over K is order-independent up to fp32 rounding,
            // and the order is fixed per tile, so results stay deterministic).
            const int seg_lo = wk.seg_lo, seg_hi = wk.seg_hi, nkb = seg_hi - seg_lo;
            int kb = seg_lo + (mp.k_rot ? static_cast<int>((static_cast<long long>(PAIR ? m_tile >> 1 : m_tile) * mp.k_rot) % nkb) : 0);
            for (int it = 0; it < nkb; ++it, ++g, kb = (kb + 1 == seg_hi ? seg_lo : kb + 1)) {
                const int s = g % STAGES;
                const uint32_t ph = (g / STAGES) & 1;
                mbar_wait(&empty_bar[s], ph ^ 1);
                if (PAIR) {
                    if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * Cfg::kStageBytes);     // both CTAs' A + B halves
                    const uint32_t fb = leader_full + s * 8;
                    if (kb >= mp.kb_a2) {                       // K-concatenated second operand: plain rows [m0, m0+128) of A2
                        tma_load_3d_pair(sA + s * kABytes, &tmA2, fb, (kb - mp.kb_a2) * kBlockK, m0, bz);
                    } else if (mp.conv) {
                        const int tap = kb / mp.cin_blocks;
                        const int cb = kb - tap * mp.cin_blocks;
                        // 3x3 taps, or (sub-pixel form of upsample + conv) the 2x2 source taps of output parity bz = 2*py + px
                        const int ky = mp.up2 ? (tap >> 1) + (bz >> 1) : tap / 3, kx = mp.up2 ? (tap & 1) + (bz & 1) : tap - ky * 3;
                        tma_load_4d_pair(sA + s * kABytes, &tmA, fb, cb * kBlockK, mp.cstride * x0 + kx - 1, mp.cstride * y0 + ky - 1, b0);
                    } else {
                        tma_load_3d_pair(sA + s * kABytes, &tmA, fb, kb * kBlockK, m0, bz);
                    }
                    tma_load_3d_pair(sB + s * Cfg::kBBytes, &tmW, fb, kb * kBlockK, n0 + static_cast<int>(rank) * (BN / 2), mp.w_batched ? bz : 0);
                    continue;
                }
                // weight-stationary walk: after this CTA's first tile the B half of every stage already holds the right K block of its one N tile
                const bool load_b = !(PERSIST && mp.b_resident && wk.iter > 0);
                mbar_arrive_expect_tx(&full_bar[s], load_b ? Cfg::kStageBytes : kABytes);
                if (kb >= mp.kb_a2) {
                    tma_load_3d(sA + s * kABytes, &tmA2, &full_bar[s], (kb - mp.kb_a2) * kBlockK, m0, bz);
                } else if (mp.conv) {
                    const int tap = kb / mp.cin_blocks;
                    const int cb = kb - tap * mp.cin_blocks;
                    const int ky = mp.up2 ? (tap >> 1) + (bz >> 1) : tap / 3, kx = mp.up2 ? (tap & 1) + (bz & 1) : tap - ky * 3;
                    tma_load_4d(sA + s * kABytes, &tmA, &full_bar[s], cb * kBlockK, mp.cstride * x0 + kx - 1, mp.cstride * y0 + ky - 1, b0);
                } else {
                    tma_load_3d(sA + s * kABytes, &tmA, &full_bar[s], kb * kBlockK, m0, bz);
                }
                if (load_b) tma_load_3d(sB + s * Cfg::kBBytes, &tmW, &full_bar[s], kb * kBlockK, n0, mp.w_batched ? bz : 0);
            }
            }   // tile loop
        }
    } else if (warp == 1) {
        if (lane == 0 && rank == 0) {
            // ---------------------------------------------------------------- MMA issuer (pair: the leader drives both SMs)
            constexpr uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, BN);
            int g = 0;
            SDOD_TILE_LOOP {
            const int iter = wk.iter;
            const int ab = PERSIST ? (iter % naccs) : 0;             // TMEM accumulator buffer of this tile
            const uint32_t tmem_acc = tmem_base + ab * BN;
            if (PERSIST && iter >= naccs) {                           // the epilogue of tile iter - naccs has drained this buffer
                mbar_wait(&tmem_empty_bar[ab], ((iter / naccs) - 1) & 1);
                tc_fence_after();
            }
            for (int kb = wk.seg_lo; kb < wk.seg_hi; ++kb, ++g) {
                const int it = kb - wk.seg_lo;
                const int s = g % STAGES;
                const uint32_t ph = (g / STAGES) & 1;
                mbar_wait(&full_bar[s], ph);
                if (g == 0) tstamp(mp, 3);                                 // first operand stage landed
                tc_fence_after();
                const uint32_t a_addr = smem_u32(sA + s * kABytes);
                const uint32_t b_addr = smem_u32(sB + s * Cfg::kBBytes);
#pragma unroll
                for (int k = 0; k < kBlockK / 16; ++k) {
                    if (PAIR) tc_mma_bf16_pair(tmem_acc, make_kmajor_sw128_desc(a_addr + k * 32), make_kmajor_sw128_desc(b_addr + k * 32),
                                               idesc, (it | k) != 0);
                    else tc_mma_bf16(tmem_acc, make_kmajor_sw128_desc(a_addr + k * 32), make_kmajor_sw128_desc(b_addr + k * 32),
                                     idesc, (it | k) != 0);
                }
                if (PAIR) tc_commit_pair(&empty_bar[s]);   // frees the slot in both CTAs when these MMAs retire
                else tc_commit(&empty_bar[s]);
            }
            if (PAIR) tc_commit_pair(&tmem_full_bar[ab]);       // accumulator complete (each CTA holds its own 128 rows)
            else tc_commit(&tmem_full_bar[ab]);
            if (iter == 0) tstamp(mp, 4);                                  // last MMA issued
            }   // tile loop
        }
    } else {
        // -------------------------------------------------------------------- epilogue
        int res_uses = 0;                              // residual-barrier phases consumed so far (stream-K head segments load no residual)
        // Two epilogue groups (persistent variant, mp.epi_groups == 2): the 16 epilogue warps split into two groups of 8 that take ALTERNATE tiles —
        // group g always drains TMEM accumulator g and stages in half g of the epilogue region — so one group's TMEM-load / math / store latencies
        // overlap the other group's instead of all 16 warps walking one tile in lockstep (ncu, GEGLU projection at batch 8: issue slots 48 %,
        // tensor pipe 26 %, L2 24 %: nothing saturated, the per-tile epilogue chain was the critical path).
        const bool two = PERSIST && mp.epi_groups == 2;
        const int half_raw = (warp - 2) >> 2;          // which of the quarter's NPART warps
        const int grp = two ? (half_raw >> 1) : 0;
        const int half = two ? (half_raw & 1) : half_raw;
        const int npart = two ? NPART / 2 : NPART;
        const int gthreads = two ? 16 * EW : 32 * EW;  // threads of this group
        const int gbar = 5 + grp;                      // the group's named barrier (5 | 6); per-quarter barriers: 1..4 (group 0), 7..10 (group 1)
        auto group_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(gbar), "r"(gthreads) : "memory"); };
        SDOD_TILE_LOOP {
        if (two && (wk.iter & 1) != grp) continue;     // the other group's tile
        SDOD_TILE_COORDS
        const int ab = PERSIST ? (iter & 1) : 0;       // bias / staging buffer
        const int acc = PERSIST ? (iter % naccs) : 0;  // TMEM accumulator
        float* s_bias = s_bias_base + ab * BN;         // double-buffered: a fast warp may already stage the next tile's bias
        // Bias tile: staged while the mainloop runs.  A persistent CTA stages tile i+1's bias during tile i's epilogue (the global
        // load is issued here and parked in a register; it lands in the other buffer at the end of the iteration), so no tile starts
        // with an exposed L2 round trip.
        const int te = static_cast<int>(threadIdx.x) - 64 - grp * gthreads;
        float bias_next = 0.f;
        const bool stage_next = PERSIST && !two && !wk.sk && mp.tma_epi && t + static_cast<int>(gridDim.x) < wk.t_end && te < BN;
        // stream-K role of this segment: a head segment (starts inside the tile) only publishes its partial; the segment that starts the tile
        // folds the partials of the CTAs that follow it (their shares begin inside this tile) and runs the epilogue
        const bool sk_publish = PERSIST && wk.sk && wk.seg_lo > 0;
        int sk_parts = 0;
        if (PERSIST && wk.sk && wk.seg_lo == 0 && wk.seg_hi < wk.kblocks) {
            const long long tile_end = static_cast<long long>(t + 1) * wk.kblocks;
            while (static_cast<int>(blockIdx.x) + sk_parts + 1 < static_cast<int>(gridDim.x) &&
                   WorkWalk::share_begin(mp, blockIdx.x + sk_parts + 1) < tile_end) ++sk_parts;
        }
        if (mp.tma_epi) {
            if (!PERSIST || iter == 0 || wk.sk || two)
                for (int i = te; i < BN; i += gthreads) s_bias[i] = (ep.bias && n0 + i < mp.N) ? ep.bias[n0 + i] : 0.f;
            group_sync();                                                // (persistent: also orders the previous tile's staging reads/stores)
            if (stage_next) {
                const int n_next = ((t + static_cast<int>(gridDim.x)) % mp.n_tiles) * BN + te;
                bias_next = (ep.bias && n_next < mp.N) ? __ldg(ep.bias + n_next) : 0.f;
            }
        }
        mbar_wait(&tmem_full_bar[acc], PERSIST ? ((iter / naccs) & 1) : 0);
        tc_fence_after();
        if (threadIdx.x == 64 && iter == 0) tstamp(mp, 5);                 // accumulator complete
        const uint32_t res_parity = PERSIST ? (res_uses & 1) : 0;
        const int q = warp & 3;                       // TMEM lane quarter this warp may access
        // this warp's columns, in 16-column units: [j_lo, j_hi) of the BN-wide tile, [g_lo, g_hi) of a GEGLU half tile
        const int j_lo = 16 * (((BN / 16) * half) / npart), j_hi = 16 * (((BN / 16) * (half + 1)) / npart);
        const int g_lo = 16 * (((BN / 32) * half) / npart), g_hi = 16 * (((BN / 32) * (half + 1)) / npart);
        auto quarter_sync = [&]() { asm volatile("bar.sync %0, %1;" ::"r"(q + 1 + 6 * grp), "r"(32 * npart) : "memory"); };
        const int row = q * 32 + lane;
        const int m = m0 + row;
        const uint32_t taddr = tmem_base + acc * BN + (static_cast<uint32_t>(q * 32) << 16);
        if (PERSIST && sk_parts) {
            // the CTAs whose shares begin inside this tile published their partials at the very start of their walks; the check is a formality
            if (threadIdx.x == 64) {
                for (int c = 1; c <= sk_parts; ++c) {
                    unsigned int seen = 0, spins = 0;
                    do {
                        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(mp.counters + blockIdx.x + c) : "memory");
                        if (!seen && ++spins > (1u << 22)) __trap();
                    } while (!seen);
                }
            }
            group_sync();
        }
        // 16 accumulator columns [j, j+16) of this thread's row: TMEM, plus (stream-K) the published partials of the later K ranges
        auto acc_ld16 = [&](int j, uint32_t (&acc)[16]) {
            tmem_ld16(taddr + j, acc);
            tmem_ld_wait();
            if (PERSIST && sk_parts) {
                const float4* p = reinterpret_cast<const float4*>(mp.ws + static_cast<long long>(blockIdx.x + 1) * (BN * kBlockM) + ((j >> 4) * kBlockM + row) * 16);
                for (int c = 0; c < sk_parts; ++c, p += BN * kBlockM / 4) {
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4) {
                        const float4 v = __ldcg(p + q4);
                        acc[4 * q4] = __float_as_uint(__uint_as_float(acc[4 * q4]) + v.x);
                        acc[4 * q4 + 1] = __float_as_uint(__uint_as_float(acc[4 * q4 + 1]) + v.y);
                        acc[4 * q4 + 2] = __float_as_uint(__uint_as_float(acc[4 * q4 + 2]) + v.z);
                        acc[4 * q4 + 3] = __float_as_uint(__uint_as_float(acc[4 * q4 + 3]) + v.w);
                    }
                }
            }
        };
        if (PERSIST && sk_publish) {
            // stream-K head segment: publish the fp32 partial tile ([chunk16][row][16]) in this CTA's slot, then raise its flag
            float* mine = mp.ws + static_cast<long long>(blockIdx.x) * (BN * kBlockM);
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                uint32_t acc[16];
                tmem_ld16(taddr + j, acc);
                tmem_ld_wait();
                float4* dst = reinterpret_cast<float4*>(mine + ((j >> 4) * kBlockM + row) * 16);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    dst[q4] = make_float4(__uint_as_float(acc[4 * q4]), __uint_as_float(acc[4 * q4 + 1]), __uint_as_float(acc[4 * q4 + 2]),
                                          __uint_as_float(acc[4 * q4 + 3]));
            }
            __threadfence();
            group_sync();
            if (threadIdx.x == 64) {
                const unsigned int one = 1u;
                asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(mp.counters + blockIdx.x), "r"(one) : "memory");
            }
        } else if (mp.split > 1) {
            // split-K: publish this CTA's fp32 partial tile ([chunk16][row][16], coalesced); splitk_reduce_kernel folds
            // the partials in fixed order (deterministic) and applies the epilogue.
            const long long tile_id = static_cast<long long>(m_tile) * (PAIR ? gridDim.y : gridDim.x) + n_tile;
            float* mine = mp.ws + (tile_id * mp.split + zs) * (BN * kBlockM);
            if (!PAIR && !PERSIST && mp.split_cluster == 2) {
                // In-cluster reduction over distributed shared memory: the `split` CTAs of this tile are one (1,1,split) cluster.  Warp task u
                // (16 rows x 16 columns; u = chunk * 8 + row / 16) is folded by CTA u % split, so every CTA PUSHES each of its task partials into
                // the owner's idle operand ring with st.async (bytes counted on the owner's mbarrier): no scratch round trip through L2, no
                // release/acquire barrier (MEMBAR.ALL.GPU behind 80 KB of stores), no second launch.  The one cluster barrier below carries no
                // data: it says that every CTA's MMAs have retired, i.e. every ring may be overwritten.
                asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory");
                asm volatile("barrier.cluster.wait;" ::: "memory");
                constexpr int kTasks = (BN / 16) * 8;
                const int T = (kTasks + mp.split - 1) / mp.split;
                if (threadIdx.x == 64) mbar_arrive_expect_tx(&ln_bar[0], static_cast<uint32_t>((kTasks - zs + mp.split - 1) / mp.split) * mp.split * 1024u);
                float* rx = reinterpret_cast<float*>(sA);                        // [split][T][16 rows][16 columns]
#pragma unroll 1
                for (int j = j_lo; j < j_hi; j += 16) {
                    uint32_t acc[16];
                    tmem_ld16(taddr + j, acc);
                    tmem_ld_wait();
                    const int u = (j >> 4) * 8 + (row >> 4);
                    const int t = u / mp.split, owner = u - t * mp.split;
                    float* dst = rx + ((zs * T + t) * 16 + (row & 15)) * 16;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        st_async_f32x4(dst + 4 * q4, &ln_bar[0], owner, __uint_as_float(acc[4 * q4]), __uint_as_float(acc[4 * q4 + 1]),
                                       __uint_as_float(acc[4 * q4 + 2]), __uint_as_float(acc[4 * q4 + 3]));
                }
            } else
#pragma unroll 1
            for (int j = j_lo; j < j_hi && m0 < mp.M; j += 16) {    // (m0 >= M: padding CTA of an odd pair grid)
                uint32_t acc[16];
                tmem_ld16(taddr + j, acc);
                tmem_ld_wait();
                float4* dst = reinterpret_cast<float4*>(mine + ((j >> 4) * kBlockM + row) * 16);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    dst[q4] = make_float4(__uint_as_float(acc[4 * q4]), __uint_as_float(acc[4 * q4 + 1]), __uint_as_float(acc[4 * q4 + 2]),
                                          __uint_as_float(acc[4 * q4 + 3]));
            }
        } else if (mp.tma_epi && ep.act == SDOD_ACT_GEGLU) {
            // GEGLU through the TMA epilogue: value half = tile columns [0,BN/2), gate half = [BN/2,BN); the output tile is
            // BN/2 bf16 columns wide (64-B swizzled 32x32 boxes).
            constexpr int HALF = BN / 2;
            constexpr int NBG = HALF / 32 > 0 ? HALF / 32 : 1;
            // (persistent: two staging buffers, so this tile's stores may still be reading while the next tile is staged)
            const bool dbuf = PERSIST && 8 * NBG * 4096 <= Cfg::kEpiBytes;
            uint8_t* qbase = stage + (dbuf ? (iter & 1) * (Cfg::kEpiBytes / 2) : 0) + q * (NBG * 4096);
#pragma unroll 1
            for (int j = g_lo; j < g_hi; j += 16) {
                uint32_t a[16], g[16];
                acc_ld16(j, a);
                acc_ld16(HALF + j, g);
                float v[16];
                const uint64_t alpha2 = pack_f32x2(ep.alpha, ep.alpha);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 ba = *reinterpret_cast<const float4*>(s_bias + j + 4 * q4);
                    const float4 bg = *reinterpret_cast<const float4*>(s_bias + HALF + j + 4 * q4);
                    // value * gelu(gate), two columns per packed-fp32 instruction
                    const uint64_t a01 = fma_f32x2(pack_f32x2(__uint_as_float(a[4 * q4]), __uint_as_float(a[4 * q4 + 1])), alpha2, pack_f32x2(ba.x, ba.y));
                    const uint64_t a23 = fma_f32x2(pack_f32x2(__uint_as_float(a[4 * q4 + 2]), __uint_as_float(a[4 * q4 + 3])), alpha2, pack_f32x2(ba.z, ba.w));
                    const uint64_t g01 = fma_f32x2(pack_f32x2(__uint_as_float(g[4 * q4]), __uint_as_float(g[4 * q4 + 1])), alpha2, pack_f32x2(bg.x, bg.y));
                    const uint64_t g23 = fma_f32x2(pack_f32x2(__uint_as_float(g[4 * q4 + 2]), __uint_as_float(g[4 * q4 + 3])), alpha2, pack_f32x2(bg.z, bg.w));
                    unpack_f32x2(mul_f32x2(a01, mp.geglu_tanh ? gelu_tanh_f32x2(g01) : gelu_f32x2(g01)), v[4 * q4], v[4 * q4 + 1]);
                    unpack_f32x2(mul_f32x2(a23, mp.geglu_tanh ? gelu_tanh_f32x2(g23) : gelu_f32x2(g23)), v[4 * q4 + 2], v[4 * q4 + 3]);
                }
                uint8_t* rowp = qbase + (j >> 5) * 4096 + lane * 64;
                const int u0 = (j & 31) >> 3;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    uint4 w;
                    w.x = pack_bf16x2(v[8 * k], v[8 * k + 1]); w.y = pack_bf16x2(v[8 * k + 2], v[8 * k + 3]);
                    w.z = pack_bf16x2(v[8 * k + 4], v[8 * k + 5]); w.w = pack_bf16x2(v[8 * k + 6], v[8 * k + 7]);
                    *reinterpret_cast<uint4*>(rowp + (((u0 + k) ^ ((lane >> 1) & 3)) << 4)) = w;
                }
            }
            fence_proxy_async_smem();
            quarter_sync();
            if (half == 0 && lane == 0 && m0 + q * 32 < mp.M) {
                for (int bx = 0; bx < NBG; ++bx)
                    if (n_tile * HALF + bx * 32 < mp.N / 2) tma_store_3d(&tmC, qbase + bx * 4096, n_tile * HALF + bx * 32, m0 + q * 32, bz);
                bulk_commit();
                if (dbuf && !two) bulk_wait_read_but_last();   // the buffer staged next was stored a whole tile ago
                else bulk_wait_read_all();                     // (two groups: this group re-stages the same half on its next tile)
            }
        } else if (ep.act == SDOD_ACT_GEGLU && !(ep.out_mode == SDOD_OUT_BF16 && mp.N % 8 == 0 && ep.ldc % 4 == 0)) {
            constexpr int HALF = BN / 2;
#pragma unroll 1
            for (int j = g_lo; j < g_hi; j += 16) {
                uint32_t a[16], g[16];
                acc_ld16(j, a);
                acc_ld16(HALF + j, g);
                epilogue_geglu16<BN>(ep, mp, bz, m, n_tile, j, a, g);
            }
        } else if (BN == 160 && mp.tma_epi == 3) {
            // Attention operand layouts through TMA stores.  SD head dims (40 / 80 / 160) are multiples of 40, so a 160-wide
            // tile is four 40-column boxes, each inside one head, and the whole tile is Q, K or V (heads*head_dim % 160 == 0).
            // Q/K box: [32 tokens][40 d] (80-B rows: conflict-free 16-B stores) -> HEADS [B*heads, tokens, dpad] at (d0, tok, bh).
            // V box  : [40 d][32 tokens] (thread = token writes a column)       -> HEADS_T [B*heads, vt_rows, tok_pad] at (tok, d0, bh).
            constexpr int BOXB = 40 * 32 * 2;
            const bool dbuf = PERSIST && 8 * 4 * BOXB <= Cfg::kEpiBytes;
            uint8_t* qbase = stage + (dbuf ? (iter & 1) * (Cfg::kEpiBytes / 2) : 0) + q * (4 * BOXB);
            const int Cw = ep.heads * ep.head_dim;
            const int which = ep.out_mode == SDOD_OUT_QKV ? n0 / Cw : (ep.out_mode == SDOD_OUT_HEADS_T ? 2 : 0);
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                uint32_t acc[16];
                acc_ld16(j, acc);
                float v[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + j + 4 * q4);
                    v[4 * q4] = fmaf(__uint_as_float(acc[4 * q4]), ep.alpha, b4.x);
                    v[4 * q4 + 1] = fmaf(__uint_as_float(acc[4 * q4 + 1]), ep.alpha, b4.y);
                    v[4 * q4 + 2] = fmaf(__uint_as_float(acc[4 * q4 + 2]), ep.alpha, b4.z);
                    v[4 * q4 + 3] = fmaf(__uint_as_float(acc[4 * q4 + 3]), ep.alpha, b4.w);
                }
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const int g8 = (j >> 3) + k;                 // 8-column group of the tile; 5 groups per box
                    const int bx = g8 / 5, unit = g8 - bx * 5;
                    uint8_t* box = qbase + bx * BOXB;
                    if (which < 2) {
                        uint4 w;
                        w.x = pack_bf16x2(v[8 * k], v[8 * k + 1]); w.y = pack_bf16x2(v[8 * k + 2], v[8 * k + 3]);
                        w.z = pack_bf16x2(v[8 * k + 4], v[8 * k + 5]); w.w = pack_bf16x2(v[8 * k + 6], v[8 * k + 7]);
                        *reinterpret_cast<uint4*>(box + lane * 80 + unit * 16) = w;
                    } else {
                        bf16* col = reinterpret_cast<bf16*>(box + unit * 8 * 64) + lane;
#pragma unroll
                        for (int i = 0; i < 8; ++i) col[i * 32] = __float2bfloat16(v[8 * k + i]);
                    }
                }
            }
            fence_proxy_async_smem();
            quarter_sync();
            const int mq = m0 + q * 32;
            if (half == 0 && lane == 0 && mq < mp.M) {
                const int b = mq / ep.tokens, tok = mq - b * ep.tokens;
                for (int bx = 0; bx < 4; ++bx) {
                    const int n = n0 + bx * 40;
                    if (n >= mp.N) break;
                    const int nn = ep.out_mode == SDOD_OUT_QKV ? n - which * Cw : n;
                    const int h = nn / ep.head_dim, d0 = nn - h * ep.head_dim;
                    const int bh = b * ep.heads + h;
                    if (which == 0) tma_store_3d(&tmC, qbase + bx * BOXB, d0, tok, bh);
                    else if (which == 1) tma_store_3d(&tmR, qbase + bx * BOXB, d0, tok, bh);
                    else tma_store_3d(&tmC2, qbase + bx * BOXB, tok, d0, bh);
                }
                bulk_commit();
                if (dbuf && !two) bulk_wait_read_but_last();   // the buffer staged next was stored a whole tile ago
                else bulk_wait_read_all();                     // (two groups: this group re-stages the same half on its next tile)
            }
        } else if (mp.ln_fuse) {
            // fp32 TMA epilogue + LayerNorm of the finished rows (SpatialTransformer norm1/2/3 folded into the GEMM that produces
            // their input; removes a kernel and a read of the fp32 stream per norm).  The N tiles of this row block are the CTAs of one
            // thread-block cluster (n_tiles,1,1).  Each thread owns one output row x half of the tile's columns: the final values
            // (alpha, bias, row bias, residual) are staged in the fp32 boxes as in the plain TMA epilogue; row sums, then sums of
            // squared deviations (two-pass: no cancellation), are combined across the two column halves through shared memory and
            // across the cluster through distributed shared memory; the normalised bf16 rows go out through a second set of boxes.
            constexpr int NB = BN / 32;
            uint8_t* qbase = stage + q * (NB * 4096);                       // fp32 boxes of this lane quarter
            uint8_t* hbase = stage + 4 * NB * 4096 + q * (NB * 2048);       // bf16 boxes (LayerNorm output)
            float* s_gamma = reinterpret_cast<float*>(stage + 4 * NB * 6144);
            float* s_beta = s_gamma + BN;
            float* s_halfsum = s_beta + BN;                                 // [NPART][128] partial row moments of the column parts
            // [2][8][128]: row sums / squared deviations of every CTA of the cluster, PUSHED here by their owners (st.async, bytes counted on
            // ln_bar): no release/acquire cluster barrier (a MEMBAR.ALL.GPU per warp each) and no exit barrier
            float* ln_peer = s_halfsum + NPART * kBlockM;
            if (mp.tma_epi == 2 && half == 0 && lane == 0) {
                mbar_arrive_expect_tx(&res_bar[q], NB * 4096);
                for (int bx = 0; bx < NB; ++bx) tma_load_3d(qbase + bx * 4096, &tmR, &res_bar[q], n0 + bx * 32, m0 + q * 32, bz);
            }
            for (int i = te; i < BN; i += 32 * EW) {
                s_gamma[i] = ep.ln_weight ? __ldg(ep.ln_weight + n0 + i) : 1.f;
                s_beta[i] = ep.ln_bias ? __ldg(ep.ln_bias + n0 + i) : 0.f;
            }
            const float* rb = (ep.row_bias && m < mp.M) ? ep.row_bias + static_cast<long long>(m / ep.rows_per_group) * (ep.ld_row_bias ? ep.ld_row_bias : mp.N) : nullptr;
            bool res_waited = (mp.tma_epi != 2);
            float part = 0.f;
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                uint32_t acc[16];
                acc_ld16(j, acc);
                float v[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + j + 4 * q4);
                    v[4 * q4] = fmaf(__uint_as_float(acc[4 * q4]), ep.alpha, b4.x);
                    v[4 * q4 + 1] = fmaf(__uint_as_float(acc[4 * q4 + 1]), ep.alpha, b4.y);
                    v[4 * q4 + 2] = fmaf(__uint_as_float(acc[4 * q4 + 2]), ep.alpha, b4.z);
                    v[4 * q4 + 3] = fmaf(__uint_as_float(acc[4 * q4 + 3]), ep.alpha, b4.w);
                }
                if (rb) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] += __ldg(rb + n0 + j + i);
                }
                if (!res_waited) { mbar_wait(&res_bar[q], res_parity); res_waited = true; }
                uint8_t* rowp = qbase + (j >> 5) * 4096 + lane * 128;
                const int u0 = (j & 31) >> 2;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    float4* cell = reinterpret_cast<float4*>(rowp + (((u0 + k) ^ (lane & 7)) << 4));
                    float4 o = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                    if (mp.tma_epi == 2) { const float4 r4 = *cell; o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w; }
                    *cell = o;
                    part += (o.x + o.y) + (o.z + o.w);
                }
            }
            auto epi_sync = [&]() { asm volatile("bar.sync 5, %0;" ::"n"(32 * EW) : "memory"); };
            const uint32_t n_cta = cluster_nctarank(), my_cta = cluster_ctarank();
            const float inv_n = 1.0f / static_cast<float>(mp.N);
            asm volatile("barrier.cluster.wait;" ::: "memory");      // every CTA of the cluster has initialised its ln_bar
            if (threadIdx.x == 64) {
                mbar_arrive_expect_tx(&ln_bar[0], n_cta * kBlockM * 4u);
                mbar_arrive_expect_tx(&ln_bar[1], n_cta * kBlockM * 4u);
            }
            // ---- mean
            s_halfsum[half * kBlockM + row] = part;
            epi_sync();
            if (half == 0) {
                float t = 0.f;
#pragma unroll
                for (int h = 0; h < NPART; ++h) t += s_halfsum[h * kBlockM + row];
                for (uint32_t r = 0; r < n_cta; ++r) st_async_f32(ln_peer + my_cta * kBlockM + row, &ln_bar[0], r, t);
            }
            mbar_wait(&ln_bar[0], 0);
            float mean = 0.f;
            for (uint32_t r = 0; r < n_cta; ++r) mean += ln_peer[r * kBlockM + row];      // fixed order: identical in every CTA
            mean *= inv_n;
            // ---- variance (second pass over the staged row)
            part = 0.f;
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                const uint8_t* rowp = qbase + (j >> 5) * 4096 + lane * 128;
                const int u0 = (j & 31) >> 2;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const float4 o = *reinterpret_cast<const float4*>(rowp + (((u0 + k) ^ (lane & 7)) << 4));
                    const float d0 = o.x - mean, d1 = o.y - mean, d2 = o.z - mean, d3 = o.w - mean;
                    part += (d0 * d0 + d1 * d1) + (d2 * d2 + d3 * d3);
                }
            }
            epi_sync();                                   // every reader of s_halfsum (mean pass) is done
            s_halfsum[half * kBlockM + row] = part;
            epi_sync();
            if (half == 0) {
                float t = 0.f;
#pragma unroll
                for (int h = 0; h < NPART; ++h) t += s_halfsum[h * kBlockM + row];
                for (uint32_t r = 0; r < n_cta; ++r) st_async_f32(ln_peer + (8 + my_cta) * kBlockM + row, &ln_bar[1], r, t);
            }
            mbar_wait(&ln_bar[1], 0);
            float var = 0.f;
            for (uint32_t r = 0; r < n_cta; ++r) var += ln_peer[(8 + r) * kBlockM + row];
            const float rstd = rsqrtf(var * inv_n + ep.ln_eps);
            // ---- normalised bf16 rows into the second set of boxes (64-B swizzled 32x32 bf16)
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                const uint8_t* rowp = qbase + (j >> 5) * 4096 + lane * 128;
                uint8_t* hrow = hbase + (j >> 5) * 2048 + lane * 64;
                const int u0 = (j & 31) >> 2, h0 = (j & 31) >> 3;
#pragma unroll
                for (int k = 0; k < 2; ++k) {
                    const float4 a = *reinterpret_cast<const float4*>(rowp + (((u0 + 2 * k) ^ (lane & 7)) << 4));
                    const float4 b = *reinterpret_cast<const float4*>(rowp + (((u0 + 2 * k + 1) ^ (lane & 7)) << 4));
                    const float4 g0 = *reinterpret_cast<const float4*>(s_gamma + j + 8 * k), g1 = *reinterpret_cast<const float4*>(s_gamma + j + 8 * k + 4);
                    const float4 e0 = *reinterpret_cast<const float4*>(s_beta + j + 8 * k), e1 = *reinterpret_cast<const float4*>(s_beta + j + 8 * k + 4);
                    uint4 w;
                    w.x = pack_bf16x2(fmaf((a.x - mean) * rstd, g0.x, e0.x), fmaf((a.y - mean) * rstd, g0.y, e0.y));
                    w.y = pack_bf16x2(fmaf((a.z - mean) * rstd, g0.z, e0.z), fmaf((a.w - mean) * rstd, g0.w, e0.w));
                    w.z = pack_bf16x2(fmaf((b.x - mean) * rstd, g1.x, e1.x), fmaf((b.y - mean) * rstd, g1.y, e1.y));
                    w.w = pack_bf16x2(fmaf((b.z - mean) * rstd, g1.z, e1.z), fmaf((b.w - mean) * rstd, g1.w, e1.w));
                    *reinterpret_cast<uint4*>(hrow + (((h0 + k) ^ ((lane >> 1) & 3)) << 4)) = w;
                }
            }
            fence_proxy_async_smem();
            quarter_sync();
            if (half == 0 && lane == 0 && m0 + q * 32 < mp.M) {
                for (int bx = 0; bx < NB; ++bx) {
                    tma_store_3d(&tmC, qbase + bx * 4096, n0 + bx * 32, m0 + q * 32, bz);
                    tma_store_3d(&tmC2, hbase + bx * 2048, n0 + bx * 32, m0 + q * 32, bz);
                }
                bulk_commit();
                bulk_wait_read_all();
            }
        } else if (mp.tma_epi) {
            // TMA epilogue.  The idle TMA ring becomes a staging area of 32x32-element boxes (128-B or 64-B swizzled rows).
            // If there is a residual of the output's dtype it is TMA-loaded straight into those boxes while the accumulator
            // is read; each thread (one accumulator row) applies alpha/bias/timestep-bias/activation in registers, adds the
            // staged residual in place, and one elected thread per lane quarter hands the boxes to TMA stores: no per-row
            // global load/store instructions, tails clipped by the tensor map.
            constexpr int NB = BN / 32;
            const bool f32 = (mp.c_bytes == 4);
            const uint32_t box_bytes = f32 ? 4096u : 2048u;
            const bool dbuf = PERSIST && !f32;           // bf16 boxes are 2 KiB: two staging buffers fit the region sized for fp32
            const uint32_t bstr = dbuf ? 2048u : 4096u;  // box stride
            uint8_t* qbase = stage + (dbuf ? (iter & 1) * (Cfg::kEpiBytes / 2) : 0) + q * (NB * bstr);
            if (mp.tma_epi == 2 && half == 0 && lane == 0) {
                mbar_arrive_expect_tx(&res_bar[q], NB * box_bytes);
                for (int bx = 0; bx < NB; ++bx) tma_load_3d(qbase + bx * bstr, &tmR, &res_bar[q], n0 + bx * 32, m0 + q * 32, bz);
            }
            const float* rb = (ep.row_bias && m < mp.M) ? ep.row_bias + static_cast<long long>(m / ep.rows_per_group) * (ep.ld_row_bias ? ep.ld_row_bias : mp.N) : nullptr;
            bool res_waited = (mp.tma_epi != 2);
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                uint32_t acc[16];
                acc_ld16(j, acc);
                float v[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const float4 b4 = *reinterpret_cast<const float4*>(s_bias + j + 4 * q4);
                    v[4 * q4] = fmaf(__uint_as_float(acc[4 * q4]), ep.alpha, b4.x);
                    v[4 * q4 + 1] = fmaf(__uint_as_float(acc[4 * q4 + 1]), ep.alpha, b4.y);
                    v[4 * q4 + 2] = fmaf(__uint_as_float(acc[4 * q4 + 2]), ep.alpha, b4.z);
                    v[4 * q4 + 3] = fmaf(__uint_as_float(acc[4 * q4 + 3]), ep.alpha, b4.w);
                }
                if (rb) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (n0 + j + i < mp.N) v[i] += __ldg(rb + n0 + j + i);
                }
                if (ep.act == SDOD_ACT_SILU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = silu_f(v[i]);
                } else if (ep.act == SDOD_ACT_GELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = gelu_f(v[i]);
                } else if (ep.act == SDOD_ACT_QUICK_GELU) {
#pragma unroll
                    for (int i = 0; i < 16; ++i) v[i] = quick_gelu_f(v[i]);
                }
                if (!res_waited) { mbar_wait(&res_bar[q], res_parity); res_waited = true; if (threadIdx.x == 64 && iter == 0) tstamp(mp, 9); }
                uint8_t* box = qbase + (j >> 5) * bstr;
                if (f32) {
                    uint8_t* rowp = box + lane * 128;
                    const int u0 = (j & 31) >> 2;
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        float4* cell = reinterpret_cast<float4*>(rowp + (((u0 + k) ^ (lane & 7)) << 4));
                        float4 o = make_float4(v[4 * k], v[4 * k + 1], v[4 * k + 2], v[4 * k + 3]);
                        if (mp.tma_epi == 2) { const float4 r4 = *cell; o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w; }
                        *cell = o;
                    }
                } else {
                    uint8_t* rowp = box + lane * 64;
                    const int u0 = (j & 31) >> 3;
#pragma unroll
                    for (int k = 0; k < 2; ++k) {
                        uint4* cell = reinterpret_cast<uint4*>(rowp + (((u0 + k) ^ ((lane >> 1) & 3)) << 4));
                        float o[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) o[i] = v[8 * k + i];
                        if (mp.tma_epi == 2) {
                            const uint4 r4 = *cell;
                            float2 t;
                            t = unpack_bf16x2(r4.x); o[0] += t.x; o[1] += t.y;
                            t = unpack_bf16x2(r4.y); o[2] += t.x; o[3] += t.y;
                            t = unpack_bf16x2(r4.z); o[4] += t.x; o[5] += t.y;
                            t = unpack_bf16x2(r4.w); o[6] += t.x; o[7] += t.y;
                        }
                        uint4 w;
                        w.x = pack_bf16x2(o[0], o[1]); w.y = pack_bf16x2(o[2], o[3]); w.z = pack_bf16x2(o[4], o[5]); w.w = pack_bf16x2(o[6], o[7]);
                        *cell = w;
                    }
                }
            }
            if (threadIdx.x == 64 && iter == 0) tstamp(mp, 10);       // tile staged
            fence_proxy_async_smem();                  // staged tile (generic-proxy writes) -> visible to the TMA engine
            quarter_sync();
            if (half == 0 && lane == 0 && m0 + q * 32 < mp.M) {
                if (mp.up2) {
                    // sub-pixel form: the quarter's 32 source pixels (x fastest, then y, then image) go to output pixels (2y+py, 2x+px): one
                    // 5-D box {32 ch, px, min(bw,32) x's, py, 32/min(bw,32) merged (image, y) rows} per 32-column box
                    const int rq = q * 32, hw = mp.H * mp.W;
                    int b0 = 0, y0 = 0, x0 = 0;
                    if (mp.bb > 1) { b0 = m_tile * mp.bb; } else { b0 = m0 / hw; const int rem = m0 - b0 * hw; y0 = rem / mp.W; x0 = rem - y0 * mp.W; }
                    const int xq = x0 + (mp.bw >= 32 ? rq % mp.bw : 0);
                    const int biq = (b0 + rq / (mp.bw * mp.bh)) * mp.H + y0 + (rq / mp.bw) % mp.bh;
                    for (int bx = 0; bx < NB; ++bx)
                        if (n0 + bx * 32 < mp.N) tma_store_5d(&tmC, qbase + bx * bstr, n0 + bx * 32, bz & 1, xq, bz >> 1, biq);
                } else {
                    for (int bx = 0; bx < NB; ++bx)
                        if (n0 + bx * 32 < mp.N) {
                            // tma_epi 4: the fp32 residual IS the output tensor (x += f(x), in place): the add happens in the TMA reduction at L2,
                            // the SM neither loads the residual nor re-reads it from shared memory
                            if (mp.tma_epi == 4) tma_reduce_add_3d(&tmC, qbase + bx * bstr, n0 + bx * 32, m0 + q * 32, bz);
                            else tma_store_3d(&tmC, qbase + bx * bstr, n0 + bx * 32, m0 + q * 32, bz);
                        }
                }
                bulk_commit();
                if (dbuf && !two) bulk_wait_read_but_last();
                else bulk_wait_read_all();             // smem may be released once the stores have read it
            }
        } else if (ep.out_mode == SDOD_OUT_BF16 || ep.out_mode == SDOD_OUT_F32) {
            // Coalesced epilogue.  Phase 1: each thread owns one accumulator row (tcgen05.ld 32x32b) and copies it with 16-B
            // st.shared into a padded staging tile (the TMA ring is idle: tmem_full fired after every MMA retired).
            // Phase 2: the warp walks its 32 rows, lanes run along the columns: bias is lane-constant (loaded once), row-bias
            // and residual loads and the output stores are full 128-B lines; 8 rows of loads are in flight before any store.
            constexpr int LDS = BN + 4;                       // (BN+4) % 32 == 4 words: conflict-free 16-B row-strided stores
            float* stg = reinterpret_cast<float*>(stage) + q * (32 * LDS);
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                uint32_t acc[16];
                acc_ld16(j, acc);
                float4* dst = reinterpret_cast<float4*>(stg + lane * LDS + j);
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4)
                    dst[q4] = make_float4(__uint_as_float(acc[4 * q4]), __uint_as_float(acc[4 * q4 + 1]), __uint_as_float(acc[4 * q4 + 2]),
                                          __uint_as_float(acc[4 * q4 + 3]));
            }
            // both warps of this lane quarter have staged their columns (named barrier per quarter, 64 threads)
            quarter_sync();
            const long long zc = static_cast<long long>(bz) * ep.strideC, zr = static_cast<long long>(bz) * ep.strideR;
            const bool vec_ok = (mp.N % 4 == 0) && (ep.ldc % 4 == 0) && (!ep.residual || ep.ldr % 4 == 0) &&
                                (!ep.row_bias || (ep.ld_row_bias ? ep.ld_row_bias : mp.N) % 4 == 0);
            const int rows_here = min(32, mp.M - (m0 + q * 32));     // valid rows of this quarter; this warp takes [r_lo, r_hi)
            const int r_lo = half * (32 / NPART), r_hi = min(rows_here, r_lo + 32 / NPART);
            const long long ldrb = ep.ld_row_bias ? ep.ld_row_bias : mp.N;
            if (ep.act == SDOD_ACT_GEGLU) {
                // value half = tile columns [0,BN/2), gate half = [BN/2,BN); lanes run along the BN/2 output columns
                constexpr int HALF = BN / 2;
                constexpr int NG4 = (HALF / 4 + 31) / 32;
                const int n_out_total = mp.N / 2;
                float4 bv[NG4], bg[NG4];
                bool ok[NG4];
#pragma unroll
                for (int ci = 0; ci < NG4; ++ci) {
                    const int c = (lane + ci * 32) * 4;
                    ok[ci] = (c < HALF) && (n_tile * HALF + c < n_out_total);
                    bv[ci] = (ep.bias && ok[ci]) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                    bg[ci] = (ep.bias && ok[ci]) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n0 + HALF + c)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                for (int r = r_lo; r < r_hi; ++r) {
                    const long long mr = m0 + q * 32 + r;
#pragma unroll
                    for (int ci = 0; ci < NG4; ++ci) {
                        if (!ok[ci]) continue;
                        const int c = (lane + ci * 32) * 4;
                        const float4 a4 = *reinterpret_cast<const float4*>(stg + r * LDS + c);
                        const float4 g4 = *reinterpret_cast<const float4*>(stg + r * LDS + HALF + c);
                        uint2 w;
                        w.x = pack_bf16x2(fmaf(a4.x, ep.alpha, bv[ci].x) * gelu_f(fmaf(g4.x, ep.alpha, bg[ci].x)),
                                          fmaf(a4.y, ep.alpha, bv[ci].y) * gelu_f(fmaf(g4.y, ep.alpha, bg[ci].y)));
                        w.y = pack_bf16x2(fmaf(a4.z, ep.alpha, bv[ci].z) * gelu_f(fmaf(g4.z, ep.alpha, bg[ci].z)),
                                          fmaf(a4.w, ep.alpha, bv[ci].w) * gelu_f(fmaf(g4.w, ep.alpha, bg[ci].w)));
                        *reinterpret_cast<uint2*>(reinterpret_cast<bf16*>(ep.C) + zc + mr * ep.ldc + n_tile * HALF + c) = w;
                    }
                }
            } else if (vec_ok) {
                // Compact code matters here: the fully unrolled, activation-generic form of this loop was ~5k SASS instructions
                // executed once per tile, i.e. pure instruction-cache misses (8 us per tile).  One non-unrolled row-group loop
                // per activation, selected outside the loop.
                constexpr int NC4 = (BN / 4 + 31) / 32;
                constexpr int RG = 4;
                float4 bias4[NC4];
                bool col_ok[NC4];
#pragma unroll
                for (int ci = 0; ci < NC4; ++ci) {
                    const int c4 = lane + ci * 32;
                    const int n = n0 + c4 * 4;
                    col_ok[ci] = (c4 < BN / 4) && (n < mp.N);
                    bias4[ci] = (ep.bias && col_ok[ci]) ? __ldg(reinterpret_cast<const float4*>(ep.bias + n)) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
                const bool has_res = ep.residual != nullptr, res32 = ep.residual_f32 != 0, out32 = (ep.out_mode == SDOD_OUT_F32);
                const float* res_f = reinterpret_cast<const float*>(ep.residual) + zr;
                const bf16* res_h = reinterpret_cast<const bf16*>(ep.residual) + zr;
                float* out_f = reinterpret_cast<float*>(ep.C) + zc;
                bf16* out_h = reinterpret_cast<bf16*>(ep.C) + zc;
                auto run_rows = [&](auto act_tag) {
                    constexpr int ACT = decltype(act_tag)::value;
#pragma unroll 1
                    for (int r0 = r_lo; r0 < r_hi; r0 += RG) {
                        float4 add[RG][NC4];
#pragma unroll
                        for (int rr = 0; rr < RG; ++rr) {
                            const long long mr = m0 + q * 32 + r0 + rr;
#pragma unroll
                            for (int ci = 0; ci < NC4; ++ci) {
                                const int n = n0 + (lane + ci * 32) * 4;
                                float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                                if (has_res && (r0 + rr < r_hi) && col_ok[ci]) {
                                    if (res32) {
                                        t = __ldg(reinterpret_cast<const float4*>(res_f + mr * ep.ldr + n));
                                    } else {
                                        const uint2 u2 = __ldg(reinterpret_cast<const uint2*>(res_h + mr * ep.ldr + n));
                                        const float2 lo = unpack_bf16x2(u2.x), hi = unpack_bf16x2(u2.y);
                                        t = make_float4(lo.x, lo.y, hi.x, hi.y);
                                    }
                                }
                                add[rr][ci] = t;
                            }
                        }
#pragma unroll
                        for (int rr = 0; rr < RG; ++rr) {
                            const long long mr = m0 + q * 32 + r0 + rr;
#pragma unroll
                            for (int ci = 0; ci < NC4; ++ci) {
                                if ((r0 + rr < r_hi) && col_ok[ci]) {
                                    const int c4 = lane + ci * 32;
                                    const int n = n0 + c4 * 4;
                                    const float4 a4 = *reinterpret_cast<const float4*>(stg + (r0 + rr) * LDS + c4 * 4);
                                    float v[4] = {fmaf(a4.x, ep.alpha, bias4[ci].x), fmaf(a4.y, ep.alpha, bias4[ci].y), fmaf(a4.z, ep.alpha, bias4[ci].z),
                                                  fmaf(a4.w, ep.alpha, bias4[ci].w)};
                                    if (ep.row_bias) {
                                        const float4 t = __ldg(reinterpret_cast<const float4*>(ep.row_bias + (mr / ep.rows_per_group) * ldrb + n));
                                        v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
                                    }
                                    if (ACT == SDOD_ACT_SILU) {
#pragma unroll
                                        for (int i = 0; i < 4; ++i) v[i] = silu_f(v[i]);
                                    } else if (ACT == SDOD_ACT_GELU) {
#pragma unroll
                                        for (int i = 0; i < 4; ++i) v[i] = gelu_f(v[i]);
                                    } else if (ACT == SDOD_ACT_QUICK_GELU) {
#pragma unroll
                                        for (int i = 0; i < 4; ++i) v[i] = quick_gelu_f(v[i]);
                                    }
                                    const float4 o = make_float4(v[0] + add[rr][ci].x, v[1] + add[rr][ci].y, v[2] + add[rr][ci].z, v[3] + add[rr][ci].w);
                                    if (out32) {
                                        *reinterpret_cast<float4*>(out_f + mr * ep.ldc + n) = o;
                                    } else {
                                        uint2 w;
                                        w.x = pack_bf16x2(o.x, o.y); w.y = pack_bf16x2(o.z, o.w);
                                        *reinterpret_cast<uint2*>(out_h + mr * ep.ldc + n) = w;
                                    }
                                }
                            }
                        }
                    }
                };
                if (ep.act == SDOD_ACT_SILU) run_rows(std::integral_constant<int, SDOD_ACT_SILU>{});
                else if (ep.act == SDOD_ACT_GELU) run_rows(std::integral_constant<int, SDOD_ACT_GELU>{});
                else if (ep.act == SDOD_ACT_QUICK_GELU) run_rows(std::integral_constant<int, SDOD_ACT_QUICK_GELU>{});
                else run_rows(std::integral_constant<int, SDOD_ACT_NONE>{});
            } else {
                for (int r = r_lo; r < r_hi; ++r) {
                    const long long mr = m0 + q * 32 + r;
                    const float* rbp = ep.row_bias ? ep.row_bias + (mr / ep.rows_per_group) * ldrb : nullptr;
                    for (int c = lane; c < BN; c += 32) {
                        const int n = n0 + c;
                        if (n >= mp.N) break;
                        float x = stg[r * LDS + c] * ep.alpha;
                        if (ep.bias) x += ep.bias[n];
                        if (rbp) x += rbp[n];
                        x = apply_act(x, ep.act);
                        if (ep.residual) x += ep.residual_f32 ? reinterpret_cast<const float*>(ep.residual)[zr + mr * ep.ldr + n]
                                                              : __bfloat162float(reinterpret_cast<const bf16*>(ep.residual)[zr + mr * ep.ldr + n]);
                        if (ep.out_mode == SDOD_OUT_F32) reinterpret_cast<float*>(ep.C)[zc + mr * ep.ldc + n] = x;
                        else reinterpret_cast<bf16*>(ep.C)[zc + mr * ep.ldc + n] = __float2bfloat16(x);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int j = j_lo; j < j_hi; j += 16) {
                uint32_t acc[16];
                acc_ld16(j, acc);
                epilogue_plain16(ep, mp, bz, m, n0 + j, acc);
            }
        }
        if (threadIdx.x == 64 && iter == 0) tstamp(mp, 6);                 // epilogue of the first tile done (stores issued and read)
        if (mp.tma_epi == 2 && !(PERSIST && sk_publish)) ++res_uses;
        if (PERSIST && sk_parts) {                                         // every thread has folded the partials: re-arm their flags
            group_sync();
            if (threadIdx.x == 64)
                for (int c = 1; c <= sk_parts; ++c) mp.counters[blockIdx.x + c] = 0u;
        }
        tc_fence_before();
        if (stage_next) s_bias_base[(ab ^ 1) * BN + te] = bias_next;
        if (PERSIST) mbar_arrive(&tmem_empty_bar[acc]);     // every TMEM read of this buffer has completed (tcgen05.wait::ld above)
        }   // tile loop
        if (PERSIST) bulk_wait_read_all();                 // (storing threads) shared memory stays valid until the last store has read it
    }
    if (!PAIR && !PERSIST && mp.ln_fuse && warp < 2) asm volatile("barrier.cluster.wait;" ::: "memory");   // (completes the split barrier of the setup)
    if (!PAIR && !PERSIST && mp.split_cluster) {
        // Split-K inside one launch: the `split` CTAs of this output tile are one thread-block cluster (1,1,split).  Each has published its
        // fp32 partial tile to the L2-resident scratch above; after the cluster barrier (release/acquire at cluster scope) every CTA folds
        // 1/split of the tile — warp tasks of 16 rows x 16 columns, dealt round-robin over the CTAs and their 8 epilogue warps — in fixed
        // z order (deterministic, independent of arrival order) and runs the shared epilogue on it.
        if (mp.split_cluster == 2) {
            if (warp < 2) {                                                // (the epilogue warps took part right after their accumulator wait)
                asm volatile("barrier.cluster.arrive.relaxed;" ::: "memory");
                asm volatile("barrier.cluster.wait;" ::: "memory");
            } else {
                mbar_wait(&ln_bar[0], 0);                                  // every partial of the tasks this CTA folds has landed in its ring
                constexpr int kTasks = (BN / 16) * 8;
                const int T = (kTasks + mp.split - 1) / mp.split;
                const float* rx = reinterpret_cast<const float*>(sA);
                const int n_tile = static_cast<int>(blockIdx.x), m_tile = static_cast<int>(blockIdx.y);
                const int row_in = lane >> 1, half = lane & 1;
                for (int u = zs + mp.split * (warp - 2); u < kTasks; u += mp.split * 8) {
                    const int chunk = u >> 3, row = (u & 7) * 16 + row_in, t = u / mp.split;
                    float4 s0 = make_float4(0.f, 0.f, 0.f, 0.f), s1 = s0;
                    for (int z = 0; z < mp.split; ++z) {                   // fixed z order: deterministic
                        const float4* src = reinterpret_cast<const float4*>(rx + ((z * T + t) * 16 + row_in) * 16 + half * 8);
                        const float4 a = src[0], b = src[1];
                        s0.x += a.x; s0.y += a.y; s0.z += a.z; s0.w += a.w; s1.x += b.x; s1.y += b.y; s1.z += b.z; s1.w += b.w;
                    }
                    float acc[8] = {s0.x, s0.y, s0.z, s0.w, s1.x, s1.y, s1.z, s1.w};
                    epilogue_plain8(ep, mp, 0, m_tile * kBlockM + row, n_tile * BN + chunk * 16 + half * 8, acc);
                }
            }
        } else {
        asm volatile("barrier.cluster.arrive.release;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire;" ::: "memory");
        if (threadIdx.x == 64) tstamp(mp, 7);                              // split-K: all partials published
        if (warp >= 2) {
            const int n_tile = static_cast<int>(blockIdx.x), m_tile = static_cast<int>(blockIdx.y);
            const long long tile_id = static_cast<long long>(m_tile) * gridDim.x + n_tile;
            const float* base = mp.ws + tile_id * mp.split * (BN * kBlockM);
            const long long zstride4 = BN * kBlockM / 4;
            constexpr int kTasks = (BN / 16) * 8;
            const int row_in = lane >> 1, half = lane & 1;
            for (int u = zs + mp.split * (warp - 2); u < kTasks; u += mp.split * 8) {
                const int chunk = u >> 3, row = (u & 7) * 16 + row_in;
                float acc[8];
                fold_partials8(reinterpret_cast<const float4*>(base + (chunk * kBlockM + row) * 16 + half * 8), mp.split, zstride4, acc);
                epilogue_plain8(ep, mp, 0, m_tile * kBlockM + row, n_tile * BN + chunk * 16 + half * 8, acc);
            }
        }
        }
    }
    __syncthreads();
    if (threadIdx.x == 0) tstamp(mp, 8);                                   // all roles done
    // neither CTA's shared / tensor memory goes away while the peer may still touch it (execution barrier: every cross-CTA access was already
    // ordered by the mbarriers the MMAs commit to)
    if (PAIR) { if (pair_relaxed) { cluster_arrive_relaxed(); cluster_wait(); } else cluster_sync_all(); }
    // (LayerNorm epilogue: no exit barrier — peers only ever WRITE into this CTA, and all of those writes had landed before its ln_bar waits returned)
    if (warp == 2) {
        tc_fence_after();
        if (PAIR) tmem_dealloc_pair<Cfg::kTmemCols>(tmem_base);
        else tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

// ------------------------------------------------------------------------------------------ host
template <int BN, bool DEEP, bool PAIR, bool PERSIST = false>
static int launch_gemm_cfg(cudaStream_t stream, const GemmLaunch& g, dim3 grid) {
    using Cfg = GemmCfg<BN, DEEP, PAIR, PERSIST>;
    static bool configured = false;
    if (!configured) {
        SDOD_TRY(check_cuda(cudaFuncSetAttribute(gemm_tcgen05_kernel<BN, DEEP, PAIR, PERSIST>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes),
                            "cudaFuncSetAttribute(gemm)"));
        configured = true;
    }
    if (!PAIR && !(g.mp.split > 1 && g.mp.split_cluster) && !g.mp.ln_fuse) {
        return check_cuda(launch_pdl(gemm_tcgen05_kernel<BN, DEEP, PAIR, PERSIST>, grid, dim3(PERSIST ? kGemmThreadsPersist : kGemmThreads),
                                     Cfg::kSmemBytes, stream, g.tmA, g.tmW, g.tmC, g.tmR, g.tmC2, g.tmA2, g.mp, g.ep),
                          "launch gemm_tcgen05_kernel");
    }
    if (!PAIR) {
        // split-K cluster: the `split` CTAs of a tile (grid.z) are co-scheduled as one cluster and reduce in-kernel;
        // LayerNorm epilogue: the N tiles of a row block form the cluster
        cudaLaunchConfig_t cfg{};
        cfg.gridDim = grid; cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
        cudaLaunchAttribute attr[2];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = static_cast<unsigned>(g.mp.split);
        if (g.mp.ln_fuse) {   // LayerNorm epilogue: the N tiles of a row block (grid.x) are one cluster
            attr[0].val.clusterDim.x = static_cast<unsigned>(g.n_tiles); attr[0].val.clusterDim.z = 1;
        }
        attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr; cfg.numAttrs = pdl_enabled() ? 2 : 1;
        return check_cuda(cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, DEEP, PAIR, PERSIST>, g.tmA, g.tmW, g.tmC, g.tmR, g.tmC2, g.tmA2, g.mp, g.ep),
                          "launch gemm_tcgen05_kernel (split-K cluster)");
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid; cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = Cfg::kSmemBytes; cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;   // two consecutive M tiles
    cudaLaunchAttribute attr_pdl[2];
    if (g.mp.pair_pdl && pdl_enabled()) {    // experiment switch SDOD_PAIR_PDL=1 (default: no programmatic serialization for pairs, see the kernel)
        attr_pdl[0] = attr[0];
        attr_pdl[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr_pdl[1].val.programmaticStreamSerializationAllowed = 1;
        cfg.attrs = attr_pdl; cfg.numAttrs = 2;
    } else {
        cfg.attrs = attr; cfg.numAttrs = 1;
    }
    cudaError_t e = cudaLaunchKernelEx(&cfg, gemm_tcgen05_kernel<BN, DEEP, PAIR, PERSIST>, g.tmA, g.tmW, g.tmC, g.tmR, g.tmC2, g.tmA2, g.mp, g.ep);
    if (e != cudaSuccess) {
        int nc = -1;
        cudaError_t e2 = cudaOccupancyMaxActiveClusters(&nc, gemm_tcgen05_kernel<BN, DEEP, PAIR, PERSIST>, &cfg);
        return fail(kCudaError, std::string("cudaLaunchKernelEx(gemm pair bn=") + std::to_string(BN) + " deep=" + std::to_string(DEEP) + " grid=" +
                                    std::to_string(grid.x) + "x" + std::to_string(grid.y) + "x" + std::to_string(grid.z) + " smem=" +
                                    std::to_string(Cfg::kSmemBytes) + "): " + cudaGetErrorName(e) + "; max active clusters " + std::to_string(nc) +
                                    " (" + cudaGetErrorName(e2) + ")");
    }
    return kOk;
}

template <int BN>
static int launch_gemm(cudaStream_t stream, const GemmLaunch& g) {
    const MainloopParams& mp = g.mp;
    dim3 grid(g.n_tiles, g.m_tiles, mp.split > 1 ? mp.split : g.batch);
    if (g.pair) grid = dim3((g.m_tiles + 1) & ~1, g.n_tiles, grid.z);
    const long long ctas = static_cast<long long>(grid.x) * grid.y * grid.z;
    if (grid.y > 65535u || grid.z > 65535u)
        return fail(kUnsupported, "gemm / conv: " + std::to_string(g.m_tiles) + " row tiles exceed one launch (65,535): split the batch");
    if constexpr (BN == 128 || BN == 160) {
        if (g.persist) {
            const int sms = device_sm_count();
            int pgrid = mp.streamk ? g.sk_grid : (mp.tiles_total < sms ? mp.tiles_total : sms);
            if (mp.b_resident) pgrid = (sms / g.n_tiles) * g.n_tiles;      // a multiple of the N tile count: tile t + k*grid keeps t's N tile
            SDOD_TRY((launch_gemm_cfg<BN, true, false, true>(stream, g, dim3(pgrid))));
            count_launch();
            return check_launch("gemm_tcgen05_kernel (persistent)");
        }
    }
    if constexpr (BN >= 128) {
        if (g.pair) {
            if (ctas <= 148) SDOD_TRY((launch_gemm_cfg<BN, true, true>(stream, g, grid)));
            else SDOD_TRY((launch_gemm_cfg<BN, false, true>(stream, g, grid)));
        }
    }
    if (!g.pair) {
        if (ctas <= 148 || mp.ln_fuse) SDOD_TRY((launch_gemm_cfg<BN, true, false>(stream, g, grid)));    // (the LayerNorm epilogue stages in the deep ring)
        else SDOD_TRY((launch_gemm_cfg<BN, false, false>(stream, g, grid)));
    }
    count_launch();
    SDOD_TRY(check_launch("gemm_tcgen05_kernel"));
    if (mp.split > 1 && !mp.split_cluster) {
        dim3 rgrid(BN / 16, g.m_tiles * g.n_tiles);
        SDOD_TRY(check_cuda(launch_pdl(splitk_reduce_kernel<BN>, rgrid, dim3(256), 0, stream, mp, g.ep, g.n_tiles), "launch splitk_reduce_kernel"));
        count_launch();
        return check_launch("splitk_reduce_kernel");
    }
    return kOk;
}

static int dispatch_gemm(const GemmLaunch& g, cudaStream_t stream) {
    const int bn = g.bn;
    switch (bn) {
        case 32: return launch_gemm<32>(stream, g);
        case 64: return launch_gemm<64>(stream, g);
        case 128: return launch_gemm<128>(stream, g);
        case 160: return launch_gemm<160>(stream, g);
        case 256: return launch_gemm<256>(stream, g);
    }
    return fail(kInvalidArgument, "unsupported block_n " + std::to_string(bn));
}

// CTA pairs (cta_group::2) for multi-wave implicit-GEMM convolutions (K = 9*Cin >= 45 blocks): measured on B200 (r1,
// tools/conv_sweep.py) they gain 5-14 % there (B8 32x32 1280->640: 99 -> 87 us) because each SM pulls half the weight bytes
// per K block; single-wave grids and short-K linear layers lose 3-15 % to the cluster launch / sync cost, so they stay
// unpaired.  SDOD_GEMM_PAIR=0 disables pairs, =2 forces them wherever legal (A/B measurements).
static bool use_pair(int bn, int m_tiles, int n_tiles, bool conv) {
    static const int env = [] { const char* e = std::getenv("SDOD_GEMM_PAIR"); return e ? std::atoi(e) : 1; }();
    if (!env || bn < 128 || m_tiles < 2) return false;
    if (env == 2) return (m_tiles % 2 == 0) || m_tiles >= 9;
    if (env == 3) return conv && (m_tiles % 2 == 0);                     // (experiment: single-wave convs as pairs too)
    return conv && (m_tiles % 2 == 0) && static_cast<long long>(m_tiles) * n_tiles > 148;
}

// Tile width: the candidate (160 / 128 / 64) that wastes the fewest padded columns, wider first.
// (Measured on B200: 128- and 160-wide tiles with 2 CTAs/SM beat 256-wide by ~5 % on large GEMMs, and
// divide the SD channel counts 320/640/960/1280/1920/3840 exactly.)  GEGLU tiles are 128 wide (64 value | 64 gate).
int pick_block_n(int M, int N, int batch, int act) { return pick_block_n_k(M, N, batch, act, 1 << 30); }

// With a short K loop (<= 24 blocks) a split-K pair of kernels costs more than it saves: prefer a narrower tile that
// yields >= 100 CTAs in a single kernel with the TMA epilogue (measured r1: M2048 N640 K640 15.6 us as bn160 split 2).
int pick_block_n_k(int M, int N, int batch, int act, int k_blocks) {
    if (act == SDOD_ACT_GEGLU) return 128;
    if (N <= 32) return 32;
    if (N <= 64) return 64;
    const int cands[3] = {160, 128, 64};
    int best = 128;
    long long best_pad = 1LL << 60;
    for (int c = 0; c < 3; ++c) {
        const int bn = cands[c];
        const long long pad = static_cast<long long>((N + bn - 1) / bn) * bn - N;
        if (pad < best_pad) { best_pad = pad; best = bn; }
    }
    const long long m_tiles = (M + kBlockM - 1) / kBlockM;
    auto tiles = [&](int bn) { return m_tiles * ((N + bn - 1) / bn) * batch; };
    if (k_blocks <= 24 && tiles(best) < 100) {
        if (best > 128 && tiles(128) >= 100) return 128;
        if (tiles(64) >= 60) return 64;
    }
    return best;
}

// TMA-store epilogue eligibility: row-major bf16/fp32 output, 16-B aligned rows, no GEGLU / split-K, and a residual
// (if any) of the output's own dtype so it can be staged and added in place.
static int setup_tma_epilogue(GemmLaunch* out, const sdod_epilogue& ep, int M, int N, int batch) {   // N by value: halved for GEGLU
    MainloopParams& mp = out->mp;
    mp.tma_epi = 0;
    mp.c_bytes = ep.out_mode == SDOD_OUT_F32 ? 4 : 2;
    std::memset(&out->tmC, 0, sizeof(CUtensorMap));
    std::memset(&out->tmR, 0, sizeof(CUtensorMap));
    if (mp.split > 1) return kOk;
    if (ep.out_mode != SDOD_OUT_BF16 && ep.out_mode != SDOD_OUT_F32) return kOk;
    if (ep.act == SDOD_ACT_GEGLU) {
        if (ep.out_mode != SDOD_OUT_BF16 || ep.residual || (out->bn / 2) % 32 != 0) return kOk;
        N = N / 2;                                   // the output tensor holds the value*gelu(gate) half
    }
    const int es = mp.c_bytes;
    if ((reinterpret_cast<uintptr_t>(ep.C) & 15) || (ep.ldc * es) % 16 || (batch > 1 && (ep.strideC * es) % 16)) return kOk;
    if (ep.residual) {
        const bool r32 = ep.residual_f32 != 0;
        if (r32 != (es == 4)) return kOk;
        if ((reinterpret_cast<uintptr_t>(ep.residual) & 15) || (ep.ldr * es) % 16 || (batch > 1 && (ep.strideR * es) % 16)) return kOk;
    }
    const uint32_t box[3] = {32, 32, 1};
    {
        uint64_t dims[3] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M), static_cast<uint64_t>(batch)};
        uint64_t strides[2] = {static_cast<uint64_t>(ep.ldc) * es, static_cast<uint64_t>(batch > 1 ? ep.strideC : static_cast<long long>(M) * ep.ldc) * es};
        SDOD_TRY(encode_tmap(&out->tmC, ep.C, es, 3, dims, strides, box, es == 4 ? 128 : 64));
    }
    mp.tma_epi = 1;
    if (ep.residual) {
        uint64_t dims[3] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M), static_cast<uint64_t>(batch)};
        uint64_t strides[2] = {static_cast<uint64_t>(ep.ldr) * es, static_cast<uint64_t>(batch > 1 ? ep.strideR : static_cast<long long>(M) * ep.ldr) * es};
        SDOD_TRY(encode_tmap(&out->tmR, ep.residual, es, 3, dims, strides, box, es == 4 ? 128 : 64));
        mp.tma_epi = 2;
    }
    return kOk;
}

// Head-layout outputs (attention operands) through TMA stores: see the `tma_epi == 3` branch of the kernel.
static bool heads_tma_eligible(const sdod_epilogue& ep, int M, int N, int batch) {
    if (ep.out_mode != SDOD_OUT_HEADS && ep.out_mode != SDOD_OUT_HEADS_T && ep.out_mode != SDOD_OUT_QKV) return false;
    if (batch != 1 || ep.residual || ep.row_bias || ep.act != SDOD_ACT_NONE) return false;
    if (ep.head_dim % 40 != 0 || ep.tokens % 32 != 0 || M % ep.tokens != 0 || N % 40 != 0) return false;
    const int Cw = ep.heads * ep.head_dim;
    if (ep.out_mode == SDOD_OUT_QKV ? (Cw % 160 != 0) : (N != Cw)) return false;
    const bool need_q = ep.out_mode != SDOD_OUT_HEADS_T, need_v = ep.out_mode != SDOD_OUT_HEADS;
    if (need_q && (ep.dpad % 8 != 0 || ep.dpad < ep.head_dim)) return false;
    if (need_v && (ep.tok_pad % 8 != 0 || ep.tok_pad < ep.tokens)) return false;
    if (reinterpret_cast<uintptr_t>(ep.C) & 15) return false;
    if (ep.out_mode == SDOD_OUT_QKV && ((reinterpret_cast<uintptr_t>(ep.C2) & 15) || (reinterpret_cast<uintptr_t>(ep.C3) & 15))) return false;
    return true;
}

static int setup_heads_epilogue(GemmLaunch* out, const sdod_epilogue& ep, int M, int N, int batch) {
    std::memset(&out->tmC2, 0, sizeof(CUtensorMap));
    MainloopParams& mp = out->mp;
    if (mp.split > 1 || out->bn != 160 || !heads_tma_eligible(ep, M, N, batch)) return kOk;
    const uint64_t BH = static_cast<uint64_t>(M / ep.tokens) * ep.heads;
    auto q_map = [&](CUtensorMap* tm, void* base) {
        uint64_t dims[3] = {static_cast<uint64_t>(ep.dpad), static_cast<uint64_t>(ep.tokens), BH};
        uint64_t strides[2] = {static_cast<uint64_t>(ep.dpad) * 2, static_cast<uint64_t>(ep.tokens) * ep.dpad * 2};
        const uint32_t box[3] = {40, 32, 1};
        return encode_tmap(tm, base, 2, 3, dims, strides, box, 0);
    };
    auto v_map = [&](CUtensorMap* tm, void* base) {
        uint64_t dims[3] = {static_cast<uint64_t>(ep.tok_pad), static_cast<uint64_t>(ep.vt_rows), BH};
        uint64_t strides[2] = {static_cast<uint64_t>(ep.tok_pad) * 2, static_cast<uint64_t>(ep.vt_rows) * ep.tok_pad * 2};
        const uint32_t box[3] = {32, 40, 1};
        return encode_tmap(tm, base, 2, 3, dims, strides, box, 0);
    };
    if (ep.out_mode == SDOD_OUT_HEADS) {
        SDOD_TRY(q_map(&out->tmC, ep.C));
    } else if (ep.out_mode == SDOD_OUT_HEADS_T) {
        SDOD_TRY(v_map(&out->tmC2, ep.C));
    } else {
        SDOD_TRY(q_map(&out->tmC, ep.C));
        SDOD_TRY(q_map(&out->tmR, ep.C2));
        SDOD_TRY(v_map(&out->tmC2, ep.C3));
    }
    mp.tma_epi = 3;
    return kOk;
}

// LayerNorm fused into the epilogue (see the kernel's ln_fuse branch): fp32 TMA-store output, every N tile full, the N tiles of a row
// block small enough for one portable cluster, a single-wave grid (the DEEP variant, whose idle ring holds both staging areas).
static int setup_ln_epilogue(GemmLaunch* out, const sdod_epilogue& ep, int M, int N, int batch, bool pair) {
    MainloopParams& mp = out->mp;
    mp.ln_fuse = 0;
    if (!ep.ln_out) return kOk;
    const int bn = out->bn, n_tiles = N / bn;
    const long long ctas = static_cast<long long>((M + kBlockM - 1) / kBlockM) * n_tiles;
    // single-wave grids only: a multi-wave attempt (one DEEP CTA per SM, B200 r2) faulted and would forfeit the two-CTA overlap anyway
    const long long max_ctas = 148;
    // few-CTA layers (16x16 / 8x8 levels at batch 2) are better served by split-K + a LayerNorm launch than by a fused epilogue that rules split-K out
    static const long long min_ctas = [] { const char* e = std::getenv("SDOD_LN_FUSE_MINCTAS"); return e ? std::atoll(e) : 0LL; }();
    if (pair || batch != 1 || mp.split > 1 || (mp.tma_epi != 1 && mp.tma_epi != 2) || mp.c_bytes != 4 || ep.act != SDOD_ACT_NONE || N % bn != 0 ||
        n_tiles > 8 || ctas > max_ctas || ctas < min_ctas || (bn != 128 && bn != 160) || M % kBlockM != 0)
        return fail(kUnsupported, "gemm: this shape cannot take the fused LayerNorm epilogue");
    if ((reinterpret_cast<uintptr_t>(ep.ln_out) & 15) || (ep.ld_ln * 2) % 16) return fail(kInvalidArgument, "gemm: ln_out must be 16-byte aligned");
    const uint32_t box[3] = {32, 32, 1};
    uint64_t dims[3] = {static_cast<uint64_t>(N), static_cast<uint64_t>(M), 1};
    uint64_t strides[2] = {static_cast<uint64_t>(ep.ld_ln) * 2, static_cast<uint64_t>(M) * ep.ld_ln * 2};
    SDOD_TRY(encode_tmap(&out->tmC2, ep.ln_out, 2, 3, dims, strides, box, 64));
    mp.ln_fuse = 1;
    return kOk;
}

// Persistent scheduling for short-K, multi-wave GEMMs with a TMA epilogue: the per-tile fill/drain chain dominates there.
// SDOD_GEMM_PERSIST=0 disables it, =2 uses it wherever legal (A/B measurements).
static void choose_persist(GemmLaunch* out) {
    static const int env = [] { const char* e = std::getenv("SDOD_GEMM_PERSIST"); return e ? std::atoi(e) : 1; }();
    MainloopParams& mp = out->mp;
    mp.n_tiles = out->n_tiles; mp.m_tiles = out->m_tiles;
    mp.tiles_total = out->n_tiles * out->m_tiles * out->batch;
    static const int pair_relaxed_env = [] { const char* e = std::getenv("SDOD_PAIR_RELAXED"); return e ? std::atoi(e) : 1; }();
    mp.pair_relaxed = pair_relaxed_env;
    static const int pair_pdl_env = [] { const char* e = std::getenv("SDOD_PAIR_PDL"); return e ? std::atoi(e) : 0; }();
    mp.pair_pdl = pair_pdl_env;
    out->persist = 0;
    if (mp.streamk) { out->persist = 1; return; }     // stream-K runs on the persistent variant, grid = out->sk_grid
    if (!env || out->pair || mp.split > 1 || !mp.tma_epi || (out->bn != 128 && out->bn != 160)) return;
    const int sms = device_sm_count();
    if (env == 2) { out->persist = mp.tiles_total > sms; return; }
    // measured (tools/hot_kernels.py, B200 r1): K=320 GEGLU 111.6 -> 92.2 us, QKV 87.0 -> 82.9 us; K=1280 (20 blocks) 46.1 -> 58.3 us,
    // where two co-resident CTAs with their own rings hide the loads better than one 4-stage ring
    out->persist = (mp.tiles_total >= 3 * sms && mp.k_blocks <= 10) ? 1 : 0;
    // fp32 output with an fp32 residual (the transformer blocks' to_out / proj_out at K = 320): the staging region holds ONE fp32 tile, so a persistent
    // CTA runs residual load -> add -> store strictly in sequence per tile, while two co-resident one-tile CTAs overlap theirs
    // (B200 r2, M131072 N320 K320: 98.7 us persistent, 90.4 us not; K = 640: 60.3 vs 61.0 us, left persistent)
    static const int res_env = [] { const char* e = std::getenv("SDOD_PERSIST_RES"); return e ? std::atoi(e) : 0; }();
    if (!res_env && mp.tma_epi == 2 && mp.k_blocks <= 5) out->persist = 0;
    // two epilogue groups on alternate tiles wherever the staging region holds two tiles (bf16 / GEGLU / head-layout outputs, no residual to
    // pre-load): SDOD_EPI_GROUPS=1 restores the single group of 16 warps (A/B measurements)
    // weight-stationary walk (see MainloopParams::b_resident): with the ring exactly one tile deep the weight half never has to be re-fetched, which
    // halves the L2 -> SM operand bytes of a K = 320 layer.  Opt-in (SDOD_B_RESIDENT=1): measured on B200 (r2) it is SLOWER — GEGLU projection at
    // batch 32 383.9 vs 369.7 us, batch 8 106.5 vs 103.5 us — so operand traffic is not what bounds these layers (and 140 of 148 SMs are used).
    static const int bres_env = [] { const char* e = std::getenv("SDOD_B_RESIDENT"); return e ? std::atoi(e) : 0; }();
    mp.b_resident = 0;
    if (out->persist && bres_env && out->batch == 1 && !mp.w_batched && !mp.k_rot && mp.kb_a2 >= mp.k_blocks && out->n_tiles <= sms &&
        mp.tiles_total >= (sms / out->n_tiles) * out->n_tiles) {
        const int stages = out->bn == 128 ? GemmCfg<128, true, false, true>::kStages : GemmCfg<160, true, false, true>::kStages;
        if (mp.k_blocks == stages) mp.b_resident = 1;
    }
    static const int groups_env = [] { const char* e = std::getenv("SDOD_EPI_GROUPS"); return e ? std::atoi(e) : 2; }();
    static const int accs_env = [] { const char* e = std::getenv("SDOD_TMEM_ACCS"); return e ? std::atoi(e) : 0; }();
    mp.tmem_accs = accs_env;
    mp.epi_groups = (out->persist && groups_env == 2 && (mp.tma_epi == 3 || (mp.tma_epi == 1 && mp.c_bytes == 2))) ? 2 : 1;
}

static int k_rotation(int k_blocks) {
    static const int env = [] { const char* e = std::getenv("SDOD_GEMM_KROT"); return e ? std::atoi(e) : 0; }();
    return k_blocks >= 8 ? env : 0;
}

static thread_local SplitKWorkspace g_splitk;
void set_splitk_workspace(const SplitKWorkspace& w) { g_splitk = w; }
static thread_local unsigned long long* g_tlog = nullptr;
void set_gemm_timeline(unsigned long long* buf) { g_tlog = buf; }

// Split-K factor: small-M layers (8x8 / 16x16 levels at batch 2) have too few output tiles to fill 148 SMs and
// are bound by streaming the weights; splitting K spreads that stream over the whole chip.
static void choose_split(MainloopParams* mp, int bn, int m_tiles, int n_tiles, int batch) {
    mp->split = 1; mp->kb_per_split = mp->k_blocks; mp->ws = nullptr; mp->counters = nullptr;
    if (batch != 1 || !g_splitk.ws) return;
    const long long tiles = static_cast<long long>(m_tiles) * n_tiles;
    if (tiles >= 120) return;
    if (tiles >= 60 && mp->k_blocks <= 24) return;             // short K: one kernel with the TMA epilogue beats split + reduce
    // One CTA per SM with the deep ring (tiles x split <= #SMs) rather than two shallow-ring CTAs per SM with twice the split: the mainloop is
    // bound by each SM's operand ingest either way (timeline, profiles/r02_*), half the partial tiles go through L2, and a grid that
    // overshoots the resident slots (64 tiles x 5 = 320 CTAs on 296 slots) no longer pays a second, nearly empty wave.
    // SDOD_SPLIT_POLICY=0 restores the round-1 rule (A/B measurements).
    static const int policy = [] { const char* e = std::getenv("SDOD_SPLIT_POLICY"); return e ? std::atoi(e) : 1; }();
    const int sms = device_sm_count();
    int split = policy ? static_cast<int>(sms / tiles) : static_cast<int>((2 * 148 + tiles - 1) / tiles);
    const int max_by_k = mp->k_blocks / 4;                    // at least 4 K blocks (256 deep) per split
    if (split > max_by_k) split = max_by_k;
    const int cap = (policy && m_tiles == 1) ? 16 : 8;        // bounds the partial-tile traffic (L2-resident); single-row-block layers are pure weight streams
    if (split > cap) split = cap;
    const size_t per_tile = static_cast<size_t>(bn) * kBlockM * sizeof(float);
    while (split > 1 && static_cast<size_t>(tiles) * split * per_tile > g_splitk.ws_bytes) --split;
    if (split < 2) return;
    const int kbps = (mp->k_blocks + split - 1) / split;
    split = (mp->k_blocks + kbps - 1) / kbps;                 // no empty splits
    if (split < 2) return;
    mp->split = split; mp->kb_per_split = kbps; mp->ws = g_splitk.ws; mp->counters = g_splitk.counters;
}

// Stream-K decision (see WorkWalk): worthwhile when whole-tile scheduling would leave the chip badly quantised — few tiles, long K — i.e.
// when an equal share of the (tile, K block) space per SM is clearly shorter than ceil(tiles / SMs) full K loops.  SDOD_STREAMK=0 disables it.
// Returns the grid size (0 = do not use stream-K).
static int streamk_grid(int bn, long long tiles, int k_blocks, int batch, bool pair) {
    // Off by default: measured on B200 (r2, profiles/r02_streamk_notes.txt) the batch-2 step ran 6.8 ms with stream-K against 5.7 ms with split-K —
    // the single owner folding up to 18 published partials serialises what the separate reduce kernel does with 300 CTAs.  SDOD_STREAMK=1 enables it.
    static const int env = [] { const char* e = std::getenv("SDOD_STREAMK"); return e ? std::atoi(e) : 0; }();
    if (!env || pair || batch != 1 || (bn != 128 && bn != 160) || !g_splitk.ws || k_blocks < 16) return 0;
    const int sms = device_sm_count();
    const long long total = tiles * k_blocks;
    if (tiles > 4LL * sms || total < 8LL * 16) return 0;
    int grid = static_cast<int>(std::min<long long>(sms, total / 8));                 // at least 8 K blocks per CTA
    if (grid < 2) return 0;
    if (static_cast<size_t>(grid) * bn * kBlockM * sizeof(float) > g_splitk.ws_bytes || grid > g_splitk.n_counters) return 0;
    const long long sk_blocks = (total + grid - 1) / grid + 2;                          // + publish / fold overhead
    const long long classic = ((tiles + sms - 1) / sms) * k_blocks;
    if (env != 2 && sk_blocks * 5 > classic * 4) return 0;                              // needs a >= 20 % shorter critical path
    return grid;
}

// In-kernel reduction over a (1,1,split) cluster (one launch instead of two) with SDOD_SPLITK_CLUSTER=1 (partials through the L2 scratch, release/
// acquire cluster barrier) or =2 (partials pushed into the owner CTA's idle ring over distributed shared memory, st.async on mbarriers: no scratch
// round trip, no fence).  Off by default: measured on B200 (r2) the batch-2 step is slower either way — 4.92 ms (=1) and 4.94 ms (=2) against 4.80 ms
// with the separate reduce kernel, at 299 instead of 340 launches.  The exchange mechanism is not what costs: a (1,1,split) cluster can only start
// when `split` SMs of one GPC are free at once, so these launches lose the programmatic overlap with their predecessor's tail that the plain
// launches (and the reduce kernel's 80 x tiles small CTAs) get.
static int split_cluster_mode(const MainloopParams& mp, bool pair, int act) {
    static const int env = [] { const char* e = std::getenv("SDOD_SPLITK_CLUSTER"); return e ? std::atoi(e) : 0; }();
    return (env && mp.split > 1 && mp.split <= 8 && !pair && act != SDOD_ACT_GEGLU) ? (env == 2 ? 2 : 1) : 0;
}

// Second A operand (K-concatenated): bf16 [M, K2] rows of ld2 elements, loaded as plain {64, 128} boxes after tmA's K blocks.
static int setup_second_operand(GemmLaunch* out, const void* A2, long long ld2, int K2, int M, int kb_main) {
    std::memset(&out->tmA2, 0, sizeof(CUtensorMap));
    out->mp.kb_a2 = 1 << 30;
    if (!A2 || K2 <= 0) return kOk;
    if (K2 % kBlockK != 0 || ld2 % 8 != 0) return fail(kInvalidArgument, "second operand: K2 must be a multiple of 64 and its row stride of 8 elements");
    uint64_t dims[3] = {static_cast<uint64_t>(K2), static_cast<uint64_t>(M), 1};
    uint64_t strides[2] = {static_cast<uint64_t>(ld2) * 2, static_cast<uint64_t>(M) * ld2 * 2};
    uint32_t box[3] = {kBlockK, kBlockM, 1};
    SDOD_TRY(encode_tmap_bf16(&out->tmA2, A2, 3, dims, strides, box, true));
    out->mp.kb_a2 = kb_main;
    return kOk;
}

static int validate_epilogue(const sdod_epilogue& ep, int N) {
    if (!ep.C) return fail(kInvalidArgument, "epilogue: C is NULL");
    if (ep.row_bias && ep.rows_per_group <= 0) return fail(kInvalidArgument, "epilogue: rows_per_group must be > 0 with row_bias");
    if (ep.out_mode >= SDOD_OUT_HEADS) {
        if (ep.heads <= 0 || ep.head_dim <= 0 || ep.tokens <= 0 || (ep.dpad <= 0 && ep.out_mode != SDOD_OUT_HEADS_T))
            return fail(kInvalidArgument, "epilogue: heads/head_dim/tokens/dpad required for head layouts");
        if (ep.out_mode == SDOD_OUT_QKV && (!ep.C2 || !ep.C3 || N != 3 * ep.heads * ep.head_dim))
            return fail(kInvalidArgument, "epilogue: QKV mode needs C2, C3 and N == 3*heads*head_dim");
        if ((ep.out_mode == SDOD_OUT_HEADS_T || ep.out_mode == SDOD_OUT_QKV) && (ep.tok_pad <= 0 || ep.vt_rows < ep.head_dim))
            return fail(kInvalidArgument, "epilogue: tok_pad and vt_rows >= head_dim required for V^T layout");
    }
    if (ep.act == SDOD_ACT_GEGLU && (N % 2 != 0)) return fail(kInvalidArgument, "epilogue: GEGLU needs even N");
    return kOk;
}

static int gemm_prepare_impl(const sdod_gemm_desc& d, GemmLaunch* out, bool try_streamk, bool* used_streamk);
int gemm_prepare(const sdod_gemm_desc& d, GemmLaunch* out) {
    bool sk = false;
    SDOD_TRY(gemm_prepare_impl(d, out, true, &sk));
    if (sk && !out->mp.tma_epi) return gemm_prepare_impl(d, out, false, &sk);    // this epilogue has no persistent form: classic scheduling
    return kOk;
}
static int gemm_prepare_impl(const sdod_gemm_desc& d, GemmLaunch* out, bool try_streamk, bool* used_streamk) {
    if (!d.A || !d.W) return fail(kInvalidArgument, "gemm: NULL operand");
    if (d.M <= 0 || d.N <= 0 || d.K <= 0 || d.batch <= 0) return fail(kInvalidArgument, "gemm: non-positive extent");
    if (d.K % kBlockK != 0) return fail(kInvalidArgument, "gemm: K must be a multiple of 64 (pad the operand)");
    if (d.lda % 8 != 0 || d.ldw % 8 != 0) return fail(kInvalidArgument, "gemm: lda/ldw must be multiples of 8 elements");
    const int K2 = d.A2 ? d.K2 : 0;
    if (K2 && d.batch != 1) return fail(kInvalidArgument, "gemm: a second operand needs batch == 1");
    const int Ktot = d.K + K2;
    SDOD_TRY(validate_epilogue(d.epi, d.N));
    int bn = d.block_n ? d.block_n : pick_block_n_k(d.M, d.N, d.batch, d.epi.act, Ktot / kBlockK);
    if (!d.block_n && d.N % 160 == 0 && heads_tma_eligible(d.epi, d.M, d.N, d.batch)) bn = 160;   // four 40-column head boxes per tile
    if (!d.block_n && d.epi.ln_out) bn = d.N % 160 == 0 ? 160 : 128;                              // whole tiles only (fused LayerNorm)
    int sk_grid = 0;
    *used_streamk = false;
    if (try_streamk && !d.epi.ln_out) {
        const int bn_sk = d.block_n ? d.block_n : pick_block_n(d.M, d.N, d.batch, d.epi.act);      // the widest well-fitting tile: tile count is no concern
        const long long tiles_sk = static_cast<long long>((d.M + kBlockM - 1) / kBlockM) * ((d.N + bn_sk - 1) / bn_sk);
        sk_grid = streamk_grid(bn_sk, tiles_sk, Ktot / kBlockK, d.batch, false);
        if (sk_grid) { bn = bn_sk; *used_streamk = true; }
    }
    const bool pair = !sk_grid && use_pair(bn, (d.M + kBlockM - 1) / kBlockM, (d.N + bn - 1) / bn, false);

    CUtensorMap& tmA = out->tmA;
    CUtensorMap& tmW = out->tmW;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(d.K), static_cast<uint64_t>(d.M), static_cast<uint64_t>(d.batch)};
        uint64_t strides[2] = {static_cast<uint64_t>(d.lda) * 2, static_cast<uint64_t>(d.batch > 1 ? d.strideA : static_cast<long long>(d.M) * d.lda) * 2};
        uint32_t box[3] = {kBlockK, kBlockM, 1};
        SDOD_TRY(encode_tmap_bf16(&tmA, d.A, 3, dims, strides, box, true));
    }
    const bool wb = d.batch > 1 && d.strideW != 0;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(Ktot), static_cast<uint64_t>(d.N), static_cast<uint64_t>(wb ? d.batch : 1)};
        uint64_t strides[2] = {static_cast<uint64_t>(d.ldw) * 2, static_cast<uint64_t>(wb ? d.strideW : static_cast<long long>(d.N) * d.ldw) * 2};
        uint32_t box[3] = {kBlockK, static_cast<uint32_t>(pair ? bn / 2 : bn), 1};
        SDOD_TRY(encode_tmap_bf16(&tmW, d.W, 3, dims, strides, box, true));
    }
    out->pair = pair ? 1 : 0;
    MainloopParams mp{};
    mp.M = d.M; mp.N = d.N; mp.k_blocks = Ktot / kBlockK; mp.conv = 0; mp.w_batched = wb ? 1 : 0;
    mp.k_rot = k_rotation(mp.k_blocks);
    choose_split(&mp, bn, (d.M + kBlockM - 1) / kBlockM, (d.N + bn - 1) / bn, d.batch);
    if (d.epi.ln_out || sk_grid) { mp.split = 1; mp.kb_per_split = mp.k_blocks; }      // the LayerNorm epilogue needs the finished rows in one CTA row
    if (sk_grid) { mp.streamk = 1; mp.ws = g_splitk.ws; mp.counters = g_splitk.counters; }
    out->sk_grid = sk_grid;
    mp.split_cluster = split_cluster_mode(mp, pair, d.epi.act);
    {
        // Off by default: measured on B200 (r2) the tanh form takes the isolated 64x64 GEGLU projection at batch 32 from 373.8 to 353.2 us but the
        // whole batch-32 pass does not move (40.38 vs 40.33 ms) — not worth leaving the exact erf form for.  SDOD_GEGLU_TANH=1 enables it (A/B).
        static const int env = [] { const char* e = std::getenv("SDOD_GEGLU_TANH"); return e ? std::atoi(e) : 0; }();
        mp.geglu_tanh = env;
    }
    out->mp = mp; out->ep = d.epi; out->bn = bn;
    out->mp.tlog = g_tlog;
    SDOD_TRY(setup_second_operand(out, d.A2, d.lda2, K2, d.M, d.K / kBlockK));
    out->w_ptr = d.W; out->w_bytes = static_cast<long long>(d.N) * d.ldw * 2 * (wb ? d.batch : 1);
    SDOD_TRY(setup_tma_epilogue(out, d.epi, d.M, d.N, d.batch));
    SDOD_TRY(setup_heads_epilogue(out, d.epi, d.M, d.N, d.batch));
    SDOD_TRY(setup_ln_epilogue(out, d.epi, d.M, d.N, d.batch, pair));
    {
        // in-place residual (fp32 stream: x += f(x)) without the LayerNorm epilogue: TMA reduce-add store instead of residual load + add
        static const int env = [] { const char* e = std::getenv("SDOD_TMA_REDUCE"); return e ? std::atoi(e) : 1; }();
        if (env && out->mp.tma_epi == 2 && !out->mp.ln_fuse && out->mp.c_bytes == 4 && d.epi.residual == d.epi.C && d.epi.ldr == d.epi.ldc &&
            (d.batch == 1 || d.epi.strideR == d.epi.strideC))
            out->mp.tma_epi = 4;
    }
    out->m_tiles = (d.M + kBlockM - 1) / kBlockM;
    out->n_tiles = (d.N + bn - 1) / bn;
    out->batch = d.batch;
    choose_persist(out);
    if (out->mp.ln_fuse) out->persist = 0;
    return kOk;
}

int gemm_launch(const GemmLaunch& g, cudaStream_t stream) {
    return dispatch_gemm(g, stream);
}

int gemm_bf16(cudaStream_t stream, const sdod_gemm_desc& d) {
    GemmLaunch g;
    SDOD_TRY(gemm_prepare(d, &g));
    return gemm_launch(g, stream);
}

static int conv3x3_prepare_impl(const sdod_conv_desc& d, GemmLaunch* out, bool try_streamk, bool* used_streamk);
int conv3x3_prepare(const sdod_conv_desc& d, GemmLaunch* out) {
    bool sk = false;
    SDOD_TRY(conv3x3_prepare_impl(d, out, true, &sk));
    if (sk && !out->mp.tma_epi) return conv3x3_prepare_impl(d, out, false, &sk);
    return kOk;
}
static int conv3x3_prepare_impl(const sdod_conv_desc& d, GemmLaunch* out, bool try_streamk, bool* used_streamk) {
    if (!d.X || !d.Wt) return fail(kInvalidArgument, "conv3x3: NULL operand");
    if (d.B <= 0 || d.H <= 0 || d.W <= 0 || d.Cin <= 0 || d.Cout <= 0) return fail(kInvalidArgument, "conv3x3: non-positive extent");
    if (d.Cin % kBlockK != 0) return fail(kInvalidArgument, "conv3x3: Cin must be a multiple of 64 (use im2col + gemm otherwise)");
    const int cstride = d.stride == 2 ? 2 : 1;
    if (d.stride != 0 && d.stride != 1 && d.stride != 2) return fail(kInvalidArgument, "conv3x3: stride must be 1 or 2");
    if (cstride == 2 && (d.H % 2 != 0 || d.W % 2 != 0 || d.upsample2x || d.X2)) return fail(kUnsupported, "conv3x3 (stride 2): even H, W; no upsample / second operand");
    const int Ho = d.H / cstride, Wo = d.W / cstride;            // tiles are laid out on the OUTPUT grid
    const int bw = Wo < 128 ? Wo : 128;
    if (128 % bw != 0 || Wo % bw != 0) return fail(kInvalidArgument, "conv3x3: W must be a power of two (or a multiple of 128)");
    int bh = 128 / bw;
    if (bh > Ho) bh = Ho;
    if (Ho % bh != 0 || 128 % (bw * bh) != 0) return fail(kInvalidArgument, "conv3x3: H*W must tile into 128-pixel boxes");
    const int bb = 128 / (bw * bh);
    const int M = d.B * Ho * Wo;
    SDOD_TRY(validate_epilogue(d.epi, d.Cout));
    if (d.epi.ln_out) return fail(kUnsupported, "conv3x3: no fused LayerNorm epilogue");
    const int Cin2 = d.X2 ? d.Cin2 : 0;
    const bool up2 = d.upsample2x != 0;                 // sub-pixel form of conv3x3(nearest_upsample_2x(X)): 4 output parities x 2x2 source taps
    const int taps = up2 ? 4 : 9;
    if (up2) {
        try_streamk = false;
        if (d.X2 || d.epi.residual || d.epi.row_bias || d.epi.act == SDOD_ACT_GEGLU || (d.epi.out_mode != SDOD_OUT_BF16 && d.epi.out_mode != SDOD_OUT_F32))
            return fail(kUnsupported, "conv3x3 (upsample2x): no second operand, residual, row bias, GEGLU or head layouts");
        if (d.Cout % 8 != 0 || (reinterpret_cast<uintptr_t>(d.epi.C) & 15)) return fail(kInvalidArgument, "conv3x3 (upsample2x): Cout % 8 and a 16-byte aligned output required");
    }
    int bn = d.block_n ? d.block_n : pick_block_n_k(M, d.Cout, up2 ? 4 : 1, d.epi.act, (taps * d.Cin + Cin2) / kBlockK);
    int sk_grid = 0;
    *used_streamk = false;
    if (try_streamk) {
        const int bn_sk = d.block_n ? d.block_n : pick_block_n(M, d.Cout, 1, d.epi.act);
        const long long tiles_sk = static_cast<long long>((M + kBlockM - 1) / kBlockM) * ((d.Cout + bn_sk - 1) / bn_sk);
        sk_grid = streamk_grid(bn_sk, tiles_sk, (taps * d.Cin + Cin2) / kBlockK, 1, false);
        if (sk_grid) { bn = bn_sk; *used_streamk = true; }
    }
    const bool pair = !sk_grid && use_pair(bn, (M + kBlockM - 1) / kBlockM, (d.Cout + bn - 1) / bn, true);

    CUtensorMap& tmA = out->tmA;
    CUtensorMap& tmW = out->tmW;
    {
        uint64_t dims[4] = {static_cast<uint64_t>(d.Cin), static_cast<uint64_t>(d.W), static_cast<uint64_t>(d.H), static_cast<uint64_t>(d.B)};
        uint64_t strides[3] = {static_cast<uint64_t>(d.Cin) * 2, static_cast<uint64_t>(d.W) * d.Cin * 2,
                               static_cast<uint64_t>(d.H) * d.W * d.Cin * 2};
        uint32_t box[4] = {kBlockK, static_cast<uint32_t>(cstride * bw), static_cast<uint32_t>(cstride * bh), static_cast<uint32_t>(bb)};
        const uint32_t estr[4] = {1, static_cast<uint32_t>(cstride), static_cast<uint32_t>(cstride), 1};
        SDOD_TRY(encode_tmap(&tmA, d.X, 2, 4, dims, strides, box, 128, cstride == 2 ? estr : nullptr));
    }
    const int K = taps * d.Cin + Cin2;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(K), static_cast<uint64_t>(d.Cout), static_cast<uint64_t>(up2 ? 4 : 1)};
        uint64_t strides[2] = {static_cast<uint64_t>(K) * 2, static_cast<uint64_t>(K) * d.Cout * 2};
        uint32_t box[3] = {kBlockK, static_cast<uint32_t>(pair ? bn / 2 : bn), 1};
        SDOD_TRY(encode_tmap_bf16(&tmW, d.Wt, 3, dims, strides, box, true));
    }
    out->pair = pair ? 1 : 0;
    MainloopParams mp{};
    mp.M = M; mp.N = d.Cout; mp.k_blocks = K / kBlockK; mp.conv = 1; mp.cin_blocks = d.Cin / kBlockK;
    mp.H = Ho; mp.W = Wo; mp.cstride = cstride; mp.bw = bw; mp.bh = bh; mp.bb = bb; mp.w_batched = up2 ? 1 : 0; mp.up2 = up2 ? 1 : 0;
    mp.k_rot = k_rotation(mp.k_blocks);
    choose_split(&mp, bn, (M + kBlockM - 1) / kBlockM, (d.Cout + bn - 1) / bn, up2 ? 4 : 1);      // (batch != 1: grid.z is taken by the parities)
    if (sk_grid) { mp.split = 1; mp.kb_per_split = mp.k_blocks; mp.streamk = 1; mp.ws = g_splitk.ws; mp.counters = g_splitk.counters; }
    out->sk_grid = sk_grid;
    mp.split_cluster = split_cluster_mode(mp, pair, d.epi.act);
    out->mp = mp; out->ep = d.epi; out->bn = bn;
    out->mp.tlog = g_tlog;
    SDOD_TRY(setup_second_operand(out, d.X2, d.ldx2, Cin2, M, taps * d.Cin / kBlockK));
    out->w_ptr = d.Wt; out->w_bytes = static_cast<long long>(d.Cout) * K * 2 * (up2 ? 4 : 1);
    if (!up2) {
        SDOD_TRY(setup_tma_epilogue(out, d.epi, M, d.Cout, 1));
    } else {
        // output [B, 2H, 2W, Cout] seen as (c, px, x, py, (b, y)): a tile row = source pixel (b, y, x) lands at output pixel (2y+py, 2x+px)
        out->mp.tma_epi = 1;
        out->mp.c_bytes = d.epi.out_mode == SDOD_OUT_F32 ? 4 : 2;
        std::memset(&out->tmR, 0, sizeof(CUtensorMap));
        const uint64_t es = static_cast<uint64_t>(out->mp.c_bytes), C = static_cast<uint64_t>(d.Cout);
        uint64_t dims[5] = {C, 2, static_cast<uint64_t>(d.W), 2, static_cast<uint64_t>(d.B) * d.H};
        uint64_t strides[4] = {C * es, 2 * C * es, 2 * static_cast<uint64_t>(d.W) * C * es, 4 * static_cast<uint64_t>(d.W) * C * es};
        const uint32_t bwq = static_cast<uint32_t>(bw < 32 ? bw : 32);
        const uint32_t box[5] = {32, 1, bwq, 1, 32 / bwq};
        SDOD_TRY(encode_tmap(&out->tmC, d.epi.C, static_cast<int>(es), 5, dims, strides, box, es == 4 ? 128 : 64));
    }
    std::memset(&out->tmC2, 0, sizeof(CUtensorMap));
    out->m_tiles = (M + kBlockM - 1) / kBlockM;
    out->n_tiles = (d.Cout + bn - 1) / bn;
    out->batch = up2 ? 4 : 1;
    choose_persist(out);
    if (up2) out->persist = 0;
    return kOk;
}

int conv3x3_bf16(cudaStream_t stream, const sdod_conv_desc& d) {
    GemmLaunch g;
    SDOD_TRY(conv3x3_prepare(d, &g));
    return gemm_launch(g, stream);
}

}  // namespace sdod

extern "C" {
SDOD_API int sdod_set_splitk_workspace(float* ws, size_t ws_bytes, unsigned int* counters, int n_counters) {
    sdod::SplitKWorkspace w;
    if (ws && counters && n_counters > 0) { w.ws = ws; w.ws_bytes = ws_bytes; w.counters = counters; w.n_counters = n_counters; }
    sdod::set_splitk_workspace(w);
    return sdod::kOk;
}
SDOD_API int sdod_set_gemm_timeline(unsigned long long* buf) {
    sdod::set_gemm_timeline(buf);
    return sdod::kOk;
}
SDOD_API int sdod_gemm_bf16(sdod_stream_t stream, const sdod_gemm_desc* d) {
    if (!d) return sdod::fail(sdod::kInvalidArgument, "sdod_gemm_bf16: desc is NULL");
    return sdod::gemm_bf16(static_cast<cudaStream_t>(stream), *d);
}
SDOD_API int sdod_conv3x3_bf16(sdod_stream_t stream, const sdod_conv_desc* d) {
    if (!d) return sdod::fail(sdod::kInvalidArgument, "sdod_conv3x3_bf16: desc is NULL");
    return sdod::conv3x3_bf16(static_cast<cudaStream_t>(stream), *d);
}
}
