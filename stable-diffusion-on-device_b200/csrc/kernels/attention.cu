// Fused flash-style attention for sm_100a (SpatialTransformer attn1 / attn2).
//
//   O = softmax(scale * Q K^T) V      per (batch*head), 128 query rows per CTA, KV tiles of 64 keys.
//
// Warp roles (320 threads):
//   warp 0     : TMA producer — Q once, then K / V^T tiles through an mbarrier ring
//   warp 1     : MMA issuer   — S = Q K^T and O += P V on tcgen05, accumulators in TMEM
//   warps 2..9 : softmax      — two threads per query row; thread `half` owns 32 of each tile's 64 keys and its own
//                O accumulator (PV runs as two K=32 MMAs), so the halves are independent online softmaxes in the exp2
//                domain (packed-fp32 FFMA2 + MUFU.EX2), each with a lazy O rescale in TMEM; P (bf16) goes into
//                128B-swizzled smem as the A operand of the PV MMAs; the halves are merged once at the end:
//                O = (O_0 w_0 + O_1 w_1) -> bf16 [B, Nq, heads*head_dim].
// TMEM columns: S at [0,128) (two 64-key buffers), O_0 at [128, 128+dv), O_1 at [128+dv, 128+2dv).  Everything is K-major + SWIZZLE_128B: K^T comes
// for free from the HEADS layout, V is stored transposed (HEADS_T) by the producing GEMM epilogue.
#include <cstdlib>
#include "../common.cuh"
#include "../host_common.h"
#include "../launch_count.h"
#include "launchers.h"
#include "sdod_kernels.h"

namespace sdod {

constexpr int kAttThreads = 320;   // warp 0 TMA, warp 1 MMA, warps 2..9 softmax (two threads per query row)
constexpr int kBQ = 128;    // query rows per CTA
constexpr int kBKV = 64;    // keys per tile: one thread holds a whole S row (64 fp32) in registers -> S is read from TMEM once

// MODE 0: P through shared memory (round 1).  MODE 1: P in TMEM (TS-form PV).  MODE 2 (head dims with a spare V^T row, i.e. d = 40):
// MODE 1 plus row sums on the tensor pipe — row DH of every V^T tile in shared memory is a row of ones (written once per CTA; the TMA box
// covers rows [0, DH) only), so column DH of the O accumulator is sum_k P[q,k] of exactly the rounded P that multiplied V, and the 16
// packed adds per tile disappear.  Tried on top and dropped (B200 r2, B8 x 8 heads x 4096^2, d = 40; profiles/r02_attention_notes.txt):
// ex2.approx.ftz.bf16x2 for the exponentials (ptxas splits it into two MUFU.EX2.BF16 plus a PRMT: no MUFU slots saved, precision lost);
// fetching S_{j+1} from TMEM in the middle of tile j's exponentials (444 us against 376: a warp then needs S_{j+1} half a tile earlier, and
// S_{j+1} is only issued once EVERY warp has handed over P_{j-1} — the slack between fast and slow warps halves).
template <int DH, int MODE = 1>
struct AttCfg {
    static constexpr bool PT = MODE >= 1;
    static constexpr bool kOnes = MODE >= 2 && MODE <= 5;     // ones row at V^T row DH
    // MODE 6 (d = 80): ONE S buffer + a separate P region.  S double buffer + two 80-column accumulators = 288 TMEM columns round up to 512, i.e.
    // one CTA (8 softmax warps) per SM; S is consumed at the START of a tile's softmax (one tcgen05.ld into registers), so a single buffer released
    // right after that load (s_free) still lets S_{j+1} run under tile j's exponentials — 64 (S) + 32 (P) + 160 (O) = 256 columns, two CTAs per SM.
    static constexpr bool SB = MODE == 6;
    // MODE 3 / 4 / 5: 2 / 1 / 3 of every 8 exponential pairs run on the FMA pipe (ex2_poly2) instead of MUFU — bit i of the mask = pair i & 7
    static constexpr unsigned kPolyMask = MODE == 3 ? 0x88u : (MODE == 4 ? 0x80u : (MODE == 5 ? 0xA4u : 0u));
    static constexpr int kVBoxRows = kOnes ? DH : ((DH + 15) / 16) * 16;
    static constexpr int kNK = (DH + 63) / 64;            // 64-wide d chunks of Q / K
    static constexpr int kKSteps = (DH + 15) / 16;        // UMMA K steps for S = Q K^T
    static constexpr int kDV = ((DH + 15) / 16) * 16;     // N of the PV MMA (rows of the V^T tile)
    static constexpr int kStages = SB ? 2 : (DH > 80 ? 3 : 4);    // (SB: two CTAs share the SM's shared memory)
    static constexpr int kQBytes = kNK * kBQ * 128;
    static constexpr int kQBufs = (DH <= 80 && !SB) ? 2 : 1;       // two Q buffers: a CTA that walks several query tiles (qpc > 1) loads tile i+1 while it works on tile i
    static constexpr int kKBytes = kNK * kBKV * 128;
    static constexpr int kVBytes = kVBoxRows * 128;       // one 64-key chunk: the TMA box, kVBoxRows rows of 128 B
    static constexpr int kVChunk = ((kDV * 128 + 1023) / 1024) * 1024;
    static constexpr int kStageBytes = kKBytes + kVChunk;
    static constexpr int kPBytes = kBQ * 128;             // one P buffer: 128 rows x 64 keys bf16
    static constexpr int kSCols = SB ? 96 : 128;          // S double buffer (P written over it) | one S buffer + the P region
    static constexpr int kTmemCols = (kSCols + 2 * kDV) <= 256 ? 256 : 512;   // + two O accumulators
    static constexpr int kCtasPerSm = (DH <= 40 || SB) ? 2 : 1;
    static constexpr int kSmemBytes = kQBufs * kQBytes + kStages * kStageBytes + (PT ? 0 : 2 * kPBytes) + 1024 + 256 + 4096;   // + row-max / row-sum exchange
};

SDOD_DEVICE float ex2(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}

// Pipeline (per CTA, 128 queries):  the MMA thread issues  S_0, S_1, PV_0, S_2, PV_1, ...  so QK^T of the next tile runs
// on the tensor pipe while the softmax warps work on the current one (S double-buffered in TMEM columns [0,64) / [64,128),
// P double-buffered in smem).  The O rescale is lazy: a row keeps its old reference max until the true max has grown by
// more than 2^8, so most tiles skip the TMEM round trip and never wait for the previous PV.
// exp2 on the FMA/ALU pipes (Cody-Waite split + degree-3 minimax polynomial, max rel err 7.7e-5 — well inside bf16 P).
// Round 1 (P through shared memory, shared-memory-bound) lost 20 % with 3 of every 8 exponentials moved here; with P in TMEM the d = 40 kernel is
// MUFU-bound (XU pipe 70 %) and moving ONE pair in eight (packed form below) gains 5.5 %; more than that and the issue slots bind instead.
SDOD_DEVICE float ex2_poly(float x) {
    x = fmaxf(x, -120.0f);
    const float t = x + 12582912.0f;                 // 1.5 * 2^23: round to nearest integer in the low mantissa bits
    const float f = x - (t - 12582912.0f);           // [-0.5, 0.5]
    float p = fmaf(0.05508868396282196f, f, 0.24260404706001282f);
    p = fmaf(p, f, 0.6932762265205383f);
    p = fmaf(p, f, 0.9999289512634277f);
    return __int_as_float(__float_as_int(p) + (__float_as_int(t) << 23));
}

// Two exponentials at once in packed fp32 (FADD2 / FFMA2): ~11 issue slots per pair and no MUFU slot.
SDOD_DEVICE void ex2_poly2(float x0, float x1, float& p0, float& p1) {
    const uint64_t x = pack_f32x2(fmaxf(x0, -120.0f), fmaxf(x1, -120.0f));
    const uint64_t t = add_f32x2(x, pack_f32x2(12582912.0f, 12582912.0f));
    const uint64_t r = add_f32x2(t, pack_f32x2(-12582912.0f, -12582912.0f));
    const uint64_t f = fma_f32x2(r, pack_f32x2(-1.0f, -1.0f), x);
    uint64_t p = fma_f32x2(pack_f32x2(0.05508868396282196f, 0.05508868396282196f), f, pack_f32x2(0.24260404706001282f, 0.24260404706001282f));
    p = fma_f32x2(p, f, pack_f32x2(0.6932762265205383f, 0.6932762265205383f));
    p = fma_f32x2(p, f, pack_f32x2(0.9999289512634277f, 0.9999289512634277f));
    float t0, t1, q0, q1;
    unpack_f32x2(t, t0, t1);
    unpack_f32x2(p, q0, q1);
    p0 = __int_as_float(__float_as_int(q0) + (__float_as_int(t0) << 23));
    p1 = __int_as_float(__float_as_int(q1) + (__float_as_int(t1) << 23));
}

// PT = true: P never touches shared memory.  Each softmax thread writes its 32 probabilities (bf16 pairs, 16 columns) over the first half of
// the S columns it has just read (tcgen05.st) and the PV MMAs take A from TMEM (tcgen05.mma TS form).  Per 128x64 tile that removes 16 KB of
// shared-memory stores, 16 KB of UMMA operand reads and the generic->async proxy fence; shared-memory bandwidth (~70 KB per tile at
// 128 B/clk) was what bound the d=40 kernel (profiles/r01_attention_notes.txt: neither removing the exponentials nor the TMEM reads helped).
// S_{j+2} overwrites the buffer that holds P_j: it is issued after PV_j and tcgen05.mma instructions execute in issue order, so the separate
// "S buffer drained" barrier goes away too (p_full(j) already says every thread has read S_j).
template <int DH, int MODE>
__global__ void __launch_bounds__(kAttThreads, AttCfg<DH, MODE>::kCtasPerSm) attention_kernel(const __grid_constant__ CUtensorMap tmQ,
                                                                 const __grid_constant__ CUtensorMap tmK,
                                                                 const __grid_constant__ CUtensorMap tmV, bf16* __restrict__ O,
                                                                 int heads, int Nq, int Nkv, float scale_log2, int causal, int qpc) {
    using Cfg = AttCfg<DH, MODE>;
    constexpr bool PT = Cfg::PT;
    constexpr bool SB = Cfg::SB;
    constexpr int STAGES = Cfg::kStages;
    static_assert(!Cfg::kOnes || (DH % 8 == 0 && Cfg::kDV >= DH + 8), "the ones row needs a spare 8-row group in the V^T tile");
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);   // 1024-B aligned shared-space pointer
    uint8_t* sQ = smem;
    uint8_t* sKV = sQ + (qpc > 1 ? 2 : 1) * Cfg::kQBytes;        // (the second Q buffer exists only in multi-tile launches: same footprint as before otherwise)
    uint8_t* sP = sKV + STAGES * Cfg::kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sP + (PT ? 0 : 2 * Cfg::kPBytes));
    uint64_t* q_full = bars;                      // [2]
    uint64_t* kv_full = bars + 2;                 // [STAGES]
    uint64_t* kv_empty = kv_full + STAGES;        // [STAGES]
    uint64_t* s_full = kv_empty + STAGES;         // [2]
    uint64_t* s_free = s_full + 2;                // [2]
    uint64_t* p_full = s_free + 2;                // [2]
    uint64_t* p_free = p_full + 2;                // [2]
    uint64_t* o_ready = p_free + 2;
    uint64_t* o_final = o_ready + 1;
    uint64_t* q_free = o_final + 1;               // [2] every S MMA that reads Q buffer i has retired
    uint64_t* o_free = q_free + 2;                // the epilogue has read both O accumulators (multi-tile CTAs)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_free + 1);
    static_assert(2 + 2 * STAGES + 8 + 2 + 2 + 1 < 32, "barrier block");
    float* xch = reinterpret_cast<float*>(bars + 32);          // [2 query-tile parities][max | sum][2 halves][128 rows] partner exchange

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // A CTA walks qpc consecutive query tiles of one (batch, head).  qpc > 1 is the cross-attention form (host: all keys fit the K/V ring, not
    // causal): K / V are loaded once and stay resident, Q is double-buffered, and the flat tile counter G = it * n_tiles + j drives the S / P
    // buffers and their barrier phases, so tile i+1's Q load and QK^T overlap tile i's softmax tail and epilogue.  With 77 keys a CTA's life was
    // one long latency chain (setup, TMEM allocation, three TMA round trips, two tiles, epilogue: 5.4 us per 128 queries at 2 CTAs/SM).
    const int bh = blockIdx.y;
    const int qt0 = blockIdx.x * qpc;
    const int n_qt = (Nq + kBQ - 1) / kBQ;
    const int n_it = min(qpc, n_qt - qt0);
    // causal (the CLIP text encoder): key k is visible to query q iff k <= q, so a query tile needs the key tiles up to its last row only
    const int n_tiles = causal ? (min(Nkv, qt0 * kBQ + kBQ) + kBKV - 1) / kBKV : (Nkv + kBKV - 1) / kBKV;
    const bool kv_resident = qpc > 1;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmQ); tma_prefetch_desc(&tmK); tma_prefetch_desc(&tmV);
    }
    if (warp == 1 && lane == 0) {
        for (int b = 0; b < 2; ++b) { mbar_init(&q_full[b], 1); mbar_init(&q_free[b], 1); }
        mbar_init(o_free, 256);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&s_full[b], 1); mbar_init(&s_free[b], 256); mbar_init(&p_full[b], 256); mbar_init(&p_free[b], 1); }
        mbar_init(o_ready, 1);
        mbar_init(o_final, 1);
        fence_mbar_init();
    }
    if (warp == 2) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
    if (Cfg::kOnes && warp >= 3) {
        // rows [DH, DH+8) of every stage's V^T tile are never written by TMA: row DH = ones (row sums on the tensor pipe), the rest zeros.
        // A row of equal 16-byte chunks is invariant under the 128-byte swizzle.
        for (int i = static_cast<int>(threadIdx.x) - 96; i < STAGES * 8 * 8; i += kAttThreads - 96) {
            const int st = i >> 6, r = (i >> 3) & 7, ch = i & 7;
            const uint32_t w = r == 0 ? 0x3F803F80u : 0u;      // bf16 1.0 pairs
            *reinterpret_cast<uint4*>(sKV + st * Cfg::kStageBytes + Cfg::kKBytes + (DH + r) * 128 + ch * 16) = make_uint4(w, w, w, w);
        }
        fence_proxy_async_smem();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
    const uint32_t tmem_O = tmem_base + Cfg::kSCols;
    griddep_wait();                    // PDL: the setup above overlapped the projection GEMM's tail
    griddep_launch();

    if (warp == 0) {
        if (lane == 0) {
            // ------------------------------------------------------------ TMA producer
            for (int it = 0; it < n_it; ++it) {
                const int qb = Cfg::kQBufs == 2 ? (it & 1) : 0;
                if (it >= 2) mbar_wait(&q_free[qb], ((it >> 1) - 1) & 1);      // (qpc > 1 implies two Q buffers)
                mbar_arrive_expect_tx(&q_full[qb], Cfg::kQBytes);
                for (int c = 0; c < Cfg::kNK; ++c)
                    tma_load_3d(sQ + qb * Cfg::kQBytes + c * (kBQ * 128), &tmQ, &q_full[qb], c * 64, (qt0 + it) * kBQ, bh);
                if (it > 0) continue;                                          // K / V resident for the whole walk
                for (int j = 0; j < n_tiles; ++j) {
                    const int st = j % STAGES;
                    const uint32_t ph = (j / STAGES) & 1;
                    mbar_wait(&kv_empty[st], ph ^ 1);
                    mbar_arrive_expect_tx(&kv_full[st], Cfg::kKBytes + Cfg::kVBytes);
                    uint8_t* sK = sKV + st * Cfg::kStageBytes;
                    uint8_t* sV = sK + Cfg::kKBytes;
                    for (int c = 0; c < Cfg::kNK; ++c) tma_load_3d(sK + c * (kBKV * 128), &tmK, &kv_full[st], c * 64, j * kBKV, bh);
                    tma_load_3d(sV, &tmV, &kv_full[st], j * kBKV, 0, bh);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            // ------------------------------------------------------------ MMA issuer
            constexpr uint32_t idesc_s = make_idesc_bf16(kBQ, kBKV);
            constexpr uint32_t idesc_o = make_idesc_bf16(kBQ, Cfg::kDV);
            const uint32_t p_addr = smem_u32(sP);
            if (qpc == 1) {
                // one query tile per CTA (self-attention): the round-2 loop, kept apart.  Routing self-attention through the general loop below cost
                // 15 % (B8 d=40: 350 -> 401 us; 471 us with two integer divisions per tile); taking the next tile's K / V wait before the P wait
                // changed nothing (365.5 vs 365.6 us with the switch compiled in) — the kernel is sensitive to code generation, not to that chain
                const uint32_t q_addr = smem_u32(sQ);
                mbar_wait(&q_full[0], 0);
                for (int j = 0; j <= n_tiles; ++j) {
                    if (j < n_tiles) {                                   // S_j = Q K_j^T -> S buffer j&1
                        const int st = j % STAGES;
                        mbar_wait(&kv_full[st], (j / STAGES) & 1);
                        if (!PT && j >= 2) mbar_wait(&s_free[j & 1], ((j >> 1) - 1) & 1);
                        if (SB && j >= 1) mbar_wait(&s_free[0], (j - 1) & 1);      // every softmax thread holds S_{j-1} in registers
                        tc_fence_after();
                        const uint32_t k_addr = smem_u32(sKV + st * Cfg::kStageBytes);
#pragma unroll
                        for (int k = 0; k < Cfg::kKSteps; ++k) {
                            const uint32_t qoff = (k >> 2) * (kBQ * 128) + (k & 3) * 32;
                            const uint32_t koff = (k >> 2) * (kBKV * 128) + (k & 3) * 32;
                            tc_mma_bf16(tmem_base + (SB ? 0 : (j & 1) * 64), make_kmajor_sw128_desc(q_addr + qoff), make_kmajor_sw128_desc(k_addr + koff), idesc_s, k != 0);
                        }
                        tc_commit(&s_full[SB ? 0 : (j & 1)]);
                    }
                    if (j >= 1) {                                        // O += P_i V_i for the previous tile
                        const int i = j - 1, st = i % STAGES;
                        mbar_wait(&p_full[SB ? 0 : (i & 1)], (SB ? i : (i >> 1)) & 1);
                        tc_fence_after();
                        const uint32_t v_addr = smem_u32(sKV + st * Cfg::kStageBytes) + Cfg::kKBytes;
                        const uint32_t pb = p_addr + (i & 1) * Cfg::kPBytes;
#pragma unroll
                        for (int k = 0; k < kBKV / 16; ++k) {  // keys [0,32) -> O_0, keys [32,64) -> O_1 (one accumulator per softmax thread of a row)
                            if (SB) tc_mma_bf16_ts(tmem_O + (k >> 1) * Cfg::kDV, tmem_base + 64 + (k >> 1) * 16 + (k & 1) * 8,
                                                   make_kmajor_sw128_desc(v_addr + k * 32), idesc_o, (i | (k & 1)) != 0);
                            else if (PT) tc_mma_bf16_ts(tmem_O + (k >> 1) * Cfg::kDV, tmem_base + (i & 1) * 64 + (k >> 1) * 32 + (k & 1) * 8,
                                                   make_kmajor_sw128_desc(v_addr + k * 32), idesc_o, (i | (k & 1)) != 0);
                            else tc_mma_bf16(tmem_O + (k >> 1) * Cfg::kDV, make_kmajor_sw128_desc(pb + k * 32), make_kmajor_sw128_desc(v_addr + k * 32), idesc_o,
                                             (i | (k & 1)) != 0);
                        }
                        tc_commit(&kv_empty[st]);
                        tc_commit(o_ready);
                    }
                }
                tc_commit(o_final);                                      // every PV has retired
            } else {
                const int total = n_it * n_tiles;
                int it = 0, j = 0, itp = 0, i = 0;                       // (query tile, key tile) of S_G and of PV_{G-1}: walked, not divided
                for (int G = 0; G <= total; ++G) {                       // flat tile counter over the CTA's query tiles
                    if (G < total) {                                     // S_G = Q K_j^T -> S buffer G&1
                        const int st = j % STAGES;
                        const int qb = Cfg::kQBufs == 2 ? (it & 1) : 0;
                        if (j == 0) mbar_wait(&q_full[qb], (Cfg::kQBufs == 2 ? (it >> 1) : it) & 1);
                        if (it == 0) mbar_wait(&kv_full[st], (j / STAGES) & 1);
                        if (!PT && G >= 2) mbar_wait(&s_free[G & 1], ((G >> 1) - 1) & 1);
                        tc_fence_after();
                        const uint32_t q_addr = smem_u32(sQ + qb * Cfg::kQBytes);
                        const uint32_t k_addr = smem_u32(sKV + st * Cfg::kStageBytes);
    #pragma unroll
                        for (int k = 0; k < Cfg::kKSteps; ++k) {
                            const uint32_t qoff = (k >> 2) * (kBQ * 128) + (k & 3) * 32;
                            const uint32_t koff = (k >> 2) * (kBKV * 128) + (k & 3) * 32;
                            tc_mma_bf16(tmem_base + (G & 1) * 64, make_kmajor_sw128_desc(q_addr + qoff), make_kmajor_sw128_desc(k_addr + koff), idesc_s, k != 0);
                        }
                        tc_commit(&s_full[G & 1]);
                        if (j == n_tiles - 1) tc_commit(&q_free[qb]);    // the last QK^T of this query tile: its Q buffer may be refilled
                        if (++j == n_tiles) { j = 0; ++it; }
                    }
                    if (G >= 1) {                                        // O += P_i V_i for the previous tile
                        const int Gp = G - 1, st = i % STAGES;
                        mbar_wait(&p_full[Gp & 1], (Gp >> 1) & 1);
                        if (i == 0 && itp > 0) mbar_wait(o_free, (itp - 1) & 1);   // the previous query tile's epilogue has read the accumulators
                        tc_fence_after();
                        const uint32_t v_addr = smem_u32(sKV + st * Cfg::kStageBytes) + Cfg::kKBytes;
                        const uint32_t pb = p_addr + (Gp & 1) * Cfg::kPBytes;
    #pragma unroll
                        for (int k = 0; k < kBKV / 16; ++k) {  // keys [0,32) -> O_0, keys [32,64) -> O_1 (one accumulator per softmax thread of a row)
                            if (PT) tc_mma_bf16_ts(tmem_O + (k >> 1) * Cfg::kDV, tmem_base + (Gp & 1) * 64 + (k >> 1) * 32 + (k & 1) * 8,
                                                   make_kmajor_sw128_desc(v_addr + k * 32), idesc_o, (i | (k & 1)) != 0);
                            else tc_mma_bf16(tmem_O + (k >> 1) * Cfg::kDV, make_kmajor_sw128_desc(pb + k * 32), make_kmajor_sw128_desc(v_addr + k * 32), idesc_o,
                                             (i | (k & 1)) != 0);
                        }
                        if (!kv_resident) tc_commit(&kv_empty[st]);
                        tc_commit(o_ready);
                        if (i == n_tiles - 1) tc_commit(o_final);        // every PV of this query tile has retired
                        if (++i == n_tiles) { i = 0; ++itp; }
                    }
                }
            }
        }
    } else {
        // ---------------------------------------------------------------- softmax / lazy correction / epilogue
        // Two threads per query row, and they never talk during the KV loop: thread `half` owns keys [32*half, 32*half+32) of
        // every tile AND its own accumulator O_half (PV is issued as two K=32 MMAs), so each keeps a private reference max /
        // row sum — no per-tile max exchange, no pair barrier.  The two partial results are merged once, in the epilogue:
        //   O = (O_0 2^(m_0-m) + O_1 2^(m_1-m)) / (l_0 2^(m_0-m) + l_1 2^(m_1-m)),  m = max(m_0, m_1).
        const int q = warp & 3;                       // TMEM lane quarter
        const int half = (warp - 2) >> 2;
        const int row = q * 32 + lane;
        const uint32_t lane_off = static_cast<uint32_t>(q * 32) << 16;
        const uint32_t tmem_mine = tmem_O + lane_off + half * Cfg::kDV;
        const int sw = row & 7;
        const uint64_t scale2 = pack_f32x2(scale_log2, scale_log2);
        for (int it = 0; it < n_it; ++it) {
        const int q0 = (qt0 + it) * kBQ;
        float m_run = -1.0e30f, l_run = 0.f;          // finite "minus infinity": a half with no valid key yet keeps p = 0, alpha = 1
        for (int j = 0; j < n_tiles; ++j) {
            const int G = it * n_tiles + j;           // flat tile counter: S / P buffer and barrier phases
            const int b = SB ? 0 : (G & 1), u = SB ? G : (G >> 1);
            const int kv_valid = min(kBKV, Nkv - j * kBKV) - half * 32;      // valid keys among this thread's 32
            mbar_wait(&s_full[b], u & 1);
            tc_fence_after();
            uint32_t sv[32];
            tmem_ld32(tmem_base + lane_off + b * 64 + half * 32, sv);
            tmem_ld_wait();
            if (!PT || SB) {
                tc_fence_before();
                mbar_arrive(&s_free[b]);             // S buffer b may be overwritten by S_{j+2} (SB: by S_{j+1})
            }
            if (kv_valid < 32) {
#pragma unroll
                for (int i = 0; i < 32; ++i)
                    if (i >= kv_valid) sv[i] = 0xff800000u;          // -inf
            }
            if (causal) {
                const int visible = q0 + row - (j * kBKV + half * 32);      // keys [0, visible] of this thread's 32 may be attended
                if (visible < 31) {
#pragma unroll
                    for (int i = 0; i < 32; ++i)
                        if (i > visible) sv[i] = 0xff800000u;
                }
            }
            float mx4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
            for (int i = 0; i < 32; i += 4) {
#pragma unroll
                for (int c = 0; c < 4; ++c) mx4[c] = fmaxf(mx4[c], __uint_as_float(sv[i + c]));
            }
            const float mx = fmaxf(fmaxf(mx4[0], mx4[1]), fmaxf(mx4[2], mx4[3])) * scale_log2;   // scale_log2 > 0
            // lazy reference max: only move it when the true max grew by more than 2^8 (P stays <= 256, exact in the ratio O/l)
            const bool need = (mx > m_run + 8.0f);
            const float m_new = need ? mx : m_run;
            const float alpha = ex2(m_run - m_new);   // 1 when unchanged, 0 on the first real tile
            const uint64_t negm2 = pack_f32x2(-m_new, -m_new);
            uint64_t ls2[2] = {0ull, 0ull};
            uint32_t pk[16];
            if (Cfg::kOnes) {
#pragma unroll
                for (int i = 0; i < 16; ++i) {
                    float x0, x1;
                    unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), scale2, negm2), x0, x1);
                    if ((Cfg::kPolyMask >> (i & 7)) & 1u) {
                        float p0, p1;
                        ex2_poly2(x0, x1, p0, p1);
                        pk[i] = pack_bf16x2(p0, p1);
                    } else {
                        pk[i] = pack_bf16x2(ex2(x0), ex2(x1));
                    }
                }
            } else
#pragma unroll
            for (int i = 0; i < 16; ++i) {
                // x = s * scale - m for two keys in one FFMA2 (sm_100 packed fp32), then MUFU.EX2 each
                float x0, x1;
                unpack_f32x2(fma_f32x2(pack_f32x2(__uint_as_float(sv[2 * i]), __uint_as_float(sv[2 * i + 1])), scale2, negm2), x0, x1);
                // (measured r1: moving 3/8 of these onto ex2_poly made the kernel 20 % slower — profiles/r01_attention_notes.txt)
                const float p0 = ex2(x0);
                const float p1 = ex2(x1);
                ls2[i & 1] = add_f32x2(ls2[i & 1], pack_f32x2(p0, p1));
                pk[i] = pack_bf16x2(p0, p1);
            }
            if (!Cfg::kOnes) {
                float a0, a1, a2, a3;
                unpack_f32x2(ls2[0], a0, a1);
                unpack_f32x2(ls2[1], a2, a3);
                l_run = fmaf(l_run, alpha, (a0 + a1) + (a2 + a3));
            }
            m_run = m_new;
            // P buffer b is free: the MMA thread issues S_j after PV_{j-2}, and a tcgen05.commit covers every MMA issued before it, so
            // having seen s_full for tile j already implies PV_{j-2} has consumed this buffer.  (A satisfied mbarrier wait still costs
            // ~200 clk here — clock64 phase timing, profiles/r01_attention_notes.txt — so the loop keeps exactly one per tile.)
            if (SB) {
                // the one P region is free once PV_{j-1} has retired (it was issued as soon as the slowest warp handed over P_{j-1})
                if (j > 0) { mbar_wait(o_ready, (G - 1) & 1); tc_fence_after(); }
                tmem_st16(tmem_base + lane_off + 64 + half * 16, pk);
            } else if (PT) {
                // keys (2i, 2i+1) of this thread's 32 -> column i of its own S half: lane = query row, 32-bit column = two consecutive K elements
                tmem_st16(tmem_base + lane_off + b * 64 + half * 32, pk);
            } else {
                uint8_t* prow = sP + b * Cfg::kPBytes + row * 128;
#pragma unroll
                for (int un = 0; un < 4; ++un)
                    *reinterpret_cast<uint4*>(prow + (((half * 4 + un) ^ sw) << 4)) = make_uint4(pk[4 * un], pk[4 * un + 1], pk[4 * un + 2], pk[4 * un + 3]);
            }
            if (j > 0 && __any_sync(0xffffffffu, need)) {            // rare: this warp's accumulator must be rescaled
                // PV_{j-1} must have completed.  s_full(j) above proves PV_{j-2} has, so o_ready is either in phase j-1 (pending) or
                // already past it: the parity test is unambiguous even though this wait is not taken every tile.
                mbar_wait(o_ready, (G - 1) & 1);
                tc_fence_after();
#pragma unroll 1
                for (int c = 0; c < Cfg::kDV / 16; ++c) {
                    uint32_t o[16];
                    tmem_ld16(tmem_mine + c * 16, o);
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
                    tmem_st16(tmem_mine + c * 16, o);
                }
            }
            if (PT) tmem_st_wait();              // P (and a rescaled O) have landed in TMEM
            else if (j > 0) tmem_st_wait();   // (MODE 0: only a rescale stores to TMEM)
            tc_fence_before();
            if (!PT) fence_proxy_async_smem();   // P (generic-proxy stores) -> visible to the UMMA async proxy
            mbar_arrive(&p_full[b]);
        }
        // epilogue: merge the two halves of each row
        float* xc = xch + (it & 1) * 512;            // double-buffered across query tiles: the partner may still be reading the previous one
        xc[half * 128 + row] = m_run;
        xc[(2 + half) * 128 + row] = l_run;
        asm volatile("bar.sync %0, 64;" ::"r"(q + 1) : "memory");
        const float m_o = xc[(half ^ 1) * 128 + row];
        float l_o = xc[(2 + (half ^ 1)) * 128 + row];
        const float m_all = fmaxf(m_run, m_o);
        const float f_me = ex2(m_run - m_all), f_ot = ex2(m_o - m_all);
        // All PVs retired: a barrier of its own, committed once after the last PV.  (o_ready flips every tile and is no longer waited in
        // lockstep, so its parity cannot tell "all done" from "two behind" here.)
        mbar_wait(o_final, it & 1);
        tc_fence_after();
        if (Cfg::kOnes) {                                   // the row sums are column DH of the two accumulators
            uint32_t a[8], b[8];
            tmem_ld8(tmem_O + lane_off + half * Cfg::kDV + DH, a);
            tmem_ld8(tmem_O + lane_off + (half ^ 1) * Cfg::kDV + DH, b);
            tmem_ld_wait();
            l_run = __uint_as_float(a[0]);
            l_o = __uint_as_float(b[0]);
        }
        const float inv_l = 1.0f / fmaf(l_run, f_me, l_o * f_ot);
        const float w_me = f_me * inv_l, w_ot = f_ot * inv_l;
        const int bb = bh / heads, h = bh - bb * heads;
        const int qi = q0 + row;
        bf16* orow = O + (static_cast<long long>(bb) * Nq + qi) * (heads * DH) + h * DH;
        const uint32_t tmem_other = tmem_O + lane_off + (half ^ 1) * Cfg::kDV;
#pragma unroll 1
        for (int c = half; c < Cfg::kDV / 16; c += 2) {
            uint32_t o[16], o2[16];
            tmem_ld16(tmem_mine + c * 16, o);
            tmem_ld16(tmem_other + c * 16, o2);
            tmem_ld_wait();
            if (qi < Nq) {
#pragma unroll
                for (int h8 = 0; h8 < 2; ++h8) {
                    const int d0 = c * 16 + h8 * 8;
                    if (d0 < DH) {     // DH % 8 == 0
                        float r[8];
#pragma unroll
                        for (int i = 0; i < 8; ++i) r[i] = fmaf(__uint_as_float(o[h8 * 8 + i]), w_me, __uint_as_float(o2[h8 * 8 + i]) * w_ot);
                        uint4 w;
                        w.x = pack_bf16x2(r[0], r[1]); w.y = pack_bf16x2(r[2], r[3]);
                        w.z = pack_bf16x2(r[4], r[5]); w.w = pack_bf16x2(r[6], r[7]);
                        *reinterpret_cast<uint4*>(orow + d0) = w;
                    }
                }
            }
        }
        tc_fence_before();
        if (it + 1 < n_it) mbar_arrive(o_free);      // both accumulators read: the next query tile's PV may overwrite them
        }   // query tiles
    }
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        tmem_dealloc<Cfg::kTmemCols>(tmem_base);
    }
}

static int attention_mode(int DH) {
    // SDOD_ATTN_MODE = 0 | 1 | 2 overrides the default (A/B measurements); MODE 2 exists for head dims with a spare V^T row group only
    static const int env = [] { const char* e = std::getenv("SDOD_ATTN_MODE"); return e ? std::atoi(e) : -1; }();
    // d = 40: MODE 4 = MODE 2 with one exponential pair in eight on the FMA pipe (B200 r2, B8 x 8 heads x 4096^2: 370.7 us as MODE 2, 350.2 as MODE 4,
    // 354.3 with two pairs in eight (MODE 3), 372.8 with three (MODE 5); B2: 118.8 -> 108.5 us)
    const int best = DH == 40 ? 4 : 1;
    return env < 0 ? best : (env > (DH == 40 ? 5 : 1) ? best : env);
}

template <int DH>
static int prepare_attention(AttnLaunch* out, const void* Qh, const void* Kh, const void* Vt, void* O, int B, int heads, int Nq,
                             int Nkv, int dpad, int kv_pad, float scale) {
    using Cfg = AttCfg<DH>;
    if (dpad != Cfg::kNK * 64) return fail(kInvalidArgument, "attention: dpad must be 64*ceil(head_dim/64)");
    if (kv_pad % 8 != 0 || kv_pad < Nkv) return fail(kInvalidArgument, "attention: kv_pad must be a multiple of 8 and >= Nkv");
    const int BH = B * heads;
    CUtensorMap& tmQ = out->tmQ;
    CUtensorMap& tmK = out->tmK;
    CUtensorMap& tmV = out->tmV;
    {
        uint64_t dims[3] = {static_cast<uint64_t>(dpad), static_cast<uint64_t>(Nq), static_cast<uint64_t>(BH)};
        uint64_t strides[2] = {static_cast<uint64_t>(dpad) * 2, static_cast<uint64_t>(Nq) * dpad * 2};
        uint32_t box[3] = {64, kBQ, 1};
        SDOD_TRY(encode_tmap_bf16(&tmQ, Qh, 3, dims, strides, box, true));
    }
    {
        uint64_t dims[3] = {static_cast<uint64_t>(dpad), static_cast<uint64_t>(Nkv), static_cast<uint64_t>(BH)};
        uint64_t strides[2] = {static_cast<uint64_t>(dpad) * 2, static_cast<uint64_t>(Nkv) * dpad * 2};
        uint32_t box[3] = {64, kBKV, 1};
        SDOD_TRY(encode_tmap_bf16(&tmK, Kh, 3, dims, strides, box, true));
    }
    {
        uint64_t dims[3] = {static_cast<uint64_t>(kv_pad), static_cast<uint64_t>(Cfg::kDV), static_cast<uint64_t>(BH)};
        uint64_t strides[2] = {static_cast<uint64_t>(kv_pad) * 2, static_cast<uint64_t>(Cfg::kDV) * kv_pad * 2};
        uint32_t box[3] = {64, static_cast<uint32_t>(attention_mode(DH) >= 2 ? DH : Cfg::kDV), 1};
        SDOD_TRY(encode_tmap_bf16(&tmV, Vt, 3, dims, strides, box, true));
    }
    out->O = O; out->heads = heads; out->Nq = Nq; out->Nkv = Nkv; out->head_dim = DH; out->BH = BH; out->causal = 0;
    out->scale_log2 = scale * 1.4426950408889634f;
    return kOk;
}

template <int DH, int MODE>
static int launch_attention_v(const AttnLaunch& a, cudaStream_t stream) {
    using Cfg = AttCfg<DH, MODE>;
    static bool configured = false;
    if (!configured) {
        SDOD_TRY(check_cuda(cudaFuncSetAttribute(attention_kernel<DH, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes),
                            "cudaFuncSetAttribute(attention)"));
        configured = true;
    }
    // Query tiles per CTA: cross-attention shapes (every key tile fits the K/V ring, so K / V stay resident) walk several query tiles per CTA
    // as long as the grid still holds >= 4 CTAs per resident slot.  SDOD_ATTN_QPC overrides (1 = one tile per CTA as in round 1).
    const int n_qt = (a.Nq + kBQ - 1) / kBQ, n_kt = (a.Nkv + kBKV - 1) / kBKV;
    int qpc = 1;
    if (Cfg::kQBufs == 2 && !Cfg::SB && !a.causal && n_kt <= Cfg::kStages && n_qt > 1) {
        static const int env = [] { const char* e = std::getenv("SDOD_ATTN_QPC"); return e ? std::atoi(e) : 0; }();
        const long long slots = static_cast<long long>(device_sm_count()) * (DH <= 40 ? 2 : 1);
        for (int c = 8; c >= 2 && qpc == 1; c >>= 1)
            if (static_cast<long long>(a.BH) * ((n_qt + c - 1) / c) >= 4 * slots) qpc = c;
        if (env >= 1) qpc = env > n_qt ? n_qt : env;
    }
    dim3 grid((n_qt + qpc - 1) / qpc, a.BH);
    const size_t smem = Cfg::kSmemBytes - (qpc > 1 ? 0 : (Cfg::kQBufs - 1) * Cfg::kQBytes);
    SDOD_TRY(check_cuda(launch_pdl(attention_kernel<DH, MODE>, grid, dim3(kAttThreads), smem, stream, a.tmQ, a.tmK, a.tmV,
                                   static_cast<bf16*>(a.O), a.heads, a.Nq, a.Nkv, a.scale_log2, a.causal, qpc), "launch attention_kernel"));
    count_launch();
    return check_launch("attention_kernel");
}

template <int DH>
static int launch_attention(const AttnLaunch& a, cudaStream_t stream) {
    const int mode = attention_mode(DH);
    if constexpr (DH == 40) {
        if (mode == 2) return launch_attention_v<DH, 2>(a, stream);
        if (mode == 3) return launch_attention_v<DH, 3>(a, stream);
        if (mode == 4) return launch_attention_v<DH, 4>(a, stream);
        if (mode == 5) return launch_attention_v<DH, 5>(a, stream);
    }
    if constexpr (DH == 80) {
        // self-attention (more key tiles than the ring holds): the single-S-buffer layout that fits two CTAs per SM; cross-attention keeps MODE 1
        // and walks several query tiles per CTA instead
        static const int sb = [] { const char* e = std::getenv("SDOD_ATTN_SB"); return e ? std::atoi(e) : 1; }();
        // (multi-wave grids only: B32 x 8 heads x 1024^2 222.9 -> 197.4 us; a single-wave grid has one CTA per SM either way and loses to the
        // shallower K/V ring and the extra waits: batch 2 20.8 -> 29.0 us)
        const long long ctas = static_cast<long long>(a.BH) * ((a.Nq + kBQ - 1) / kBQ);
        if (mode && sb && (a.Nkv + kBKV - 1) / kBKV > AttCfg<DH, 1>::kStages && ctas > device_sm_count()) return launch_attention_v<DH, 6>(a, stream);
    }
    return mode ? launch_attention_v<DH, 1>(a, stream) : launch_attention_v<DH, 0>(a, stream);
}

int attention_prepare(AttnLaunch* out, const void* Qh, const void* Kh, const void* Vt, void* O, int B, int heads, int Nq, int Nkv,
                      int head_dim, int dpad, int kv_pad, float scale) {
    if (!Qh || !Kh || !Vt || !O) return fail(kInvalidArgument, "attention: NULL tensor");
    if (B <= 0 || heads <= 0 || Nq <= 0 || Nkv <= 0) return fail(kInvalidArgument, "attention: non-positive extent");
    switch (head_dim) {
        case 40: return prepare_attention<40>(out, Qh, Kh, Vt, O, B, heads, Nq, Nkv, dpad, kv_pad, scale);
        case 64: return prepare_attention<64>(out, Qh, Kh, Vt, O, B, heads, Nq, Nkv, dpad, kv_pad, scale);
        case 80: return prepare_attention<80>(out, Qh, Kh, Vt, O, B, heads, Nq, Nkv, dpad, kv_pad, scale);
        case 160: return prepare_attention<160>(out, Qh, Kh, Vt, O, B, heads, Nq, Nkv, dpad, kv_pad, scale);
    }
    return fail(kUnsupported, "attention: head_dim must be one of 40, 64, 80, 160 (SD v1.x)");
}

int attention_launch(const AttnLaunch& a, cudaStream_t stream) {
    switch (a.head_dim) {
        case 40: return launch_attention<40>(a, stream);
        case 64: return launch_attention<64>(a, stream);
        case 80: return launch_attention<80>(a, stream);
        case 160: return launch_attention<160>(a, stream);
    }
    return fail(kUnsupported, "attention: bad plan");
}

int attention_bf16(cudaStream_t stream, const void* Qh, const void* Kh, const void* Vt, void* O, int B, int heads, int Nq, int Nkv,
                   int head_dim, int dpad, int kv_pad, float scale, int causal) {
    AttnLaunch a;
    SDOD_TRY(attention_prepare(&a, Qh, Kh, Vt, O, B, heads, Nq, Nkv, head_dim, dpad, kv_pad, scale));
    a.causal = causal ? 1 : 0;
    return attention_launch(a, stream);
}

}  // namespace sdod

extern "C" SDOD_API int sdod_attention_bf16(sdod_stream_t stream, const void* Qh, const void* Kh, const void* Vt, void* O, int B, int heads,
                                            int Nq, int Nkv, int head_dim, int dpad, int kv_pad, float scale) {
    return sdod::attention_bf16(static_cast<cudaStream_t>(stream), Qh, Kh, Vt, O, B, heads, Nq, Nkv, head_dim, dpad, kv_pad, scale, 0);
}
extern "C" SDOD_API int sdod_attention_causal_bf16(sdod_stream_t stream, const void* Qh, const void* Kh, const void* Vt, void* O, int B, int heads,
                                                   int Nq, int Nkv, int head_dim, int dpad, int kv_pad, float scale) {
    return sdod::attention_bf16(static_cast<cudaStream_t>(stream), Qh, Kh, Vt, O, B, heads, Nq, Nkv, head_dim, dpad, kv_pad, scale, 1);
}
