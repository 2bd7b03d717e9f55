// Prepared launches: tensor maps + parameters are encoded once (plan build time) and replayed
// every step, so the hot loop does no host-side encoding and is CUDA-graph capturable.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include "sdod_kernels.h"

namespace sdod {

struct MainloopParams {
    int M, N;          // logical extents (masking)
    int k_blocks;      // K / 64
    int conv;          // 0: A is a 3-D map (K, M, batch); 1: A is a 4-D map (C, W, H, B)
    int cin_blocks;    // conv: Cin / 64
    int H, W;          // conv: output spatial size
    int bw, bh, bb;    // conv: TMA box (pixels); bw*bh*bb == 128
    int w_batched;     // W has a batch dimension
    int split;         // split-K factor (grid.z when batch == 1); 1 = off
    int kb_per_split;  // K blocks per split
    float* ws;         // split-K partial tiles: [tile][split][BN/16][128][16] fp32
    unsigned int* counters;   // one ticket per output tile, self-resetting
    int tma_epi;       // 0: LSU epilogue; 1: TMA-store epilogue; 2: TMA-store + residual TMA-loaded and added in place
    int c_bytes;       // output element size for the TMA epilogue (2 | 4)
    int split_cluster; // 1: the `split` CTAs of a tile form a thread-block cluster (1,1,split) and fold the partials in-kernel after a
                       //    cluster barrier; 0: partials are folded by splitk_reduce_kernel (a second launch)
    int kb_a2;         // first K block served by the second A operand (tmA2: plain [K2, M] rows, the K-concatenated skip / concat
                       //    source); >= k_blocks when there is none
    int streamk;       // persistent variant only: CTAs take equal contiguous shares of the (tile, K block) space (see WorkWalk); partial tiles go
                       //    through `ws` (one slot per CTA), `counters` are the per-CTA "partial published" flags
    int ln_fuse;       // 1: LayerNorm of the output rows fused into the epilogue (cluster over the N tiles, see the kernel); tmC2 = bf16 LN output
    unsigned long long* tlog; // profiling hook (tools/gemm_timeline.py): per-CTA globaltimer stamps of the kernel's phases, 16 slots per CTA; NULL = off
    const char* pf_ptr;       // L2 prefetch of the NEXT layer's weights (constant data, issued before griddepcontrol.wait); NULL = none
    long long pf_bytes;
    int b_resident;    // persistent variant, k_blocks == ring stages and grid % n_tiles == 0: a CTA keeps one N tile for its whole walk and stage s always
                       //    holds K block s, so the weight tile is loaded once and only the A rows are streamed (half the L2 -> SM operand bytes)
    int epi_groups;    // persistent variant: 2 = two epilogue warp groups on alternate tiles / TMEM accumulators (see the kernel); else one group of 16 warps
    int pair_pdl;      // CTA-pair variant: take part in programmatic dependent launch (experiment switch, default 0)
    int pair_relaxed;  // CTA-pair variant: setup / exit cluster barriers without release/acquire memory ordering (0 = the round-1 form, A/B switch)
    int tmem_accs;     // persistent variant: TMEM accumulators in rotation (0 = all that fit: 4 x 128 / 3 x 160 columns; 2 = the round-2 double buffer)
    int geglu_tanh;    // GEGLU TMA epilogue: gate GELU in tanh form (1 MUFU + 5 packed ops instead of 2 MUFU + 14; see common.cuh)
    int cstride;       // conv stride (1 | 2): tile pixel (x, y) reads input (cstride*x + kx - 1, cstride*y + ky - 1); H, W, bw, bh are OUTPUT extents
    int up2;           // conv on the nearest-2x-upsampled image in sub-pixel form: blockIdx.z = output parity (py, px); 2x2 taps on the SOURCE image
                       //    with pre-summed weights (tmW batch = parity), output rows scattered through a 5-D tensor map (tmC)
    int k_rot;         // K-loop start rotation per M tile (in K blocks); 0 = every tile starts at block 0
    int n_tiles, m_tiles, tiles_total;   // persistent scheduling: tile t -> (t % n_tiles, (t / n_tiles) % m_tiles, t / (n_tiles*m_tiles))
};

struct GemmLaunch {
    CUtensorMap tmA, tmW, tmC, tmR, tmC2;   // tmC2: V^T boxes of the head-layout TMA epilogue
    CUtensorMap tmA2;                       // second A operand (K-concatenated after tmA's blocks); zeroed when unused
    const void* w_ptr;                      // this launch's weight matrix and its size: the previous launch prefetches it into L2
    long long w_bytes;
    MainloopParams mp;
    sdod_epilogue ep;
    int bn, m_tiles, n_tiles, batch;
    int pair;   // 1: launched as 2-CTA clusters running tcgen05 cta_group::2
    int persist;   // 1: one CTA per SM walks the tile list (TMEM double-buffered accumulator)
    int sk_grid;   // stream-K: number of persistent CTAs (0 = off)
};

struct AttnLaunch {
    CUtensorMap tmQ, tmK, tmV;
    void* O;
    int heads, Nq, Nkv, head_dim, BH;
    float scale_log2;
    int causal;      // 1: key k is visible to query q iff k <= q (CLIP text encoder); set after attention_prepare
};

int pick_block_n(int M, int N, int batch, int act);
int pick_block_n_k(int M, int N, int batch, int act, int k_blocks);
// Split-K scratch shared by all GEMM/conv launches on one stream (launches are serialised by stream order).
struct SplitKWorkspace { float* ws = nullptr; size_t ws_bytes = 0; unsigned int* counters = nullptr; int n_counters = 0; };
void set_splitk_workspace(const SplitKWorkspace& w);     // thread-local "current" workspace used by *_prepare
void set_gemm_timeline(unsigned long long* buf);          // thread-local: launches prepared afterwards record their phase stamps into buf
int gemm_prepare(const sdod_gemm_desc& d, GemmLaunch* out);
int conv3x3_prepare(const sdod_conv_desc& d, GemmLaunch* out);
int gemm_launch(const GemmLaunch& g, cudaStream_t stream);
int gemm_bf16(cudaStream_t stream, const sdod_gemm_desc& d);
int conv3x3_bf16(cudaStream_t stream, const sdod_conv_desc& d);

int attention_prepare(AttnLaunch* out, const void* Qh, const void* Kh, const void* Vt, void* O, int B, int heads, int Nq, int Nkv,
                      int head_dim, int dpad, int kv_pad, float scale);
int attention_launch(const AttnLaunch& a, cudaStream_t stream);
int attention_bf16(cudaStream_t stream, const void* Qh, const void* Kh, const void* Vt, void* O, int B, int heads, int Nq, int Nkv,
                   int head_dim, int dpad, int kv_pad, float scale, int causal);

int group_norm(cudaStream_t stream, const void* x, void* y, const float* weight, const float* bias, const float* add_nc, int N, int C,
               int HW, int G, float eps, int dtype, int layout, int fuse_silu, void* ws, size_t ws_bytes);

int group_norm_nhwc(cudaStream_t stream, const void* x, int in_dtype, void* y, int out_dtype, const float* weight, const float* bias,
                    const float* add_nc, int N, int C, int HW, int G, float eps, int fuse_silu, void* ws, size_t ws_bytes);

// Single-launch NHWC GroupNorm over the channel concatenation [xa | xb] (xb may be NULL) with an optional bf16 copy of the raw
// concatenation; only for L2-resident tensors (group_norm_fused_eligible).
bool group_norm_fused_eligible(int N, int Ca, int Cb, int HW, int G, int in_dtype);
int group_norm_nhwc2(cudaStream_t stream, const void* xa, int Ca, const void* xb, int Cb, int in_dtype, void* y, int out_dtype, void* raw_bf16,
                     const float* weight, const float* bias, int N, int HW, int G, float eps, int fuse_silu, void* ws, size_t ws_bytes);

}  // namespace sdod
