/*
 * sdod_kernels.h — flat C ABI of libsdod_b200.so, kernel level.
 *
 * Plain pointers and sizes only (device pointers unless stated otherwise); every call
 * enqueues hand-written sm_100a kernels on `stream` and returns 0 on success or a
 * negative status (sdod_last_error() holds the message).  There is no CPU fallback:
 * without a CUDA device every compute entry point returns an error.
 *
 * Each entry point cites the reference interface it replaces (paths relative to the
 * reference repo vaenyr/stable-diffusion-on-device).
 */
#ifndef SDOD_KERNELS_H
#define SDOD_KERNELS_H

#include <stddef.h>
#include <stdint.h>

#ifndef SDOD_API
#define SDOD_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

typedef void* sdod_stream_t; /* cudaStream_t */

enum sdod_dtype { SDOD_F32 = 0, SDOD_BF16 = 1 };
enum sdod_layout { SDOD_NCHW = 0, SDOD_NHWC = 1 };
enum sdod_act { SDOD_ACT_NONE = 0, SDOD_ACT_SILU = 1, SDOD_ACT_GELU = 2, SDOD_ACT_GEGLU = 3, SDOD_ACT_QUICK_GELU = 4 /* x*sigmoid(1.702x): CLIP */ };
/* GEMM/conv output placement */
enum sdod_out_mode {
    SDOD_OUT_BF16 = 0,      /* C[m*ldc + n] bf16                                                     */
    SDOD_OUT_F32 = 1,       /* C[m*ldc + n] fp32                                                     */
    SDOD_OUT_HEADS = 2,     /* bf16 [B*heads, tokens, dpad]  : attention Q / K operand layout        */
    SDOD_OUT_HEADS_T = 3,   /* bf16 [B*heads, dpad, tok_pad] : attention V^T operand layout          */
    SDOD_OUT_QKV = 4        /* N = 3*heads*head_dim: Q,K -> HEADS (C, C2), V -> HEADS_T (C3)         */
};

SDOD_API const char* sdod_last_error(void);
SDOD_API int sdod_abi_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
SDOD_API unsigned long long sdod_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * GroupNorm (+SiLU) (+per-(n,c) additive term, i.e. the timestep-embedding add).
 * Replaces: sdod/efficient_gn.py:9-12 (EfficientGNFun.forward == F.group_norm), :29-30, :61-70 and
 * the op contract of csrc/sdod_ops/config/group_norm.xml:16-107 / group_norm.json:6-21
 * (inputs input/weight/bias, attrs num_groups, eps; NCHW per the XML, NHWC per the UDO JSON).
 *   y = act( ((x + add[n,c]) - mu_g) * rsqrt(var_g + eps) * weight[c] + bias[c] )
 * x,y: [N,C,HW] (NCHW) or [N,HW,C] (NHWC), dtype f32 or bf16; weight/bias/add fp32 or NULL.
 * workspace: >= sdod_group_norm_workspace() bytes (NHWC path only; may be NULL for NCHW).
 * ------------------------------------------------------------------------------------------- */
SDOD_API size_t sdod_group_norm_workspace(int N, int C, int HW, int num_groups, int layout);
SDOD_API int sdod_group_norm(sdod_stream_t stream, const void* x, void* y, const float* weight, const float* bias,
                             const float* add_nc, int N, int C, int HW, int num_groups, float eps, int dtype,
                             int layout, int fuse_silu, void* workspace, size_t workspace_bytes);
/* NHWC only: input and output dtypes may differ (fp32 residual stream in, bf16 GEMM operand out). */
SDOD_API int sdod_group_norm_nhwc(sdod_stream_t stream, const void* x, int in_dtype, void* y, int out_dtype, const float* weight,
                                  const float* bias, const float* add_nc, int N, int C, int HW, int num_groups, float eps,
                                  int fuse_silu, void* workspace, size_t workspace_bytes);

/* Single-launch NHWC form for L2-resident tensors (the batch-2 UNet step): GroupNorm over the channel concatenation
 * x = [xa | xb] (xa [N,HW,Ca], xb [N,HW,Cb] or NULL; Ca % 8 == 0) -> y [N,HW,Ca+Cb]; raw_bf16 (optional) receives the bf16
 * copy of the un-normalised concatenation (the operand of the ResBlock's 1x1 skip_connection).  One cooperative grid
 * (<= one CTA per SM) computes the statistics and applies them.  sdod_group_norm_nhwc2_supported() tells whether a
 * shape qualifies (<= 48 MB input, N <= 128, C % 8 == 0); sdod_group_norm_nhwc() uses this form automatically when it does. */
SDOD_API int sdod_group_norm_nhwc2(sdod_stream_t stream, const void* xa, int Ca, const void* xb, int Cb, int in_dtype, void* y,
                                   int out_dtype, void* raw_bf16, const float* weight, const float* bias, int N, int HW,
                                   int num_groups, float eps, int fuse_silu, void* workspace, size_t workspace_bytes);
SDOD_API int sdod_group_norm_nhwc2_supported(int N, int Ca, int Cb, int HW, int num_groups, int in_dtype);

/* LayerNorm over the last dim (SpatialTransformer norm1/2/3). x [rows, width] f32|bf16 -> y bf16. */
SDOD_API int sdod_layer_norm(sdod_stream_t stream, const void* x, int in_dtype, void* y, const float* weight, const float* bias,
                             int rows, int width, float eps);

/* ---------------------------------------------------------------------------------------------
 * Fused classifier-free-guidance combine + eps->x0 + DPM-Solver++(2M) update (one kernel).
 * Replaces: csrc/libsdod/src/context.cpp:359-378 (CFG via QnnTensor::get_data scale/accum,
 * qnn_context.cpp:1065-1081) and DPMSolver::update, dpm_solver.cpp:136-181.
 *   e  = (g == 1) ? eps_c : g*eps_c + (1-g)*eps_u          (separate roundings, as the host loops)
 *   y0 = (x + (-sigma_s)*e) / alpha_s
 *   x  = c_x*x  [+ c_prev*y_prev]  + c_y0*y0 ;  y_prev = y0
 * x, y_prev fp32 (updated in place); eps_* f32 or bf16; x_copy optional second fp32 destination for
 * the new x (the uncond batch slot of the next UNet input).  order 1 ignores y_prev on input.
 * Bit-exact vs the reference in fp32.
 * ------------------------------------------------------------------------------------------- */
SDOD_API int sdod_cfg_dpm_step(sdod_stream_t stream, float* x, float* y_prev, const void* eps_c, const void* eps_u,
                               int eps_dtype, size_t n, float guidance, float sigma_s, float alpha_s, float c_x,
                               float c_prev, float c_y0, int order, float* x_copy);

/* CFG split over a GPU pair (the reference evaluates cond and uncond as two sequential graph executions, context.cpp:352,366):
 * this rank holds eps of ONE guidance half (role 0 = cond, 1 = uncond).  One kernel stores it into the peer's exchange slot (peer_slot / peer_flag
 * are IPC-mapped pointers into the other GPU's memory: plain remote stores over NVLink), raises the peer's flag to `seq`, waits for its own
 * flag to reach `seq`, then applies the same fused CFG + DPM update as sdod_cfg_dpm_step on both ranks, so x / y_prev stay replicated bit for
 * bit.  done_counter: one zero-initialised uint32 in local memory.  Callers alternate two slots / flags with the parity of seq. */
SDOD_API int sdod_cfg_dpm_step_pair(sdod_stream_t stream, float* x, float* y_prev, const float* eps_local, float* peer_slot, const float* recv_slot,
                                    unsigned int* peer_flag, const unsigned int* my_flag, unsigned int* done_counter, unsigned int seq, int role,
                                    size_t n, float guidance, float sigma_s, float alpha_s, float c_x, float c_prev, float c_y0, int order);

/* Host-side schedule tables. Replaces DPMSolver::DPMSolver / ::prepare, dpm_solver.cpp:84-131.
 * Each out array has steps+1 floats (may be NULL). Returns 0 or a negative status. */
SDOD_API int sdod_dpm_schedule(unsigned timesteps, float lin_start, float lin_end, unsigned steps, float* ts,
                               float* log_alphas, float* lambdas, float* sigmas, float* alphas, float* phis, float* i2rs,
                               float* model_ts);
/* Per-step update coefficients derived from the tables exactly as dpm_solver.cpp:137,153-154,168-170. */
SDOD_API int sdod_dpm_coeffs(unsigned timesteps, float lin_start, float lin_end, unsigned steps, unsigned step,
                             float* sigma_s, float* alpha_s, float* c_x, float* c_prev, float* c_y0, int* order);
/* DDIM (eta 0) tables for the same fused step kernel (order 1): model_ts[steps] (descending integer timesteps, as floats) and
 * coeffs[steps][5] = {sigma_s, alpha_s, c_x, c_prev(=0), c_y0}.  Not in the reference tree: restated from the public
 * CompVis ddim.py ("uniform" timesteps i*(T/steps)+1, float64 alphas_cumprod of the scaled-linear betas); parity unpinned. */
SDOD_API int sdod_ddim_schedule(unsigned timesteps, float lin_start, float lin_end, unsigned steps, float* model_ts, float* coeffs);
/* PLMS / linear-multistep variant of the fused step (public CompVis plms.py; parity unpinned, the reference tree has no PLMS):
 *   e = CFG(eps_c, eps_u); if (e_out) e_out = e; e' = w[0] e + w[1] h1 + w[2] h2 + w[3] h3 (NULL history terms are skipped);
 *   x0 = ((x_from ? x_from : x) - s_t e') / a_t;  x = a_prev x0 + s_prev e'   (a = sqrt(alphas_cumprod), s = sqrt(1 - alphas_cumprod)); x_copy as above. */
SDOD_API int sdod_cfg_lms_step(sdod_stream_t stream, float* x, const float* x_from, const void* eps_c, const void* eps_u, int eps_dtype, size_t n,
                               float guidance, const float* w, const float* h1, const float* h2, const float* h3, float* e_out, float a_t,
                               float s_t, float a_prev, float s_prev, float* x_copy);

/* Sinusoidal timestep features, cos first (context.cpp:257-275). out: fp32 [n_t, dim] on device. */
SDOD_API int sdod_timestep_sinusoid(sdod_stream_t stream, const float* t_dev, int n_t, int dim, float max_period, float* out);

/* Initial noise x_T ~ N(0,1) (context.cpp:333-334); Philox4x32-10 + Box-Muller on device. */
SDOD_API int sdod_randn(sdod_stream_t stream, float* out, size_t n, unsigned long long seed, unsigned long long offset);

/* Image quantise: out[i] = uint8(clamp(255*f,0,255)), truncating (context.cpp:392-395).
 * img: [N,HW,C] f32 or bf16 (NHWC == the reference's [H,W,C] output order, libsdod.h:89). */
SDOD_API int sdod_image_to_u8(sdod_stream_t stream, const void* img, int dtype, uint8_t* out, size_t n);

/* ---------------------------------------------------------------------------------------------
 * tcgen05/TMEM GEMM:  C[M,N] = epilogue( alpha * A[M,K] * W[N,K]^T )   (bf16 in, fp32 accumulate)
 * The reference runs these layers inside its opaque `unet`/`decoder` graphs
 * (context.cpp:352,366,387); layer list: analyze_results.py:25-87.
 * ------------------------------------------------------------------------------------------- */
typedef struct sdod_epilogue {
    void* C;                 /* output                                                             */
    void* C2;                /* SDOD_OUT_QKV: K destination                                        */
    void* C3;                /* SDOD_OUT_QKV: V^T destination                                      */
    long long ldc;           /* row stride (elements) for BF16/F32 modes                           */
    long long strideC;       /* batch stride (elements)                                            */
    const float* bias;       /* [N] or NULL                                                        */
    const float* row_bias;   /* [M/rows_per_group, ld_row_bias] fp32 or NULL (timestep-embedding add) */
    int rows_per_group;
    long long ld_row_bias;   /* row stride of row_bias in elements; 0 = N                           */
    const void* residual;    /* bf16 [M,N] row-major or NULL, added last; may alias C (x += f(x)): an fp32 in-place residual leaves through a TMA reduce-add store */
    long long ldr;
    long long strideR;
    float alpha;
    int act;                 /* sdod_act; GEGLU: tile columns [0,BN/2) = value, [BN/2,BN) = gate (default BN 128) */
    int out_mode;            /* sdod_out_mode                                                      */
    int heads, head_dim, tokens, dpad, tok_pad;   /* HEADS / HEADS_T / QKV modes                   */
    int vt_rows;             /* HEADS_T / QKV: rows allocated per head in the V^T buffer           */
    int residual_f32;        /* residual is fp32 (fp32 residual stream) instead of bf16            */
    /* Fused LayerNorm of the output rows (SpatialTransformer norm1/2/3 folded into the GEMM that produces their input):
     * with out_mode F32 and N a multiple of the tile width, ln_out receives bf16 LayerNorm(C[m,:]) * ln_weight + ln_bias
     * (row stride ld_ln) next to the fp32 output.  The N tiles of a row block run as one thread-block cluster and
     * exchange their row moments over distributed shared memory (two-pass mean / variance).  Returns an error status
     * when the shape does not allow it (N / tile width > 8, split-K, batched, multi-wave grid); NULL = off. */
    void* ln_out; long long ld_ln; const float* ln_weight; const float* ln_bias; float ln_eps;
} sdod_epilogue;

typedef struct sdod_gemm_desc {
    const void* A; long long lda; long long strideA;   /* bf16 [batch][M,K], K contiguous           */
    const void* W; long long ldw; long long strideW;   /* bf16 [batch|1][N,K]; strideW==0: shared   */
    int M, N, K, batch;
    int block_n;             /* 0 = auto                                                           */
    sdod_epilogue epi;
    /* optional second operand, K-concatenated:  C = epilogue(alpha * [A | A2] * W^T), W bf16 [N, K + K2]          */
    const void* A2; long long lda2; int K2;            /* bf16 [M, K2] (batch == 1 only), K2 % 64 == 0; NULL = none */
} sdod_gemm_desc;
SDOD_API int sdod_gemm_bf16(sdod_stream_t stream, const sdod_gemm_desc* d);
/* Optional split-K scratch for the calling thread's subsequent sdod_gemm_bf16 / sdod_conv3x3_bf16 calls (small-M
 * layers spread their K loop over the chip: the `split` CTAs of an output tile form a thread-block cluster, publish
 * fp32 partial tiles to this L2-resident scratch and, after a cluster barrier, each folds 1/split of the tile in fixed
 * order (deterministic) and runs the epilogue on it — one launch).  ws: fp32 scratch; counters: reserved (n_counters
 * zero-initialised uint32).  Pass NULLs to disable. */
SDOD_API int sdod_set_splitk_workspace(float* ws, size_t ws_bytes, unsigned int* counters, int n_counters);
/* Profiling hook for the calling thread's subsequent GEMM / conv launches: every CTA records globaltimer stamps of its phases
 * (16 x uint64 per CTA: 0 start, 1 set-up done, 2 predecessor complete, 3 first operands landed, 4 last MMA issued, 5 accumulator
 * complete, 6 epilogue done, 7 split-K partials published, 8 all roles done, 9 residual landed, 10 tile staged) into buf
 * (device memory, >= 16 * 8 * #CTAs bytes).  Results are unaffected.  NULL turns it off.  tools/gemm_timeline.py. */
SDOD_API int sdod_set_gemm_timeline(unsigned long long* buf);

/* Implicit-GEMM conv3x3, stride 1, pad 1, NHWC:  Y[B,H,W,Cout] = epilogue( X (*) Wt )
 * X bf16 [B,H,W,Cin] (Cin % 64 == 0), Wt bf16 [Cout, 9*Cin] with k = (ky*3+kx)*Cin + c.
 * X2 != NULL fuses a 1x1 convolution of a second tensor into the same accumulator (the ResBlock skip_connection):
 * Y = epilogue( X (*) Wt[:, :9*Cin] + X2 * Wt[:, 9*Cin:]^T ), X2 bf16 [B*H*W, Cin2] rows of ldx2 elements, Cin2 % 64 == 0,
 * Wt then has 9*Cin + Cin2 columns. */
typedef struct sdod_conv_desc {
    const void* X; const void* Wt;
    int B, H, W, Cin, Cout;
    int block_n;
    sdod_epilogue epi;       /* M = B*H*W rows, N = Cout                                            */
    const void* X2; long long ldx2; int Cin2;
    /* != 0: Y = conv3x3(nearest_upsample_2x(X)) without materialising the upsampled tensor (UNet / VAE Upsample blocks).  Sub-pixel form: each of
     * the four output parities (y%2, x%2) is a 2x2 convolution over X with pre-summed weights — 4/9 of the multiply-adds.  Wt is then
     * [4 parities][Cout][4*Cin] as packed by sdod_pack_conv3x3_up2_weight; epi.C is [B, 2H, 2W, Cout]; no residual / row_bias / X2. */
    int upsample2x;
    /* 0 / 1: stride 1.  2: stride-2 convolution with padding 1 (the UNet's Downsample blocks): H, W are the INPUT size (even), the output
     * is [B, H/2, W/2, Cout]; the A tiles are TMA boxes with element strides {1, 2, 2, 1} on the input — no im2col buffer. */
    int stride;
} sdod_conv_desc;
SDOD_API int sdod_conv3x3_bf16(sdod_stream_t stream, const sdod_conv_desc* d);
/* OIHW fp32 [Cout,Cin,3,3] -> bf16 [4][Cout][4*Cin] for sdod_conv_desc.upsample2x: parity p = 2*py+px, k = (2a+b)*Cin + c, weight = sum of the
 * 3x3 taps (ky,kx) that read source pixel (i+a-1+py, j+b-1+px) (fp32 sums, rounded once). */
SDOD_API int sdod_pack_conv3x3_up2_weight(sdod_stream_t stream, const float* w_oihw, void* out, int Cout, int Cin);

/* Fused flash-style attention on tcgen05 (S and O accumulators in TMEM, online softmax, P written back to
 * TMEM and consumed as the A operand of the PV MMA).  SpatialTransformer attn1/attn2 (analyze_results.py:60-75).
 * Qh [BH, Nq, dpad], Kh [BH, Nkv, dpad] (HEADS layout, dpad = 64*ceil(head_dim/64), zero padded),
 * Vt [BH, vt_rows, kv_pad] (HEADS_T layout, vt_rows = 16*ceil(head_dim/16), kv_pad % 8 == 0, zero padded).
 * O bf16 [B, Nq, heads*head_dim] (token-major, heads concatenated).  softmax(scale * Q K^T) V. */
SDOD_API int sdod_attention_bf16(sdod_stream_t stream, const void* Qh, const void* Kh, const void* Vt, void* O,
                                 int B, int heads, int Nq, int Nkv, int head_dim, int dpad, int kv_pad, float scale);
/* The same with a causal mask (key k visible to query q iff k <= q): the CLIP text encoder's self-attention (head_dim 64). */
SDOD_API int sdod_attention_causal_bf16(sdod_stream_t stream, const void* Qh, const void* Kh, const void* Vt, void* O,
                                 int B, int heads, int Nq, int Nkv, int head_dim, int dpad, int kv_pad, float scale);

/* Row softmax (unfused attention for the VAE mid block, d=512). x,y bf16 [rows, cols], ld in elements. */
SDOD_API int sdod_softmax_rows(sdod_stream_t stream, const void* x, void* y, long long rows, int cols, long long ld, float scale);

/* Layout / data-movement helpers of the hot path (all bf16 NHWC unless stated). */
SDOD_API int sdod_nchw_f32_to_nhwc_bf16(sdod_stream_t stream, const float* x, void* y, int N, int C, int HW);
SDOD_API int sdod_nhwc_to_nchw_f32(sdod_stream_t stream, const void* x, int dtype, float* y, int N, int C, int HW);
/* x f32|bf16 -> y bf16 */
SDOD_API int sdod_upsample2x_nhwc(sdod_stream_t stream, const void* x, int in_dtype, void* y, int N, int H, int W, int C);
/* a, b, y share one dtype (f32|bf16) */
SDOD_API int sdod_concat_channels(sdod_stream_t stream, const void* a, int Ca, const void* b, int Cb, void* y, long long rows, int dtype);
/* x f32|bf16 -> y bf16 [N*Ho*Wo, Kpad] */
SDOD_API int sdod_im2col3x3(sdod_stream_t stream, const void* x, int in_dtype, void* y, int N, int H, int W, int C, int stride, int Kpad);
SDOD_API int sdod_cast_f32_to_bf16(sdod_stream_t stream, const float* x, void* y, size_t n);
SDOD_API int sdod_silu_bf16(sdod_stream_t stream, const void* x, void* y, size_t n);
/* weight packing (device): OIHW fp32 -> [O][ky][kx][I] bf16 ; GEGLU row interleave per block_n tile */
SDOD_API int sdod_pack_conv3x3_weight(sdod_stream_t stream, const float* w_oihw, void* out, int Cout, int Cin, int Kpad);

#ifdef __cplusplus
}
#endif
#endif /* SDOD_KERNELS_H */
