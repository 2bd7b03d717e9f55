/*
 * sdod_model.h — C ABI of libsdod_b200.so, graph level: the UNet denoiser, the time-embedding MLP
 * and the VAE decoder as pre-planned launch sequences of the kernels in sdod_kernels.h.
 *
 * These replace the reference's opaque serialized graphs and their execute() calls:
 *   unet    : csrc/libsdod/src/context.cpp:352,366  (graph inputs 0=x, 1=t (1280-d), 2=p: context.cpp:214-218)
 *   temb    : context.cpp:257-279                    (sinusoid -> `temb` graph -> 1280 vector per step)
 *   decoder : context.cpp:386-395                    (+ the host uint8 map)
 * Weight names are the public ldm state_dict keys (analyze_results.py:25-87 shows the same names), so a
 * real SD-v1 checkpoint loads by key.  Latents cross this ABI as fp32 NHWC [B,H,W,4].
 */
#ifndef SDOD_MODEL_H
#define SDOD_MODEL_H

#include "sdod_kernels.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct sdod_weights sdod_weights; /* named fp32 tensors resident on the device */
typedef struct sdod_unet sdod_unet;
typedef struct sdod_vae sdod_vae;

SDOD_API int sdod_weights_create(sdod_weights** out);
/* host fp32 data in PyTorch layout (conv: OIHW, linear: [out,in]); copied to the device */
SDOD_API int sdod_weights_set_f32(sdod_weights* w, const char* name, const float* host_data, int ndim, const long long* shape);
/* flat file: "SDODW001", u32 count, then per tensor: u32 name_len, name, u32 ndim, i64 shape[ndim], f32 data */
SDOD_API int sdod_weights_load_file(sdod_weights* w, const char* path);
SDOD_API long long sdod_weights_count(const sdod_weights* w);
SDOD_API void sdod_weights_destroy(sdod_weights* w);

/* weights == NULL: random-init with `seed` (PyTorch-default-like scales; benchmarking without checkpoints).
 * latent_hw: latent spatial size (64 for 512x512).  max_batch: largest B a forward may use (CFG: 2 x images). */
SDOD_API int sdod_unet_create(sdod_unet** out, const sdod_weights* weights, unsigned long long seed, int latent_hw, int max_batch);
SDOD_API void sdod_unet_destroy(sdod_unet* u);
/* emb_out[n,1280] = time_embed(sinusoid(t[n]))  (device pointers) */
SDOD_API int sdod_unet_time_embed(sdod_unet* u, sdod_stream_t stream, const float* t, int n, float* emb_out);
/* context [B,77,768] f32 or bf16 (device): projects and caches the cross-attention K / V^T of every block */
SDOD_API int sdod_unet_set_context(sdod_unet* u, sdod_stream_t stream, const void* context, int dtype, int B);
/* eps[B,H,W,4] (fp32) = UNet(x[B,H,W,4] (fp32 NHWC), emb[B,1280] (fp32), cached context).  use_graph != 0
 * replays a CUDA graph captured on first use for this B. */
SDOD_API int sdod_unet_forward(sdod_unet* u, sdod_stream_t stream, const float* x, const float* emb, float* eps, int B, int use_graph);
SDOD_API unsigned long long sdod_unet_launches_per_forward(const sdod_unet* u, int B);
/* Per-op device times of the forward plan (eager, CUDA events between ops), written as "ms<TAB>name" lines into buf. */
SDOD_API int sdod_unet_profile(sdod_unet* u, sdod_stream_t stream, int B, int iters, char* buf, size_t buf_bytes);

SDOD_API int sdod_vae_create(sdod_vae** out, const sdod_weights* weights, unsigned long long seed, int latent_hw, int max_batch);
SDOD_API void sdod_vae_destroy(sdod_vae* v);
/* z [B,H,W,4] fp32 NHWC latent -> image_u8 [B,8H,8W,3] (may be NULL) and image_f32 [B,8H,8W,3] in [0,1] (may be NULL) */
SDOD_API int sdod_vae_decode(sdod_vae* v, sdod_stream_t stream, const float* z, uint8_t* image_u8, float* image_f32, int B, int use_graph);

/* ---- prompt path: CLIP byte-pair tokenizer + CLIP ViT-L/14 text encoder.
 * Replaces libsdod::Tokenizer (csrc/libsdod/src/tokenizer.{h,cpp}; loaded at context.cpp:180-187, called at context.cpp:235,325) and the
 * `cond_model` graph ("text_encoder.serialized", context.cpp:143,170; executed at context.cpp:237,327). */
typedef struct sdod_tokenizer sdod_tokenizer;
typedef struct sdod_text_encoder sdod_text_encoder;

/* bpe_file: a ctokenizer.txt in the reference's format (gen_tokenizer_file.py:27-42).  NULL: byte-level vocabulary without merges
 * (the 512 byte symbols in the same id order) — what random-init contexts use. */
SDOD_API int sdod_tokenizer_create(sdod_tokenizer** out, const char* bpe_file);
SDOD_API void sdod_tokenizer_destroy(sdod_tokenizer* t);
/* tokens_out[context_len] = [start, ids..., end padding]; context_len = 77 for SD (tokenizer.h:24).  Invalid UTF-8 -> status -1 (invalid argument).
 * The ids equal the reference's for every prompt on which the reference terminates.  *deviated (may be NULL) is set to 1 for the prompts on which
 * it does not: its merge pass (tokenizer.cpp:339-356) repeats for ever when the best-ranked pair (a, b) only occurs as "a a b"; there the
 * textbook BPE merge is applied instead. */
SDOD_API int sdod_tokenizer_encode(const sdod_tokenizer* t, const char* utf8, unsigned short* tokens_out, unsigned context_len, int* deviated);
SDOD_API int sdod_tokenizer_vocab_size(const sdod_tokenizer* t);

/* weights == NULL: random-init with `seed`.  Weight names: the HuggingFace CLIPTextModel state_dict keys under "text_model."
 * (ldm checkpoints: "cond_stage_model.transformer.text_model.", stripped by sdod/checkpoint.py). */
SDOD_API int sdod_text_encoder_create(sdod_text_encoder** out, const sdod_weights* weights, unsigned long long seed, int max_batch);
SDOD_API void sdod_text_encoder_destroy(sdod_text_encoder* e);
/* tokens [B,77] int32 (device) -> context_out [B,77,768] (device; dtype SDOD_F32 or SDOD_BF16): last_hidden_state after final_layer_norm */
SDOD_API int sdod_text_encoder_forward(sdod_text_encoder* e, sdod_stream_t stream, const int* tokens, int B, void* context_out, int out_dtype);

#ifdef __cplusplus
}
#endif
#endif /* SDOD_MODEL_H */
