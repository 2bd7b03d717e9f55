/*
 * libsdod.h — the libsdod C API, served by libsdod_b200.so on NVIDIA B200 (sm_100a).
 *
 * Drop-in for the reference's public API (reference: csrc/libsdod/api/libsdod.h): the same eight
 * symbols, signatures, status codes and ownership rules, so a caller such as the reference's
 * test/simple_app.cpp (or the planned JNI app, README.md:15) relinks unchanged.  Behind them the
 * QNN graph executions are replaced by hand-written CUDA kernels (include/sdod_model.h,
 * include/sdod_kernels.h).  Each declaration cites the reference line it replaces.
 *
 * The `libsdod_b200_*` functions at the bottom are extensions (explicit seed / latent /
 * conditioning and batched generation) used by the parity tests and bench.py; the reference has no
 * counterpart (its set_seed is not exported: context.cpp:285-289).
 */
#ifndef LIBSDOD_API
#define LIBSDOD_API __attribute__((visibility("default")))
#endif

#ifdef __cplusplus
extern "C" {
#endif

/* reference api/libsdod.h:10-18 (== src/errors.h:12-19) */
enum libsdod_status_code {
    LIBSDOD_NO_ERROR,
    LIBSDOD_INVALID_CONTEXT,
    LIBSDOD_INVALID_ARGUMENT,
    LIBSDOD_FAILED_ALLOCATION,
    LIBSDOD_RUNTIME_ERROR,
    LIBSDOD_INTERNAL_ERROR,
};

/* reference api/libsdod.h:21-27 */
enum libsdod_log_level {
    LIBSDOD_LOG_NOTHING,
    LIBSDOD_LOG_ERROR,
    LIBSDOD_LOG_INFO,
    LIBSDOD_LOG_DEBUG,
    LIBSDOD_LOG_ABUSIVE
};

/* reference api/libsdod.h:47.  *context must be NULL on entry; it may be set even on failure and must then
 * still be released.  latent_channels / latent_spatial / upscale_factor: SD1.5 uses 4 / 64 / 8
 * (latent_channels != 4 or upscale_factor != 8 -> LIBSDOD_INVALID_ARGUMENT; latent_spatial must be a
 * power of two >= 8).  models_dir: a directory holding unet.sdodw and vae_decoder.sdodw (flat named-tensor
 * files, see sdod_model.h), or "random-init[:<seed>]" for random-init weights (no checkpoint can be
 * shipped offline).  use_htp is accepted and ignored (there is one backend: sm_100a).
 * steps: the reference accepts only 20 (context.cpp:250-251); any 1..1000 is accepted here. */
LIBSDOD_API int libsdod_setup(void** context, const char* models_dir, unsigned int latent_channels, unsigned int latent_spatial,
                              unsigned int upscale_factor, unsigned int steps, unsigned int log_level, int use_htp);

/* reference api/libsdod.h:57 */
LIBSDOD_API int libsdod_set_steps(void* context, unsigned int steps);

/* reference api/libsdod.h:67 */
LIBSDOD_API int libsdod_set_log_level(void* context, unsigned int log_level);

/* reference api/libsdod.h:75 */
LIBSDOD_API int libsdod_ref_context(void* context);

/* reference api/libsdod.h:81.  At ref-count 0 the engine is destroyed; the handle shell stays allocated so
 * stale use is reported as LIBSDOD_INVALID_CONTEXT (reference src/libsdod.cpp:53-60,154-158). */
LIBSDOD_API int libsdod_release(void* context);

/* reference api/libsdod.h:117.  guidance_scale g: e = g*eps(x,t,prompt) + (1-g)*eps(x,t,"") (context.cpp:362,373);
 * g == 1.0 skips the unconditional pass (context.cpp:359).  *image_out == NULL: the library malloc()s
 * 3*(latent_spatial*upscale)^2 bytes and the caller free()s them; otherwise the caller's buffer of
 * *image_buffer_size bytes is used and *image_buffer_size is overwritten with the bytes written.
 * Output: RGB, [H, W, C], uint8(clamp(255*f, 0, 255)) (context.cpp:392-395).
 * Text conditioning: the CLIP encoder + tokenizer are outside this hot path; the prompt is mapped
 * to a deterministic pseudo-embedding (seeded by its bytes) unless conditioning is supplied through
 * libsdod_b200_generate below. */
LIBSDOD_API int libsdod_generate_image(void* context, const char* prompt, float guidance_scale, unsigned char** image_out,
                                       unsigned int* image_buffer_size);

/* reference api/libsdod.h:124 */
LIBSDOD_API const char* libsdod_get_error_description(int errorcode);

/* reference api/libsdod.h:138 */
LIBSDOD_API const char* libsdod_get_last_error_extra_info(int errorcode, void* context);

/* ------------------------------------------------------------------------------------------------ extensions */

/* Seed of the initial-noise generator (the reference's Context::set_seed, context.cpp:285-289, not exported there). */
LIBSDOD_API int libsdod_b200_set_seed(void* context, unsigned long long seed);

/* Sampler of the denoising loop.  DPM-Solver++(2M) is the reference's only sampler (dpm_solver.cpp) and the default; DDIM (eta 0) runs on the
 * same fused CFG + update kernel with its own timestep / coefficient tables (sdod_ddim_schedule; SURVEY §8 row f4, parity unpinned: the
 * reference tree has no DDIM).  Re-prepares the schedule for the current step count. */
#define LIBSDOD_B200_SAMPLER_DPM 0
#define LIBSDOD_B200_SAMPLER_DDIM 1
#define LIBSDOD_B200_SAMPLER_PLMS 2   /* CompVis plms.py on the DDIM timesteps; first step = two UNet evaluations */
LIBSDOD_API int libsdod_b200_set_sampler(void* context, int sampler);

/* Batched generation with explicit inputs (all HOST pointers; copies are part of the call):
 *   cond, uncond : [n_images, 77, 768] fp32 text-encoder outputs (uncond may be NULL when guidance == 1)
 *   latents      : [n_images, 4, S, S] fp32 NCHW initial noise, or NULL to draw from the context's generator
 *   images_out   : [n_images, 8S, 8S, 3] uint8
 *   latents_out  : optional [n_images, 4, S, S] fp32 final latents (parity tests), may be NULL
 * n_images must not exceed the max_images given at setup through libsdod_b200_setup (default 1). */
LIBSDOD_API int libsdod_b200_generate(void* context, unsigned int n_images, const float* cond, const float* uncond, const float* latents,
                                      float guidance_scale, unsigned char* images_out, float* latents_out);

/* Same loop with DEVICE pointers and no host copies: cond/uncond [n,77,768] fp32, latents [n,S,S,4] fp32 NHWC (or NULL),
 * images_out [n,8S,8S,3] uint8 on the device.  Used to time the loop with inputs already resident in HBM. */
LIBSDOD_API int libsdod_b200_generate_device(void* context, unsigned int n_images, const float* cond_dev, const float* uncond_dev,
                                             const float* latents_nhwc_dev, float guidance_scale, unsigned char* images_out_dev);

/* libsdod_setup with a batch capacity (> 1 image per call) and an explicit CUDA device. */
LIBSDOD_API int libsdod_b200_setup(void** context, const char* models_dir, unsigned int latent_spatial, unsigned int steps,
                                   unsigned int log_level, unsigned int max_images, int device);

/* ---- CFG split over a GPU pair (one process per GPU).  The reference runs the conditional and the unconditional UNet pass one
 * after the other on one device (context.cpp:352,366); here two contexts on two GPUs each run ONE of them per step and exchange the noise
 * prediction over NVLink: libsdod_b200_pair_export() returns a 64-byte CUDA IPC handle of this context's exchange buffer; the application
 * hands it to the peer process (any transport: a torch.distributed all_gather, a pipe, ...); libsdod_b200_pair_connect() maps the peer's buffer
 * and fixes this context's role (0 = conditional half, 1 = unconditional half).  libsdod_b200_generate_pair() then runs the same loop as
 * libsdod_b200_generate with a UNet batch of n_images instead of 2*n_images; per step one fused kernel stores this half's eps into the peer's
 * buffer, waits for the peer's and applies the identical CFG + solver update on both ranks (the sampler state stays replicated bit for bit).
 * The VAE decode is split by image: this rank decodes and returns images [*first_image, *first_image + *n_decoded) of images_out
 * ([n_images, 8S, 8S, 3] on both ranks; the rest of the buffer is left untouched).  Both contexts must be set up alike (steps, sampler, seed when
 * latents == NULL) and call generate_pair the same number of times.  DPM-Solver++ and DDIM only; guidance must differ from 1. */
LIBSDOD_API int libsdod_b200_pair_export(void* context, unsigned char handle_out[64]);
LIBSDOD_API int libsdod_b200_pair_connect(void* context, const unsigned char peer_handle[64], int role);
LIBSDOD_API int libsdod_b200_generate_pair(void* context, unsigned int n_images, const float* conditioning_half, const float* latents,
                                           float guidance_scale, unsigned char* images_out, unsigned int* first_image, unsigned int* n_decoded,
                                           float* latents_out);

/* The prompt path on its own (context.cpp:325-329: tokenizer -> cond_model): embedding_out[77*768] (host, fp32) = last_hidden_state of the CLIP text
 * encoder for `prompt` — what libsdod_generate_image feeds the UNet as conditioning, and what libsdod_b200_generate takes as cond / uncond.
 * tokens_out (may be NULL) receives the 77 token ids.  Fails with LIBSDOD_RUNTIME_ERROR when the context has no tokenizer / text encoder
 * (a models_dir without ctokenizer.txt + text_encoder.sdodw), LIBSDOD_INVALID_ARGUMENT on invalid UTF-8 (tokenizer.cpp:77). */
LIBSDOD_API int libsdod_b200_encode_prompt(void* context, const char* prompt, float* embedding_out, unsigned short* tokens_out);

/* libsdod_generate_image for n prompts in one call (n <= max_images): the prompts go through the tokenizer and ONE text-encoder batch, every image's
 * unconditional half is the cached empty prompt (context.cpp:233-239), the latents are drawn from the context's generator, and images_out
 * ([n_images, 8S, 8S, 3] uint8, HOST) receives the images.  Same status codes as libsdod_generate_image. */
LIBSDOD_API int libsdod_b200_generate_images(void* context, unsigned int n_images, const char* const* prompts, float guidance_scale,
                                             unsigned char* images_out);

/* Milliseconds of the last generate call as the reference logs them (context.cpp:309-314,331,381,398,402):
 * out[0] conditioning, out[1] mean single iteration, out[2] decoding, out[3] total.  Device-event timed. */
LIBSDOD_API int libsdod_b200_last_timings(void* context, float out[4]);

#ifdef __cplusplus
}
#endif
