// The libsdod C API on B200: context handle, error tables, and the generate loop.
//
// Behavioural contract restated from the reference (csrc/libsdod/src/):
//   handle / ref-count / validation     libsdod.cpp:16-27, 48-63, 146-161
//   error recording & strings           libsdod.cpp:29-45, errors.cpp:8-15, 37-66
//   setup sequence                      libsdod.cpp:66-111, context.cpp:49-80, 191-282
//   generate loop                       context.cpp:292-403   (the hot path)
//   output buffer ownership             context.cpp:406-421, buffer.cpp:12-18
// The loop keeps x, y_prev, eps and the image on the device; per step it launches one CUDA-graph
// replay of the batch-2n UNet (cond + uncond stacked) and one fused CFG+DPM kernel — no host
// round trip until the uint8 image is copied out.
#include "libsdod.h"

#include <array>
#include <chrono>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <new>
#include <optional>
#include <random>
#include <string>
#include <vector>

#include "../kernels/glue.h"
#include "clip_text.h"
#include "dpm_schedule.h"
#include "sdod_kernels.h"
#include "tokenizer.h"
#include "unet.h"
#include "vae.h"

namespace {

constexpr unsigned kMagic = 0x00534443;      // reference libsdod.cpp:16
constexpr unsigned kHandleVersion = 1;        // reference libsdod.cpp:17
constexpr int kNumErrors = 6;

const char* kErrorStrings[kNumErrors] = {     // reference errors.cpp:8-15
    "No error",
    "Invalid context",
    "Invalid argument",
    "Failed to allocate memory or initialise an object",
    "Runtime error occurred",
    "Internal error occurred"};

using ErrorTable = std::array<std::optional<std::string>, kNumErrors>;
ErrorTable g_contextless_errors;

struct ApiError {
    int code;
    std::string msg;
    const char* func;
    const char* file;
    int line;
};
#define API_THROW(code, msg) throw ApiError{code, msg, __func__, __FILE__, __LINE__}
#define CU(expr)                                                                                                   \
    do {                                                                                                           \
        cudaError_t _e = (expr);                                                                                   \
        if (_e != cudaSuccess) API_THROW(LIBSDOD_RUNTIME_ERROR, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
    } while (0)
#define SD(expr)                                                                            \
    do {                                                                                    \
        if ((expr) != 0) API_THROW(LIBSDOD_RUNTIME_ERROR, std::string(sdod::last_error())); \
    } while (0)

class Logger {                                  // reference logging.h:12-18, logging.cpp:23-41,82-102
public:
    explicit Logger(unsigned level) : level_(level), t0_(std::chrono::steady_clock::now()) {}
    void set_level(unsigned l) { level_ = l; }
    void log(unsigned level, const char* fmt, ...) const {
        if (level > level_ || level == LIBSDOD_LOG_NOTHING) return;
        static const char* names[] = {"", "ERROR", "INFO", "DEBUG", "ABUSIVE"};
        const double secs = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0_).count();
        FILE* out = level == LIBSDOD_LOG_ERROR ? stderr : stdout;
        std::fprintf(out, "[+%.3f]:[%s] ", secs, names[level]);
        va_list ap;
        va_start(ap, fmt);
        std::vfprintf(out, fmt, ap);
        va_end(ap);
        std::fputc('\n', out);
    }

private:
    unsigned level_;
    std::chrono::steady_clock::time_point t0_;
};

class Engine {
public:
    // Like the reference's Context (context.cpp:24-47) the constructor only records its arguments and cannot fail; everything that can
    // (device, weights, plans) happens in init(), which libsdod_setup runs AFTER publishing the handle (libsdod.cpp:77-87), so a failed
    // setup leaves a context whose error table can be queried and which must be released.
    Engine(const std::string& models_dir, unsigned latent_spatial, unsigned log_level, unsigned max_images, int device)
        : log_(log_level), S_(static_cast<int>(latent_spatial)), max_images_(static_cast<int>(max_images)), device_(device), sched_(1000, 0.00085f, 0.0120f), ddim_(1000, 0.00085f, 0.0120f),   // context.cpp:196
          models_dir_(models_dir) {}

    void init() {
        std::string models_dir = models_dir_;
        if (const char* over = std::getenv("LIBSDOD_B200_MODELS_DIR")) {      // deployment override: the reference's callers hard-code a relative path (simple_app.cpp:11)
            models_dir = over;
            log_.log(LIBSDOD_LOG_INFO, "models_dir overridden by LIBSDOD_B200_MODELS_DIR: %s", over);
        }
        int ndev = 0;
        if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
            cudaGetLastError();
            API_THROW(LIBSDOD_RUNTIME_ERROR, "no CUDA device available (libsdod_b200 has no CPU fallback)");
        }
        if (device_ >= 0) CU(cudaSetDevice(device_));
        else CU(cudaGetDevice(&device_));                                 // libsdod_setup: pin the caller's current device for every later entry
        CU(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking));
        for (auto& e : ev_) CU(cudaEventCreate(&e));
        seed_ = std::random_device{}();                                   // context.cpp:16
        const sdod::WeightStore* uw = nullptr;
        const sdod::WeightStore* vw = nullptr;
        unsigned long long wseed = 0;
        if (models_dir.rfind("random-init", 0) == 0) {
            const auto pos = models_dir.find(':');
            if (pos != std::string::npos) wseed = std::strtoull(models_dir.c_str() + pos + 1, nullptr, 10);
            log_.log(LIBSDOD_LOG_INFO, "Using random-init weights (seed %llu)", wseed);
            tokenizer_ = std::make_unique<sdod::Tokenizer>();             // byte-level vocabulary: there is no vocabulary file to read
            have_text_ = true;
        } else {
            std::string dir = models_dir.empty() ? "." : models_dir;          // context.cpp:19-22
            if (dir.back() == '/') dir.pop_back();
            unet_w_ = std::make_unique<sdod::WeightStore>();
            vae_w_ = std::make_unique<sdod::WeightStore>();
            SD(unet_w_->load_file(dir + "/unet.sdodw"));
            SD(vae_w_->load_file(dir + "/vae_decoder.sdodw"));
            uw = unet_w_.get();
            vw = vae_w_.get();
            log_.log(LIBSDOD_LOG_INFO, "Loaded %zu UNet and %zu decoder tensors from %s", unet_w_->size(), vae_w_->size(), dir.c_str());
            // prompt path (context.cpp:180-187 tokenizer, :143,170 cond_model): <models_dir>/ctokenizer.txt + text_encoder.sdodw.  Deviation: the
            // reference fails setup without them; here a models_dir without the text model still serves libsdod_b200_generate (conditioning
            // supplied by the caller) and libsdod_generate_image reports the missing files when it is called.
            const std::string tok_path = dir + "/ctokenizer.txt", te_path = dir + "/text_encoder.sdodw";
            FILE* tf = std::fopen(tok_path.c_str(), "rb");
            FILE* ef = std::fopen(te_path.c_str(), "rb");
            if (tf) std::fclose(tf);
            if (ef) std::fclose(ef);
            if (tf && ef) {
                try {
                    tokenizer_ = std::make_unique<sdod::Tokenizer>(tok_path);
                } catch (const std::exception& e) {
                    API_THROW(LIBSDOD_INVALID_ARGUMENT, e.what());
                }
                text_w_ = std::make_unique<sdod::WeightStore>();
                SD(text_w_->load_file(te_path));
                have_text_ = true;
                log_.log(LIBSDOD_LOG_INFO, "Loaded the tokenizer (%zu tokens, %zu merges) and %zu text-encoder tensors", tokenizer_->vocab_size(), tokenizer_->merges(), text_w_->size());
            } else {
                log_.log(LIBSDOD_LOG_INFO, "No %s in %s: prompts cannot be encoded (libsdod_generate_image will fail; libsdod_b200_generate takes embeddings)",
                         tf ? "text_encoder.sdodw" : "ctokenizer.txt", dir.c_str());
            }
        }
        try {
            unet_ = std::make_unique<sdod::UNet>(uw, wseed, S_, 2 * max_images_);
            vae_ = std::make_unique<sdod::VaeDecoder>(vw, wseed + 1, S_, max_images_);
            if (have_text_) {
                text_ = std::make_unique<sdod::ClipTextEncoder>(text_w_.get(), wseed + 2, max_images_);
                if (tokenizer_->vocab_size() > static_cast<size_t>(sdod::ClipTextEncoder::kVocab))
                    API_THROW(LIBSDOD_INVALID_ARGUMENT, "ctokenizer.txt defines more tokens than the text encoder's embedding table holds");
            }
        } catch (const ApiError&) {
            throw;
        } catch (const std::exception& e) {
            API_THROW(LIBSDOD_RUNTIME_ERROR, e.what());
        }
        const size_t lat = static_cast<size_t>(max_images_) * S_ * S_ * 4;
        CU(cudaMalloc(reinterpret_cast<void**>(&y_prev_), lat * sizeof(float)));
        CU(cudaMemset(y_prev_, 0, lat * sizeof(float)));
        CU(cudaMalloc(reinterpret_cast<void**>(&ctx_dev_), static_cast<size_t>(2 * max_images_) * 77 * 768 * sizeof(float)));
        CU(cudaMallocHost(reinterpret_cast<void**>(&pin_ctx_), static_cast<size_t>(2 * max_images_) * 77 * 768 * sizeof(float)));
        CU(cudaMallocHost(reinterpret_cast<void**>(&pin_lat_), lat * sizeof(float)));
        CU(cudaMallocHost(reinterpret_cast<void**>(&pin_img_), static_cast<size_t>(max_images_) * image_bytes()));
        // cached empty-prompt conditioning (context.cpp:233-239)
        if (have_text_) {
            CU(cudaMalloc(reinterpret_cast<void**>(&tok_dev_), static_cast<size_t>(max_images_) * 77 * sizeof(int)));
            uncond_default_.resize(77 * 768);
            prompt_embedding("", uncond_default_.data());
        }
        ready_ = true;
        log_.log(LIBSDOD_LOG_INFO, "Models and buffers prepared!");
    }

    ~Engine() {                                                           // safe on a partially initialised engine (failed init)
        if (device_ >= 0) cudaSetDevice(device_);
        if (stream_) cudaStreamSynchronize(stream_);
        unet_.reset();
        vae_.reset();
        text_.reset();
        cudaFree(tok_dev_);
        cudaFree(y_prev_); cudaFree(ctx_dev_); cudaFree(temb_); cudaFree(plms_buf_);
        if (peer_exch_) cudaIpcCloseMemHandle(peer_exch_);
        cudaFree(exch_); cudaFree(pair_done_);
        if (pin_ctx_) cudaFreeHost(pin_ctx_);
        if (pin_lat_) cudaFreeHost(pin_lat_);
        if (pin_img_) cudaFreeHost(pin_img_);
        for (auto& e : ev_) if (e) cudaEventDestroy(e);
        if (stream_) cudaStreamDestroy(stream_);
        cudaGetLastError();
    }
    void require_ready() const {
        if (!ready_) API_THROW(LIBSDOD_RUNTIME_ERROR, "context is not initialised (libsdod_setup failed; query its error and release the context)");
    }

    size_t image_bytes() const { return static_cast<size_t>(3) * S_ * 8 * S_ * 8; }     // context.cpp:406-409
    Logger& logger() { return log_; }
    ErrorTable& errors() { return errors_; }
    void set_sampler(int sampler) {
        if (sampler != LIBSDOD_B200_SAMPLER_DPM && sampler != LIBSDOD_B200_SAMPLER_DDIM && sampler != LIBSDOD_B200_SAMPLER_PLMS)
            API_THROW(LIBSDOD_INVALID_ARGUMENT, "unknown sampler id: " + std::to_string(sampler));
        require_ready();
        sampler_ = sampler;
        if (steps_) prepare_schedule(steps_);
        log_.log(LIBSDOD_LOG_INFO, "Sampler: %s", sampler == LIBSDOD_B200_SAMPLER_DDIM ? "DDIM (eta 0)" : sampler == LIBSDOD_B200_SAMPLER_PLMS ? "PLMS" : "DPM-Solver++(2M)");
    }
    void set_seed(unsigned long long s) { seed_ = s; rng_offset_ = 0; log_.log(LIBSDOD_LOG_INFO, "Using seed: %llu", s); }
    const float* timings() const { return timings_; }
    int max_images() const { return max_images_; }

    void prepare_schedule(unsigned steps) {                               // context.cpp:245-282
        require_ready();
        if (steps < 1 || steps > 1000) API_THROW(LIBSDOD_INVALID_ARGUMENT, "steps must be in [1, 1000], got: " + std::to_string(steps));
        if (device_ >= 0) CU(cudaSetDevice(device_));
        sched_.prepare(steps);
        if (sampler_ != LIBSDOD_B200_SAMPLER_DPM) ddim_.prepare(steps);     // PLMS runs on the DDIM timesteps / alphas (plms.py)
        const float* model_ts = sampler_ != LIBSDOD_B200_SAMPLER_DPM ? ddim_.model_ts.data() : sched_.model_ts.data();
        cudaFree(temb_);
        temb_ = nullptr;
        float* t_dev = nullptr;
        CU(cudaMalloc(reinterpret_cast<void**>(&temb_), static_cast<size_t>(steps) * 1280 * sizeof(float)));
        CU(cudaMalloc(reinterpret_cast<void**>(&t_dev), steps * sizeof(float)));
        CU(cudaMemcpyAsync(t_dev, model_ts, steps * sizeof(float), cudaMemcpyHostToDevice, stream_));   // first `steps` of steps+1 (context.cpp:267)
        int st = unet_->time_embed(stream_, t_dev, static_cast<int>(steps), temb_);
        CU(cudaStreamSynchronize(stream_));
        cudaFree(t_dev);
        if (st != 0) API_THROW(LIBSDOD_RUNTIME_ERROR, std::string(sdod::last_error()));
        steps_ = steps;
        log_.log(LIBSDOD_LOG_INFO, "Time schedule prepared for %u steps!", steps);
    }

    // context.cpp:325-329: tokenize -> cond_model -> p_cond.  77 ids go up, the [77,768] last_hidden_state comes back to the host buffer the
    // generate loop stages from (a prompt is encoded once per image, not per step).
    void prompt_embedding(const char* prompt, float* out_host, unsigned short* tokens_out = nullptr) {
        prompt_embeddings(1, &prompt, out_host, tokens_out);
    }
    // n prompts as ONE text-encoder batch: out_host [n, 77, 768] fp32
    void prompt_embeddings(unsigned n, const char* const* prompts, float* out_host, unsigned short* tokens_out = nullptr) {
        if (!have_text_) API_THROW(LIBSDOD_RUNTIME_ERROR, "this context has no tokenizer / text encoder (models_dir lacks ctokenizer.txt or text_encoder.sdodw)");
        if (n < 1 || static_cast<int>(n) > max_images_) API_THROW(LIBSDOD_INVALID_ARGUMENT, "number of prompts out of range (max_images = " + std::to_string(max_images_) + ")");
        std::vector<int> tok(static_cast<size_t>(n) * 77);
        for (unsigned i = 0; i < n; ++i) {
            if (!prompts[i]) API_THROW(LIBSDOD_INVALID_ARGUMENT, "prompt is nullptr");
            std::vector<sdod::Tokenizer::token_type> ids;
            try {
                ids = tokenizer_->encode(prompts[i], 77);
            } catch (const sdod::TokenizerError& e) {
                API_THROW(LIBSDOD_INVALID_ARGUMENT, e.what());              // tokenizer.cpp:77
            }
            for (int j = 0; j < 77; ++j) tok[i * 77 + static_cast<size_t>(j)] = ids[static_cast<size_t>(j)];
            if (tokens_out) std::memcpy(tokens_out + i * 77, ids.data(), 77 * sizeof(unsigned short));
        }
        const size_t count = static_cast<size_t>(n) * 77 * 768;
        float* d = nullptr;
        CU(cudaMalloc(reinterpret_cast<void**>(&d), count * sizeof(float)));
        cudaError_t e = cudaMemcpyAsync(tok_dev_, tok.data(), tok.size() * sizeof(int), cudaMemcpyHostToDevice, stream_);
        int st = e == cudaSuccess ? text_->forward(stream_, tok_dev_, static_cast<int>(n), d, SDOD_F32) : 0;
        if (e == cudaSuccess && st == 0) e = cudaMemcpyAsync(out_host, d, count * sizeof(float), cudaMemcpyDeviceToHost, stream_);
        cudaStreamSynchronize(stream_);
        cudaFree(d);
        if (st != 0) API_THROW(LIBSDOD_RUNTIME_ERROR, std::string(sdod::last_error()));
        CU(e);
    }

    // context.cpp:292-403 for n prompts in one call: one text-encoder batch, the cached empty prompt as every image's unconditional half
    void generate_prompts(unsigned n, const char* const* prompts, float guidance, unsigned char* images_out) {
        require_ready();
        if (!prompts) API_THROW(LIBSDOD_INVALID_ARGUMENT, "prompts is nullptr");
        if (device_ >= 0) CU(cudaSetDevice(device_));
        std::vector<float> cond(static_cast<size_t>(n ? n : 1) * 77 * 768), uncond;
        prompt_embeddings(n, prompts, cond.data());
        uncond.reserve(cond.size());
        for (unsigned i = 0; i < n; ++i) uncond.insert(uncond.end(), uncond_default_.begin(), uncond_default_.end());
        generate(n, cond.data(), uncond.data(), nullptr, guidance, images_out, nullptr);
    }

    void generate_prompt(const char* prompt, float guidance, unsigned char* out) {
        require_ready();
        log_.log(LIBSDOD_LOG_INFO, "Starting image generation for prompt: \"%s\" and guidance %g", prompt, guidance);
        std::vector<float> cond(77 * 768);
        prompt_embedding(prompt, cond.data());
        generate(1, cond.data(), uncond_default_.data(), nullptr, guidance, out, nullptr);
    }

    // PLMS (pseudo linear multistep, public CompVis plms.py p_sample_plms; row f4, parity unpinned): 4-term Adams-Bashforth over the CFG-combined
    // eps, kept in a 3-slot device ring written by the fused step kernel itself; the first step is a pseudo improved-Euler step with a second
    // UNet evaluation at the next timestep.
    void plms_loop(float* x, float* eps, int B, size_t lat, bool cfg, float guidance) {
        if (plms_cap_ < lat) {
            cudaFree(plms_buf_);
            plms_buf_ = nullptr;
            CU(cudaMalloc(reinterpret_cast<void**>(&plms_buf_), 4 * lat * sizeof(float)));     // 3 history slots + the saved x_t
            plms_cap_ = lat;
        }
        float* hist[3] = {plms_buf_, plms_buf_ + lat, plms_buf_ + 2 * lat};
        float* x_saved = plms_buf_ + 3 * lat;
        const float* eu = cfg ? eps + lat : nullptr;
        float* xc = cfg ? x + lat : nullptr;
        auto run_unet = [&](unsigned step) {
            SD(sdod::broadcast_rows(stream_, unet_->emb_in(), temb_ + static_cast<size_t>(step) * 1280, B, 1280));
            SD(unet_->forward(stream_, x, unet_->emb_in(), eps, B, true));
        };
        for (unsigned step = 0; step < steps_; ++step) {
            const sdod::DpmStep k = ddim_.step(step);            // sigma_s = sqrt(1 - a_t), alpha_s = sqrt(a_t)
            const float a_prev = ddim_.sqrt_a_prev[step], s_prev = ddim_.sqrt_1m_a_prev[step];
            float* slot = hist[step % 3];
            const float* h1 = step >= 1 ? hist[(step - 1) % 3] : nullptr;
            const float* h2 = step >= 2 ? hist[(step - 2) % 3] : nullptr;
            const float* h3 = step >= 3 ? hist[(step - 3) % 3] : nullptr;
            run_unet(step);
            if (step == 0 && steps_ > 1) {
                const float w_euler[4] = {1.f, 0.f, 0.f, 0.f}, w_avg[4] = {0.5f, 0.5f, 0.f, 0.f};
                CU(cudaMemcpyAsync(x_saved, x, lat * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
                SD(sdod_cfg_lms_step(stream_, x, nullptr, eps, eu, SDOD_F32, lat, guidance, w_euler, nullptr, nullptr, nullptr, slot, k.alpha_s, k.sigma_s,
                                     a_prev, s_prev, xc));                                     // x_prev from e_t alone; e_t -> history
                run_unet(1);                                                                    // e_t_next = model(x_prev, t_next)
                SD(sdod_cfg_lms_step(stream_, x, x_saved, eps, eu, SDOD_F32, lat, guidance, w_avg, slot, nullptr, nullptr, nullptr, k.alpha_s, k.sigma_s,
                                     a_prev, s_prev, xc));                                     // e' = (e_t + e_t_next) / 2 applied to the saved x_t
                continue;
            }
            const float w1[4] = {1.f, 0.f, 0.f, 0.f};
            const float w2[4] = {1.5f, -0.5f, 0.f, 0.f};
            const float w3[4] = {23.f / 12.f, -16.f / 12.f, 5.f / 12.f, 0.f};
            const float w4[4] = {55.f / 24.f, -59.f / 24.f, 37.f / 24.f, -9.f / 24.f};
            const float* w = step == 0 ? w1 : step == 1 ? w2 : step == 2 ? w3 : w4;
            // the ring slot being written (step % 3) is the one holding e_{t-3}: read it (h3) and write it in the same pass, element by element
            SD(sdod_cfg_lms_step(stream_, x, nullptr, eps, eu, SDOD_F32, lat, guidance, w, h1, h2, h3, slot, k.alpha_s, k.sigma_s, a_prev, s_prev, xc));
        }
    }

    // ---- CFG split over a GPU pair (libsdod.h: libsdod_b200_pair_*)
    // exchange buffer: [2 flags, padded to 256 B][slot 0][slot 1], a slot = max_images latents (fp32); the peer writes into it over NVLink
    size_t slot_floats() const { return static_cast<size_t>(max_images_) * S_ * S_ * 4; }
    void pair_export(unsigned char out[64]) {
        require_ready();
        if (device_ >= 0) CU(cudaSetDevice(device_));
        if (!exch_) {
            const size_t bytes = 256 + 2 * slot_floats() * sizeof(float);
            CU(cudaMalloc(reinterpret_cast<void**>(&exch_), bytes));
            CU(cudaMemset(exch_, 0, bytes));
            CU(cudaMalloc(reinterpret_cast<void**>(&pair_done_), sizeof(unsigned int)));
            CU(cudaMemset(pair_done_, 0, sizeof(unsigned int)));
            CU(cudaDeviceSynchronize());
        }
        cudaIpcMemHandle_t h;
        CU(cudaIpcGetMemHandle(&h, exch_));
        static_assert(sizeof(h) == 64, "CUDA IPC handles are 64 bytes");
        std::memcpy(out, &h, 64);
    }
    void pair_connect(const unsigned char peer[64], int role) {
        require_ready();
        if (role != 0 && role != 1) API_THROW(LIBSDOD_INVALID_ARGUMENT, "role must be 0 (conditional half) or 1 (unconditional half)");
        if (!exch_) API_THROW(LIBSDOD_INVALID_ARGUMENT, "call libsdod_b200_pair_export first (the peer needs this context's handle too)");
        if (device_ >= 0) CU(cudaSetDevice(device_));
        if (peer_exch_) { cudaIpcCloseMemHandle(peer_exch_); peer_exch_ = nullptr; }
        cudaIpcMemHandle_t h;
        std::memcpy(&h, peer, 64);
        CU(cudaIpcOpenMemHandle(reinterpret_cast<void**>(&peer_exch_), h, cudaIpcMemLazyEnablePeerAccess));
        pair_role_ = role;
        pair_seq_ = 0;
        log_.log(LIBSDOD_LOG_INFO, "CFG split: this context runs the %s half; eps exchange over a peer-mapped buffer", role == 0 ? "conditional" : "unconditional");
    }
    void generate_pair(unsigned n, const float* cond_half, const float* latents, float guidance, unsigned char* images_out, unsigned* first_image,
                       unsigned* n_decoded, float* latents_out) {
        require_ready();
        if (!peer_exch_) API_THROW(LIBSDOD_INVALID_ARGUMENT, "pair not connected (libsdod_b200_pair_export / _connect)");
        if (n < 1 || static_cast<int>(n) > max_images_) API_THROW(LIBSDOD_INVALID_ARGUMENT, "n_images out of range (max_images = " + std::to_string(max_images_) + ")");
        if (!cond_half || !images_out || !first_image || !n_decoded) API_THROW(LIBSDOD_INVALID_ARGUMENT, "nullptr argument");
        if (guidance == 1.0f) API_THROW(LIBSDOD_INVALID_ARGUMENT, "guidance 1 skips the unconditional pass (context.cpp:359): nothing to split");
        if (sampler_ == LIBSDOD_B200_SAMPLER_PLMS) API_THROW(LIBSDOD_INVALID_ARGUMENT, "the pair loop supports DPM-Solver++ and DDIM");
        if (steps_ == 0) API_THROW(LIBSDOD_RUNTIME_ERROR, "schedule not prepared");
        if (device_ >= 0) CU(cudaSetDevice(device_));
        const int B = static_cast<int>(n);
        const size_t per = static_cast<size_t>(S_) * S_ * 4, lat = per * n, ctx1 = static_cast<size_t>(77) * 768;
        float* x = unet_->x_in();
        float* eps = unet_->eps_out();
        unsigned int* my_flags = reinterpret_cast<unsigned int*>(exch_);
        unsigned int* peer_flags = reinterpret_cast<unsigned int*>(peer_exch_);
        float* my_slots = reinterpret_cast<float*>(reinterpret_cast<char*>(exch_) + 256);
        float* peer_slots = reinterpret_cast<float*>(reinterpret_cast<char*>(peer_exch_) + 256);
        CU(cudaEventRecord(ev_[0], stream_));
        std::memcpy(pin_ctx_, cond_half, n * ctx1 * sizeof(float));
        CU(cudaMemcpyAsync(ctx_dev_, pin_ctx_, n * ctx1 * sizeof(float), cudaMemcpyHostToDevice, stream_));
        SD(unet_->set_context(stream_, ctx_dev_, SDOD_F32, B));
        CU(cudaEventRecord(ev_[1], stream_));
        if (latents) {
            for (unsigned i = 0; i < n; ++i)
                for (int c = 0; c < 4; ++c)
                    for (int p = 0; p < S_ * S_; ++p) pin_lat_[i * per + static_cast<size_t>(p) * 4 + c] = latents[i * per + static_cast<size_t>(c) * S_ * S_ + p];
            CU(cudaMemcpyAsync(x, pin_lat_, lat * sizeof(float), cudaMemcpyHostToDevice, stream_));
        } else {
            SD(sdod_randn(stream_, x, lat, seed_, rng_offset_));            // both ranks: same seed, same offsets -> same x_T
            rng_offset_ += (lat + 3) / 4;
        }
        for (unsigned step = 0; step < steps_; ++step) {
            SD(sdod::broadcast_rows(stream_, unet_->emb_in(), temb_ + static_cast<size_t>(step) * 1280, B, 1280));
            SD(unet_->forward(stream_, x, unet_->emb_in(), eps, B, true));
            const sdod::DpmStep k = sampler_ == LIBSDOD_B200_SAMPLER_DDIM ? ddim_.step(step) : sched_.step(step);
            ++pair_seq_;
            const unsigned par = pair_seq_ & 1u;
            SD(sdod_cfg_dpm_step_pair(stream_, x, y_prev_, eps, peer_slots + par * slot_floats(), my_slots + par * slot_floats(), peer_flags + par,
                                      my_flags + par, pair_done_, pair_seq_, pair_role_, lat, guidance, k.sigma_s, k.alpha_s, k.c_x, k.c_prev, k.c_y0, k.order));
        }
        CU(cudaEventRecord(ev_[2], stream_));
        // decode split by image: role 0 takes the first ceil(n/2) images, role 1 the rest
        const unsigned first = pair_role_ == 0 ? 0u : (n + 1) / 2, cnt = pair_role_ == 0 ? (n + 1) / 2 : n / 2;
        if (cnt) SD(vae_->decode(stream_, x + first * per, reinterpret_cast<uint8_t*>(pin_img_), nullptr, static_cast<int>(cnt), true));
        if (latents_out) CU(cudaMemcpyAsync(pin_lat_, x, lat * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        CU(cudaEventRecord(ev_[3], stream_));
        CU(cudaStreamSynchronize(stream_));
        if (cnt) std::memcpy(images_out + first * image_bytes(), pin_img_, cnt * image_bytes());
        *first_image = first; *n_decoded = cnt;
        if (latents_out) {
            for (unsigned i = 0; i < n; ++i)
                for (int c = 0; c < 4; ++c)
                    for (int p = 0; p < S_ * S_; ++p) latents_out[i * per + static_cast<size_t>(c) * S_ * S_ + p] = pin_lat_[i * per + static_cast<size_t>(p) * 4 + c];
        }
        float ms[3];
        for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&ms[i], ev_[i], ev_[i + 1]);
        timings_[0] = ms[0]; timings_[1] = ms[1] / steps_; timings_[2] = ms[2]; timings_[3] = ms[0] + ms[1] + ms[2];
        log_.log(LIBSDOD_LOG_INFO, "CFG-split generate: %u images, iteration %.3fms, decode of %u images %.3fms, total %.3fms", n, timings_[1], cnt, timings_[2], timings_[3]);
    }

    // context.cpp:292-403, batched over n images
    void generate(unsigned n, const float* cond, const float* uncond, const float* latents, float guidance, unsigned char* images_out,
                  float* latents_out, bool device_ptrs = false) {
        require_ready();
        if (n < 1 || static_cast<int>(n) > max_images_) API_THROW(LIBSDOD_INVALID_ARGUMENT, "n_images out of range (max_images = " + std::to_string(max_images_) + ")");
        if (!cond) API_THROW(LIBSDOD_INVALID_ARGUMENT, "cond is nullptr");
        if (!images_out) API_THROW(LIBSDOD_INVALID_ARGUMENT, "images_out is nullptr");
        if (steps_ == 0) API_THROW(LIBSDOD_RUNTIME_ERROR, "schedule not prepared");
        const bool cfg = !(guidance == 1.0f);                             // exact compare, context.cpp:359
        if (cfg && !uncond) API_THROW(LIBSDOD_INVALID_ARGUMENT, "uncond is nullptr with guidance != 1");
        if (device_ >= 0) CU(cudaSetDevice(device_));
        const int B = cfg ? 2 * static_cast<int>(n) : static_cast<int>(n);
        const size_t per = static_cast<size_t>(S_) * S_ * 4, lat = per * n, ctx1 = static_cast<size_t>(77) * 768;
        float* x = unet_->x_in();             // slots [0,n): cond batch (the sampler state), [n,2n): uncond copy
        float* eps = unet_->eps_out();

        CU(cudaEventRecord(ev_[0], stream_));
        // ---- conditioning (context.cpp:325-330): host -> pinned -> device, then cross-attention K/V projection
        if (device_ptrs) {
            CU(cudaMemcpyAsync(ctx_dev_, cond, n * ctx1 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
            if (cfg) CU(cudaMemcpyAsync(ctx_dev_ + n * ctx1, uncond, n * ctx1 * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
        } else {
            std::memcpy(pin_ctx_, cond, n * ctx1 * sizeof(float));
            if (cfg) std::memcpy(pin_ctx_ + n * ctx1, uncond, n * ctx1 * sizeof(float));
            CU(cudaMemcpyAsync(ctx_dev_, pin_ctx_, static_cast<size_t>(B) * ctx1 * sizeof(float), cudaMemcpyHostToDevice, stream_));
        }
        SD(unet_->set_context(stream_, ctx_dev_, SDOD_F32, B));
        CU(cudaEventRecord(ev_[1], stream_));
        // ---- x_T (context.cpp:333-334)
        if (latents && device_ptrs) {
            CU(cudaMemcpyAsync(x, latents, lat * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
        } else if (latents) {
            for (unsigned i = 0; i < n; ++i)                               // NCHW -> NHWC
                for (int c = 0; c < 4; ++c)
                    for (int p = 0; p < S_ * S_; ++p) pin_lat_[i * per + static_cast<size_t>(p) * 4 + c] = latents[i * per + static_cast<size_t>(c) * S_ * S_ + p];
            CU(cudaMemcpyAsync(x, pin_lat_, lat * sizeof(float), cudaMemcpyHostToDevice, stream_));
        } else {
            SD(sdod_randn(stream_, x, lat, seed_, rng_offset_));
            rng_offset_ += (lat + 3) / 4;
        }
        if (cfg) CU(cudaMemcpyAsync(x + lat, x, lat * sizeof(float), cudaMemcpyDeviceToDevice, stream_));
        // ---- denoising loop (context.cpp:342-382)
        if (sampler_ == LIBSDOD_B200_SAMPLER_PLMS) {
            plms_loop(x, eps, B, lat, cfg, guidance);
        } else
        for (unsigned step = 0; step < steps_; ++step) {
            SD(sdod::broadcast_rows(stream_, unet_->emb_in(), temb_ + static_cast<size_t>(step) * 1280, B, 1280));
            SD(unet_->forward(stream_, x, unet_->emb_in(), eps, B, true));
            const sdod::DpmStep k = sampler_ == LIBSDOD_B200_SAMPLER_DDIM ? ddim_.step(step) : sched_.step(step);
            SD(sdod_cfg_dpm_step(stream_, x, y_prev_, eps, cfg ? eps + lat : nullptr, SDOD_F32, lat, guidance, k.sigma_s, k.alpha_s, k.c_x,
                                 k.c_prev, k.c_y0, k.order, cfg ? x + lat : nullptr));
        }
        CU(cudaEventRecord(ev_[2], stream_));
        // ---- decode (context.cpp:386-395)
        SD(vae_->decode(stream_, x, device_ptrs ? images_out : reinterpret_cast<uint8_t*>(pin_img_), nullptr, static_cast<int>(n), true));
        if (latents_out) CU(cudaMemcpyAsync(pin_lat_, x, lat * sizeof(float), cudaMemcpyDeviceToHost, stream_));
        CU(cudaEventRecord(ev_[3], stream_));
        CU(cudaStreamSynchronize(stream_));
        if (!device_ptrs) std::memcpy(images_out, pin_img_, n * image_bytes());
        if (latents_out) {
            for (unsigned i = 0; i < n; ++i)
                for (int c = 0; c < 4; ++c)
                    for (int p = 0; p < S_ * S_; ++p) latents_out[i * per + static_cast<size_t>(c) * S_ * S_ + p] = pin_lat_[i * per + static_cast<size_t>(p) * 4 + c];
        }
        float ms[3];
        for (int i = 0; i < 3; ++i) cudaEventElapsedTime(&ms[i], ev_[i], ev_[i + 1]);
        timings_[0] = ms[0]; timings_[1] = ms[1] / steps_; timings_[2] = ms[2]; timings_[3] = ms[0] + ms[1] + ms[2];
        log_.log(LIBSDOD_LOG_INFO, "Conditioning took %.3fms", timings_[0]);                 // context.cpp:331
        log_.log(LIBSDOD_LOG_INFO, "Single iteration took %.3fms (mean of %u)", timings_[1], steps_);   // :381
        log_.log(LIBSDOD_LOG_INFO, "Decoding took %.3fms", timings_[2]);                     // :398
        log_.log(LIBSDOD_LOG_INFO, "Image successfully generated!");
        log_.log(LIBSDOD_LOG_INFO, "Image generation took %.3fms", timings_[3]);             // :402
    }

private:
    Logger log_;
    ErrorTable errors_;
    int S_, max_images_, device_;
    sdod::DpmSchedule sched_;
    sdod::DdimSchedule ddim_;
    float* plms_buf_ = nullptr;
    size_t plms_cap_ = 0;
    int sampler_ = LIBSDOD_B200_SAMPLER_DPM;
    unsigned steps_ = 0;
    std::unique_ptr<sdod::WeightStore> unet_w_, vae_w_, text_w_;
    std::unique_ptr<sdod::Tokenizer> tokenizer_;
    std::unique_ptr<sdod::ClipTextEncoder> text_;
    bool have_text_ = false;
    int* tok_dev_ = nullptr;
    std::unique_ptr<sdod::UNet> unet_;
    std::unique_ptr<sdod::VaeDecoder> vae_;
    cudaStream_t stream_ = nullptr;
    cudaEvent_t ev_[4] = {};
    float *y_prev_ = nullptr, *ctx_dev_ = nullptr, *temb_ = nullptr;
    float *pin_ctx_ = nullptr, *pin_lat_ = nullptr;
    unsigned char* pin_img_ = nullptr;
    std::vector<float> uncond_default_;
    unsigned long long seed_ = 0, rng_offset_ = 0;
    float timings_[4] = {0, 0, 0, 0};
    std::string models_dir_;
    bool ready_ = false;
    void* exch_ = nullptr;            // this context's exchange buffer (exported over CUDA IPC)
    void* peer_exch_ = nullptr;       // the peer's, mapped into this process
    unsigned int* pair_done_ = nullptr;
    int pair_role_ = 0;
    unsigned int pair_seq_ = 0;
};

struct Handle {                                  // reference libsdod.cpp:22-27
    unsigned magic_info = kMagic;
    unsigned context_version = kHandleVersion;
    unsigned ref_count = 0;
    Engine* cptr = nullptr;
};

int record(ErrorTable* tab, int code, const std::string& msg, const char* func, const char* file, int line) {   // libsdod.cpp:29-45
    const char* slash = std::strrchr(file, '/');
    if (slash) file = slash + 1;
    std::string full = std::string(func) + ": " + msg + " [" + file + ":" + std::to_string(line) + "]";
    (tab ? *tab : g_contextless_errors)[code] = std::move(full);
    return code;
}
#define ERR(tab, code, msg) record(tab, code, msg, __func__, __FILE__, __LINE__)

// libsdod.cpp:48-63 (TRY_RETRIEVE_CONTEXT)
int retrieve(void* context, Handle** hnd_out, const char* func) {
    auto bad = [&](const std::string& m) { return record(nullptr, LIBSDOD_INVALID_CONTEXT, m, func, __FILE__, __LINE__); };
    if (context == nullptr) return bad("context is nullptr");
    auto* hnd = reinterpret_cast<Handle*>(context);
    if (hnd->magic_info != kMagic) return bad("context magic header mismatch! got: " + std::to_string(hnd->magic_info));
    if (hnd->context_version != kHandleVersion) return bad("context version mismatch! got: " + std::to_string(hnd->context_version));
    if (hnd->ref_count == 0) return bad("context has been released!");
    if (hnd->cptr == nullptr) return bad("corrupted context, internal pointer is nullptr");
    *hnd_out = hnd;
    return LIBSDOD_NO_ERROR;
}

template <class F>
int guarded(ErrorTable* tab, const char* func, F&& f) {                // libsdod.cpp:102-108
    try {
        f();
    } catch (const ApiError& e) {
        return record(tab, e.code, e.msg, e.func, e.file, e.line);
    } catch (const std::bad_alloc&) {
        return record(tab, LIBSDOD_FAILED_ALLOCATION, "allocation failed", func, __FILE__, __LINE__);
    } catch (const std::exception& e) {
        return record(tab, LIBSDOD_INTERNAL_ERROR, e.what(), func, __FILE__, __LINE__);
    } catch (...) {
        return record(tab, LIBSDOD_INTERNAL_ERROR, "Unspecified error", func, __FILE__, __LINE__);
    }
    return LIBSDOD_NO_ERROR;
}

int setup_common(void** context, const char* models_dir, unsigned latent_spatial, unsigned steps, unsigned log_level, unsigned max_images,
                 int device, const char* func) {
    if (context == nullptr) return record(nullptr, LIBSDOD_INVALID_ARGUMENT, "Context argument should not be nullptr!", func, __FILE__, __LINE__);
    if (*context != nullptr) return record(nullptr, LIBSDOD_INVALID_ARGUMENT, "Context should point to a nullptr-initialized variable!", func, __FILE__, __LINE__);
    if (log_level > LIBSDOD_LOG_ABUSIVE) return record(nullptr, LIBSDOD_INVALID_ARGUMENT, "Invalid log_level", func, __FILE__, __LINE__);
    if (models_dir == nullptr) return record(nullptr, LIBSDOD_INVALID_ARGUMENT, "models_dir is nullptr", func, __FILE__, __LINE__);
    if (latent_spatial < 8 || latent_spatial > 128 || (latent_spatial & (latent_spatial - 1)) != 0)
        return record(nullptr, LIBSDOD_INVALID_ARGUMENT, "latent_spatial must be a power of two in [8, 128], got: " + std::to_string(latent_spatial), func, __FILE__, __LINE__);
    if (max_images < 1 || max_images > 128) return record(nullptr, LIBSDOD_INVALID_ARGUMENT, "max_images must be in [1, 128]", func, __FILE__, __LINE__);
    Handle* hnd = new (std::nothrow) Handle;
    if (hnd == nullptr) return record(nullptr, LIBSDOD_FAILED_ALLOCATION, "Could not create a new CAPI_Context_Handler object", func, __FILE__, __LINE__);
    hnd->ref_count += 1;
    hnd->cptr = new (std::nothrow) Engine(models_dir, latent_spatial, log_level, max_images, device);     // records its arguments only
    if (hnd->cptr == nullptr) {
        delete hnd;
        return record(nullptr, LIBSDOD_FAILED_ALLOCATION, "Could not create a new Context object", func, __FILE__, __LINE__);
    }
    // As in the reference (libsdod.cpp:77-87) the handle is published before initialisation can fail: on error *context is
    // non-NULL, the message sits in the context's own table (libsdod_get_last_error_extra_info(code, ctx)) and the caller must
    // still release the context (libsdod.h:42-45).
    *context = hnd;
    return guarded(&hnd->cptr->errors(), func, [&] { hnd->cptr->init(); hnd->cptr->prepare_schedule(steps); });
}

}  // namespace

extern "C" {

LIBSDOD_API int libsdod_setup(void** context, const char* models_dir, unsigned int latent_channels, unsigned int latent_spatial,
                              unsigned int upscale_factor, unsigned int steps, unsigned int log_level, int use_htp) {
    (void)use_htp;
    if (latent_channels != 4) return ERR(nullptr, LIBSDOD_INVALID_ARGUMENT, "latent_channels must be 4 (SD v1.x), got: " + std::to_string(latent_channels));
    if (upscale_factor != 8) return ERR(nullptr, LIBSDOD_INVALID_ARGUMENT, "upscale_factor must be 8 (SD v1.x), got: " + std::to_string(upscale_factor));
    return setup_common(context, models_dir, latent_spatial, steps, log_level, 1, -1, __func__);
}

LIBSDOD_API int libsdod_b200_setup(void** context, const char* models_dir, unsigned int latent_spatial, unsigned int steps,
                                   unsigned int log_level, unsigned int max_images, int device) {
    return setup_common(context, models_dir, latent_spatial, steps, log_level, max_images, device, __func__);
}

LIBSDOD_API int libsdod_set_steps(void* context, unsigned int steps) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    return guarded(&hnd->cptr->errors(), __func__, [&] { hnd->cptr->prepare_schedule(steps); });
}

LIBSDOD_API int libsdod_set_log_level(void* context, unsigned int log_level) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    if (log_level > LIBSDOD_LOG_ABUSIVE) return ERR(&hnd->cptr->errors(), LIBSDOD_INVALID_ARGUMENT, "Invalid log_level");
    hnd->cptr->logger().set_level(log_level);
    return LIBSDOD_NO_ERROR;
}

LIBSDOD_API int libsdod_ref_context(void* context) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    ++hnd->ref_count;
    return LIBSDOD_NO_ERROR;
}

LIBSDOD_API int libsdod_release(void* context) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    if (--hnd->ref_count == 0) {
        delete hnd->cptr;          // the shell stays allocated: stale handles are detected, not dereferenced (libsdod.cpp:154-158)
        hnd->cptr = nullptr;
    }
    return LIBSDOD_NO_ERROR;
}

LIBSDOD_API int libsdod_generate_image(void* context, const char* prompt, float guidance_scale, unsigned char** image_out,
                                       unsigned int* image_buffer_size) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    if (image_out == nullptr) return ERR(&e->errors(), LIBSDOD_INVALID_ARGUMENT, "image_out is nullptr");
    if (image_buffer_size == nullptr) return ERR(&e->errors(), LIBSDOD_INVALID_ARGUMENT, "image_buffer_size is nullptr");
    if (prompt == nullptr) return ERR(&e->errors(), LIBSDOD_INVALID_ARGUMENT, "prompt is nullptr");
    const size_t need = e->image_bytes();
    unsigned char* buf = *image_out;
    bool owned = false;
    if (buf == nullptr) {                                                  // context.cpp:406-409, buffer.cpp:12-18
        buf = static_cast<unsigned char*>(std::malloc(need));
        if (!buf) return ERR(&e->errors(), LIBSDOD_FAILED_ALLOCATION, "Could not allocate the output image");
        owned = true;
    } else if (*image_buffer_size < need) {                                // context.cpp:416-417
        return ERR(&e->errors(), LIBSDOD_INVALID_ARGUMENT, "Provided buffer is too small, missing " + std::to_string(need - *image_buffer_size) + " bytes");
    }
    int st = guarded(&e->errors(), __func__, [&] { e->generate_prompt(prompt, guidance_scale, buf); });
    if (st != LIBSDOD_NO_ERROR) {
        if (owned) std::free(buf);
        return st;
    }
    *image_out = buf;
    *image_buffer_size = static_cast<unsigned int>(need);
    return LIBSDOD_NO_ERROR;
}

LIBSDOD_API const char* libsdod_get_error_description(int errorcode) {
    if (errorcode < 0 || errorcode >= kNumErrors) return nullptr;
    return kErrorStrings[errorcode];
}

LIBSDOD_API const char* libsdod_get_last_error_extra_info(int errorcode, void* context) {   // libsdod.cpp:194-209
    if (errorcode < 0 || errorcode >= kNumErrors) return nullptr;
    ErrorTable* tab = &g_contextless_errors;
    if (context && errorcode != LIBSDOD_INVALID_CONTEXT) {
        auto* hnd = reinterpret_cast<Handle*>(context);
        if (hnd->magic_info == kMagic && hnd->context_version == kHandleVersion && hnd->ref_count > 0 && hnd->cptr != nullptr)
            tab = &hnd->cptr->errors();
    }
    auto& v = (*tab)[errorcode];
    return v ? v->c_str() : nullptr;
}

LIBSDOD_API int libsdod_b200_set_seed(void* context, unsigned long long seed) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    hnd->cptr->set_seed(seed);
    return LIBSDOD_NO_ERROR;
}

LIBSDOD_API int libsdod_b200_encode_prompt(void* context, const char* prompt, float* embedding_out, unsigned short* tokens_out) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    return guarded(&e->errors(), __func__, [&] {
        if (!prompt) API_THROW(LIBSDOD_INVALID_ARGUMENT, "prompt is nullptr");
        if (!embedding_out) API_THROW(LIBSDOD_INVALID_ARGUMENT, "embedding_out is nullptr");
        e->require_ready();
        e->prompt_embedding(prompt, embedding_out, tokens_out);
    });
}

LIBSDOD_API int libsdod_b200_generate_images(void* context, unsigned int n_images, const char* const* prompts, float guidance_scale, unsigned char* images_out) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    return guarded(&e->errors(), __func__, [&] {
        if (!images_out) API_THROW(LIBSDOD_INVALID_ARGUMENT, "images_out is nullptr");
        e->generate_prompts(n_images, prompts, guidance_scale, images_out);
    });
}

LIBSDOD_API int libsdod_b200_set_sampler(void* context, int sampler) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    return guarded(&e->errors(), __func__, [&] { e->set_sampler(sampler); });
}

LIBSDOD_API int libsdod_b200_generate(void* context, unsigned int n_images, const float* cond, const float* uncond, const float* latents,
                                      float guidance_scale, unsigned char* images_out, float* latents_out) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    return guarded(&e->errors(), __func__, [&] { e->generate(n_images, cond, uncond, latents, guidance_scale, images_out, latents_out); });
}

LIBSDOD_API int libsdod_b200_generate_device(void* context, unsigned int n_images, const float* cond_dev, const float* uncond_dev,
                                             const float* latents_nhwc_dev, float guidance_scale, unsigned char* images_out_dev) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    return guarded(&e->errors(), __func__, [&] { e->generate(n_images, cond_dev, uncond_dev, latents_nhwc_dev, guidance_scale, images_out_dev, nullptr, true); });
}

LIBSDOD_API int libsdod_b200_pair_export(void* context, unsigned char handle_out[64]) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    if (!handle_out) return ERR(&e->errors(), LIBSDOD_INVALID_ARGUMENT, "handle_out is nullptr");
    return guarded(&e->errors(), __func__, [&] { e->pair_export(handle_out); });
}

LIBSDOD_API int libsdod_b200_pair_connect(void* context, const unsigned char peer_handle[64], int role) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    if (!peer_handle) return ERR(&e->errors(), LIBSDOD_INVALID_ARGUMENT, "peer_handle is nullptr");
    return guarded(&e->errors(), __func__, [&] { e->pair_connect(peer_handle, role); });
}

LIBSDOD_API int libsdod_b200_generate_pair(void* context, unsigned int n_images, const float* conditioning_half, const float* latents,
                                           float guidance_scale, unsigned char* images_out, unsigned int* first_image, unsigned int* n_decoded,
                                           float* latents_out) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    Engine* e = hnd->cptr;
    return guarded(&e->errors(), __func__, [&] { e->generate_pair(n_images, conditioning_half, latents, guidance_scale, images_out, first_image, n_decoded, latents_out); });
}

LIBSDOD_API int libsdod_b200_last_timings(void* context, float out[4]) {
    Handle* hnd = nullptr;
    if (int st = retrieve(context, &hnd, __func__)) return st;
    if (!out) return ERR(&hnd->cptr->errors(), LIBSDOD_INVALID_ARGUMENT, "out is nullptr");
    for (int i = 0; i < 4; ++i) out[i] = hnd->cptr->timings()[i];
    return LIBSDOD_NO_ERROR;
}

}  // extern "C"
