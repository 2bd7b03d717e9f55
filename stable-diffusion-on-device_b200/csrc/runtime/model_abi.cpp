// extern "C" surface of include/sdod_model.h
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "sdod_model.h"
#include "clip_text.h"
#include "tokenizer.h"
#include "unet.h"
#include "vae.h"

struct sdod_weights { sdod::WeightStore store; };
struct sdod_unet { sdod::UNet* net; };
struct sdod_vae { sdod::VaeDecoder* net; };
struct sdod_tokenizer { sdod::Tokenizer tok; };
struct sdod_text_encoder { sdod::ClipTextEncoder* net; };

using sdod::fail;
using sdod::kCudaError;
using sdod::kInvalidArgument;
using sdod::kOk;

static int have_device() {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
        cudaGetLastError();
        return fail(sdod::kNoDevice, "no CUDA device: libsdod_b200 has no CPU fallback");
    }
    return kOk;
}

extern "C" {

SDOD_API int sdod_weights_create(sdod_weights** out) {
    if (!out) return fail(kInvalidArgument, "weights_create: out is NULL");
    *out = new (std::nothrow) sdod_weights;
    return *out ? kOk : fail(kInvalidArgument, "weights_create: allocation failed");
}
SDOD_API int sdod_weights_set_f32(sdod_weights* w, const char* name, const float* host_data, int ndim, const long long* shape) {
    if (!w || !name || !host_data || ndim < 0 || ndim > 8 || (ndim && !shape)) return fail(kInvalidArgument, "weights_set_f32: bad arguments");
    SDOD_TRY(have_device());
    return w->store.set(name, host_data, std::vector<long long>(shape, shape + ndim));
}
SDOD_API int sdod_weights_load_file(sdod_weights* w, const char* path) {
    if (!w || !path) return fail(kInvalidArgument, "weights_load_file: bad arguments");
    SDOD_TRY(have_device());
    return w->store.load_file(path);
}
SDOD_API long long sdod_weights_count(const sdod_weights* w) { return w ? static_cast<long long>(w->store.size()) : -1; }
SDOD_API void sdod_weights_destroy(sdod_weights* w) { delete w; }

SDOD_API int sdod_unet_create(sdod_unet** out, const sdod_weights* weights, unsigned long long seed, int latent_hw, int max_batch) {
    if (!out) return fail(kInvalidArgument, "unet_create: out is NULL");
    *out = nullptr;
    SDOD_TRY(have_device());
    try {
        auto* u = new sdod_unet{new sdod::UNet(weights ? &weights->store : nullptr, seed, latent_hw, max_batch)};
        *out = u;
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("unet_create: ") + e.what());
    }
    return kOk;
}
SDOD_API void sdod_unet_destroy(sdod_unet* u) {
    if (u) { delete u->net; delete u; }
}
SDOD_API int sdod_unet_time_embed(sdod_unet* u, sdod_stream_t stream, const float* t, int n, float* emb_out) {
    if (!u || !t || !emb_out) return fail(kInvalidArgument, "unet_time_embed: bad arguments");
    return u->net->time_embed(static_cast<cudaStream_t>(stream), t, n, emb_out);
}
SDOD_API int sdod_unet_set_context(sdod_unet* u, sdod_stream_t stream, const void* context, int dtype, int B) {
    if (!u || !context) return fail(kInvalidArgument, "unet_set_context: bad arguments");
    return u->net->set_context(static_cast<cudaStream_t>(stream), context, dtype, B);
}
SDOD_API int sdod_unet_forward(sdod_unet* u, sdod_stream_t stream, const float* x, const float* emb, float* eps, int B, int use_graph) {
    if (!u || !x || !emb || !eps) return fail(kInvalidArgument, "unet_forward: bad arguments");
    return u->net->forward(static_cast<cudaStream_t>(stream), x, emb, eps, B, use_graph != 0);
}
SDOD_API unsigned long long sdod_unet_launches_per_forward(const sdod_unet* u, int B) {
    return u ? u->net->launches_per_forward(B) : 0;
}

SDOD_API int sdod_unet_profile(sdod_unet* u, sdod_stream_t stream, int B, int iters, char* buf, size_t buf_bytes) {
    if (!u || !buf || buf_bytes == 0) return fail(kInvalidArgument, "unet_profile: bad arguments");
    std::vector<std::pair<std::string, float>> rows;
    SDOD_TRY(u->net->profile_forward(static_cast<cudaStream_t>(stream), B, iters, &rows));
    std::string out;
    for (auto& r : rows) out += std::to_string(r.second) + "\t" + r.first + "\n";
    if (out.size() + 1 > buf_bytes) out.resize(buf_bytes - 1);
    std::memcpy(buf, out.c_str(), out.size() + 1);
    return kOk;
}

SDOD_API int sdod_vae_create(sdod_vae** out, const sdod_weights* weights, unsigned long long seed, int latent_hw, int max_batch) {
    if (!out) return fail(kInvalidArgument, "vae_create: out is NULL");
    *out = nullptr;
    SDOD_TRY(have_device());
    try {
        auto* v = new sdod_vae{new sdod::VaeDecoder(weights ? &weights->store : nullptr, seed, latent_hw, max_batch)};
        *out = v;
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("vae_create: ") + e.what());
    }
    return kOk;
}
SDOD_API void sdod_vae_destroy(sdod_vae* v) {
    if (v) { delete v->net; delete v; }
}
SDOD_API int sdod_vae_decode(sdod_vae* v, sdod_stream_t stream, const float* z, uint8_t* image_u8, float* image_f32, int B, int use_graph) {
    if (!v || !z) return fail(kInvalidArgument, "vae_decode: bad arguments");
    return v->net->decode(static_cast<cudaStream_t>(stream), z, image_u8, image_f32, B, use_graph != 0);
}

SDOD_API int sdod_tokenizer_create(sdod_tokenizer** out, const char* bpe_file) {
    if (!out) return fail(kInvalidArgument, "tokenizer_create: out is NULL");
    *out = nullptr;
    try {
        *out = bpe_file ? new sdod_tokenizer{sdod::Tokenizer(std::string(bpe_file))} : new sdod_tokenizer{sdod::Tokenizer()};
    } catch (const std::exception& e) {
        return fail(kInvalidArgument, std::string("tokenizer_create: ") + e.what());
    }
    return kOk;
}
SDOD_API void sdod_tokenizer_destroy(sdod_tokenizer* t) { delete t; }
SDOD_API int sdod_tokenizer_encode(const sdod_tokenizer* t, const char* utf8, unsigned short* tokens_out, unsigned context_len, int* deviated) {
    if (!t || !utf8 || !tokens_out || context_len < 2) return fail(kInvalidArgument, "tokenizer_encode: bad arguments");
    try {
        bool dev = false;
        const std::vector<sdod::Tokenizer::token_type> ids = t->tok.encode(utf8, context_len, &dev);
        std::memcpy(tokens_out, ids.data(), sizeof(unsigned short) * context_len);
        if (deviated) *deviated = dev ? 1 : 0;
    } catch (const std::exception& e) {
        return fail(kInvalidArgument, std::string("tokenizer_encode: ") + e.what());
    }
    return kOk;
}
SDOD_API int sdod_tokenizer_vocab_size(const sdod_tokenizer* t) { return t ? static_cast<int>(t->tok.vocab_size()) : -1; }

SDOD_API int sdod_text_encoder_create(sdod_text_encoder** out, const sdod_weights* weights, unsigned long long seed, int max_batch) {
    if (!out) return fail(kInvalidArgument, "text_encoder_create: out is NULL");
    *out = nullptr;
    SDOD_TRY(have_device());
    try {
        *out = new sdod_text_encoder{new sdod::ClipTextEncoder(weights ? &weights->store : nullptr, seed, max_batch)};
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("text_encoder_create: ") + e.what());
    }
    return kOk;
}
SDOD_API void sdod_text_encoder_destroy(sdod_text_encoder* e) {
    if (e) { delete e->net; delete e; }
}
SDOD_API int sdod_text_encoder_forward(sdod_text_encoder* e, sdod_stream_t stream, const int* tokens, int B, void* context_out, int out_dtype) {
    if (!e) return fail(kInvalidArgument, "text_encoder_forward: bad arguments");
    return e->net->forward(static_cast<cudaStream_t>(stream), tokens, B, context_out, out_dtype);
}
}
