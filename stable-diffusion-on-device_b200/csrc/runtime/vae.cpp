// SD-v1 VAE decoder (AutoencoderKL.decode) as a pre-planned launch sequence, NHWC bf16.
//
// Replaces the reference's opaque `decoder` graph + host post-processing (csrc/libsdod/src/context.cpp:
// 386-395): z/0.18215 -> post_quant_conv -> Decoder -> clamp((x+1)/2,0,1) -> uint8(clamp(255 f,0,255)),
// output [H,W,C] (api/libsdod.h:89).  Layer names = ldm state_dict keys (first_stage_model.* stripped).
// The single-head d=512 mid-block attention runs unfused (QK^T GEMM -> row softmax -> PV GEMM): its
// O accumulator (128x512 fp32) would fill TMEM on its own and it is 1.4 % of the decoder FLOPs.
#include "vae.h"

#include <algorithm>

#include <cmath>
#include <stdexcept>

#include "../kernels/glue.h"

namespace sdod {

namespace {
constexpr int kCh = 128;
const int kMult[4] = {1, 2, 4, 4};
constexpr float kLatentScale = 0.18215f;
}  // namespace

VaeDecoder::VaeDecoder(const WeightStore* ws, unsigned long long seed, int latent_hw, int max_batch)
    : NetBase(ws, seed), hw_(latent_hw), max_batch_(max_batch) {
    if (latent_hw < 8 || (latent_hw & (latent_hw - 1)) != 0) throw std::runtime_error("vae: latent_hw must be a power of two >= 8");
    const size_t mb = static_cast<size_t>(max_batch), px = static_cast<size_t>(8 * hw_) * 8 * hw_ * 3;
    z_in_ = static_cast<float*>(dev_alloc(mb * hw_ * hw_ * 4 * sizeof(float), true));
    conv_out_ = static_cast<float*>(dev_alloc(mb * px * sizeof(float), false));
    u8_out_ = static_cast<uint8_t*>(dev_alloc(mb * px, false));
    img_out_ = static_cast<float*>(dev_alloc(mb * px * sizeof(float), false));
}

VaeDecoder::~VaeDecoder() = default;

Act VaeDecoder::res_block(const Act& x, const std::string& prefix, int cout) {
    Act h1 = gn(x, prefix + ".norm1", 1e-6f, true);
    Act h2 = conv3(h1, prefix + ".conv1", cout, nullptr, 0, nullptr);
    release(h1);
    Act h3 = gn(h2, prefix + ".norm2", 1e-6f, true);
    release(h2);
    Act out;
    if (x.C != cout) {
        Act s = conv1x1(x, prefix + ".nin_shortcut", cout, nullptr);
        out = conv3(h3, prefix + ".conv2", cout, nullptr, 0, &s);
        release(s);
    } else {
        out = conv3(h3, prefix + ".conv2", cout, nullptr, 0, &x);
    }
    release(h3);
    return out;
}

Act VaeDecoder::attn_block(const Act& x, const std::string& prefix) {
    const int C = x.C, N = x.H * x.W, B = x.B;
    Act hn = gn(x, prefix + ".norm", 1e-6f, false);
    Act q = conv1x1(hn, prefix + ".q", C, nullptr);
    Act k = conv1x1(hn, prefix + ".k", C, nullptr);
    // V^T [B, C, N] via the HEADS_T epilogue (heads = 1)
    Act vt = new_act(B, C, 1, N);
    {
        sdod_gemm_desc d{};
        d.A = hn.p; d.lda = C; d.W = pack_linear(prefix + ".v.weight", C, C); d.ldw = C; d.M = hn.M(); d.N = C; d.K = C; d.batch = 1;
        d.epi.C = vt.p; d.epi.bias = w32(prefix + ".v.bias", {C}, kInitBias); d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_HEADS_T;
        d.epi.heads = 1; d.epi.head_dim = C; d.epi.tokens = N; d.epi.tok_pad = N; d.epi.vt_rows = C;
        gemm_into(d);
    }
    release(hn);
    // S[b] = scale * q[b] k[b]^T   -> bf16 [B, N, N]
    Act s = new_act(B, N, 1, N);
    {
        sdod_gemm_desc d{};
        d.A = q.p; d.lda = C; d.strideA = static_cast<long long>(N) * C;
        d.W = k.p; d.ldw = C; d.strideW = static_cast<long long>(N) * C;
        d.M = N; d.N = N; d.K = C; d.batch = B;
        d.epi.C = s.p; d.epi.ldc = N; d.epi.strideC = static_cast<long long>(N) * N;
        d.epi.alpha = 1.0f / std::sqrt(static_cast<float>(C)); d.epi.out_mode = SDOD_OUT_BF16;
        gemm_into(d);
    }
    release(q);
    release(k);
    Act p = new_act(B, N, 1, N);
    {
        const void* sp = s.p;
        void* pp = p.p;
        const long long rows = static_cast<long long>(B) * N;
        plan_->push([=](cudaStream_t st) { return sdod_softmax_rows(st, sp, pp, rows, N, N, 1.0f); });
    }
    release(s);
    // O[b] = P[b] V[b]  (W operand = V^T[b], K-major)
    Act o = new_act(B, x.H, x.W, C);
    {
        sdod_gemm_desc d{};
        d.A = p.p; d.lda = N; d.strideA = static_cast<long long>(N) * N;
        d.W = vt.p; d.ldw = N; d.strideW = static_cast<long long>(C) * N;
        d.M = N; d.N = C; d.K = N; d.batch = B;
        d.epi.C = o.p; d.epi.ldc = C; d.epi.strideC = static_cast<long long>(N) * C;
        d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_BF16;
        gemm_into(d);
    }
    release(p);
    release(vt);
    Act out = conv1x1(o, prefix + ".proj_out", C, &x);
    release(o);
    return out;
}

std::unique_ptr<Plan> VaeDecoder::build(int B) {
    auto plan = std::make_unique<Plan>();
    begin_plan(plan.get(), false);
    Act z0 = new_act(B, hw_, hw_, 4);
    {
        const float* zin = z_in_;
        void* zp = z0.p;
        const float* w = w32("post_quant_conv.weight", {4, 4, 1, 1}, kInitWeight);
        const float* b = w32("post_quant_conv.bias", {4}, kInitBias);
        const size_t rows = static_cast<size_t>(B) * hw_ * hw_;
        plan_->push([=](cudaStream_t st) { return latent_prequant(st, zin, zp, rows, w, b, 1.0f / kLatentScale); });
    }
    int ch = kCh * kMult[3];
    Act h = conv3_im2col(z0, "decoder.conv_in", ch, 1);
    release(z0);
    {
        Act a = res_block(h, "decoder.mid.block_1", ch); release(h);
        Act b = attn_block(a, "decoder.mid.attn_1"); release(a);
        h = res_block(b, "decoder.mid.block_2", ch); release(b);
    }
    for (int level = 3; level >= 0; --level) {
        const int cout = kCh * kMult[level];
        for (int i = 0; i < 3; ++i) {
            Act r = res_block(h, "decoder.up." + std::to_string(level) + ".block." + std::to_string(i), cout);
            release(h);
            h = r;
        }
        if (level != 0) {
            Act c = conv3_up2(h, "decoder.up." + std::to_string(level) + ".upsample.conv", h.C);     // nearest 2x + conv3x3, sub-pixel form
            release(h);
            h = c;
        }
    }
    Act hn = gn(h, "decoder.norm_out", 1e-6f, true);
    release(h);
    conv3(hn, "decoder.conv_out", 3, nullptr, 0, nullptr, false, conv_out_);
    release(hn);
    {
        const float* co = conv_out_;
        uint8_t* u8 = u8_out_;
        float* img = img_out_;
        const size_t n = static_cast<size_t>(B) * 8 * hw_ * 8 * hw_ * 3;
        plan_->push([=](cudaStream_t st) { return vae_post(st, co, u8, img, n); });
    }
    end_plan();
    return plan;
}

int VaeDecoder::decode(cudaStream_t s, const float* z, uint8_t* image_u8, float* image_f32, int B, bool use_graph) {
    if (B < 1 || B > max_batch_) return fail(kInvalidArgument, "vae decode: batch exceeds max_batch");
    // Large batches are decoded in chunks: at 512x512 one image is 2,048 row tiles of the last level's convolutions, and a launch addresses at
    // most 65,535 row tiles (gridDim.y) — 32 images per launch would be 65,536.  16 images per chunk keeps every level multi-wave anyway.
    int chunk_max = 1;
    while (2LL * chunk_max * 64 * hw_ * hw_ / 128 <= 65535 && 2 * chunk_max <= max_batch_) chunk_max *= 2;     // largest power of two that fits
    try {
        const size_t lat1 = static_cast<size_t>(hw_) * hw_ * 4, img1 = static_cast<size_t>(8) * hw_ * 8 * hw_ * 3;
        for (int b0 = 0; b0 < B; b0 += chunk_max) {
            const int nb = std::min(chunk_max, B - b0);
            auto it = plans_.find(nb);
            if (it == plans_.end()) it = plans_.emplace(nb, build(nb)).first;
            const float* zc = z + b0 * lat1;
            if (zc != z_in_) SDOD_TRY(check_cuda(cudaMemcpyAsync(z_in_, zc, nb * lat1 * sizeof(float), cudaMemcpyDeviceToDevice, s), "copy z"));
            SDOD_TRY(it->second->run(s, use_graph));
            if (image_u8) SDOD_TRY(check_cuda(cudaMemcpyAsync(image_u8 + b0 * img1, u8_out_, nb * img1, cudaMemcpyDefault, s), "copy image u8"));
            if (image_f32) SDOD_TRY(check_cuda(cudaMemcpyAsync(image_f32 + b0 * img1, img_out_, nb * img1 * sizeof(float), cudaMemcpyDefault, s), "copy image f32"));
        }
        return kOk;
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("vae decode: ") + e.what());
    }
}

}  // namespace sdod
