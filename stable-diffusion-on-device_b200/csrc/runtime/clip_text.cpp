// CLIP text encoder plan — see clip_text.h.  Weight names are the HuggingFace CLIPTextModel state_dict keys ("text_model...."), which is what an
// SD v1 checkpoint holds under "cond_stage_model.transformer." (sdod/checkpoint.py strips that prefix).
#include "clip_text.h"

#include <cmath>
#include <stdexcept>

#include "../kernels/glue.h"

namespace sdod {

namespace {
int pad8(int n) { return (n + 7) / 8 * 8; }
}

ClipTextEncoder::ClipTextEncoder(const WeightStore* ws, unsigned long long seed, int max_batch) : NetBase(ws, seed), max_batch_(max_batch) {
    if (max_batch < 1 || max_batch > 256) throw std::runtime_error("text encoder: max_batch out of range");
    const size_t mb = static_cast<size_t>(max_batch);
    tok_in_ = static_cast<int*>(dev_alloc(mb * kTokens * sizeof(int), true));
    out_bf16_ = dev_alloc(mb * kTokens * kWidth * 2, true);
    qh_ = dev_alloc(mb * kHeads * kTokens * kHeadDim * 2, true);
    kh_ = dev_alloc(mb * kHeads * kTokens * kHeadDim * 2, true);
    vt_ = dev_alloc(mb * kHeads * kHeadDim * pad8(kTokens) * 2, true);
}

ClipTextEncoder::~ClipTextEncoder() = default;

std::unique_ptr<Plan> ClipTextEncoder::build(int B) {
    auto plan = std::make_unique<Plan>();
    begin_plan(plan.get(), false);
    const std::string tm = "text_model.";
    // token + position embeddings -> fp32 residual stream [B*77, 768]
    Act h = new_act(B, kTokens, 1, kWidth, true);
    {
        const float* te = w32(tm + "embeddings.token_embedding.weight", {kVocab, kWidth}, kInitBias);      // N(0, 0.02) like CLIP's init
        const float* pe = w32(tm + "embeddings.position_embedding.weight", {kTokens, kWidth}, kInitBias);
        const int* tok = tok_in_;
        float* hp = static_cast<float*>(h.p);
        const int rows = B * kTokens;
        plan_->push([=](cudaStream_t st) { return clip_embed(st, tok, te, pe, hp, rows, kTokens, kWidth, kVocab); }, 1, "clip_embed");
    }
    for (int l = 0; l < kLayers; ++l) {
        const std::string p = tm + "encoder.layers." + std::to_string(l);
        Act n1 = ln(h, p + ".layer_norm1");
        {
            void* wqkv = pack_concat(p + ".qkv", {p + ".self_attn.q_proj.weight", p + ".self_attn.k_proj.weight", p + ".self_attn.v_proj.weight"},
                                     {kWidth, kWidth, kWidth}, kWidth);
            sdod_gemm_desc d{};
            d.A = n1.p; d.lda = kWidth; d.W = wqkv; d.ldw = kWidth; d.M = n1.M(); d.N = 3 * kWidth; d.K = kWidth; d.batch = 1;
            d.epi.C = qh_; d.epi.C2 = kh_; d.epi.C3 = vt_;
            d.epi.bias = concat_bias(p + ".qkv", {p + ".self_attn.q_proj.bias", p + ".self_attn.k_proj.bias", p + ".self_attn.v_proj.bias"}, {kWidth, kWidth, kWidth});
            d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_QKV;
            d.epi.heads = kHeads; d.epi.head_dim = kHeadDim; d.epi.tokens = kTokens; d.epi.dpad = kHeadDim; d.epi.tok_pad = pad8(kTokens); d.epi.vt_rows = kHeadDim;
            gemm_into(d);
        }
        release(n1);
        Act a = new_act(B, kTokens, 1, kWidth);
        {
            auto al = std::make_shared<AttnLaunch>();
            check(attention_prepare(al.get(), qh_, kh_, vt_, a.p, B, kHeads, kTokens, kTokens, kHeadDim, kHeadDim, pad8(kTokens), 1.0f / std::sqrt(static_cast<float>(kHeadDim))));
            al->causal = 1;
            plan_->push([al](cudaStream_t st) { return attention_launch(*al, st); }, 1, "attn causal Nq77 Nkv77 d64");
        }
        LinearOpts oo;
        oo.bias = w32(p + ".self_attn.out_proj.bias", {kWidth}, kInitBias);
        oo.residual = &h;
        oo.out_f32 = true;
        Act h2 = linear(a, pack_linear(p + ".self_attn.out_proj.weight", kWidth, kWidth), kWidth, oo);
        release(a);
        release(h);
        Act n2 = ln(h2, p + ".layer_norm2");
        LinearOpts o1;
        o1.bias = w32(p + ".mlp.fc1.bias", {kMlp}, kInitBias);
        o1.act = SDOD_ACT_QUICK_GELU;
        Act f = linear(n2, pack_linear(p + ".mlp.fc1.weight", kMlp, kWidth), kMlp, o1);
        release(n2);
        LinearOpts o2;
        o2.bias = w32(p + ".mlp.fc2.bias", {kWidth}, kInitBias);
        o2.residual = &h2;
        o2.out_f32 = true;
        Act h3 = linear(f, pack_linear(p + ".mlp.fc2.weight", kWidth, kMlp), kWidth, o2);
        release(f);
        release(h2);
        h = h3;
    }
    {
        const float* w = w32(tm + "final_layer_norm.weight", {kWidth}, kInitOnes);
        const float* b = w32(tm + "final_layer_norm.bias", {kWidth}, kInitZeros);
        const void* xp = h.p;
        void* yp = out_bf16_;
        const int rows = h.M();
        plan_->push([=](cudaStream_t st) { return sdod_layer_norm(st, xp, SDOD_F32, yp, w, b, rows, kWidth, 1e-5f); }, 1, "ln final");
    }
    release(h);
    end_plan();
    return plan;
}

int ClipTextEncoder::forward(cudaStream_t s, const int* tokens, int B, void* out, int out_dtype) {
    if (B < 1 || B > max_batch_) return fail(kInvalidArgument, "text encoder: batch exceeds max_batch");
    if (!tokens || !out) return fail(kInvalidArgument, "text encoder: NULL tensor");
    if (out_dtype != SDOD_F32 && out_dtype != SDOD_BF16) return fail(kInvalidArgument, "text encoder: out dtype must be f32 or bf16");
    try {
        auto it = plans_.find(B);
        if (it == plans_.end()) it = plans_.emplace(B, build(B)).first;
        const size_t n = static_cast<size_t>(B) * kTokens;
        if (tokens != tok_in_) SDOD_TRY(check_cuda(cudaMemcpyAsync(tok_in_, tokens, n * sizeof(int), cudaMemcpyDeviceToDevice, s), "copy tokens"));
        SDOD_TRY(it->second->run(s, false));
        if (out_dtype == SDOD_BF16) return check_cuda(cudaMemcpyAsync(out, out_bf16_, n * kWidth * 2, cudaMemcpyDeviceToDevice, s), "copy context");
        return cast_bf16_to_f32(s, out_bf16_, static_cast<float*>(out), n * kWidth);
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("text encoder forward: ") + e.what());
    }
}

}  // namespace sdod
