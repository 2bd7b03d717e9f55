#pragma once
#include "net.h"

namespace sdod {

// CLIP ViT-L/14 text encoder (the conditioning model of SD v1.x) as a pre-planned launch sequence on the GEMM / attention / LayerNorm kernels.
// Replaces the reference's `cond_model` graph ("text_encoder.serialized": context.cpp:143,170), executed once per prompt (context.cpp:327) and once
// at setup for the cached empty prompt (context.cpp:233-239).  Architecture = the public CLIPTextModel of openai/clip-vit-large-patch14: 49408 x 768
// token embedding + 77 learned positions, 12 pre-LN layers (12 heads of 64, causal self-attention, MLP 768 -> 3072 -> 768 with QuickGELU),
// final LayerNorm; the output is last_hidden_state [B, 77, 768].
class ClipTextEncoder : public NetBase {
public:
    static constexpr int kTokens = 77, kWidth = 768, kHeads = 12, kHeadDim = 64, kLayers = 12, kMlp = 3072, kVocab = 49408;
    ClipTextEncoder(const WeightStore* ws, unsigned long long seed, int max_batch);
    ~ClipTextEncoder() override;
    // tokens [B,77] int32 (device) -> out [B,77,768] fp32 or bf16 (device)
    int forward(cudaStream_t s, const int* tokens, int B, void* out, int out_dtype);
    int max_batch() const { return max_batch_; }

private:
    std::unique_ptr<Plan> build(int B);
    int max_batch_;
    int* tok_in_ = nullptr;          // [maxB, 77]
    void* out_bf16_ = nullptr;       // [maxB, 77, 768]
    void *qh_ = nullptr, *kh_ = nullptr, *vt_ = nullptr;      // attention operand layouts (zero-filled once; pad lanes never written)
    std::map<int, std::unique_ptr<Plan>> plans_;
};

}  // namespace sdod
