#pragma once
#include "net.h"

namespace sdod {

class UNet : public NetBase {
public:
    UNet(const WeightStore* ws, unsigned long long seed, int latent_hw, int max_batch);
    ~UNet() override;

    int time_embed(cudaStream_t s, const float* t, int n, float* out);
    int set_context(cudaStream_t s, const void* context, int dtype, int B);
    int forward(cudaStream_t s, const float* x, const float* emb, float* eps, int B, bool use_graph);
    unsigned long long launches_per_forward(int B);
    int profile_forward(cudaStream_t s, int B, int iters, std::vector<std::pair<std::string, float>>* out);

    // persistent I/O buffers (fp32): callers may write/read these directly to skip the staging copies
    float* x_in() { return x_in_; }
    float* emb_in() { return emb_in_; }
    float* eps_out() { return eps_out_; }
    int latent_hw() const { return hw_; }
    int max_batch() const { return max_batch_; }

private:
    struct LevelBufs { void *qh, *kh, *vt; };
    struct CtxBufs { void *kh, *vt; int C; };

    Act res_block(const Act& x, const Act* x_cat, const std::string& prefix, int cout, int emb_index);
    Act spatial_transformer(const Act& x, const std::string& prefix, int level);
    void attention_op(const void* qh, const void* kh, const void* vt, void* out, int B, int tokens, int n_kv, int C);
    CtxBufs& ctx_bufs(const std::string& prefix, int C);
    std::unique_ptr<Plan> build_forward(int B);
    std::unique_ptr<Plan> build_context(int B);
    std::unique_ptr<Plan> build_time_embed(int n);
    Plan* forward_plan(int B);

    int hw_, max_batch_;
    std::vector<std::string> emb_names_;
    std::vector<int> emb_couts_, emb_offsets_;
    int emb_total_ = 0;
    float *x_in_ = nullptr, *emb_in_ = nullptr, *eps_out_ = nullptr, *emb_proj_ = nullptr;
    float *temb_sin_ = nullptr, *temb_out_ = nullptr;
    void* ctx_bf16_ = nullptr;
    std::vector<LevelBufs> level_bufs_;
    std::map<std::string, CtxBufs> ctx_;
    std::vector<std::string> ctx_order_;
    std::map<int, std::unique_ptr<Plan>> fwd_, ctxp_, temb_;
};

}  // namespace sdod
