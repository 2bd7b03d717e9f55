// Runtime plumbing shared by the UNet and VAE executors: named device weights, a device buffer pool,
// pre-planned launch sequences (optionally replayed as a CUDA graph) and the layer builders that
// turn ldm layer names into prepared kernel launches.
#pragma once
#include <cuda_runtime.h>

#include <functional>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#include "../host_common.h"
#include "../kernels/launchers.h"
#include "sdod_kernels.h"

namespace sdod {

struct DevTensor {
    float* p = nullptr;
    std::vector<long long> shape;
    size_t numel = 0;
};

class WeightStore {
public:
    ~WeightStore();
    int set(const std::string& name, const float* host, const std::vector<long long>& shape);
    int load_file(const std::string& path);
    const DevTensor* find(const std::string& name) const;
    size_t size() const { return t_.size(); }

private:
    std::unordered_map<std::string, DevTensor> t_;
};

// Caching pool of device buffers; plan-build-time alloc/release gives buffer reuse across layers
// (stream order makes reuse safe), which keeps the step's working set L2-friendly.
class Pool {
public:
    ~Pool();
    void* get(size_t bytes);
    void put(void* p);
    size_t total_bytes() const { return total_; }

private:
    struct Blk { void* p; size_t sz; bool free_; };
    std::vector<Blk> blks_;
    size_t total_ = 0;
};

struct Act {
    void* p = nullptr;
    int B = 0, H = 0, W = 0, C = 0;
    bool f32 = false;            // fp32 residual-stream tensor (otherwise bf16 GEMM operand)
    int M() const { return B * H * W; }
    size_t bytes() const { return static_cast<size_t>(M()) * C * (f32 ? 4 : 2); }
    int dtype() const { return f32 ? SDOD_F32 : SDOD_BF16; }
};

class Plan {
public:
    ~Plan();
    void push(std::function<int(cudaStream_t)> f, unsigned launches = 1, std::string name = "op") {
        ops_.push_back(std::move(f));
        names_.push_back(std::move(name));
        launches_ += launches;
    }
    // Eager pass with a CUDA event between consecutive ops (hot caches, real launch gaps): per-op milliseconds, averaged.
    int profile(cudaStream_t s, int iters, std::vector<std::pair<std::string, float>>* out);
    int run(cudaStream_t s, bool use_graph);
    unsigned long long launches() const { return launches_; }
    size_t size() const { return ops_.size(); }

private:
    int run_eager(cudaStream_t s);
    std::vector<std::function<int(cudaStream_t)>> ops_;
    std::vector<std::string> names_;
    unsigned long long launches_ = 0;
    bool warmed_ = false;
    cudaGraphExec_t exec_ = nullptr;
};

enum InitKind { kInitWeight = 0, kInitBias = 1, kInitOnes = 2, kInitZeros = 3 };

class NetBase {
public:
    NetBase(const WeightStore* ws, unsigned long long seed);
    virtual ~NetBase();

protected:
    // ---- weights
    const float* w32(const std::string& name, const std::vector<long long>& shape, InitKind kind);
    void* pack_linear(const std::string& wname, int N, int K, int Kpad = 0, const std::vector<int>* rowmap = nullptr);
    void* pack_conv3(const std::string& wname, int Cout, int Cin, int Kpad = 0);
    // [Cout, 9*Cin + Cskip]: the conv3x3 weight followed by a 1x1 weight (K-concatenated: ResBlock out conv + skip_connection)
    void* pack_conv3_skip(const std::string& wname, const std::string& skip_wname, int Cout, int Cin, int Cskip);
    const float* sum_bias(const std::string& a, const std::string& b, int N);
    // rows of several [N_i, K] matrices stacked into one bf16 [sum N_i, K] matrix (fused QKV, all emb_layers)
    void* pack_concat(const std::string& key, const std::vector<std::string>& wnames, const std::vector<int>& Ns, int K);
    const float* concat_bias(const std::string& key, const std::vector<std::string>& bnames, const std::vector<int>& Ns);
    const float* gather_bias(const std::string& bname, int N, const std::vector<int>& rowmap);
    void* dev_alloc(size_t bytes, bool zero);           // owned by the net, freed in the destructor
    // ---- activations
    Act new_act(int B, int H, int W, int C, bool f32 = false);
    void release(Act& a);
    // ---- layer builders (append prepared launches to *plan_)
    Act gn(const Act& x, const std::string& prefix, float eps, bool silu);
    // GroupNorm over the channel concatenation [x | x2] (x2 may be NULL); raw_out (optional) receives the bf16 un-normalised concatenation
    Act gn2(const Act& x, const Act* x2, const std::string& prefix, float eps, bool silu, Act* raw_out);
    Act ln(const Act& x, const std::string& prefix);
    struct LinearOpts {
        const float* bias = nullptr;
        const Act* residual = nullptr;
        int act = SDOD_ACT_NONE;
        float alpha = 1.0f;
        int block_n = 0;
        bool out_f32 = false;    // write the fp32 residual stream
        bool in_place = false;   // fp32 stream with a residual: write the result over the residual (x += f(x)); the returned Act IS *residual
        // LayerNorm of the output rows (fp32 stream outputs only): *ln_out receives bf16 LN(y) * ln_w + ln_b — fused into the GEMM's
        // epilogue where the shape allows (gemm_tcgen05.cu, ln_fuse), else a separate layer_norm launch
        Act* ln_out = nullptr;
        const float* ln_w = nullptr;
        const float* ln_b = nullptr;
    };
    Act linear(const Act& x, const void* w_bf16, int N, const LinearOpts& o);          // out [M, N or N/2 (GEGLU)]
    int gemm_into(const sdod_gemm_desc& d, bool may_fail = false);                    // fully custom epilogue; may_fail: return the status instead of throwing
    Act conv3(const Act& x, const std::string& prefix, int cout, const float* row_bias, long long ld_row_bias, const Act* residual,
              bool stream_out = false, float* out_f32 = nullptr);
    // out = conv3x3(x) + conv1x1(x_skip) + biases, one launch (skip weights K-concatenated); fp32 stream output
    Act conv3_skip(const Act& x, const std::string& prefix, const std::string& skip_prefix, const Act& x_skip, int cout);
    // conv3x3(nearest_upsample_2x(x)) in sub-pixel form: four 2x2 convolutions over x with pre-summed weights, output [B, 2H, 2W, cout]
    // written in place through a 5-D tensor map (no upsampled tensor, 4/9 of the multiply-adds).  x fp32 is cast to bf16 first.
    Act conv3_up2(const Act& x, const std::string& prefix, int cout, bool stream_out = false);
    void* pack_conv3_up2(const std::string& wname, int Cout, int Cin);
    // stride-2 conv3x3, padding 1 (UNet Downsample): implicit GEMM on TMA boxes with element strides 2 — no im2col buffer.  x fp32 is cast first.
    Act conv3_s2(const Act& x, const std::string& prefix, int cout, bool stream_out = false);
    Act conv3_im2col(const Act& x, const std::string& prefix, int cout, int stride, bool stream_out = false);
    Act conv1x1(const Act& x, const std::string& prefix, int cout, const Act* residual, bool stream_out = false, bool in_place = false);
    Act to_bf16(const Act& x);
    Act upsample(const Act& x);
    Act concat(const Act& a, const Act& b);
    void check(int status);                              // throws std::runtime_error with sdod last error
    // plan-build bookkeeping: every GEMM/conv launch of the plan in order, so that each one can prefetch its successor's weights into L2
    void begin_plan(Plan* p, bool prefetch_weights);
    void end_plan();
    void note_gemm(const std::shared_ptr<GemmLaunch>& g) { if (plan_) plan_gemms_.push_back(g); }
    std::vector<std::shared_ptr<GemmLaunch>> plan_gemms_;
    bool plan_prefetch_ = false;

    const WeightStore* ws_;
    bool random_;
    unsigned long long seed_, rng_ctr_ = 0;
    std::vector<void*> owned_;
    std::unordered_map<std::string, void*> cache_;       // packed / random-init tensors, shared by all plans
    Pool pool_;
    Plan* plan_ = nullptr;
    void* gn_ws_ = nullptr;
    size_t gn_ws_bytes_ = 0;
    SplitKWorkspace skw_;                                 // split-K partial tiles + tickets (zeroed once, self-resetting)
};

}  // namespace sdod
