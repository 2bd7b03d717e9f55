// SD-v1 UNet denoiser as a pre-planned launch sequence (NHWC bf16 activations, fp32 accumulation).
//
// Replaces the reference's opaque `unet` graph and its execute() calls (csrc/libsdod/src/context.cpp:
// 352,366; inputs 0=x, 1=t(1280-d), 2=p per context.cpp:214-218) and the `temb` graph (context.cpp:
// 257-279).  Architecture = public CompVis v1 config, layer names = ldm state_dict
// keys, which are also the names the reference's profiler reports (analyze_results.py:25-87).
//
// Fusions: GroupNorm+SiLU one op; conv bias + timestep-embedding add + residual in the conv epilogue;
// all 22 emb_layers Linear in one GEMM per step; q/k/v projections in one GEMM whose epilogue writes
// the attention operand layouts directly; GEGLU gate in the GEMM epilogue; every residual add in the
// producing GEMM's epilogue; cross-attention K/V of the (constant) prompt projected once per prompt.
#include "unet.h"

#include <cmath>
#include <stdexcept>

#include "../kernels/glue.h"

namespace sdod {

namespace {
constexpr int kHeads = 8;
constexpr int kCtxTokens = 77;
constexpr int kCtxDim = 768;
constexpr int kTed = 1280;
constexpr int kMc = 320;
const int kMult[4] = {1, 2, 4, 4};

struct HeadGeom { int dh, dpad, vt_rows; };
HeadGeom head_geom(int C) {
    HeadGeom g;
    g.dh = C / kHeads;
    g.dpad = 64 * ((g.dh + 63) / 64);
    g.vt_rows = 16 * ((g.dh + 15) / 16);
    return g;
}
int pad8(int n) { return (n + 7) / 8 * 8; }
}  // namespace

UNet::UNet(const WeightStore* ws, unsigned long long seed, int latent_hw, int max_batch)
    : NetBase(ws, seed), hw_(latent_hw), max_batch_(max_batch) {
    if (latent_hw < 8 || (latent_hw & (latent_hw - 1)) != 0) throw std::runtime_error("unet: latent_hw must be a power of two >= 8");
    if (max_batch < 1 || max_batch > 256) throw std::runtime_error("unet: max_batch out of range");
    // enumerate ResBlocks (emb_layers) in execution order -> one fused projection
    {
        int ch = kMc, k = 1;
        auto add = [&](const std::string& prefix, int cout) {
            emb_names_.push_back(prefix + ".emb_layers.1");
            emb_couts_.push_back(cout);
            emb_offsets_.push_back(emb_total_);
            emb_total_ += cout;
        };
        for (int level = 0; level < 4; ++level) {
            for (int i = 0; i < 2; ++i) { add("input_blocks." + std::to_string(k++) + ".0", kMult[level] * kMc); ch = kMult[level] * kMc; }
            if (level != 3) k++;
        }
        add("middle_block.0", ch);
        add("middle_block.2", ch);
        k = 0;
        for (int level = 3; level >= 0; --level)
            for (int i = 0; i < 3; ++i) add("output_blocks." + std::to_string(k++) + ".0", kMult[level] * kMc);
    }
    const size_t mb = static_cast<size_t>(max_batch_);
    x_in_ = static_cast<float*>(dev_alloc(mb * hw_ * hw_ * 4 * sizeof(float), true));
    emb_in_ = static_cast<float*>(dev_alloc(mb * kTed * sizeof(float), true));
    eps_out_ = static_cast<float*>(dev_alloc(mb * hw_ * hw_ * 4 * sizeof(float), true));
    emb_proj_ = static_cast<float*>(dev_alloc(mb * emb_total_ * sizeof(float), true));
    ctx_bf16_ = dev_alloc(mb * kCtxTokens * kCtxDim * 2, true);
    // per-level attention operand buffers (zero-filled once; the pad lanes are never written afterwards)
    for (int level = 0; level < 4; ++level) {
        const int C = kMult[level] * kMc, side = hw_ >> level, tokens = side * side;
        const HeadGeom g = head_geom(C);
        LevelBufs lb;
        lb.qh = dev_alloc(mb * kHeads * tokens * g.dpad * 2, true);
        lb.kh = dev_alloc(mb * kHeads * tokens * g.dpad * 2, true);
        lb.vt = dev_alloc(mb * kHeads * g.vt_rows * pad8(tokens) * 2, true);
        level_bufs_.push_back(lb);
    }
}

UNet::~UNet() = default;

UNet::CtxBufs& UNet::ctx_bufs(const std::string& prefix, int C) {
    auto it = ctx_.find(prefix);
    if (it != ctx_.end()) return it->second;
    const HeadGeom g = head_geom(C);
    CtxBufs cb;
    cb.C = C;
    cb.kh = dev_alloc(static_cast<size_t>(max_batch_) * kHeads * kCtxTokens * g.dpad * 2, true);
    cb.vt = dev_alloc(static_cast<size_t>(max_batch_) * kHeads * g.vt_rows * pad8(kCtxTokens) * 2, true);
    ctx_order_.push_back(prefix);
    return ctx_.emplace(prefix, cb).first->second;
}

// ---------------------------------------------------------------------------------------------- layers
// ResBlock over the channel concatenation [x | x_cat] (x_cat = the popped skip tensor of an output block, else NULL).  The concatenation
// never exists in fp32: the first GroupNorm reads both sources and, when the block changes the channel count, also emits the bf16 copy
// that the 1x1 skip_connection consumes; that 1x1 convolution runs inside out_layers.3's K loop (K-concatenated weights, one accumulator).
Act UNet::res_block(const Act& x, const Act* x_cat, const std::string& prefix, int cout, int emb_index) {
    const int cin = x.C + (x_cat ? x_cat->C : 0);
    const bool project = cin != cout;
    if (x_cat && !project) throw std::runtime_error("res_block: a concatenated input needs a skip_connection");
    Act raw;
    Act h1 = gn2(x, x_cat, prefix + ".in_layers.0", 1e-5f, true, project ? &raw : nullptr);
    Act h2 = conv3(h1, prefix + ".in_layers.2", cout, emb_proj_ + emb_offsets_[emb_index], emb_total_, nullptr);
    release(h1);
    Act h3 = gn(h2, prefix + ".out_layers.0", 1e-5f, true);
    release(h2);
    Act out;
    if (project) {
        out = conv3_skip(h3, prefix + ".out_layers.3", prefix + ".skip_connection", raw, cout);
        release(raw);
    } else {
        out = conv3(h3, prefix + ".out_layers.3", cout, nullptr, 0, &x, true);
    }
    release(h3);
    return out;
}

void UNet::attention_op(const void* qh, const void* kh, const void* vt, void* out, int B, int tokens, int n_kv, int C) {
    const HeadGeom g = head_geom(C);
    auto a = std::make_shared<AttnLaunch>();
    check(attention_prepare(a.get(), qh, kh, vt, out, B, kHeads, tokens, n_kv, g.dh, g.dpad, pad8(n_kv), 1.0f / std::sqrt(static_cast<float>(g.dh))));
    plan_->push([a](cudaStream_t st) { return attention_launch(*a, st); }, 1,
                "attn Nq" + std::to_string(tokens) + " Nkv" + std::to_string(n_kv) + " d" + std::to_string(g.dh));
}

Act UNet::spatial_transformer(const Act& x, const std::string& prefix, int level) {
    const int C = x.C, tokens = x.H * x.W, B = x.B;
    const HeadGeom g = head_geom(C);
    const LevelBufs& lb = level_bufs_[level];
    Act hn = gn(x, prefix + ".norm", 1e-6f, false);
    const std::string tb = prefix + ".transformer_blocks.0";
    // proj_in writes the fp32 token stream and, from the same epilogue, norm1 of it (the three LayerNorms of the block ride on the
    // GEMMs that produce their inputs: proj_in, attn1.to_out, attn2.to_out)
    Act n1;
    LinearOpts oi;
    oi.bias = w32(prefix + ".proj_in.bias", {C}, kInitBias);
    oi.out_f32 = true;
    oi.ln_out = &n1; oi.ln_w = w32(tb + ".norm1.weight", {C}, kInitOnes); oi.ln_b = w32(tb + ".norm1.bias", {C}, kInitZeros);
    Act h = linear(hn, pack_linear(prefix + ".proj_in.weight", C, C), C, oi);     // fp32 token stream
    release(hn);

    // ---- self-attention: fused q/k/v projection straight into the attention operand layouts
    {
        void* wqkv = pack_concat(tb + ".attn1.qkv", {tb + ".attn1.to_q.weight", tb + ".attn1.to_k.weight", tb + ".attn1.to_v.weight"}, {C, C, C}, C);
        sdod_gemm_desc d{};
        d.A = n1.p; d.lda = C; d.W = wqkv; d.ldw = C; d.M = n1.M(); d.N = 3 * C; d.K = C; d.batch = 1;
        d.epi.C = lb.qh; d.epi.C2 = lb.kh; d.epi.C3 = lb.vt;
        d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_QKV;
        d.epi.heads = kHeads; d.epi.head_dim = g.dh; d.epi.tokens = tokens; d.epi.dpad = g.dpad; d.epi.tok_pad = pad8(tokens); d.epi.vt_rows = g.vt_rows;
        gemm_into(d);
    }
    release(n1);
    Act a1 = new_act(B, x.H, x.W, C);
    attention_op(lb.qh, lb.kh, lb.vt, a1.p, B, tokens, tokens, C);
    LinearOpts o1;
    o1.bias = w32(tb + ".attn1.to_out.0.bias", {C}, kInitBias);
    o1.residual = &h;
    o1.out_f32 = true;
    o1.in_place = true;       // h += to_out(attention): the block's fp32 token stream is updated in place (TMA reduce-add store where no LayerNorm rides along)
    Act n2;
    o1.ln_out = &n2; o1.ln_w = w32(tb + ".norm2.weight", {C}, kInitOnes); o1.ln_b = w32(tb + ".norm2.bias", {C}, kInitZeros);
    Act h2 = linear(a1, pack_linear(tb + ".attn1.to_out.0.weight", C, C), C, o1);
    release(a1);
    if (h2.p != h.p) release(h);

    // ---- cross-attention against the cached prompt K / V^T
    {
        sdod_gemm_desc d{};
        d.A = n2.p; d.lda = C; d.W = pack_linear(tb + ".attn2.to_q.weight", C, C); d.ldw = C; d.M = n2.M(); d.N = C; d.K = C; d.batch = 1;
        d.epi.C = lb.qh; d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_HEADS;
        d.epi.heads = kHeads; d.epi.head_dim = g.dh; d.epi.tokens = tokens; d.epi.dpad = g.dpad;
        gemm_into(d);
    }
    release(n2);
    CtxBufs& cb = ctx_bufs(tb + ".attn2", C);
    Act a2 = new_act(B, x.H, x.W, C);
    attention_op(lb.qh, cb.kh, cb.vt, a2.p, B, tokens, kCtxTokens, C);
    LinearOpts o2;
    o2.bias = w32(tb + ".attn2.to_out.0.bias", {C}, kInitBias);
    o2.residual = &h2;
    o2.out_f32 = true;
    o2.in_place = true;
    Act n3;
    o2.ln_out = &n3; o2.ln_w = w32(tb + ".norm3.weight", {C}, kInitOnes); o2.ln_b = w32(tb + ".norm3.bias", {C}, kInitZeros);
    Act h3 = linear(a2, pack_linear(tb + ".attn2.to_out.0.weight", C, C), C, o2);
    release(a2);
    if (h3.p != h2.p) release(h2);

    // ---- GEGLU feed-forward (gate fused in the projection's epilogue)
    std::vector<int> rowmap(8 * C);
    {
        const int half = 64, n = 4 * C;    // block_n 128: [64 value rows | 64 gate rows] per tile (2 CTAs/SM: epilogues overlap mainloops)
        int idx = 0;
        for (int t = 0; t < n / half; ++t) {
            for (int i = 0; i < half; ++i) rowmap[idx++] = t * half + i;
            for (int i = 0; i < half; ++i) rowmap[idx++] = n + t * half + i;
        }
    }
    LinearOpts og;
    og.bias = gather_bias(tb + ".ff.net.0.proj.bias", 8 * C, rowmap);
    og.act = SDOD_ACT_GEGLU;
    og.block_n = 128;
    Act gg = linear(n3, pack_linear(tb + ".ff.net.0.proj.weight", 8 * C, C, 0, &rowmap), 8 * C, og);
    release(n3);
    LinearOpts o3;
    o3.bias = w32(tb + ".ff.net.2.bias", {C}, kInitBias);
    o3.residual = &h3;        // (bf16 output: h4 is proj_out's GEMM operand, so this one cannot update the fp32 stream in place)
    Act h4 = linear(gg, pack_linear(tb + ".ff.net.2.weight", C, 4 * C), C, o3);
    release(gg);
    release(h3);

    Act out = conv1x1(h4, prefix + ".proj_out", C, &x, true, true);     // x += proj_out(h4): in place (x is this block's private ResBlock output)
    release(h4);
    return out;
}

// ---------------------------------------------------------------------------------------------- plans
std::unique_ptr<Plan> UNet::build_forward(int B) {
    auto plan = std::make_unique<Plan>();
    begin_plan(plan.get(), B <= 4);      // small batches stream the weights from HBM every step: chain L2 prefetches through the GEMMs
    int emb_index = 0;
    // emb path: SiLU(emb) -> all 22 emb_layers Linear in one GEMM -> fp32 [B, emb_total]
    {
        Act se = new_act(B, 1, 1, kTed);
        const float* ein = emb_in_;
        void* sp = se.p;
        const size_t n = static_cast<size_t>(B) * kTed;
        plan_->push([=](cudaStream_t st) { return silu_f32_to_bf16(st, ein, sp, n, 1); });
        std::vector<std::string> wn, bn;
        for (auto& e : emb_names_) { wn.push_back(e + ".weight"); bn.push_back(e + ".bias"); }
        sdod_gemm_desc d{};
        d.A = se.p; d.lda = kTed; d.W = pack_concat("emb_layers", wn, emb_couts_, kTed); d.ldw = kTed;
        d.M = B; d.N = emb_total_; d.K = kTed; d.batch = 1;
        d.epi.C = emb_proj_; d.epi.ldc = emb_total_; d.epi.bias = concat_bias("emb_layers", bn, emb_couts_);
        d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_F32;
        gemm_into(d);
        release(se);
    }
    // x: fp32 NHWC -> bf16
    Act x0 = new_act(B, hw_, hw_, 4);
    {
        const float* xin = x_in_;
        void* xp = x0.p;
        const size_t n = static_cast<size_t>(B) * hw_ * hw_ * 4;
        plan_->push([=](cudaStream_t st) { return sdod_cast_f32_to_bf16(st, xin, xp, n); });
    }
    std::vector<Act> hs;
    Act h = conv3_im2col(x0, "input_blocks.0.0", kMc, 1, true);
    release(x0);
    hs.push_back(h);
    int k = 1;
    for (int level = 0; level < 4; ++level) {
        for (int i = 0; i < 2; ++i) {
            const std::string p = "input_blocks." + std::to_string(k++);
            Act r = res_block(h, nullptr, p + ".0", kMult[level] * kMc, emb_index++);
            // h stays alive: it is in hs (skip connection)
            if (level < 3) {
                Act t = spatial_transformer(r, p + ".1", level);
                if (t.p != r.p) release(r);
                r = t;
            }
            h = r;
            hs.push_back(h);
        }
        if (level != 3) {
            const std::string p = "input_blocks." + std::to_string(k++);
            h = conv3_s2(h, p + ".0.op", h.C, true);          // Downsample: stride-2 conv as an implicit GEMM on strided TMA boxes
            hs.push_back(h);
        }
    }
    {
        Act r = res_block(h, nullptr, "middle_block.0", h.C, emb_index++);
        Act t = spatial_transformer(r, "middle_block.1", 3);
        if (t.p != r.p) release(r);
        Act r2 = res_block(t, nullptr, "middle_block.2", t.C, emb_index++);
        release(t);
        h = r2;     // previous h is hs.back(): released when popped
    }
    k = 0;
    bool h_owned = true;   // h is not in hs
    for (int level = 3; level >= 0; --level) {
        for (int i = 0; i < 3; ++i) {
            const std::string p = "output_blocks." + std::to_string(k++);
            Act skip = hs.back();
            hs.pop_back();
            Act r = res_block(h, &skip, p + ".0", kMult[level] * kMc, emb_index++);
            if (h_owned) release(h);
            release(skip);
            int sub = 1;
            if (level < 3) {
                Act t = spatial_transformer(r, p + ".1", level);
                if (t.p != r.p) release(r);
                r = t;
                sub = 2;
            }
            if (level > 0 && i == 2) {
                Act c = conv3_up2(r, p + "." + std::to_string(sub) + ".conv", r.C, true);     // Upsample: nearest 2x + conv3x3, sub-pixel form
                release(r);
                r = c;
            }
            h = r;
            h_owned = true;
        }
    }
    Act hn = gn(h, "out.0", 1e-5f, true);
    release(h);
    conv3(hn, "out.2", 4, nullptr, 0, nullptr, false, eps_out_);
    release(hn);
    end_plan();
    return plan;
}

std::unique_ptr<Plan> UNet::build_context(int B) {
    // make sure every cross-attention block has registered its buffers
    if (ctx_order_.empty()) forward_plan(1);
    auto plan = std::make_unique<Plan>();
    begin_plan(plan.get(), false);
    Act ctx;
    ctx.p = ctx_bf16_; ctx.B = B; ctx.H = kCtxTokens; ctx.W = 1; ctx.C = kCtxDim;
    for (const std::string& name : ctx_order_) {
        CtxBufs& cb = ctx_.at(name);
        const HeadGeom g = head_geom(cb.C);
        for (int which = 0; which < 2; ++which) {
            sdod_gemm_desc d{};
            d.A = ctx.p; d.lda = kCtxDim; d.ldw = kCtxDim; d.M = B * kCtxTokens; d.N = cb.C; d.K = kCtxDim; d.batch = 1;
            d.W = pack_linear(name + (which == 0 ? ".to_k.weight" : ".to_v.weight"), cb.C, kCtxDim);
            d.epi.C = which == 0 ? cb.kh : cb.vt;
            d.epi.alpha = 1.0f; d.epi.out_mode = which == 0 ? SDOD_OUT_HEADS : SDOD_OUT_HEADS_T;
            d.epi.heads = kHeads; d.epi.head_dim = g.dh; d.epi.tokens = kCtxTokens; d.epi.dpad = g.dpad;
            d.epi.tok_pad = pad8(kCtxTokens); d.epi.vt_rows = g.vt_rows;
            gemm_into(d);
        }
    }
    end_plan();
    return plan;
}

std::unique_ptr<Plan> UNet::build_time_embed(int n) {
    auto plan = std::make_unique<Plan>();
    begin_plan(plan.get(), false);
    Act s0 = new_act(n, 1, 1, kMc);
    {
        const float* sin_in = temb_sin_;
        void* sp = s0.p;
        const size_t cnt = static_cast<size_t>(n) * kMc;
        plan_->push([=](cudaStream_t st) { return silu_f32_to_bf16(st, sin_in, sp, cnt, 0); });
    }
    LinearOpts o1;
    o1.bias = w32("time_embed.0.bias", {kTed}, kInitBias);
    o1.act = SDOD_ACT_SILU;
    Act h = linear(s0, pack_linear("time_embed.0.weight", kTed, kMc), kTed, o1);
    release(s0);
    sdod_gemm_desc d{};
    d.A = h.p; d.lda = kTed; d.W = pack_linear("time_embed.2.weight", kTed, kTed); d.ldw = kTed; d.M = n; d.N = kTed; d.K = kTed; d.batch = 1;
    d.epi.C = temb_out_; d.epi.ldc = kTed; d.epi.bias = w32("time_embed.2.bias", {kTed}, kInitBias); d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_F32;
    gemm_into(d);
    release(h);
    end_plan();
    return plan;
}

Plan* UNet::forward_plan(int B) {
    auto it = fwd_.find(B);
    if (it == fwd_.end()) it = fwd_.emplace(B, build_forward(B)).first;
    return it->second.get();
}

// ---------------------------------------------------------------------------------------------- entry points
int UNet::time_embed(cudaStream_t s, const float* t, int n, float* out) {
    if (n < 1 || n > 1024) return fail(kInvalidArgument, "time_embed: n out of range");
    try {
        if (!temb_sin_) {
            temb_sin_ = static_cast<float*>(dev_alloc(1024 * kMc * sizeof(float), true));
            temb_out_ = static_cast<float*>(dev_alloc(1024 * kTed * sizeof(float), true));
        }
        auto it = temb_.find(n);
        if (it == temb_.end()) it = temb_.emplace(n, build_time_embed(n)).first;
        SDOD_TRY(sdod_timestep_sinusoid(s, t, n, kMc, 10000.0f, temb_sin_));
        SDOD_TRY(it->second->run(s, false));
        return check_cuda(cudaMemcpyAsync(out, temb_out_, static_cast<size_t>(n) * kTed * sizeof(float), cudaMemcpyDeviceToDevice, s), "copy emb");
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("unet time_embed: ") + e.what());
    }
}

int UNet::set_context(cudaStream_t s, const void* context, int dtype, int B) {
    if (B < 1 || B > max_batch_) return fail(kInvalidArgument, "set_context: batch exceeds max_batch");
    try {
        const size_t n = static_cast<size_t>(B) * kCtxTokens * kCtxDim;
        if (dtype == SDOD_F32) SDOD_TRY(sdod_cast_f32_to_bf16(s, static_cast<const float*>(context), ctx_bf16_, n));
        else SDOD_TRY(check_cuda(cudaMemcpyAsync(ctx_bf16_, context, n * 2, cudaMemcpyDeviceToDevice, s), "copy context"));
        auto it = ctxp_.find(B);
        if (it == ctxp_.end()) it = ctxp_.emplace(B, build_context(B)).first;
        return it->second->run(s, false);
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("unet set_context: ") + e.what());
    }
}

int UNet::forward(cudaStream_t s, const float* x, const float* emb, float* eps, int B, bool use_graph) {
    if (B < 1 || B > max_batch_) return fail(kInvalidArgument, "unet forward: batch exceeds max_batch");
    try {
        Plan* p = forward_plan(B);
        const size_t nx = static_cast<size_t>(B) * hw_ * hw_ * 4 * sizeof(float);
        if (x != x_in_) SDOD_TRY(check_cuda(cudaMemcpyAsync(x_in_, x, nx, cudaMemcpyDeviceToDevice, s), "copy x"));
        if (emb != emb_in_) SDOD_TRY(check_cuda(cudaMemcpyAsync(emb_in_, emb, static_cast<size_t>(B) * kTed * sizeof(float), cudaMemcpyDeviceToDevice, s), "copy emb"));
        const bool prev = set_pdl_for_thread(B <= 4);     // baked into the graph at capture time; see host_common.h
        const int st = p->run(s, use_graph);
        set_pdl_for_thread(prev);
        SDOD_TRY(st);
        if (eps != eps_out_) SDOD_TRY(check_cuda(cudaMemcpyAsync(eps, eps_out_, nx, cudaMemcpyDeviceToDevice, s), "copy eps"));
        return kOk;
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("unet forward: ") + e.what());
    }
}

int UNet::profile_forward(cudaStream_t s, int B, int iters, std::vector<std::pair<std::string, float>>* out) {
    try {
        return forward_plan(B)->profile(s, iters, out);
    } catch (const std::exception& e) {
        return fail(kCudaError, std::string("unet profile: ") + e.what());
    }
}

unsigned long long UNet::launches_per_forward(int B) {
    try {
        return forward_plan(B)->launches();
    } catch (...) {
        return 0;
    }
}

}  // namespace sdod
