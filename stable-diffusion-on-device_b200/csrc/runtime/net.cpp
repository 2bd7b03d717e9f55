#include "net.h"

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <stdexcept>

#include "../kernels/glue.h"
#include "../launch_count.h"

namespace sdod {

// ------------------------------------------------------------------------------------------ WeightStore
WeightStore::~WeightStore() {
    for (auto& kv : t_) cudaFree(kv.second.p);
}

int WeightStore::set(const std::string& name, const float* host, const std::vector<long long>& shape) {
    size_t n = 1;
    for (long long s : shape) {
        if (s <= 0) return fail(kInvalidArgument, "weights: non-positive dimension for " + name);
        n *= static_cast<size_t>(s);
    }
    DevTensor& d = t_[name];
    if (d.p && d.numel != n) { cudaFree(d.p); d.p = nullptr; }
    if (!d.p) SDOD_TRY(check_cuda(cudaMalloc(reinterpret_cast<void**>(&d.p), n * sizeof(float)), "cudaMalloc(weight)"));
    d.shape = shape;
    d.numel = n;
    return check_cuda(cudaMemcpy(d.p, host, n * sizeof(float), cudaMemcpyHostToDevice), "cudaMemcpy(weight)");
}

int WeightStore::load_file(const std::string& path) {
    struct Closer { void operator()(FILE* f) const { if (f) std::fclose(f); } };
    std::unique_ptr<FILE, Closer> f(std::fopen(path.c_str(), "rb"));
    if (!f) return fail(kInvalidArgument, "weights: cannot open " + path);
    auto bail = [&](const std::string& m) { return fail(kInvalidArgument, "weights: " + m + " in " + path); };
    std::fseek(f.get(), 0, SEEK_END);
    const long long file_bytes = std::ftell(f.get());
    std::fseek(f.get(), 0, SEEK_SET);
    char magic[8];
    uint32_t count = 0;
    if (std::fread(magic, 1, 8, f.get()) != 8 || std::memcmp(magic, "SDODW001", 8) != 0) return bail("bad magic");
    if (std::fread(&count, 4, 1, f.get()) != 1) return bail("truncated header");
    std::vector<float> buf;
    try {
        for (uint32_t i = 0; i < count; ++i) {
            uint32_t nl = 0, nd = 0;
            if (std::fread(&nl, 4, 1, f.get()) != 1 || nl == 0 || nl > 4096) return bail("bad name length");
            std::string name(nl, '\0');
            if (std::fread(&name[0], 1, nl, f.get()) != nl) return bail("truncated name");
            if (std::fread(&nd, 4, 1, f.get()) != 1 || nd > 8) return bail("bad rank");
            std::vector<long long> shape(nd);
            if (nd && std::fread(shape.data(), 8, nd, f.get()) != nd) return bail("truncated shape");
            // every dimension positive, and the element count bounded by what is left of the file (no overflow, no huge resize)
            const long long left = (file_bytes - std::ftell(f.get())) / 4;
            long long n = 1;
            for (long long d : shape) {
                if (d <= 0 || d > left || n > left / d) return bail("bad shape for " + name);
                n *= d;
            }
            buf.resize(static_cast<size_t>(n));
            if (std::fread(buf.data(), 4, static_cast<size_t>(n), f.get()) != static_cast<size_t>(n)) return bail("truncated data for " + name);
            SDOD_TRY(set(name, buf.data(), shape));
        }
    } catch (const std::exception& e) {
        return bail(std::string("exception while reading: ") + e.what());
    }
    return kOk;
}

const DevTensor* WeightStore::find(const std::string& name) const {
    auto it = t_.find(name);
    return it == t_.end() ? nullptr : &it->second;
}

// ------------------------------------------------------------------------------------------ Pool
Pool::~Pool() {
    for (auto& b : blks_) cudaFree(b.p);
}

void* Pool::get(size_t bytes) {
    bytes = (bytes + 1023) & ~static_cast<size_t>(1023);
    int best = -1;
    for (size_t i = 0; i < blks_.size(); ++i) {
        if (blks_[i].free_ && blks_[i].sz >= bytes && blks_[i].sz <= bytes + bytes / 2 + 4096) {
            if (best < 0 || blks_[i].sz < blks_[best].sz) best = static_cast<int>(i);
        }
    }
    if (best >= 0) {
        blks_[best].free_ = false;
        return blks_[best].p;
    }
    void* p = nullptr;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        throw std::runtime_error("device pool: cudaMalloc of " + std::to_string(bytes) + " bytes failed");
    }
    blks_.push_back({p, bytes, false});
    total_ += bytes;
    return p;
}

void Pool::put(void* p) {
    for (auto& b : blks_)
        if (b.p == p) { b.free_ = true; return; }
}

// ------------------------------------------------------------------------------------------ Plan
Plan::~Plan() {
    if (exec_) cudaGraphExecDestroy(exec_);
}

int Plan::run_eager(cudaStream_t s) {
    for (auto& f : ops_) SDOD_TRY(f(s));
    return kOk;
}

int Plan::run(cudaStream_t s, bool use_graph) {
    if (!use_graph || s == nullptr) return run_eager(s);    // legacy default stream cannot be captured
    if (!warmed_) {                                         // first pass eager: sets kernel attributes, validates launches
        warmed_ = true;
        return run_eager(s);
    }
    if (!exec_) {
        cudaGraph_t g = nullptr;
        SDOD_TRY(check_cuda(cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal), "cudaStreamBeginCapture"));
        const unsigned long long before = g_launch_count.load();
        int st = run_eager(s);
        g_launch_count.store(before);                       // captured, not launched
        cudaError_t e = cudaStreamEndCapture(s, &g);
        if (st != kOk) { if (g) cudaGraphDestroy(g); return st; }
        SDOD_TRY(check_cuda(e, "cudaStreamEndCapture"));
        e = cudaGraphInstantiate(&exec_, g, 0);
        cudaGraphDestroy(g);
        SDOD_TRY(check_cuda(e, "cudaGraphInstantiate"));
    }
    SDOD_TRY(check_cuda(cudaGraphLaunch(exec_, s), "cudaGraphLaunch"));
    count_launch(launches_);
    return kOk;
}

int Plan::profile(cudaStream_t s, int iters, std::vector<std::pair<std::string, float>>* out) {
    const size_t n = ops_.size();
    std::vector<cudaEvent_t> ev(n + 1);
    for (auto& e : ev) SDOD_TRY(check_cuda(cudaEventCreate(&e), "cudaEventCreate"));
    std::vector<double> acc(n, 0.0);
    SDOD_TRY(run_eager(s));
    for (int it = 0; it < iters; ++it) {
        SDOD_TRY(check_cuda(cudaEventRecord(ev[0], s), "cudaEventRecord"));
        for (size_t i = 0; i < n; ++i) {
            SDOD_TRY(ops_[i](s));
            SDOD_TRY(check_cuda(cudaEventRecord(ev[i + 1], s), "cudaEventRecord"));
        }
        SDOD_TRY(check_cuda(cudaStreamSynchronize(s), "cudaStreamSynchronize"));
        for (size_t i = 0; i < n; ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, ev[i], ev[i + 1]);
            acc[i] += ms;
        }
    }
    out->clear();
    for (size_t i = 0; i < n; ++i) out->emplace_back(names_[i], static_cast<float>(acc[i] / iters));
    for (auto& e : ev) cudaEventDestroy(e);
    return kOk;
}

// ------------------------------------------------------------------------------------------ NetBase
NetBase::NetBase(const WeightStore* ws, unsigned long long seed) : ws_(ws), random_(ws == nullptr), seed_(seed) {
    gn_ws_bytes_ = sdod_group_norm_workspace(256, 1, 1, 64, SDOD_NHWC);
    gn_ws_ = dev_alloc(gn_ws_bytes_, true);
    skw_.ws_bytes = static_cast<size_t>(64) << 20;
    skw_.ws = static_cast<float*>(dev_alloc(skw_.ws_bytes, false));
    skw_.n_counters = 4096;
    skw_.counters = static_cast<unsigned int*>(dev_alloc(skw_.n_counters * sizeof(unsigned int), true));
}

NetBase::~NetBase() {
    for (void* p : owned_) cudaFree(p);
}

void NetBase::check(int status) {
    if (status != kOk) throw std::runtime_error(last_error());
}

void* NetBase::dev_alloc(size_t bytes, bool zero) {
    void* p = nullptr;
    if (bytes == 0) bytes = 16;
    if (cudaMalloc(&p, bytes) != cudaSuccess) {
        cudaGetLastError();
        throw std::runtime_error("cudaMalloc of " + std::to_string(bytes) + " bytes failed");
    }
    if (zero) cudaMemset(p, 0, bytes);
    owned_.push_back(p);
    return p;
}

const float* NetBase::w32(const std::string& name, const std::vector<long long>& shape, InitKind kind) {
    size_t n = 1;
    for (long long s : shape) n *= static_cast<size_t>(s);
    if (!random_) {
        const DevTensor* t = ws_->find(name);
        if (!t) throw std::runtime_error("missing weight tensor '" + name + "'");
        if (t->numel != n) throw std::runtime_error("weight tensor '" + name + "' has " + std::to_string(t->numel) + " elements, expected " + std::to_string(n));
        // same element count is not enough: a transposed tensor would load silently.  Singleton dimensions may differ (ldm stores 1x1 convs as
        // [Cout, Cin, 1, 1] where a linear layer says [Cout, Cin]).
        auto squeeze = [](const std::vector<long long>& v) { std::vector<long long> o; for (long long d : v) if (d != 1) o.push_back(d); return o; };
        if (squeeze(t->shape) != squeeze(shape)) throw std::runtime_error("weight tensor '" + name + "' has an unexpected shape");
        return t->p;
    }
    auto hit = cache_.find("r|" + name);
    if (hit != cache_.end()) return static_cast<const float*>(hit->second);
    float* p = static_cast<float*>(dev_alloc(n * sizeof(float), false));
    cache_["r|" + name] = p;
    if (kind == kInitOnes || kind == kInitZeros) {
        check(fill_f32(nullptr, p, n, kind == kInitOnes ? 1.f : 0.f));
    } else {
        const size_t fan_in = shape.empty() ? 1 : n / static_cast<size_t>(shape[0]);
        const float std_ = kind == kInitWeight ? 1.0f / std::sqrt(static_cast<float>(fan_in ? fan_in : 1)) : 0.02f;
        check(sdod_randn(nullptr, p, n, seed_, rng_ctr_));
        rng_ctr_ += (n + 3) / 4;
        check(scale_f32(nullptr, p, n, std_));
    }
    return p;
}

void* NetBase::pack_linear(const std::string& wname, int N, int K, int Kpad, const std::vector<int>* rowmap) {
    if (Kpad == 0) Kpad = K;
    const std::string key = "l|" + wname + "|" + std::to_string(Kpad) + (rowmap ? "|g" : "");
    auto hit = cache_.find(key);
    if (hit != cache_.end()) return hit->second;
    const float* src = w32(wname, {N, K}, kInitWeight);
    void* dst = dev_alloc(static_cast<size_t>(N) * Kpad * 2, false);
    cache_[key] = dst;
    int* dmap = nullptr;
    if (rowmap) {
        check(check_cuda(cudaMalloc(reinterpret_cast<void**>(&dmap), rowmap->size() * sizeof(int)), "cudaMalloc(rowmap)"));
        check(check_cuda(cudaMemcpy(dmap, rowmap->data(), rowmap->size() * sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy(rowmap)"));
    }
    check(pack_rows(nullptr, src, dst, N, K, Kpad, dmap));
    if (dmap) { cudaDeviceSynchronize(); cudaFree(dmap); }
    return dst;
}

void* NetBase::pack_conv3(const std::string& wname, int Cout, int Cin, int Kpad) {
    if (Kpad == 0) Kpad = 9 * Cin;
    const std::string key = "c|" + wname + "|" + std::to_string(Kpad);
    auto hit = cache_.find(key);
    if (hit != cache_.end()) return hit->second;
    const float* src = w32(wname, {Cout, Cin, 3, 3}, kInitWeight);
    void* dst = dev_alloc(static_cast<size_t>(Cout) * Kpad * 2, false);
    cache_[key] = dst;
    check(sdod_pack_conv3x3_weight(nullptr, src, dst, Cout, Cin, Kpad));
    return dst;
}

void* NetBase::pack_conv3_skip(const std::string& wname, const std::string& skip_wname, int Cout, int Cin, int Cskip) {
    const std::string key = "cs|" + wname + "|" + skip_wname;
    auto hit = cache_.find(key);
    if (hit != cache_.end()) return hit->second;
    const int K = 9 * Cin + Cskip;
    const float* src = w32(wname, {Cout, Cin, 3, 3}, kInitWeight);
    const float* ssrc = w32(skip_wname, {Cout, Cskip}, kInitWeight);      // ldm stores the 1x1 conv as [Cout, Cskip, 1, 1]
    void* dst = dev_alloc(static_cast<size_t>(Cout) * K * 2, false);
    cache_[key] = dst;
    check(sdod_pack_conv3x3_weight(nullptr, src, dst, Cout, Cin, K));
    check(pack_rows_window(nullptr, ssrc, dst, Cout, Cskip, K, 9 * Cin));
    return dst;
}

const float* NetBase::sum_bias(const std::string& a, const std::string& b, int N) {
    const std::string key = "sb|" + a + "|" + b;
    auto hit = cache_.find(key);
    if (hit != cache_.end()) return static_cast<const float*>(hit->second);
    const float* pa = w32(a, {N}, kInitBias);
    const float* pb = w32(b, {N}, kInitBias);
    float* dst = static_cast<float*>(dev_alloc(N * sizeof(float), false));
    cache_[key] = dst;
    check(add_f32(nullptr, pa, pb, dst, N));
    return dst;
}

void NetBase::begin_plan(Plan* p, bool prefetch_weights) {
    static const int env = [] { const char* e = std::getenv("SDOD_W_PREFETCH"); return e ? std::atoi(e) : 1; }();
    plan_ = p;
    plan_gemms_.clear();
    plan_prefetch_ = prefetch_weights && env != 0;
}

void NetBase::end_plan() {
    // Launch i pulls launch i+1's weight matrix towards L2 (cp.async.bulk.prefetch.L2, see gemm_tcgen05_kernel); capped so that the
    // prefetched bytes plus the step's activations stay well inside the 126 MB L2.
    if (plan_prefetch_) {
        for (size_t i = 0; i + 1 < plan_gemms_.size(); ++i) {
            const GemmLaunch& nx = *plan_gemms_[i + 1];
            if (!nx.w_ptr || nx.w_bytes < (64 << 10)) continue;
            plan_gemms_[i]->mp.pf_ptr = static_cast<const char*>(nx.w_ptr);
            plan_gemms_[i]->mp.pf_bytes = std::min<long long>(nx.w_bytes, 64LL << 20);
        }
    }
    plan_gemms_.clear();
    plan_ = nullptr;
    // weight init / packing ran on the legacy default stream while plans run on the caller's (possibly non-blocking) stream:
    // make every packed tensor visible before the plan's first launch
    cudaDeviceSynchronize();
}

void* NetBase::pack_concat(const std::string& key, const std::vector<std::string>& wnames, const std::vector<int>& Ns, int K) {
    auto hit = cache_.find("k|" + key);
    if (hit != cache_.end()) return hit->second;
    size_t total = 0;
    for (int n : Ns) total += n;
    char* dst = static_cast<char*>(dev_alloc(total * K * 2, false));
    cache_["k|" + key] = dst;
    size_t off = 0;
    for (size_t i = 0; i < wnames.size(); ++i) {
        const float* src = w32(wnames[i], {Ns[i], K}, kInitWeight);
        check(pack_rows(nullptr, src, dst + off * K * 2, Ns[i], K, K, nullptr));
        off += Ns[i];
    }
    return dst;
}

const float* NetBase::concat_bias(const std::string& key, const std::vector<std::string>& bnames, const std::vector<int>& Ns) {
    auto hit = cache_.find("kb|" + key);
    if (hit != cache_.end()) return static_cast<const float*>(hit->second);
    size_t total = 0;
    for (int n : Ns) total += n;
    float* dst = static_cast<float*>(dev_alloc(total * sizeof(float), false));
    cache_["kb|" + key] = dst;
    size_t off = 0;
    for (size_t i = 0; i < bnames.size(); ++i) {
        const float* src = w32(bnames[i], {Ns[i]}, kInitBias);
        check(check_cuda(cudaMemcpy(dst + off, src, Ns[i] * sizeof(float), cudaMemcpyDeviceToDevice), "cudaMemcpy(bias concat)"));
        off += Ns[i];
    }
    return dst;
}

const float* NetBase::gather_bias(const std::string& bname, int N, const std::vector<int>& rowmap) {
    auto hit = cache_.find("gb|" + bname);
    if (hit != cache_.end()) return static_cast<const float*>(hit->second);
    const float* src = w32(bname, {N}, kInitBias);
    float* dst = static_cast<float*>(dev_alloc(N * sizeof(float), false));
    cache_["gb|" + bname] = dst;
    int* dmap = nullptr;
    check(check_cuda(cudaMalloc(reinterpret_cast<void**>(&dmap), N * sizeof(int)), "cudaMalloc(rowmap)"));
    check(check_cuda(cudaMemcpy(dmap, rowmap.data(), N * sizeof(int), cudaMemcpyHostToDevice), "cudaMemcpy(rowmap)"));
    check(gather_f32(nullptr, src, dst, N, dmap));
    cudaDeviceSynchronize();
    cudaFree(dmap);
    return dst;
}

Act NetBase::new_act(int B, int H, int W, int C, bool f32) {
    Act a;
    a.B = B; a.H = H; a.W = W; a.C = C; a.f32 = f32;
    a.p = pool_.get(a.bytes());
    return a;
}

void NetBase::release(Act& a) {
    if (a.p) pool_.put(a.p);
    a.p = nullptr;
}

Act NetBase::gn(const Act& x, const std::string& prefix, float eps, bool silu) {
    const float* w = w32(prefix + ".weight", {x.C}, kInitOnes);
    const float* b = w32(prefix + ".bias", {x.C}, kInitZeros);
    Act y = new_act(x.B, x.H, x.W, x.C);
    void* ws = gn_ws_;
    const size_t wsb = gn_ws_bytes_;
    const void* xp = x.p;
    void* yp = y.p;
    const int B = x.B, C = x.C, HW = x.H * x.W, s = silu ? 1 : 0, idt = x.dtype();
    const bool one = group_norm_fused_eligible(B, C, 0, HW, 32, idt);      // group_norm_nhwc picks the single-launch kernel itself
    plan_->push([=](cudaStream_t st) { return group_norm_nhwc(st, xp, idt, yp, SDOD_BF16, w, b, nullptr, B, C, HW, 32, eps, s, ws, wsb); }, one ? 1 : 2,
                std::string(one ? "gn1 " : "gn ") + std::string(x.f32 ? "f32" : "bf16") + " HW" + std::to_string(HW) + " C" + std::to_string(C));
    return y;
}

Act NetBase::gn2(const Act& x, const Act* x2, const std::string& prefix, float eps, bool silu, Act* raw_out) {
    const int Ca = x.C, Cb = x2 ? x2->C : 0, C = Ca + Cb;
    if (x2 && (x2->f32 != x.f32 || x2->M() != x.M())) throw std::runtime_error("gn2: sources must share dtype and rows");
    if (!group_norm_fused_eligible(x.B, Ca, Cb, x.H * x.W, 32, x.dtype())) {
        // large tensors (big batches): materialise the concatenation, two-kernel GroupNorm, separate cast
        Act cat = x;
        bool own = false;
        if (x2) { cat = concat(x, *x2); own = true; }
        Act y = gn(cat, prefix, eps, silu);
        if (raw_out) {
            if (cat.f32) *raw_out = to_bf16(cat);
            else if (own) { *raw_out = cat; own = false; }
            else throw std::runtime_error("gn2: the raw copy of a single bf16 source is the source itself");
        }
        if (own) release(cat);
        return y;
    }
    const float* w = w32(prefix + ".weight", {C}, kInitOnes);
    const float* b = w32(prefix + ".bias", {C}, kInitZeros);
    Act y = new_act(x.B, x.H, x.W, C);
    void* rawp = nullptr;
    if (raw_out) { *raw_out = new_act(x.B, x.H, x.W, C); rawp = raw_out->p; }
    void* ws = gn_ws_;
    const size_t wsb = gn_ws_bytes_;
    const void *xp = x.p, *x2p = x2 ? x2->p : nullptr;
    void* yp = y.p;
    const int B = x.B, HW = x.H * x.W, s = silu ? 1 : 0, idt = x.dtype();
    plan_->push([=](cudaStream_t st) { return group_norm_nhwc2(st, xp, Ca, x2p, Cb, idt, yp, SDOD_BF16, rawp, w, b, B, HW, 32, eps, s, ws, wsb); }, 1,
                "gn1 " + std::string(x.f32 ? "f32" : "bf16") + " HW" + std::to_string(HW) + " C" + std::to_string(C) + (x2 ? "cat" : "") + (raw_out ? "+raw" : ""));
    return y;
}

Act NetBase::ln(const Act& x, const std::string& prefix) {
    const float* w = w32(prefix + ".weight", {x.C}, kInitOnes);
    const float* b = w32(prefix + ".bias", {x.C}, kInitZeros);
    Act y = new_act(x.B, x.H, x.W, x.C);
    const void* xp = x.p;
    void* yp = y.p;
    const int rows = x.M(), C = x.C, idt = x.dtype();
    plan_->push([=](cudaStream_t st) { return sdod_layer_norm(st, xp, idt, yp, w, b, rows, C, 1e-5f); }, 1, "ln rows" + std::to_string(rows) + " C" + std::to_string(C));
    return y;
}

int NetBase::gemm_into(const sdod_gemm_desc& d, bool may_fail) {
    auto g = std::make_shared<GemmLaunch>();
    set_splitk_workspace(skw_);
    const int st_prep = gemm_prepare(d, g.get());
    set_splitk_workspace(SplitKWorkspace{});              // never leave a pointer to this net's scratch behind
    if (may_fail && st_prep != kOk) return st_prep;
    check(st_prep);
    note_gemm(g);
    plan_->push([g](cudaStream_t st) { return gemm_launch(*g, st); }, (g->mp.split > 1 && !g->mp.split_cluster) ? 2 : 1,
                "gemm M" + std::to_string(d.M) + " N" + std::to_string(d.N) + " K" + std::to_string(d.K) + " bn" + std::to_string(g->bn) + (g->mp.streamk ? " sk" + std::to_string(g->sk_grid) : " split" + std::to_string(g->mp.split)) +
                    (d.batch > 1 ? " batch" + std::to_string(d.batch) : "") + (g->mp.ln_fuse ? " +ln" : ""));
    return kOk;
}

Act NetBase::linear(const Act& x, const void* w_bf16, int N, const LinearOpts& o) {
    const int n_out = o.act == SDOD_ACT_GEGLU ? N / 2 : N;
    if (x.f32) throw std::runtime_error("linear: GEMM operand must be bf16");
    const bool in_place = o.in_place && o.residual && o.out_f32 && o.residual->f32 && o.residual->C == n_out && o.residual->M() == x.M();
    Act y = in_place ? *o.residual : new_act(x.B, x.H, x.W, n_out, o.out_f32);
    sdod_gemm_desc d{};
    d.A = x.p; d.lda = x.C; d.strideA = 0;
    d.W = w_bf16; d.ldw = x.C; d.strideW = 0;
    d.M = x.M(); d.N = N; d.K = x.C; d.batch = 1; d.block_n = o.block_n;
    d.epi.C = y.p; d.epi.ldc = n_out;
    d.epi.bias = o.bias;
    if (o.residual) { d.epi.residual = o.residual->p; d.epi.ldr = o.residual->C; d.epi.residual_f32 = o.residual->f32 ? 1 : 0; }
    d.epi.alpha = o.alpha; d.epi.act = o.act; d.epi.out_mode = o.out_f32 ? SDOD_OUT_F32 : SDOD_OUT_BF16;
    if (o.ln_out) {
        static const int env = [] { const char* e = std::getenv("SDOD_LN_FUSE"); return e ? std::atoi(e) : 1; }();
        if (!o.out_f32) throw std::runtime_error("linear: a LayerNorm output needs the fp32 stream output");
        *o.ln_out = new_act(x.B, x.H, x.W, n_out, false);
        sdod_gemm_desc dl = d;
        dl.epi.ln_out = o.ln_out->p; dl.epi.ld_ln = n_out; dl.epi.ln_weight = o.ln_w; dl.epi.ln_bias = o.ln_b; dl.epi.ln_eps = 1e-5f;
        if (env && gemm_into(dl, true) == kOk) return y;
        gemm_into(d);                                        // shape not eligible: plain GEMM + a LayerNorm launch
        const void* yp = y.p;
        void* lp = o.ln_out->p;
        const float *w = o.ln_w, *b = o.ln_b;
        const int rows = y.M(), C = n_out;
        plan_->push([=](cudaStream_t st) { return sdod_layer_norm(st, yp, SDOD_F32, lp, w, b, rows, C, 1e-5f); }, 1, "ln rows" + std::to_string(rows) + " C" + std::to_string(C));
        return y;
    }
    gemm_into(d);
    return y;
}

Act NetBase::conv3(const Act& x, const std::string& prefix, int cout, const float* row_bias, long long ld_row_bias, const Act* residual,
                   bool stream_out, float* out_f32) {
    if (x.f32) throw std::runtime_error("conv3: operand must be bf16");
    void* wt = pack_conv3(prefix + ".weight", cout, x.C);
    const float* bias = w32(prefix + ".bias", {cout}, kInitBias);
    Act y;
    if (!out_f32) y = new_act(x.B, x.H, x.W, cout, stream_out);
    sdod_conv_desc d{};
    d.X = x.p; d.Wt = wt; d.B = x.B; d.H = x.H; d.W = x.W; d.Cin = x.C; d.Cout = cout;
    d.epi.C = out_f32 ? static_cast<void*>(out_f32) : y.p;
    d.epi.ldc = cout;
    d.epi.bias = bias;
    d.epi.row_bias = row_bias; d.epi.rows_per_group = x.H * x.W; d.epi.ld_row_bias = ld_row_bias;
    if (residual) { d.epi.residual = residual->p; d.epi.ldr = residual->C; d.epi.residual_f32 = residual->f32 ? 1 : 0; }
    d.epi.alpha = 1.0f; d.epi.out_mode = (out_f32 || stream_out) ? SDOD_OUT_F32 : SDOD_OUT_BF16;
    auto g = std::make_shared<GemmLaunch>();
    set_splitk_workspace(skw_);
    const int st_prep = conv3x3_prepare(d, g.get());
    set_splitk_workspace(SplitKWorkspace{});
    check(st_prep);
    note_gemm(g);
    plan_->push([g](cudaStream_t st) { return gemm_launch(*g, st); }, (g->mp.split > 1 && !g->mp.split_cluster) ? 2 : 1,
                "conv3 HW" + std::to_string(x.H * x.W) + " Cin" + std::to_string(x.C) + " Cout" + std::to_string(cout) + " bn" + std::to_string(g->bn) + (g->mp.streamk ? " sk" + std::to_string(g->sk_grid) : " split" + std::to_string(g->mp.split)));
    return y;
}

Act NetBase::conv3_skip(const Act& x, const std::string& prefix, const std::string& skip_prefix, const Act& x_skip, int cout) {
    if (x.f32 || x_skip.f32) throw std::runtime_error("conv3_skip: operands must be bf16");
    if (x_skip.M() != x.M()) throw std::runtime_error("conv3_skip: row mismatch");
    void* wt = pack_conv3_skip(prefix + ".weight", skip_prefix + ".weight", cout, x.C, x_skip.C);
    const float* bias = sum_bias(prefix + ".bias", skip_prefix + ".bias", cout);
    Act y = new_act(x.B, x.H, x.W, cout, true);
    sdod_conv_desc d{};
    d.X = x.p; d.Wt = wt; d.B = x.B; d.H = x.H; d.W = x.W; d.Cin = x.C; d.Cout = cout;
    d.X2 = x_skip.p; d.ldx2 = x_skip.C; d.Cin2 = x_skip.C;
    d.epi.C = y.p; d.epi.ldc = cout; d.epi.bias = bias; d.epi.alpha = 1.0f; d.epi.out_mode = SDOD_OUT_F32;
    auto g = std::make_shared<GemmLaunch>();
    set_splitk_workspace(skw_);
    const int st_prep = conv3x3_prepare(d, g.get());
    set_splitk_workspace(SplitKWorkspace{});
    check(st_prep);
    note_gemm(g);
    plan_->push([g](cudaStream_t st) { return gemm_launch(*g, st); }, (g->mp.split > 1 && !g->mp.split_cluster) ? 2 : 1,
                "conv3+skip HW" + std::to_string(x.H * x.W) + " Cin" + std::to_string(x.C) + "+" + std::to_string(x_skip.C) + " Cout" + std::to_string(cout) + " bn" + std::to_string(g->bn) +
                    (g->mp.streamk ? " sk" + std::to_string(g->sk_grid) : " split" + std::to_string(g->mp.split)));
    return y;
}

void* NetBase::pack_conv3_up2(const std::string& wname, int Cout, int Cin) {
    const std::string key = "cu|" + wname;
    auto hit = cache_.find(key);
    if (hit != cache_.end()) return hit->second;
    const float* src = w32(wname, {Cout, Cin, 3, 3}, kInitWeight);
    void* dst = dev_alloc(static_cast<size_t>(4) * Cout * 4 * Cin * 2, false);
    cache_[key] = dst;
    check(sdod_pack_conv3x3_up2_weight(nullptr, src, dst, Cout, Cin));
    return dst;
}

Act NetBase::conv3_up2(const Act& x, const std::string& prefix, int cout, bool stream_out) {
    static const int env = [] { const char* e = std::getenv("SDOD_CONV_UP2"); return e ? std::atoi(e) : 1; }();     // 0: upsample kernel + 3x3 conv (A/B)
    if (!env || x.C % 64 != 0 || cout % 8 != 0) {
        Act u = upsample(x);
        Act c = conv3(u, prefix, cout, nullptr, 0, nullptr, stream_out);
        release(u);
        return c;
    }
    Act xb = x;
    bool own = false;
    if (x.f32) { xb = to_bf16(x); own = true; }
    void* wt = pack_conv3_up2(prefix + ".weight", cout, x.C);
    const float* bias = w32(prefix + ".bias", {cout}, kInitBias);
    Act y = new_act(x.B, 2 * x.H, 2 * x.W, cout, stream_out);
    sdod_conv_desc d{};
    d.X = xb.p; d.Wt = wt; d.B = x.B; d.H = x.H; d.W = x.W; d.Cin = x.C; d.Cout = cout; d.upsample2x = 1;
    d.epi.C = y.p; d.epi.ldc = cout; d.epi.bias = bias; d.epi.alpha = 1.0f; d.epi.out_mode = stream_out ? SDOD_OUT_F32 : SDOD_OUT_BF16;
    auto g = std::make_shared<GemmLaunch>();
    set_splitk_workspace(skw_);
    const int st_prep = conv3x3_prepare(d, g.get());
    set_splitk_workspace(SplitKWorkspace{});
    check(st_prep);
    note_gemm(g);
    plan_->push([g](cudaStream_t st) { return gemm_launch(*g, st); }, 1,
                "conv3up2 HW" + std::to_string(x.H * x.W) + " Cin" + std::to_string(x.C) + " Cout" + std::to_string(cout) + " bn" + std::to_string(g->bn));
    if (own) release(xb);
    return y;
}

Act NetBase::conv3_s2(const Act& x, const std::string& prefix, int cout, bool stream_out) {
    static const int env = [] { const char* e = std::getenv("SDOD_CONV_S2"); return e ? std::atoi(e) : 1; }();      // 0: im2col + GEMM (A/B)
    if (!env || x.C % 64 != 0 || x.H % 2 != 0 || x.W % 2 != 0) return conv3_im2col(x, prefix, cout, 2, stream_out);
    Act xb = x;
    bool own = false;
    if (x.f32) { xb = to_bf16(x); own = true; }
    void* wt = pack_conv3(prefix + ".weight", cout, x.C);
    const float* bias = w32(prefix + ".bias", {cout}, kInitBias);
    Act y = new_act(x.B, x.H / 2, x.W / 2, cout, stream_out);
    sdod_conv_desc d{};
    d.X = xb.p; d.Wt = wt; d.B = x.B; d.H = x.H; d.W = x.W; d.Cin = x.C; d.Cout = cout; d.stride = 2;
    d.epi.C = y.p; d.epi.ldc = cout; d.epi.bias = bias; d.epi.alpha = 1.0f; d.epi.out_mode = stream_out ? SDOD_OUT_F32 : SDOD_OUT_BF16;
    auto g = std::make_shared<GemmLaunch>();
    set_splitk_workspace(skw_);
    const int st_prep = conv3x3_prepare(d, g.get());
    set_splitk_workspace(SplitKWorkspace{});
    check(st_prep);
    note_gemm(g);
    plan_->push([g](cudaStream_t st) { return gemm_launch(*g, st); }, (g->mp.split > 1 && !g->mp.split_cluster) ? 2 : 1,
                "conv3s2 HW" + std::to_string(y.H * y.W) + " Cin" + std::to_string(x.C) + " Cout" + std::to_string(cout) + " bn" + std::to_string(g->bn) +
                    (g->mp.streamk ? " sk" + std::to_string(g->sk_grid) : " split" + std::to_string(g->mp.split)));
    if (own) release(xb);
    return y;
}

Act NetBase::to_bf16(const Act& x) {
    if (!x.f32) throw std::runtime_error("to_bf16: already bf16");
    Act y = new_act(x.B, x.H, x.W, x.C, false);
    const float* xp = static_cast<const float*>(x.p);
    void* yp = y.p;
    const size_t n = static_cast<size_t>(x.M()) * x.C;
    plan_->push([=](cudaStream_t st) { return sdod_cast_f32_to_bf16(st, xp, yp, n); }, 1, "cast_f32_bf16");
    return y;
}

Act NetBase::conv3_im2col(const Act& x, const std::string& prefix, int cout, int stride, bool stream_out) {
    const int K = 9 * x.C;
    const int Kpad = (K + 63) / 64 * 64;
    void* wt = pack_conv3(prefix + ".weight", cout, x.C, Kpad);
    const float* bias = w32(prefix + ".bias", {cout}, kInitBias);
    const int Ho = (x.H + stride - 1) / stride, Wo = (x.W + stride - 1) / stride;
    Act cols = new_act(x.B, Ho, Wo, Kpad);
    {
        const void* xp = x.p;
        void* cp = cols.p;
        const int B = x.B, H = x.H, W = x.W, C = x.C, idt = x.dtype();
        plan_->push([=](cudaStream_t st) { return sdod_im2col3x3(st, xp, idt, cp, B, H, W, C, stride, Kpad); }, 1, "im2col");
    }
    LinearOpts o;
    o.bias = bias;
    o.out_f32 = stream_out;
    Act y = linear(cols, wt, cout, o);
    release(cols);
    return y;
}

Act NetBase::conv1x1(const Act& x, const std::string& prefix, int cout, const Act* residual, bool stream_out, bool in_place) {
    void* w = pack_linear(prefix + ".weight", cout, x.C);
    LinearOpts o;
    o.bias = w32(prefix + ".bias", {cout}, kInitBias);
    o.residual = residual;
    o.out_f32 = stream_out;
    o.in_place = in_place;
    return linear(x, w, cout, o);
}

Act NetBase::upsample(const Act& x) {
    Act y = new_act(x.B, 2 * x.H, 2 * x.W, x.C, false);
    const void* xp = x.p;
    void* yp = y.p;
    const int B = x.B, H = x.H, W = x.W, C = x.C, idt = x.dtype();
    plan_->push([=](cudaStream_t st) { return sdod_upsample2x_nhwc(st, xp, idt, yp, B, H, W, C); }, 1, "upsample2x");
    return y;
}

Act NetBase::concat(const Act& a, const Act& b) {
    if (a.f32 != b.f32) throw std::runtime_error("concat: dtype mismatch");
    Act y = new_act(a.B, a.H, a.W, a.C + b.C, a.f32);
    const void *ap = a.p, *bp = b.p;
    void* yp = y.p;
    const int Ca = a.C, Cb = b.C, dt = a.dtype();
    const long long rows = a.M();
    plan_->push([=](cudaStream_t st) { return sdod_concat_channels(st, ap, Ca, bp, Cb, yp, rows, dt); }, 1, "concat");
    return y;
}

}  // namespace sdod
