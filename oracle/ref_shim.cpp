// ORACLE — TEST INFRASTRUCTURE ONLY.
// extern "C" shim over the reference's own libsdod::DPMSolver
// (/root/reference/csrc/libsdod/src/dpm_solver.{h,cpp}), compiled where the
// sources lie; no reference source is copied into this repo.  Output goes to
// oracle/_ref/libdpm_ref.so (git-ignored, travels to the GPU box).
#include "dpm_solver.h"
#include <vector>
#include <cstring>

using libsdod::DPMSolver;

struct ref_handle {
    DPMSolver solver;
    std::vector<float> model_ts;
    ref_handle(unsigned t, float a, float b) : solver(t, a, b) {}
};

template <class V>
static unsigned copy_out(V const& v, float* out, unsigned cap) {
    unsigned n = static_cast<unsigned>(v.size());
    if (n > cap) n = cap;
    std::memcpy(out, v.data(), sizeof(float) * n);
    return n;
}

extern "C" {
__attribute__((visibility("default"))) void* dpm_ref_create(unsigned timesteps, float lin_start, float lin_end) {
    return new ref_handle(timesteps, lin_start, lin_end);
}
__attribute__((visibility("default"))) void dpm_ref_destroy(void* h) { delete static_cast<ref_handle*>(h); }
__attribute__((visibility("default"))) void dpm_ref_prepare(void* h, unsigned steps) {
    auto* r = static_cast<ref_handle*>(h);
    r->solver.prepare(steps, r->model_ts);
}
__attribute__((visibility("default"))) unsigned dpm_ref_table(void* h, int which, float* out, unsigned cap) {
    auto* r = static_cast<ref_handle*>(h);
    switch (which) {
        case 0: return copy_out(r->solver.get_ts(), out, cap);
        case 1: return copy_out(r->solver.get_log_alphas(), out, cap);
        case 2: return copy_out(r->solver.get_lambdas(), out, cap);
        case 3: return copy_out(r->solver.get_sigmas(), out, cap);
        case 4: return copy_out(r->solver.get_alphas(), out, cap);
        case 5: return copy_out(r->solver.get_phis(), out, cap);
        case 6: return copy_out(r->solver.get_i2rs(), out, cap);
        case 7: return copy_out(r->model_ts, out, cap);
        case 8: return copy_out(r->solver.get_all_t(), out, cap);
        case 9: return copy_out(r->solver.get_all_log_alpha(), out, cap);
    }
    return 0;
}
__attribute__((visibility("default"))) void dpm_ref_update(void* h, unsigned step, float* x, float* y, size_t n) {
    auto* r = static_cast<ref_handle*>(h);
    std::vector<float> vx(x, x + n), vy(y, y + n);
    r->solver.update(step, vx, vy);
    std::memcpy(x, vx.data(), sizeof(float) * n);
    std::memcpy(y, vy.data(), sizeof(float) * n);
}
}
