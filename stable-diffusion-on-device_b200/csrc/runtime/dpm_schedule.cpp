#include "dpm_schedule.h"

#include <cmath>
#include <limits>
#include <stdexcept>

#include "../host_common.h"
#include "sdod_kernels.h"

namespace sdod {
namespace {

// float running value advanced by a double increment (dpm_solver.cpp:12-26)
std::vector<float> ramp(float from, float to, unsigned count, unsigned skip) {
    std::vector<float> out;
    out.reserve(count - skip);
    const double inc = static_cast<double>(to - from) / (count - 1);
    float v = from;
    for (unsigned i = 0; i < count; ++i) {
        if (i >= skip) out.push_back(v);
        v = static_cast<float>(static_cast<double>(v) + inc);
    }
    return out;
}

float secant(float x, float xa, float ya, float xb, float yb) {       // dpm_solver.cpp:29-32
    const float slope = (yb - ya) / (xb - xa);
    return slope * (x - xa) + ya;
}

}  // namespace

DpmSchedule::DpmSchedule(unsigned timesteps, float lin_start, float lin_end) : timesteps_(timesteps) {
    if (timesteps < 2) throw std::invalid_argument("DpmSchedule: timesteps must be >= 2");
    all_t = ramp(0.0f, 1.0f, timesteps + 1, 1);
    all_log_alpha = ramp(std::sqrt(lin_start), std::sqrt(lin_end), timesteps, 0);   // sqrt(beta), linear
    double cumulative = 1.0;
    for (float& v : all_log_alpha) {
        const float alpha = 1 - v * v;
        cumulative *= alpha;
        v = static_cast<float>(0.5 * std::log(cumulative));
    }
}

void DpmSchedule::prepare(unsigned steps) {
    if (steps < 1) throw std::invalid_argument("DpmSchedule: steps must be >= 1");
    const unsigned n = steps + 1;
    ts = ramp(1.0f, static_cast<float>(1.0 / timesteps_), n, 0);
    for (auto* v : {&log_alphas, &lambdas, &sigmas, &alphas, &phis, &i2rs, &model_ts}) v->assign(n, 0.f);
    const float inf = std::numeric_limits<float>::infinity();
    size_t cursor = all_t.size();     // descending search position (dpm_solver.cpp:36-55)
    for (unsigned i = 0; i < n; ++i) {
        const float t = ts[i];
        model_ts[i] = static_cast<float>((static_cast<double>(t) - 1.0 / timesteps_) * 1000);
        float la;
        if (t < all_t.front() || t > all_t.back()) {
            la = secant(t, all_t.back(), all_log_alpha.back(), all_t.front(), all_log_alpha.front());
        } else {
            while (cursor > 1 && all_t[cursor - 1] > t) --cursor;
            if (cursor >= all_t.size()) cursor = all_t.size() - 1;
            la = secant(t, all_t[cursor - 1], all_log_alpha[cursor - 1], all_t[cursor], all_log_alpha[cursor]);
        }
        log_alphas[i] = la;
        const float e2 = std::exp(2 * la);
        lambdas[i] = static_cast<float>(static_cast<double>(la) - 0.5 * static_cast<double>(std::log(1 - e2)));
        sigmas[i] = std::sqrt(1 - e2);
        alphas[i] = std::exp(la);
        phis[i] = i ? std::expm1(-(lambdas[i] - lambdas[i - 1])) : inf;
        if (i >= 2) {
            const float ratio = (lambdas[i - 1] - lambdas[i - 2]) / (lambdas[i] - lambdas[i - 1]);
            i2rs[i] = static_cast<float>(1.0 / static_cast<double>(2 * ratio));
        } else {
            i2rs[i] = inf;
        }
    }
    steps_ = steps;
}

DpmStep DpmSchedule::step(unsigned s) const {
    if (s >= steps_) throw std::out_of_range("DpmSchedule::step");
    const unsigned n = steps_ + 1;
    DpmStep k{};
    // order rule, dpm_solver.cpp:137
    unsigned order = (s == 0) ? 1u : (s < 10 ? std::min(2u, n - s) : 2u);
    k.order = static_cast<int>(order);
    k.sigma_s = sigmas[s];
    k.alpha_s = alphas[s];
    k.c_x = sigmas[s + 1] / sigmas[s];
    if (order == 1) {
        k.c_prev = 0.f;
        k.c_y0 = -alphas[s + 1] * phis[s + 1];                              // :154
    } else {
        k.c_prev = alphas[s + 1] * phis[s + 1] * i2rs[s + 1];               // :169
        k.c_y0 = -alphas[s + 1] * phis[s + 1] * (1 + i2rs[s + 1]);          // :170
    }
    return k;
}

}  // namespace sdod

namespace sdod {

DdimSchedule::DdimSchedule(unsigned timesteps, float lin_start, float lin_end) {
    if (timesteps < 2) throw std::invalid_argument("DdimSchedule: timesteps must be >= 2");
    // betas = linspace(sqrt(lin_start), sqrt(lin_end), T, float64) ** 2 ; alphas_cumprod = cumprod(1 - betas)
    alphas_cumprod_.resize(timesteps);
    const double b0 = std::sqrt(static_cast<double>(lin_start)), b1 = std::sqrt(static_cast<double>(lin_end));
    double cum = 1.0;
    for (unsigned i = 0; i < timesteps; ++i) {
        const double sb = b0 + (b1 - b0) * static_cast<double>(i) / static_cast<double>(timesteps - 1);
        cum *= 1.0 - sb * sb;
        alphas_cumprod_[i] = cum;
    }
}

void DdimSchedule::prepare(unsigned steps) {
    const unsigned T = static_cast<unsigned>(alphas_cumprod_.size());
    if (steps < 1 || steps > T) throw std::invalid_argument("DdimSchedule: steps must be in [1, timesteps]");
    const unsigned c = T / steps;                       // "uniform" spacing; t_i = i*c + 1, i in [0, steps)
    model_ts.assign(steps, 0.f);
    coeffs.assign(steps, DpmStep{});
    sqrt_a_prev.assign(steps, 0.f);
    sqrt_1m_a_prev.assign(steps, 0.f);
    for (unsigned k = 0; k < steps; ++k) {              // loop step k walks the timesteps downwards
        const unsigned i = steps - 1 - k;
        const unsigned t = i * c + 1;
        const double a_t = alphas_cumprod_[t < T ? t : T - 1];
        const double a_prev = i > 0 ? alphas_cumprod_[(i - 1) * c + 1] : alphas_cumprod_[0];
        const double cx = std::sqrt(1.0 - a_prev) / std::sqrt(1.0 - a_t);
        DpmStep s{};
        s.sigma_s = static_cast<float>(std::sqrt(1.0 - a_t));
        s.alpha_s = static_cast<float>(std::sqrt(a_t));
        s.c_x = static_cast<float>(cx);
        s.c_prev = 0.f;
        s.c_y0 = static_cast<float>(std::sqrt(a_prev) - cx * std::sqrt(a_t));
        s.order = 1;
        coeffs[k] = s;
        sqrt_a_prev[k] = static_cast<float>(std::sqrt(a_prev));
        sqrt_1m_a_prev[k] = static_cast<float>(std::sqrt(1.0 - a_prev));
        model_ts[k] = static_cast<float>(t);
    }
}

}  // namespace sdod

extern "C" {
SDOD_API int sdod_ddim_schedule(unsigned timesteps, float lin_start, float lin_end, unsigned steps, float* model_ts, float* coeffs) {
    try {
        sdod::DdimSchedule sch(timesteps, lin_start, lin_end);
        sch.prepare(steps);
        for (unsigned k = 0; k < steps; ++k) {
            const sdod::DpmStep s = sch.step(k);
            if (model_ts) model_ts[k] = sch.model_ts[k];
            if (coeffs) { coeffs[5 * k] = s.sigma_s; coeffs[5 * k + 1] = s.alpha_s; coeffs[5 * k + 2] = s.c_x; coeffs[5 * k + 3] = s.c_prev; coeffs[5 * k + 4] = s.c_y0; }
        }
    } catch (const std::exception& e) {
        return sdod::fail(sdod::kInvalidArgument, std::string("ddim_schedule: ") + e.what());
    }
    return sdod::kOk;
}


SDOD_API int sdod_dpm_schedule(unsigned timesteps, float lin_start, float lin_end, unsigned steps, float* ts, float* log_alphas,
                               float* lambdas, float* sigmas, float* alphas, float* phis, float* i2rs, float* model_ts) {
    try {
        sdod::DpmSchedule sch(timesteps, lin_start, lin_end);
        sch.prepare(steps);
        auto put = [&](float* dst, const std::vector<float>& v) {
            if (dst) for (size_t i = 0; i < v.size(); ++i) dst[i] = v[i];
        };
        put(ts, sch.ts); put(log_alphas, sch.log_alphas); put(lambdas, sch.lambdas); put(sigmas, sch.sigmas);
        put(alphas, sch.alphas); put(phis, sch.phis); put(i2rs, sch.i2rs); put(model_ts, sch.model_ts);
    } catch (const std::exception& e) {
        return sdod::fail(sdod::kInvalidArgument, std::string("dpm_schedule: ") + e.what());
    }
    return sdod::kOk;
}

SDOD_API int sdod_dpm_coeffs(unsigned timesteps, float lin_start, float lin_end, unsigned steps, unsigned step, float* sigma_s,
                             float* alpha_s, float* c_x, float* c_prev, float* c_y0, int* order) {
    try {
        sdod::DpmSchedule sch(timesteps, lin_start, lin_end);
        sch.prepare(steps);
        const sdod::DpmStep k = sch.step(step);
        if (sigma_s) *sigma_s = k.sigma_s;
        if (alpha_s) *alpha_s = k.alpha_s;
        if (c_x) *c_x = k.c_x;
        if (c_prev) *c_prev = k.c_prev;
        if (c_y0) *c_y0 = k.c_y0;
        if (order) *order = k.order;
    } catch (const std::exception& e) {
        return sdod::fail(sdod::kInvalidArgument, std::string("dpm_coeffs: ") + e.what());
    }
    return sdod::kOk;
}
}
