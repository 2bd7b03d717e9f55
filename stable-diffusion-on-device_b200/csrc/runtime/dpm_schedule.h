// DPM-Solver++(2M) schedule tables and per-step coefficients (host side of the sampler).
// Behavioural contract: reference csrc/libsdod/src/dpm_solver.cpp:84-131 (tables) and :137-170
// (order rule, coefficient products).  Compiled with -ffp-contract=off: every float rounding of
// the reference's x86-64 build is preserved so the tables match bit for bit.
#pragma once
#include <vector>

namespace sdod {

struct DpmStep {
    float sigma_s, alpha_s;     // eps -> x0:  y0 = (x - sigma_s*e) / alpha_s
    float c_x, c_prev, c_y0;    // x <- c_x*x (+ c_prev*y_prev) + c_y0*y0
    int order;                  // 1 or 2
};

class DpmSchedule {
public:
    DpmSchedule(unsigned timesteps, float lin_start, float lin_end);
    void prepare(unsigned steps);
    unsigned steps() const { return steps_; }
    DpmStep step(unsigned s) const;

    std::vector<float> all_t, all_log_alpha;                                   // [timesteps]
    std::vector<float> ts, log_alphas, lambdas, sigmas, alphas, phis, i2rs, model_ts;   // [steps+1]

private:
    unsigned timesteps_;
    unsigned steps_ = 0;
};

// DDIM (eta = 0) on the same fused kernel: x0 = (x - sqrt(1-a_t) e) / sqrt(a_t), x <- sqrt(a_prev) x0 + sqrt(1-a_prev) e, which in the kernel's
// form is an order-1 step with c_x = sqrt(1-a_prev)/sqrt(1-a_t), c_y0 = sqrt(a_prev) - c_x sqrt(a_t).  Not in the reference tree (SURVEY §8 row
// f4; the reference's README only names the flag of the external ldm fork): restated from the public CompVis `ddim.py` — "uniform"
// timesteps i*(T/steps)+1, alphas_cumprod of the scaled-linear beta schedule in float64 — so its parity is unpinned.
class DdimSchedule {
public:
    DdimSchedule(unsigned timesteps, float lin_start, float lin_end);
    void prepare(unsigned steps);
    unsigned steps() const { return static_cast<unsigned>(model_ts.size()); }
    DpmStep step(unsigned s) const { return coeffs.at(s); }
    std::vector<float> model_ts;        // [steps], descending integer timesteps fed to the UNet
    std::vector<DpmStep> coeffs;        // [steps]
    std::vector<float> sqrt_a_prev, sqrt_1m_a_prev;   // [steps]  sqrt(a_prev), sqrt(1 - a_prev) (the PLMS step wants them unfolded)
private:
    std::vector<double> alphas_cumprod_;
};

}  // namespace sdod
