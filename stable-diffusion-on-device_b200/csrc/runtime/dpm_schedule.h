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

}  // namespace sdod
