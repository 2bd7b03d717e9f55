/*
 * ORACLE — TEST INFRASTRUCTURE ONLY.  Not part of the product path.
 *
 * Plain-C CPU restatement of the in-tree numerics on the reference's hot path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this.  The product (libsdod_b200.so) never links it.
 *
 * Pinned against: oracle/_ref/libdpm_ref.so (the reference's own
 * csrc/libsdod/src/dpm_solver.cpp compiled from /root/reference) and the golden
 * tables of SURVEY.md Appendix A (tests/golden/dpm_tables_20.json).
 * CFG combine / sinusoid / uint8 map are NOT pinned by any reference test
 * ("parity unpinned" in the reference; pinned here by restatement only).
 *
 * Every function cites the reference file:line it restates.  All arithmetic
 * keeps the reference's float/double promotion order; compile with
 * -ffp-contract=off so no FMA contraction changes a rounding.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdint.h>

#define ORACLE_API __attribute__((visibility("default")))

/* ---- dpm_solver.cpp:12-26  linspace<float>: float accumulator, double step ---- */
static void linspace_f(float *buf, float start, float end, unsigned num_steps, unsigned offset) {
    double step = (double)(end - start) / (double)(num_steps - 1);
    unsigned insert = 0;
    for (unsigned i = 0; i < num_steps; ++i) {
        if (!offset) buf[insert++] = start;
        else --offset;
        start = (float)((double)start + step);
    }
}

/* ---- dpm_solver.cpp:29-32  _interpolate (all float) ---- */
static float interp2(float x, float x1, float y1, float x2, float y2) {
    float a = (y2 - y1) / (x2 - x1);
    return a * (x - x1) + y1;
}

/* ---- dpm_solver.cpp:36-55  interpolate with descending hint; out-of-range -> global secant ---- */
static float interpolate_f(float x, const float *xs, const float *ys, unsigned n, unsigned *hint) {
    if (x < xs[0] || x > xs[n - 1])
        return interp2(x, xs[n - 1], ys[n - 1], xs[0], ys[0]);
    while (*hint > 1 && xs[*hint - 1] > x) --(*hint);
    return interp2(x, xs[*hint - 1], ys[*hint - 1], xs[*hint], ys[*hint]);
}

typedef struct {
    unsigned total_timesteps;
    unsigned steps;            /* 0 until prepare() */
    float *all_t, *all_log_alpha;                      /* [total_timesteps] */
    float *ts, *log_alphas, *lambdas, *sigmas, *alphas, *phis, *i2rs, *model_ts; /* [steps+1] */
    float *prev_y; size_t prev_n;                      /* persists across trajectories (dpm_solver.cpp:177-180) */
} dpm_oracle;

/* ---- dpm_solver.cpp:84-97  ctor: sqrt-linear betas -> 0.5*log(cumprod alpha) (cum in double) ---- */
ORACLE_API dpm_oracle *dpm_oracle_create(unsigned timesteps, float lin_start, float lin_end) {
    dpm_oracle *s = (dpm_oracle *)calloc(1, sizeof(dpm_oracle));
    s->total_timesteps = timesteps;
    s->all_t = (float *)malloc(sizeof(float) * timesteps);
    s->all_log_alpha = (float *)malloc(sizeof(float) * timesteps);
    linspace_f(s->all_t, 0.0f, 1.0f, timesteps + 1, 1);
    linspace_f(s->all_log_alpha, sqrtf(lin_start), sqrtf(lin_end), timesteps, 0);
    double cum = 1.0;
    for (unsigned i = 0; i < timesteps; ++i) {
        float b = s->all_log_alpha[i];
        b = 1 - b * b;
        cum *= b;
        s->all_log_alpha[i] = (float)(0.5 * log(cum));
    }
    return s;
}

ORACLE_API void dpm_oracle_destroy(dpm_oracle *s) {
    if (!s) return;
    free(s->all_t); free(s->all_log_alpha);
    free(s->ts); free(s->log_alphas); free(s->lambdas); free(s->sigmas);
    free(s->alphas); free(s->phis); free(s->i2rs); free(s->model_ts); free(s->prev_y);
    free(s);
}

/* ---- dpm_solver.cpp:100-131  prepare(steps) ---- */
ORACLE_API void dpm_oracle_prepare(dpm_oracle *s, unsigned steps) {
    unsigned n = steps + 1;
    float **arrs[] = {&s->ts, &s->log_alphas, &s->lambdas, &s->sigmas, &s->alphas, &s->phis, &s->i2rs, &s->model_ts};
    for (unsigned k = 0; k < 8; ++k) { free(*arrs[k]); *arrs[k] = (float *)malloc(sizeof(float) * n); }
    s->steps = steps;
    float first_t = 1.0f;
    float last_t = (float)(1.0 / s->total_timesteps);
    linspace_f(s->ts, first_t, last_t, n, 0);
    unsigned hint = s->total_timesteps;
    for (unsigned i = 0; i < n; ++i) {
        s->model_ts[i] = (float)(((double)s->ts[i] - 1.0 / s->total_timesteps) * 1000);
        float la = interpolate_f(s->ts[i], s->all_t, s->all_log_alpha, s->total_timesteps, &hint);
        s->log_alphas[i] = la;
        s->lambdas[i] = (float)((double)la - (0.5 * (double)logf(1 - expf(2 * la))));
        s->sigmas[i] = sqrtf(1 - expf(2 * la));
        s->alphas[i] = expf(la);
        s->phis[i] = i ? expm1f(-(s->lambdas[i] - s->lambdas[i - 1])) : INFINITY;
        if (i >= 2)
            s->i2rs[i] = (float)(1.0 / (double)(2 * ((s->lambdas[i - 1] - s->lambdas[i - 2]) / (s->lambdas[i] - s->lambdas[i - 1]))));
        else
            s->i2rs[i] = INFINITY;
    }
}

/* table id: 0 ts 1 log_alphas 2 lambdas 3 sigmas 4 alphas 5 phis 6 i2rs 7 model_ts 8 all_t 9 all_log_alpha */
ORACLE_API unsigned dpm_oracle_table(const dpm_oracle *s, int which, float *out, unsigned cap) {
    const float *src = NULL; unsigned n = s->steps + 1;
    switch (which) {
        case 0: src = s->ts; break;        case 1: src = s->log_alphas; break;
        case 2: src = s->lambdas; break;   case 3: src = s->sigmas; break;
        case 4: src = s->alphas; break;    case 5: src = s->phis; break;
        case 6: src = s->i2rs; break;      case 7: src = s->model_ts; break;
        case 8: src = s->all_t; n = s->total_timesteps; break;
        case 9: src = s->all_log_alpha; n = s->total_timesteps; break;
        default: return 0;
    }
    if (!src) return 0;
    if (n > cap) n = cap;
    memcpy(out, src, sizeof(float) * n);
    return n;
}

/* ---- dpm_solver.cpp:136-181  update(step, x, y): eps -> x0, 1st/2nd order multistep ---- */
ORACLE_API void dpm_oracle_update(dpm_oracle *s, unsigned step, float *x, float *y, size_t n) {
    unsigned tsz = s->steps + 1;
    unsigned order = (step == 0) ? 1u : (step < 10 ? ((tsz - step) < 2u ? (tsz - step) : 2u) : 2u);
    {   /* normalize(y, x, y, -sigma, alpha): y = (x + (-sigma)*y) / alpha   (:139) */
        float a = -s->sigmas[step], b = s->alphas[step];
        for (size_t i = 0; i < n; ++i) y[i] = (x[i] + a * y[i]) / b;
    }
    float sc = s->sigmas[step + 1] / s->sigmas[step];
    for (size_t i = 0; i < n; ++i) x[i] *= sc;                              /* :153 / :168 */
    if (order == 1) {
        float a = -s->alphas[step + 1] * s->phis[step + 1];                 /* :154 */
        for (size_t i = 0; i < n; ++i) x[i] += a * y[i];
    } else {
        float a1 = s->alphas[step + 1] * s->phis[step + 1] * s->i2rs[step + 1];         /* :169 */
        float a2 = -s->alphas[step + 1] * s->phis[step + 1] * (1 + s->i2rs[step + 1]);  /* :170 */
        for (size_t i = 0; i < n; ++i) x[i] += a1 * s->prev_y[i];
        for (size_t i = 0; i < n; ++i) x[i] += a2 * y[i];
    }
    if (!s->prev_y) {                                                       /* :177-180 */
        s->prev_y = (float *)malloc(sizeof(float) * n); s->prev_n = n;
        memcpy(s->prev_y, y, sizeof(float) * n);
    } else {
        for (size_t i = 0; i < n; ++i) { float t = y[i]; y[i] = s->prev_y[i]; s->prev_y[i] = t; }
    }
}

/* ---- CFG combine: context.cpp:359-373 via qnn_context.cpp:1065-1081 simple_cast<Accum,Scale> ----
 * g == 1.0f: e = eps_c (uncond pass skipped).  else e = g*eps_c ; e += (1-g)*eps_u               */
ORACLE_API void cfg_oracle_combine(float *e, const float *eps_c, const float *eps_u, float g, size_t n) {
    if (g == 1.0f) { memcpy(e, eps_c, sizeof(float) * n); return; }
    for (size_t i = 0; i < n; ++i) e[i] = eps_c[i] * g;
    float s2 = 1 - g;
    for (size_t i = 0; i < n; ++i) e[i] += eps_u[i] * s2;
}

/* ---- timestep sinusoid: context.cpp:257-275 (cos first, then sin; mode_dim even) ---- */
ORACLE_API void temb_oracle_sinusoid(float t, unsigned mode_dim, float max_period, float *mode) {
    float log_period = -logf(max_period);
    unsigned half = mode_dim / 2;
    for (unsigned j = 0; j < half; ++j) {
        float arg = t * expf(log_period * j / half);
        mode[j] = cosf(arg);
        mode[half + j] = sinf(arg);
    }
}

/* ---- image quantise: context.cpp:392-395  uint8(clamp(255*f, 0, 255)) — truncation ---- */
ORACLE_API void image_oracle_to_u8(const float *img, uint8_t *out, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float v = 255 * img[i];
        v = v < 0.0f ? 0.0f : (v > 255.0f ? 255.0f : v);
        out[i] = (uint8_t)v;
    }
}
