// Sampler-side device kernels: fused CFG combine + eps->x0 + DPM-Solver++(2M) update, sinusoidal
// timestep features, Gaussian noise init, image quantisation.
//
// Reference behaviour restated (csrc/libsdod/src/):
//   CFG     context.cpp:359-373 via QnnTensor::get_data(scale, accum) -> qnn_context.cpp:1065-1081
//   update  dpm_solver.cpp:136-181 (normalize / scale / accumulate helpers :58-75)
//   temb    context.cpp:257-275   noise context.cpp:333-334   uint8 map context.cpp:392-395
// The host code performs ~6 separate passes with one float rounding per operation; the fused kernel
// keeps every one of those roundings (__fmul_rn/__fadd_rn/__fdiv_rn forbid FMA contraction), so the
// fp32 result is bit-identical to the reference for the same eps inputs.
#include "../common.cuh"
#include "../host_common.h"
#include "../launch_count.h"
#include "sdod_kernels.h"

namespace sdod {

struct StepCoef {
    float g, one_minus_g, neg_sigma, alpha, c_x, c_prev, c_y0;
    int order, cfg;
};

template <typename E> SDOD_DEVICE float eps_load(const E* p, size_t i);
template <> SDOD_DEVICE float eps_load<float>(const float* p, size_t i) { return p[i]; }
template <> SDOD_DEVICE float eps_load<bf16>(const bf16* p, size_t i) { return __bfloat162float(p[i]); }

SDOD_DEVICE float dpm_one(float x, float ec, float eu, float& yprev, const StepCoef& k) {
    float e;
    if (k.cfg) {
        e = __fmul_rn(ec, k.g);                                  // e = g*eps_c            (context.cpp:362)
        e = __fadd_rn(e, __fmul_rn(eu, k.one_minus_g));          // e += (1-g)*eps_u       (context.cpp:373)
    } else {
        e = ec;                                                  // g == 1: uncond skipped (context.cpp:359-360)
    }
    const float y0 = __fdiv_rn(__fadd_rn(x, __fmul_rn(k.neg_sigma, e)), k.alpha);   // dpm_solver.cpp:139
    float xn = __fmul_rn(x, k.c_x);                              // scale                  (:153 / :168)
    if (k.order == 2) xn = __fadd_rn(xn, __fmul_rn(k.c_prev, yprev));               // :169
    xn = __fadd_rn(xn, __fmul_rn(k.c_y0, y0));                   // :154 / :170
    yprev = y0;                                                  // :177-180
    return xn;
}

template <typename E>
__global__ void __launch_bounds__(256) cfg_dpm_step_kernel(float* __restrict__ x, float* __restrict__ y_prev, const E* __restrict__ eps_c,
                                                           const E* __restrict__ eps_u, size_t n, StepCoef k, float* __restrict__ x_copy) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float yp = (k.order == 2) ? y_prev[i] : 0.f;
        const float ec = eps_load<E>(eps_c, i);
        const float eu = k.cfg ? eps_load<E>(eps_u, i) : 0.f;
        const float xn = dpm_one(x[i], ec, eu, yp, k);
        x[i] = xn;
        y_prev[i] = yp;
        if (x_copy) x_copy[i] = xn;
    }
}

// CFG split over a GPU pair: this rank evaluated ONE half of classifier-free guidance (role 0: cond, role 1: uncond) and the
// peer the other.  One kernel does the whole per-step exchange + update: (1) every CTA stores its part of the local eps straight into the
// PEER's exchange slot (remote stores over NVLink through an IPC-mapped pointer); the last CTA to finish, after a system-scope fence,
// raises the peer's flag to this step's sequence number; (2) every CTA waits until the peer has done the same for us, then applies the
// identical fused CFG + DPM update on both ranks (same inputs, same roundings as cfg_dpm_step_kernel), so x / y_prev stay replicated
// bit for bit and no second exchange is needed.  The grid is capped at one CTA per SM so that every CTA is resident before any of
// them waits; the wait is bounded and traps instead of hanging.  Slots and flags alternate with the sequence number's parity.
__global__ void __launch_bounds__(256) cfg_dpm_step_pair_kernel(float* __restrict__ x, float* __restrict__ y_prev, const float* __restrict__ eps_local,
                                                                float* peer_slot, const float* recv_slot, unsigned int* peer_flag,
                                                                const unsigned int* my_flag, unsigned int* done_counter, unsigned int seq, int role,
                                                                size_t n, StepCoef k) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    const size_t first = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    for (size_t i = first; i < n; i += stride) peer_slot[i] = eps_local[i];
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        if (atomicAdd(done_counter, 1u) == gridDim.x - 1) {
            *done_counter = 0;                                           // re-armed for the next step (stream order)
            __threadfence_system();
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(peer_flag), "r"(seq) : "memory");
        }
        unsigned int seen = 0, spins = 0;
        do {
            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(my_flag) : "memory");
            if (static_cast<int>(seen - seq) < 0 && ++spins > (1u << 24)) __trap();      // the peer never arrived: fail the launch, do not hang
        } while (static_cast<int>(seen - seq) < 0);
    }
    __syncthreads();
    const float* eps_c = role == 0 ? eps_local : recv_slot;
    const float* eps_u = role == 0 ? recv_slot : eps_local;
    for (size_t i = first; i < n; i += stride) {
        float yp = (k.order == 2) ? y_prev[i] : 0.f;
        const float xn = dpm_one(x[i], __ldcg(eps_c + i), __ldcg(eps_u + i), yp, k);
        x[i] = xn;
        y_prev[i] = yp;
    }
}

// PLMS / linear-multistep form of the same fusion (public CompVis plms.py, parity unpinned): CFG combine, optional store of the
// combined eps into the history ring, e' = w0 e + w1 h1 + w2 h2 + w3 h3, x0 = (x_from - s_t e') / a_t, x = a_prev x0 + s_prev e'.
// DDIM is w = (1,0,0,0); the first PLMS step calls it twice (second call: x_from = the saved x_t, h1 = the first eps, w = (1/2, 1/2)).
struct LmsCoef {
    float g, one_minus_g;
    float w0, w1, w2, w3;
    float a_t, s_t, a_prev, s_prev;
    int cfg;
};

template <typename E>
__global__ void __launch_bounds__(256) cfg_lms_step_kernel(float* __restrict__ x, const float* __restrict__ x_from, const E* __restrict__ eps_c,
                                                           const E* __restrict__ eps_u, const float* h1, const float* h2, const float* h3,
                                                           float* e_out, size_t n, LmsCoef k, float* __restrict__ x_copy) {
    // (h1..h3 / e_out are slots of one history ring and e_out may BE the oldest slot: history is read before e_out is written)
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float e = eps_load<E>(eps_c, i);
        if (k.cfg) e = __fadd_rn(__fmul_rn(e, k.g), __fmul_rn(eps_load<E>(eps_u, i), k.one_minus_g));
        float ep = __fmul_rn(k.w0, e);
        if (h1) ep = __fadd_rn(ep, __fmul_rn(k.w1, h1[i]));
        if (h2) ep = __fadd_rn(ep, __fmul_rn(k.w2, h2[i]));
        if (h3) ep = __fadd_rn(ep, __fmul_rn(k.w3, h3[i]));
        if (e_out) e_out[i] = e;
        const float xs = x_from ? x_from[i] : x[i];
        const float x0 = __fdiv_rn(__fadd_rn(xs, -__fmul_rn(k.s_t, ep)), k.a_t);
        const float xn = __fadd_rn(__fmul_rn(k.a_prev, x0), __fmul_rn(k.s_prev, ep));
        x[i] = xn;
        if (x_copy) x_copy[i] = xn;
    }
}

__global__ void timestep_sinusoid_kernel(const float* __restrict__ t, int n_t, int dim, float log_period, float* __restrict__ out) {
    const int half = dim / 2;
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_t * half) return;
    const int i = idx / half, j = idx - i * half;
    // arg = t * exp(log_period * j / half)   (context.cpp:271), cos first (:272-273)
    const float arg = __fmul_rn(t[i], expf(__fdiv_rn(__fmul_rn(log_period, static_cast<float>(j)), static_cast<float>(half))));
    out[static_cast<size_t>(i) * dim + j] = cosf(arg);
    out[static_cast<size_t>(i) * dim + half + j] = sinf(arg);
}

// Philox4x32-10 counter RNG + Box-Muller.  (The reference draws from std::mt19937 /
// std::normal_distribution, context.cpp:16,333-334 — implementation-defined bit stream, so parity
// is distributional; tests inject identical latents into both paths. SURVEY App. C.)
SDOD_DEVICE void philox4x32_10(uint32_t (&c)[4], uint32_t k0, uint32_t k1) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
        const uint32_t n0 = hi1 ^ c[1] ^ k0, n1 = lo1, n2 = hi0 ^ c[3] ^ k1, n3 = lo0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

__global__ void randn_kernel(float* __restrict__ out, size_t n, unsigned long long seed, unsigned long long offset) {
    const size_t q = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;   // one Philox block = 4 normals
    if (q * 4 >= n) return;
    const unsigned long long ctr = offset + q;
    uint32_t c[4] = {static_cast<uint32_t>(ctr), static_cast<uint32_t>(ctr >> 32), 0u, 0u};
    philox4x32_10(c, static_cast<uint32_t>(seed), static_cast<uint32_t>(seed >> 32));
    float z[4];
#pragma unroll
    for (int h = 0; h < 2; ++h) {
        const float u1 = (static_cast<float>(c[2 * h] >> 8) + 0.5f) * (1.0f / 16777216.0f);       // (0,1)
        const float u2 = (static_cast<float>(c[2 * h + 1] >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float r = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincospif(2.0f * u2, &sn, &cs);
        z[2 * h] = r * cs;
        z[2 * h + 1] = r * sn;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (q * 4 + i < n) out[q * 4 + i] = z[i];
}

template <typename T>
__global__ void image_to_u8_kernel(const T* __restrict__ img, uint8_t* __restrict__ out, size_t n) {
    const size_t stride = static_cast<size_t>(gridDim.x) * blockDim.x;
    for (size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
        float f;
        if constexpr (sizeof(T) == 4) f = img[i]; else f = __bfloat162float(img[i]);
        float v = __fmul_rn(255.0f, f);
        v = fminf(fmaxf(v, 0.0f), 255.0f);       // std::clamp(255*f, 0, 255)
        out[i] = static_cast<uint8_t>(v);        // truncation, not rounding (context.cpp:394)
    }
}

static inline int grid_for(size_t n, int block, int per_thread = 1) {
    size_t g = (n + static_cast<size_t>(block) * per_thread - 1) / (static_cast<size_t>(block) * per_thread);
    const size_t cap = 148 * 16;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    return static_cast<int>(g);
}

}  // namespace sdod

using namespace sdod;

extern "C" {

SDOD_API int sdod_cfg_dpm_step_pair(sdod_stream_t stream, float* x, float* y_prev, const float* eps_local, float* peer_slot, const float* recv_slot,
                                    unsigned int* peer_flag, const unsigned int* my_flag, unsigned int* done_counter, unsigned int seq, int role,
                                    size_t n, float guidance, float sigma_s, float alpha_s, float c_x, float c_prev, float c_y0, int order) {
    if (!x || !y_prev || !eps_local || !peer_slot || !recv_slot || !peer_flag || !my_flag || !done_counter)
        return fail(kInvalidArgument, "cfg_dpm_step_pair: NULL pointer");
    if (order != 1 && order != 2) return fail(kInvalidArgument, "cfg_dpm_step_pair: order must be 1 or 2");
    if (role != 0 && role != 1) return fail(kInvalidArgument, "cfg_dpm_step_pair: role must be 0 (cond) or 1 (uncond)");
    if (guidance == 1.0f) return fail(kInvalidArgument, "cfg_dpm_step_pair: guidance 1 has no unconditional half to split off");
    StepCoef k;
    k.g = guidance;
    k.one_minus_g = 1 - guidance;
    k.neg_sigma = -sigma_s;
    k.alpha = alpha_s;
    k.c_x = c_x; k.c_prev = c_prev; k.c_y0 = c_y0;
    k.order = order;
    k.cfg = 1;
    if (n == 0) return kOk;
    int grid = grid_for(n, 256);
    const int sms = device_sm_count();
    if (grid > sms) grid = sms;                     // every CTA resident before any of them waits for the peer
    cfg_dpm_step_pair_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, y_prev, eps_local, peer_slot, recv_slot, peer_flag, my_flag, done_counter,
                                                                                 seq, role, n, k);
    count_launch();
    return check_launch("cfg_dpm_step_pair_kernel");
}

SDOD_API int sdod_cfg_dpm_step(sdod_stream_t stream, float* x, float* y_prev, const void* eps_c, const void* eps_u, int eps_dtype,
                               size_t n, float guidance, float sigma_s, float alpha_s, float c_x, float c_prev, float c_y0, int order,
                               float* x_copy) {
    if (!x || !y_prev || !eps_c) return fail(kInvalidArgument, "cfg_dpm_step: NULL tensor");
    if (order != 1 && order != 2) return fail(kInvalidArgument, "cfg_dpm_step: order must be 1 or 2");
    StepCoef k;
    k.g = guidance;
    k.one_minus_g = 1 - guidance;            // float, as `1-guidance` at context.cpp:373
    k.neg_sigma = -sigma_s;
    k.alpha = alpha_s;
    k.c_x = c_x; k.c_prev = c_prev; k.c_y0 = c_y0;
    k.order = order;
    k.cfg = (guidance == 1.0f) ? 0 : 1;      // exact compare, context.cpp:359
    if (k.cfg && !eps_u) return fail(kInvalidArgument, "cfg_dpm_step: eps_u required when guidance != 1");
    if (n == 0) return kOk;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n, 256);
    if (eps_dtype == SDOD_F32)
        cfg_dpm_step_kernel<float><<<grid, 256, 0, s>>>(x, y_prev, static_cast<const float*>(eps_c), static_cast<const float*>(eps_u), n, k, x_copy);
    else if (eps_dtype == SDOD_BF16)
        cfg_dpm_step_kernel<bf16><<<grid, 256, 0, s>>>(x, y_prev, static_cast<const bf16*>(eps_c), static_cast<const bf16*>(eps_u), n, k, x_copy);
    else
        return fail(kInvalidArgument, "cfg_dpm_step: unknown eps dtype");
    count_launch();
    return check_launch("cfg_dpm_step_kernel");
}

SDOD_API int sdod_cfg_lms_step(sdod_stream_t stream, float* x, const float* x_from, const void* eps_c, const void* eps_u, int eps_dtype, size_t n,
                               float guidance, const float* w, const float* h1, const float* h2, const float* h3, float* e_out, float a_t,
                               float s_t, float a_prev, float s_prev, float* x_copy) {
    if (!x || !eps_c || !w) return fail(kInvalidArgument, "cfg_lms_step: NULL tensor");
    LmsCoef k;
    k.g = guidance;
    k.one_minus_g = 1 - guidance;
    k.w0 = w[0]; k.w1 = w[1]; k.w2 = w[2]; k.w3 = w[3];
    k.a_t = a_t; k.s_t = s_t; k.a_prev = a_prev; k.s_prev = s_prev;
    k.cfg = (guidance == 1.0f) ? 0 : 1;
    if (k.cfg && !eps_u) return fail(kInvalidArgument, "cfg_lms_step: eps_u required when guidance != 1");
    if (n == 0) return kOk;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    const int grid = grid_for(n, 256);
    if (eps_dtype == SDOD_F32)
        cfg_lms_step_kernel<float><<<grid, 256, 0, s>>>(x, x_from, static_cast<const float*>(eps_c), static_cast<const float*>(eps_u), h1, h2, h3, e_out, n, k, x_copy);
    else if (eps_dtype == SDOD_BF16)
        cfg_lms_step_kernel<bf16><<<grid, 256, 0, s>>>(x, x_from, static_cast<const bf16*>(eps_c), static_cast<const bf16*>(eps_u), h1, h2, h3, e_out, n, k, x_copy);
    else
        return fail(kInvalidArgument, "cfg_lms_step: unknown eps dtype");
    count_launch();
    return check_launch("cfg_lms_step_kernel");
}

SDOD_API int sdod_timestep_sinusoid(sdod_stream_t stream, const float* t_dev, int n_t, int dim, float max_period, float* out) {
    if (!t_dev || !out || n_t <= 0 || dim <= 0 || (dim & 1)) return fail(kInvalidArgument, "timestep_sinusoid: bad arguments (dim must be even)");
    const float log_period = -logf(max_period);   // context.cpp:261
    const int total = n_t * (dim / 2);
    timestep_sinusoid_kernel<<<(total + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(t_dev, n_t, dim, log_period, out);
    count_launch();
    return check_launch("timestep_sinusoid_kernel");
}

SDOD_API int sdod_randn(sdod_stream_t stream, float* out, size_t n, unsigned long long seed, unsigned long long offset) {
    if (!out) return fail(kInvalidArgument, "randn: NULL output");
    if (n == 0) return kOk;
    const size_t blocks = (n + 3) / 4;
    randn_kernel<<<static_cast<unsigned>((blocks + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(out, n, seed, offset);
    count_launch();
    return check_launch("randn_kernel");
}

SDOD_API int sdod_image_to_u8(sdod_stream_t stream, const void* img, int dtype, uint8_t* out, size_t n) {
    if (!img || !out) return fail(kInvalidArgument, "image_to_u8: NULL tensor");
    if (n == 0) return kOk;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    if (dtype == SDOD_F32) image_to_u8_kernel<float><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const float*>(img), out, n);
    else if (dtype == SDOD_BF16) image_to_u8_kernel<bf16><<<grid_for(n, 256), 256, 0, s>>>(static_cast<const bf16*>(img), out, n);
    else return fail(kInvalidArgument, "image_to_u8: unknown dtype");
    count_launch();
    return check_launch("image_to_u8_kernel");
}

}  // extern "C"
