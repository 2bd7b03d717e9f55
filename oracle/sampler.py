"""ORACLE — TEST INFRASTRUCTURE ONLY.  ctypes front-end for oracle/liboracle.so
(plain-C restatement, sampler_oracle.c) and, when built, oracle/_ref/libdpm_ref.so
(the reference's own csrc/libsdod/src/dpm_solver.cpp compiled from /root/reference).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
TABLES = ["ts", "log_alphas", "lambdas", "sigmas", "alphas", "phis", "i2rs", "model_ts", "all_t", "all_log_alpha"]
_f32p = ctypes.POINTER(ctypes.c_float)


def _fp(a):
    return a.ctypes.data_as(_f32p)


def build(ref=True):
    """Compile the C restatement (and the reference shim when /root/reference exists)."""
    subprocess.run(["make", "-s", "-C", _HERE, "oracle"], check=True)
    if ref and os.path.isdir("/root/reference/csrc/libsdod/src"):
        subprocess.run(["make", "-s", "-C", _HERE, "ref"], check=True)


def _load(path):
    lib = ctypes.CDLL(path)
    return lib


class _Solver:
    """Common driver: `prefix` selects dpm_oracle_* (restatement) or dpm_ref_* (reference)."""

    def __init__(self, lib, prefix, timesteps=1000, lin_start=0.00085, lin_end=0.0120):
        self._lib, self._p = lib, prefix
        f = lambda n: getattr(lib, prefix + n)
        f("create").restype = ctypes.c_void_p
        f("create").argtypes = [ctypes.c_uint, ctypes.c_float, ctypes.c_float]
        f("destroy").argtypes = [ctypes.c_void_p]
        f("prepare").argtypes = [ctypes.c_void_p, ctypes.c_uint]
        f("table").restype = ctypes.c_uint
        f("table").argtypes = [ctypes.c_void_p, ctypes.c_int, _f32p, ctypes.c_uint]
        f("update").argtypes = [ctypes.c_void_p, ctypes.c_uint, _f32p, _f32p, ctypes.c_size_t]
        self._h = f("create")(timesteps, lin_start, lin_end)
        self.timesteps = timesteps
        self.steps = 0

    def __del__(self):
        if getattr(self, "_h", None):
            getattr(self._lib, self._p + "destroy")(self._h)
            self._h = None

    def prepare(self, steps):
        getattr(self._lib, self._p + "prepare")(self._h, steps)
        self.steps = steps

    def table(self, name):
        which = TABLES.index(name)
        cap = self.timesteps if which >= 8 else self.steps + 1
        out = np.empty(cap, dtype=np.float32)
        n = getattr(self._lib, self._p + "table")(self._h, which, _fp(out), cap)
        return out[:n]

    def tables(self):
        return {k: self.table(k) for k in TABLES[:8]}

    def update(self, step, x, y):
        """In place on float32 arrays x (latent) and y (eps in; x0 / swapped prev out)."""
        assert x.dtype == np.float32 and y.dtype == np.float32 and x.size == y.size
        getattr(self._lib, self._p + "update")(self._h, step, _fp(x), _fp(y), x.size)


_ORACLE_SO = os.path.join(_HERE, "liboracle.so")
_REF_SO = os.path.join(_HERE, "_ref", "libdpm_ref.so")


def oracle_lib():
    if not os.path.exists(_ORACLE_SO):
        build(ref=False)
    return _load(_ORACLE_SO)


def ref_available():
    return os.path.exists(_REF_SO)


def OracleSolver(**kw):
    return _Solver(oracle_lib(), "dpm_oracle_", **kw)


def RefSolver(**kw):
    if not ref_available():
        raise FileNotFoundError(_REF_SO + " (run `make -C oracle ref` where /root/reference exists)")
    return _Solver(_load(_REF_SO), "dpm_ref_", **kw)


def cfg_combine(eps_c, eps_u, g):
    lib = oracle_lib()
    lib.cfg_oracle_combine.argtypes = [_f32p, _f32p, _f32p, ctypes.c_float, ctypes.c_size_t]
    eps_c = np.ascontiguousarray(eps_c, dtype=np.float32)
    eps_u = np.ascontiguousarray(eps_u, dtype=np.float32)
    e = np.empty_like(eps_c)
    lib.cfg_oracle_combine(_fp(e), _fp(eps_c), _fp(eps_u), g, e.size)
    return e


def sinusoid(t, mode_dim=320, max_period=10000.0):
    lib = oracle_lib()
    lib.temb_oracle_sinusoid.argtypes = [ctypes.c_float, ctypes.c_uint, ctypes.c_float, _f32p]
    out = np.empty(mode_dim, dtype=np.float32)
    lib.temb_oracle_sinusoid(t, mode_dim, max_period, _fp(out))
    return out


def to_u8(img):
    lib = oracle_lib()
    lib.image_oracle_to_u8.argtypes = [_f32p, ctypes.POINTER(ctypes.c_uint8), ctypes.c_size_t]
    img = np.ascontiguousarray(img, dtype=np.float32)
    out = np.empty(img.shape, dtype=np.uint8)
    lib.image_oracle_to_u8(_fp(img), out.ctypes.data_as(ctypes.POINTER(ctypes.c_uint8)), img.size)
    return out


def seeded_trajectory(solver, steps=20, n=8):
    """SURVEY.md Appendix A trajectory: x[i]=0.25(i-3.5); e[i]=0.1(((7i+3s) mod 11)-5)."""
    solver.prepare(steps)
    x = (0.25 * (np.arange(n, dtype=np.float64) - 3.5)).astype(np.float32)
    out = []
    for s in range(steps):
        e = (0.1 * (((7 * np.arange(n) + 3 * s) % 11) - 5).astype(np.float64)).astype(np.float32)
        solver.update(s, x, e)
        out.append(x.copy())
    return np.stack(out)


# ---------------------------------------------------------------------------------------------- DDIM (row f4)
def ddim_tables(steps=20, timesteps=1000, lin_start=0.00085, lin_end=0.0120):
    """ORACLE — DDIM (eta = 0) schedule, restated from the public CompVis stable-diffusion `ldm/models/diffusion/ddim.py` /
    `ldm/modules/diffusionmodules/util.py` (make_beta_schedule 'linear', make_ddim_timesteps 'uniform', make_ddim_sampling_parameters).
    The reference tree has no DDIM (README.md:61-62 only names the flag of its external ldm fork): PARITY UNPINNED.
    Returns (t [steps] descending ints, a_t, a_prev) in float64."""
    betas = np.linspace(lin_start ** 0.5, lin_end ** 0.5, timesteps, dtype=np.float64) ** 2
    alphas_cumprod = np.cumprod(1.0 - betas, axis=0)
    c = timesteps // steps
    ddim_t = (np.arange(steps) * c) + 1                       # first `steps` entries of range(0, T, c) + 1
    a = alphas_cumprod[np.minimum(ddim_t, timesteps - 1)]
    a_prev = np.asarray([alphas_cumprod[0]] + alphas_cumprod[ddim_t[:-1]].tolist())
    return ddim_t[::-1].copy(), a[::-1].copy(), a_prev[::-1].copy()


def ddim_update(x, e, a_t, a_prev):
    """x_{t-1} = sqrt(a_prev) * pred_x0 + sqrt(1 - a_prev) * e,  pred_x0 = (x - sqrt(1 - a_t) e) / sqrt(a_t)   (ddim.py p_sample_ddim, sigma = 0);
    float32 tensors, float64 scalars rounded to float32 as torch does when a python/numpy scalar meets a float32 tensor."""
    f = np.float32
    pred_x0 = (x - f(np.sqrt(1.0 - a_t)) * e) / f(np.sqrt(a_t))
    return f(np.sqrt(a_prev)) * pred_x0 + f(np.sqrt(1.0 - a_prev)) * e


def plms_step_weights(n_hist):
    """Adams-Bashforth weights over (e_t, e_{t-1}, e_{t-2}, e_{t-3}) used by plms.py p_sample_plms once n_hist old eps exist (n_hist >= 1)."""
    return {1: (3 / 2, -1 / 2, 0, 0), 2: (23 / 12, -16 / 12, 5 / 12, 0)}.get(n_hist, (55 / 24, -59 / 24, 37 / 24, -9 / 24))
