"""CPU: pin the oracle against the reference's golden vectors and against the compiled reference."""
import os

import numpy as np
import pytest
import torch

from oracle import ldm_oracle as L
from oracle import sampler as S

APPENDIX_A = {  # SURVEY.md Appendix A, `echo 20 | ./test_dpm` printed with 6 decimals
    "model_ts": [999.000000, 949.049988, 899.099976, 849.149963, 799.200012, 749.250000, 699.299988, 649.349976, 599.399963, 549.449951,
                 499.499969, 449.549957, 399.599945, 349.649963, 299.699951, 249.749939, 199.799942, 149.849945, 99.899933, 49.949936, -0.000066],
    "log_alphas": [-2.684436, -2.393709, -2.124387, -1.875666, -1.646717, -1.436708, -1.244813, -1.070206, -0.912062, -0.769558, -0.641875,
                   -0.528191, -0.427690, -0.339553, -0.262965, -0.197112, -0.141181, -0.094357, -0.055831, -0.024790, -0.000425],
    "sigmas": [0.997668, 0.995824, 0.992833, 0.988187, 0.981261, 0.971336, 0.957632, 0.939358, 0.915773, 0.886245, 0.850296,
               0.807644, 0.758207, 0.702090, 0.639527, 0.570786, 0.495983, 0.414701, 0.325043, 0.219934, 0.029152],
    "i2rs": [np.inf, np.inf, 0.465400, 0.465265, 0.465612, 0.466506, 0.468047, 0.470331, 0.473430, 0.477414, 0.482350,
             0.488334, 0.495517, 0.504174, 0.514808, 0.528388, 0.546887, 0.574844, 0.624714, 0.747303, 2.425075],
}


def _bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


@pytest.fixture(scope="module")
def dpm_golden(golden_dir):
    return np.load(os.path.join(golden_dir, "dpm_golden.npz"))


@pytest.mark.parametrize("steps", [20, 10, 50])
def test_tables_bit_exact_vs_reference_golden(dpm_golden, steps):
    o = S.OracleSolver()
    o.prepare(steps)
    for k in S.TABLES[:8]:
        assert np.array_equal(_bits(o.table(k)), _bits(dpm_golden["s%d_%s" % (steps, k)])), k
    assert np.array_equal(_bits(o.table("all_t")), _bits(dpm_golden["all_t"]))
    assert np.array_equal(_bits(o.table("all_log_alpha")), _bits(dpm_golden["all_log_alpha"]))


def test_tables_match_survey_appendix_a():
    o = S.OracleSolver()
    o.prepare(20)
    for k, v in APPENDIX_A.items():
        got = o.table(k).astype(np.float64)
        want = np.array(v)
        fin = np.isfinite(want)
        assert np.array_equal(np.isfinite(got), fin)
        assert np.allclose(got[fin], want[fin], atol=6e-7 * np.maximum(1, np.abs(want[fin])).max(), rtol=0), k
    # endpoint quirk (App. C.2): log_alphas[0] comes from the out-of-range secant, not the table end
    assert abs(o.table("all_log_alpha")[-1] - (-2.68441057)) < 1e-6 and abs(o.table("log_alphas")[0] - (-2.684436)) < 1e-6


@pytest.mark.parametrize("steps", [20, 10, 50])
def test_trajectory_bit_exact(dpm_golden, steps):
    t = S.seeded_trajectory(S.OracleSolver(), steps=steps)
    assert np.array_equal(_bits(t), _bits(dpm_golden["s%d_traj" % steps]))


def test_trajectory_matches_appendix_a_print():
    t = S.seeded_trajectory(S.OracleSolver())
    assert np.allclose(t[0], [-1.000997, -0.9035658, -0.4338336, -0.3364028, 0.1333294, 0.6030616, 0.7004924, 1.170225], rtol=2e-6)
    assert np.allclose(t[19], [-11.28695, -10.67749, -5.74098, -1.681036, 0.7471199, 6.408724, 9.386093, 12.8329], rtol=2e-6)


def test_prev_y_persists_across_trajectories(dpm_golden):
    o = S.OracleSolver()
    S.seeded_trajectory(o)
    assert np.array_equal(_bits(S.seeded_trajectory(o)), _bits(dpm_golden["s20_traj_second"]))


@pytest.mark.skipif(not S.ref_available(), reason="oracle/_ref not built (needs /root/reference)")
def test_oracle_vs_compiled_reference_random():
    rng = np.random.default_rng(0)
    o, r = S.OracleSolver(), S.RefSolver()
    o.prepare(20), r.prepare(20)
    xo = rng.standard_normal(4096).astype(np.float32)
    xr = xo.copy()
    for s in range(20):
        e = rng.standard_normal(4096).astype(np.float32)
        eo, er = e.copy(), e.copy()
        o.update(s, xo, eo), r.update(s, xr, er)
        assert np.array_equal(_bits(xo), _bits(xr)) and np.array_equal(_bits(eo), _bits(er))


def test_cfg_sinusoid_u8_restatements():
    rng = np.random.default_rng(1)
    ec, eu = rng.standard_normal(1000).astype(np.float32), rng.standard_normal(1000).astype(np.float32)
    g = np.float32(7.5)
    want = ec * g
    want += eu * np.float32(1 - g)
    assert np.array_equal(_bits(S.cfg_combine(ec, eu, 7.5)), _bits(want))
    assert np.array_equal(_bits(S.cfg_combine(ec, eu, 1.0)), _bits(ec))          # g==1 skips uncond (context.cpp:359)
    m = S.sinusoid(999.0)
    j = np.arange(160, dtype=np.float64)
    arg = 999.0 * np.exp(-np.log(10000.0) * j / 160)
    assert np.allclose(m[:160], np.cos(arg), atol=2e-4) and np.allclose(m[160:], np.sin(arg), atol=2e-4)   # cos first
    img = np.array([-0.1, 0.0, 0.5, 0.99999, 1.0, 1.7, 0.00392, 0.003922], dtype=np.float32)
    assert S.to_u8(img).tolist() == [0, 0, 127, 254, 255, 255, 0, 1]                 # truncation, not rounding


def test_group_norm_f64_vs_reference_golden(golden_dir):
    gn = np.load(os.path.join(golden_dir, "gn_golden.npz"))
    keys = sorted(k[:-2] for k in gn.files if k.endswith("_x"))
    assert len(keys) == 12
    for k in keys:
        G, eps, affine = gn[k + "_meta"]
        x = torch.from_numpy(gn[k + "_x"])
        w = torch.from_numpy(gn[k + "_w"]) if affine else None
        b = torch.from_numpy(gn[k + "_b"]) if affine else None
        y = L.group_norm_f64(x, int(G), w, b, float(eps)).numpy()
        assert np.allclose(y, gn[k + "_y"], atol=2e-5, rtol=1e-5), k


def test_network_restatement_param_counts():
    assert L.count_params(L.make_unet()) == 859_520_964
    assert L.count_params(L.make_vae()) == 49_490_199


def test_network_restatement_runs_small():
    u, v = L.make_unet(), L.make_vae()
    with torch.no_grad():
        e = u(torch.randn(1, 4, 8, 8), u.embed_time(torch.tensor([999.0])), torch.randn(1, 77, 768))
        img = v(torch.randn(1, 4, 8, 8))
    assert e.shape == (1, 4, 8, 8) and torch.isfinite(e).all()
    assert img.shape == (1, 3, 64, 64) and img.min() >= 0 and img.max() <= 1
