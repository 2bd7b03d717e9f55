"""Generate the committed golden fixtures from the REFERENCE itself (run in the build container only;
/root/reference does not exist on the GPU box).

  dpm_golden.npz : tables + seeded trajectories from oracle/_ref/libdpm_ref.so, i.e. the reference's own
                   csrc/libsdod/src/dpm_solver.cpp compiled where it lies (steps 20 = SURVEY App. A, plus 10/50)
  gn_golden.npz  : inputs/outputs of the reference's own sdod.EfficientGN (imported from /root/reference)
                   for impl None / 'eff' — the tests/gn_to_ln.py case plus hot-path-shaped small cases
Usage: python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)

from oracle import sampler as S  # noqa: E402


def dpm():
    S.build(ref=True)
    out = {}
    for steps in (20, 10, 50):
        r = S.RefSolver()
        r.prepare(steps)
        for k in S.TABLES[:8]:
            out["s%d_%s" % (steps, k)] = r.table(k)
        out["s%d_traj" % steps] = S.seeded_trajectory(S.RefSolver(), steps=steps, n=8)
    r = S.RefSolver()
    out["all_t"], out["all_log_alpha"] = r.table("all_t"), r.table("all_log_alpha")
    # second trajectory on a used solver: prev_y persistence across generate() calls (dpm_solver.cpp:177-180)
    r = S.RefSolver()
    S.seeded_trajectory(r)
    out["s20_traj_second"] = S.seeded_trajectory(r)
    np.savez_compressed(os.path.join(HERE, "dpm_golden.npz"), **out)
    print("dpm_golden.npz", {k: v.shape for k, v in out.items() if k.startswith("s20")})


def gn():
    sys.path.insert(0, "/root/reference")
    import importlib
    ref = importlib.import_module("sdod")            # the reference package, NOT ours
    assert ref.__file__.startswith("/root/reference"), ref.__file__
    out = {}
    g = torch.Generator().manual_seed(1234)
    cases = [("gn_to_ln", (1, 8, 2, 2), 2, 1e-5, True),           # reference tests/gn_to_ln.py:7-31
             ("c320", (2, 320, 8, 8), 32, 1e-5, True),
             ("c640_eps6", (1, 640, 4, 4), 32, 1e-6, True),
             ("noaffine", (2, 64, 5, 3), 32, 1e-5, False),
             ("rank3", (2, 32, 24), 8, 1e-5, True),
             ("ragged", (3, 12, 3, 7), 4, 1e-5, True)]
    for name, shape, G, eps, affine in cases:
        for impl in (None, "eff"):
            m = ref.EfficientGN(G, shape[1], eps=eps, affine=affine, impl=impl)
            if affine:
                with torch.no_grad():
                    m.weight.copy_(torch.randn(shape[1], generator=g))
                    m.bias.copy_(torch.randn(shape[1], generator=g))
            x = torch.randn(shape, generator=g) * 2 + 0.5
            with torch.no_grad():
                y = m(x)
            key = "%s_%s" % (name, impl or "none")
            out[key + "_x"], out[key + "_y"] = x.numpy(), y.numpy()
            out[key + "_meta"] = np.array([G, eps, int(affine)], dtype=np.float64)
            if affine:
                out[key + "_w"], out[key + "_b"] = m.weight.detach().numpy(), m.bias.detach().numpy()
    np.savez_compressed(os.path.join(HERE, "gn_golden.npz"), **out)
    print("gn_golden.npz", len(out), "arrays")


if __name__ == "__main__":
    dpm()
    gn()
