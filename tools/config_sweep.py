"""BASELINE.json configs C4 (VAE decoder alone, batch 8) and C5 (images per generate call 1..16 on one GPU): device times, CUDA events."""
import json
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import libsdod as A  # noqa: E402
from sdod import model as M  # noqa: E402

out = {}
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
vae = M.VaeDecoder(None, seed=1, latent_hw=64, max_batch=8)
z = torch.randn(8, 4, 64, 64, device="cuda")      # NCHW, as VaeDecoder.__call__ takes it
for _ in range(3):
    vae(z)
ts = []
for _ in range(5):
    flush.zero_()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    vae(z)
    b.record()
    torch.cuda.synchronize()
    ts.append(a.elapsed_time(b))
ts.sort()
ms = ts[len(ts) // 2]
out["C4_vae_decode_batch8"] = {"ms": ms, "tflops": 8 * 2.5145 / (ms * 1e-3), "gn_gbs_if_all_time_were_gn": 8 * 1.84e9 / (ms * 1e-3) * 1e-9}
del vae
for n in (1, 2, 4, 8, 16):
    with A.Context("random-init:0", latent_spatial=64, steps=20, max_images=n, device=0) as ctx:
        cond = torch.randn(n, 77, 768, device="cuda")
        unc = torch.randn(n, 77, 768, device="cuda")
        lat = torch.randn(n, 64, 64, 4, device="cuda")
        img = torch.empty(n, 512, 512, 3, dtype=torch.uint8, device="cuda")
        for _ in range(2):
            ctx.generate_device(cond, unc, lat, 7.5, img)
        tt = []
        for _ in range(3):
            flush.zero_()
            ctx.generate_device(cond, unc, lat, 7.5, img)
            tt.append(ctx.last_timings())
        tt.sort(key=lambda t: t["total_ms"])
        t = tt[1]
        out["C5_images_per_call_%d" % n] = {"images_per_s": n / (t["total_ms"] * 1e-3), "total_ms": t["total_ms"], "iteration_ms": t["iteration_ms"],
                                           "decoding_ms": t["decoding_ms"]}
print(json.dumps(out, indent=1))
