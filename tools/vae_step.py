"""Run the VAE decoder (BASELINE config C4: batch 8, 64x64x4 -> 512x512x3) a few times, eager — the command profiled by ncu for profiles/."""
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import model as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
vae = M.VaeDecoder(None, seed=1, latent_hw=64, max_batch=B)
z = torch.randn(B, 4, 64, 64, device="cuda")
torch.cuda.synchronize()
for _ in range(n):
    img = vae(z)
torch.cuda.synchronize()
img = img[0] if isinstance(img, (tuple, list)) else img
print("ok", img.float().mean().item())
