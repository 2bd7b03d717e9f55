"""Run the batch-2 UNet step a few times (eager plan, no graph) — the command profiled by ncu for profiles/."""
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import model as M  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2
B = int(sys.argv[2]) if len(sys.argv) > 2 else 2
net = M.UNet(None, seed=0, latent_hw=64, max_batch=B)
net.set_context(torch.randn(B, 77, 768, device="cuda"))
x, emb = torch.randn(B, 64, 64, 4, device="cuda"), torch.randn(B, 1280, device="cuda")
torch.cuda.synchronize()
for _ in range(n):
    eps = net.forward_nhwc(x, emb, use_graph=False)
torch.cuda.synchronize()
print("ok", eps.float().abs().mean().item(), "launches/forward", net.launches_per_forward(B))
