"""One GEMM shape through the C ABI, several launches — the target of `ncu --set full` captures."""
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import ops  # noqa: E402

M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (8192, 320, 320)
mode = sys.argv[4] if len(sys.argv) > 4 else "res32"
a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
bias = torch.randn(N, device="cuda")
res = torch.randn(M, N, device="cuda")
for _ in range(5):
    if mode == "res32":
        y = torch.ops.sdod.linear(a, w, bias, res, 0, 1.0, True)
    elif mode == "plain":
        y = torch.ops.sdod.linear(a, w)
    torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
ev[0].record()
for _ in range(20):
    y = torch.ops.sdod.linear(a, w, bias, res, 0, 1.0, True) if mode == "res32" else torch.ops.sdod.linear(a, w)
ev[1].record()
torch.cuda.synchronize()
print("M%d N%d K%d %s: %.1f us per call (back-to-back, incl. host)" % (M, N, K, mode, ev[0].elapsed_time(ev[1]) * 50))
