"""Hot conv / GEMM shapes of the UNet under the current SDOD_GEMM_* environment (CUDA events, L2 flushed)."""
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


BN = int(os.environ.get("SDOD_SWEEP_BN", "0"))
tag = "bn=%d " % BN + " ".join("%s=%s" % (k, v) for k, v in sorted(os.environ.items()) if k.startswith("SDOD_GEMM"))
for (B, HW, Cin, Cout) in [(8, 64, 320, 320), (2, 64, 320, 320), (8, 32, 640, 640), (8, 16, 1280, 1280), (2, 16, 1280, 1280), (8, 32, 1280, 640), (8, 16, 2560, 1280), (8, 32, 640, 1280)]:
    x = torch.randn(B, HW, HW, Cin, device="cuda").to(torch.bfloat16)
    w = ops.pack_conv3x3_weight((torch.randn(Cout, Cin, 3, 3, device="cuda") / (9 * Cin) ** 0.5))
    bias = torch.randn(Cout, device="cuda")
    us = timeit(lambda: torch.ops.sdod.conv3x3(x, w, bias, None, None, 0, BN))
    fl = 2.0 * B * HW * HW * Cout * 9 * Cin
    print("[%s] conv B%d HW%d %d->%d: %.1f us %.0f TF/s" % (tag, B, HW, Cin, Cout, us, fl / us * 1e-6))
for (M, N, K) in [(32768, 320, 1280), (8192, 640, 2560), (2048, 1280, 5120), (32768, 320, 320), (2048, 1280, 1280), (2048, 3840, 1280)]:
    a = torch.randn(M, K, device="cuda").to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    us = timeit(lambda: torch.ops.sdod.linear(a, w, None, None, 0, 1.0, False, None, 0, BN))
    print("[%s] gemm M%d N%d K%d: %.1f us %.0f TF/s" % (tag, M, N, K, us, 2.0 * M * N * K / us * 1e-6))
