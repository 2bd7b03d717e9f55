"""Phase timeline of single GEMM / conv launches (globaltimer stamps recorded by the kernel's profiling hook, sdod_set_gemm_timeline).
Each case runs warm inside a short chain of launches (so PDL overlap is as in the step); prints, per phase, the median and max over CTAs
in ns since the first CTA of the launch started.   python tools/gemm_timeline.py  -> gpurun_out/gemm_timeline.txt"""
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import _cabi as C  # noqa: E402
from sdod import ops  # noqa: E402

DEV = "cuda"
SLOTS = ["start", "setup", "dep_done", "ops_landed", "last_mma", "acc_done", "epi_done", "splitk_pub", "roles_done", "res_landed", "staged"]
out = []


def run(name, fn, n_cta_max=4096):
    buf = torch.zeros(n_cta_max * 16, dtype=torch.int64, device=DEV)
    for _ in range(3):
        fn()
    C.check(C.lib().sdod_set_gemm_timeline(buf.data_ptr()), "timeline")
    try:
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        fn()           # predecessor in the chain
        a.record()
        buf.zero_()
        fn()
        b.record()
        torch.cuda.synchronize()
    finally:
        C.check(C.lib().sdod_set_gemm_timeline(None), "timeline")
    t = buf.view(-1, 16).cpu()
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    line = "%-46s ctas %3d |" % (name, t.shape[0])
    for i, s in enumerate(SLOTS):
        col = t[:, i]
        col = col[col > 0]
        if len(col):
            line += " %s %d/%d" % (s, int((col - t0).float().median()), int((col - t0).max()))
    out.append(line)
    print(line)


def gemm(M, N, K, res=False, f32=False, splitk=False, a2=0, act=0):
    a, w = torch.randn(M, K, device=DEV).bfloat16(), torch.randn(N, K + a2, device=DEV).bfloat16()
    bias = torch.randn(N, device=DEV)
    r = (torch.randn(M, N, device=DEV) if f32 else torch.randn(M, N, device=DEV).bfloat16()) if res else None
    x2 = torch.randn(M, a2, device=DEV).bfloat16() if a2 else None
    ops.enable_splitk(splitk)
    return lambda: torch.ops.sdod.linear(a, w, bias, r, act, 1.0, f32, None, 0, 0, x2)


def conv(B, H, Cin, Cout, res=False, splitk=False):
    x = torch.randn(B, H, H, Cin, device=DEV).bfloat16()
    wt = torch.randn(Cout, 9 * Cin, device=DEV).bfloat16()
    bias = torch.randn(Cout, device=DEV)
    r = torch.randn(B, H, H, Cout, device=DEV).bfloat16() if res else None
    ops.enable_splitk(splitk)
    return lambda: torch.ops.sdod.conv3x3(x, wt, bias, r)


cases = [
    ("gemm M8192 N320 K320 f32res", lambda: gemm(8192, 320, 320, True, True)),
    ("gemm M8192 N320 K320 plain bf16", lambda: gemm(8192, 320, 320)),
    ("gemm M8192 N960 K320 bf16", lambda: gemm(8192, 960, 320)),
    ("gemm M8192 N320 K1280 f32res", lambda: gemm(8192, 320, 1280, True, True)),
    ("gemm M2048 N640 K640 f32res", lambda: gemm(2048, 640, 640, True, True)),
    ("gemm M512 N1280 K1280 f32res nosplit", lambda: gemm(512, 1280, 1280, True, True)),
    ("gemm M512 N1280 K1280 f32res split", lambda: gemm(512, 1280, 1280, True, True, True)),
    ("gemm M512 N1280 K5120 f32res split", lambda: gemm(512, 1280, 5120, True, True, True)),
    ("gemm M128 N1280 K1280 f32res split", lambda: gemm(128, 1280, 1280, True, True, True)),
    ("gemm M2048 N640 K2560 f32res split", lambda: gemm(2048, 640, 2560, True, True, True)),
    ("gemm M512 N10240 K1280 geglu split", lambda: gemm(512, 10240, 1280, False, False, True, 0, 3)),
    ("conv B2 64x64 320->320 bf16", lambda: conv(2, 64, 320, 320)),
    ("conv B2 32x32 640->640 split", lambda: conv(2, 32, 640, 640, False, True)),
    ("conv B2 16x16 1280->1280 split", lambda: conv(2, 16, 1280, 1280, False, True)),
    ("conv B2 8x8 1280->1280 split", lambda: conv(2, 8, 1280, 1280, False, True)),
]
for name, mk in cases:
    run(name, mk())
ops.enable_splitk(False)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
open(os.path.join(ROOT, "gpurun_out", "gemm_timeline%s.txt" % ("_" + sys.argv[1] if len(sys.argv) > 1 else "")), "w").write("\n".join(out) + "\n")
