"""Per-kernel timings on the GPU box (CUDA events, L2 flushed between iterations). Not the bench contract —
a development aid whose output is copied into profiles/ as evidence."""
import json
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import ops  # noqa: E402
from sdod import _cabi as C  # noqa: E402

DEV = "cuda"
flush = torch.empty(256 << 20, dtype=torch.uint8, device=DEV)


def timeit(fn, iters=10, warm=3):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2]


def bf(*s):
    return torch.randn(*s, device=DEV).to(torch.bfloat16)


out = []


def rec(name, ms, flops=None, bytes_=None, **kw):
    r = {"kernel": name, "ms": round(ms, 4)}
    if flops:
        r["tflops"] = round(flops / ms / 1e9, 1)
    if bytes_:
        r["gbs"] = round(bytes_ / ms / 1e6, 1)
    r.update(kw)
    out.append(r)
    print(json.dumps(r), flush=True)


which = sys.argv[1:] or ["gemm", "conv", "attn", "gn", "misc"]

if "gemm" in which:
    for (M, N, K) in [(8192, 320, 320), (8192, 2560, 320), (8192, 320, 1280), (2048, 640, 640), (2048, 5120, 640), (512, 1280, 1280),
                      (512, 10240, 1280), (128, 1280, 1280), (8192, 8192, 8192), (32768, 320, 320), (32768, 2560, 320)]:
        a, w = bf(M, K), bf(N, K)
        for bn in ([0] if N < 8192 else [256]):
            rec("gemm M%d N%d K%d bn%s" % (M, N, K, bn or "auto"), timeit(lambda: torch.ops.sdod.linear(a, w, None, None, 0, 1.0, False, None, 0, bn)), 2.0 * M * N * K)
        if (M, N, K) in [(8192, 2560, 320), (8192, 8192, 8192), (2048, 640, 640)]:
            for bn in (128, 160, 256):
                rec("gemm M%d N%d K%d bn%d" % (M, N, K, bn), timeit(lambda: torch.ops.sdod.linear(a, w, None, None, 0, 1.0, False, None, 0, bn)), 2.0 * M * N * K)
            rec("torch.matmul M%d N%d K%d" % (M, N, K), timeit(lambda: a @ w.t()), 2.0 * M * N * K)

if "conv" in which:
    for (B, H, Cin, Cout) in [(2, 64, 320, 320), (2, 64, 640, 320), (2, 64, 960, 320), (2, 32, 640, 640), (2, 32, 1280, 640), (2, 16, 1280, 1280),
                              (2, 16, 2560, 1280), (2, 8, 1280, 1280), (2, 8, 2560, 1280), (16, 64, 320, 320), (1, 512, 128, 128), (1, 256, 256, 256)]:
        x, w = bf(B, H, H, Cin), bf(Cout, 9 * Cin)
        rec("conv3x3 B%d %dx%d %d->%d" % (B, H, H, Cin, Cout), timeit(lambda: torch.ops.sdod.conv3x3(x, w)), 2.0 * B * H * H * Cout * 9 * Cin)

if "attn" in which:
    for (B, heads, dh, Nq, Nkv) in [(2, 8, 40, 4096, 4096), (2, 8, 80, 1024, 1024), (2, 8, 160, 256, 256), (2, 8, 40, 4096, 77), (16, 8, 40, 4096, 4096)]:
        q, k, v = bf(B, Nq, heads * dh), bf(B, Nkv, heads * dh), bf(B, Nkv, heads * dh)
        qh, kh, vt = ops.pack_heads(q, heads, dh), ops.pack_heads(k, heads, dh), ops.pack_heads(v, heads, dh, True)
        rec("attn B%d h%d d%d %dx%d" % (B, heads, dh, Nq, Nkv), timeit(lambda: torch.ops.sdod.attention(qh, kh, vt, B, heads, dh, Nkv, dh ** -0.5)),
            4.0 * B * heads * Nq * Nkv * dh)

if "gn" in which:
    x = torch.randn(2, 320, 64, 64, device=DEV)
    w, b = torch.randn(320, device=DEV), torch.randn(320, device=DEV)
    rec("gn+silu NCHW fp32 [2,320,64,64] (config 1)", timeit(lambda: torch.ops.sdod.group_norm(x, 32, w, b, 1e-5, True, None), 20), None, 2 * x.numel() * 4)
    rec("torch GN+silu same (library)", timeit(lambda: torch.nn.functional.silu(torch.nn.functional.group_norm(x, 32, w, b, 1e-5)), 20), None, 2 * x.numel() * 4)
    for shape in [(2, 4096, 320), (2, 4096, 960), (2, 1024, 1280), (2, 64, 2560), (16, 4096, 320), (1, 262144, 128), (8, 262144, 128)]:
        xb = bf(*shape)
        wc, bc = torch.randn(shape[2], device=DEV), torch.randn(shape[2], device=DEV)
        rec("gn+silu NHWC bf16 %s" % (shape,), timeit(lambda: ops.group_norm_nhwc(xb, 32, wc, bc, 1e-5, True)), None, 2 * xb.numel() * 2)

if "misc" in which:
    x = torch.randn(2 * 16384, device=DEV)
    yp = torch.zeros_like(x)
    ec, eu = torch.randn_like(x), torch.randn_like(x)
    rec("cfg_dpm_step n=32768", timeit(lambda: torch.ops.sdod.cfg_dpm_step(x, yp, ec, eu, 7.5, 0.9, 0.1, 0.99, 0.1, -0.2, 2), 20), None, 24 * x.numel())
    xl = bf(8192, 320)
    wl, bl = torch.randn(320, device=DEV), torch.randn(320, device=DEV)
    rec("layer_norm [8192,320]", timeit(lambda: torch.ops.sdod.layer_norm(xl, wl, bl, 1e-5)), None, 2 * xl.numel() * 2)

os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "microbench.json"), "w"), indent=1)
