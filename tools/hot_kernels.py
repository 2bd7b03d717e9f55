"""The three hottest level-0 kernels of a batch-8 UNet step, each through the C ABI with CUDA-event timing —
also the target of `ncu --set full -k regex:...` captures (profiles/).
    python tools/hot_kernels.py [geglu|qkv|attn|ffout|all] [B]
"""
import os
import sys

import torch

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import ops  # noqa: E402
from sdod import _cabi as C  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 8
TOK, CH, HEADS, DH = 4096, 320, 8, 40
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, iters=10, warm=3):
    if os.environ.get("HOT_ONCE"):
        iters, warm = 1, 0
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


M = B * TOK
x = torch.randn(M, CH, device="cuda").to(torch.bfloat16)
if which in ("geglu", "all"):
    w = (torch.randn(2 * 4 * CH, CH, device="cuda") / CH ** 0.5)
    b = torch.randn(2 * 4 * CH, device="cuda")
    wp, bp = ops.pack_geglu_weight(w.to(torch.bfloat16), b)
    us = timeit(lambda: torch.ops.sdod.linear(x, wp, bp, None, C.ACT_GEGLU))
    fl = 2.0 * M * w.shape[0] * CH
    print("geglu  M%d N%d K%d: %.1f us  %.0f TFLOP/s" % (M, w.shape[0], CH, us, fl / us * 1e-6))
if which in ("qkv", "all"):
    w = (torch.randn(3 * CH, CH, device="cuda") / CH ** 0.5).to(torch.bfloat16)
    us = timeit(lambda: ops.qkv_project(x, w, HEADS, DH, TOK))
    fl = 2.0 * M * w.shape[0] * CH
    print("qkv    M%d N%d K%d: %.1f us  %.0f TFLOP/s (includes three torch.zeros fills)" % (M, w.shape[0], CH, us, fl / us * 1e-6))
if which in ("ffout", "all"):
    a = torch.randn(M, 4 * CH, device="cuda").to(torch.bfloat16)
    w = (torch.randn(CH, 4 * CH, device="cuda") / (4 * CH) ** 0.5).to(torch.bfloat16)
    bias = torch.randn(CH, device="cuda")
    res = torch.randn(M, CH, device="cuda")
    us = timeit(lambda: torch.ops.sdod.linear(a, w, bias, res, 0, 1.0, True))
    fl = 2.0 * M * CH * 4 * CH
    print("ffout  M%d N%d K%d: %.1f us  %.0f TFLOP/s" % (M, CH, 4 * CH, us, fl / us * 1e-6))
if which in ("attn", "all"):
    dpad, vt_rows, kv_pad = ops.head_geometry(DH, TOK)
    q = torch.randn(B, TOK, HEADS * DH, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, TOK, HEADS * DH, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, TOK, HEADS * DH, device="cuda").to(torch.bfloat16)
    qh, kh, vt = ops.pack_heads(q, HEADS, DH), ops.pack_heads(k, HEADS, DH), ops.pack_heads(v, HEADS, DH, transpose=True)
    us = timeit(lambda: torch.ops.sdod.attention(qh, kh, vt, B, HEADS, DH, TOK, DH ** -0.5))
    fl = 4.0 * B * HEADS * TOK * TOK * DH
    ex = B * HEADS * TOK * TOK
    print("attn   B%d h%d N%d d%d: %.1f us  %.0f TFLOP/s  %.2f Texp/s" % (B, HEADS, TOK, DH, us, fl / us * 1e-6, ex / us * 1e-6))
if which in ("xattn", "all"):
    # cross-attention against the 77 CLIP tokens at the three UNet levels (d = 40 / 80 / 160)
    for tok, heads_dh in ((4096, 40), (1024, 80), (256, 160)):
        nkv = 77
        dpad, vt_rows, kv_pad = ops.head_geometry(heads_dh, nkv)
        q = torch.randn(B, tok, HEADS * heads_dh, device="cuda").to(torch.bfloat16)
        k = torch.randn(B, nkv, HEADS * heads_dh, device="cuda").to(torch.bfloat16)
        v = torch.randn(B, nkv, HEADS * heads_dh, device="cuda").to(torch.bfloat16)
        qh, kh, vt = ops.pack_heads(q, HEADS, heads_dh), ops.pack_heads(k, HEADS, heads_dh), ops.pack_heads(v, HEADS, heads_dh, transpose=True)
        us = timeit(lambda: torch.ops.sdod.attention(qh, kh, vt, B, HEADS, heads_dh, nkv, heads_dh ** -0.5))
        by = qh.numel() * 2 + B * tok * HEADS * heads_dh * 2
        print("xattn  B%d h%d Nq%d Nkv%d d%d: %.1f us  %.0f GB/s (Q read + O write)" % (B, HEADS, tok, nkv, heads_dh, us, by / us * 1e-3))
