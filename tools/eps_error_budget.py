"""Where the UNet step's eps error comes from: the fp32 oracle against precision-EMULATED copies of itself on the host.
Each emulation reproduces one of the product path's precision decisions inside the oracle's own PyTorch graph (fp32 accumulation everywhere, as the
tensor cores do), so the contributions can be separated without the GPU:
  W   : every conv / linear weight rounded to bf16 (the packed weights)
  A   : every conv / linear INPUT rounded to bf16 (GEMM A operands: GroupNorm / LayerNorm / GEGLU / attention outputs), weights fp32
  WA  : both (the tensor-core operand precision of the product)
  WAP : WA + attention operands and probabilities in bf16 (Q, K, V, P), softmax in fp32
Output: eps relative L2 vs the fp32 oracle for each mode, and for WAP the error of every top-level block's output (input_blocks.k / middle_block /
output_blocks.k / out), i.e. how the error grows through the network.  Same inputs as the parity tests (seeds 0 / 1 / 2, batch 2 = cond + uncond).
   python tools/eps_error_budget.py  -> profiles/r02_eps_error_budget.txt   (CPU only, a few minutes)"""
import contextlib
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, ROOT)
from oracle import ldm_oracle as L  # noqa: E402
from oracle import sampler as S  # noqa: E402


def bf(t):
    return t.to(torch.bfloat16).to(torch.float32)


@contextlib.contextmanager
def emulate(weights, acts, attn):
    conv2d, linear, einsum, softmax = F.conv2d, F.linear, torch.einsum, torch.softmax

    def c2(x, w, b=None, *a, **k):
        return conv2d(bf(x) if acts else x, bf(w) if weights else w, b, *a, **k)

    def lin(x, w, b=None):
        return linear(bf(x) if acts else x, bf(w) if weights else w, b)

    def es(eq, *ops):                      # CrossAttention's QK^T and PV
        return einsum(eq, *[bf(o) for o in ops]) if attn else einsum(eq, *ops)

    F.conv2d, F.linear, torch.einsum = c2, lin, es
    try:
        yield
    finally:
        F.conv2d, F.linear, torch.einsum, torch.softmax = conv2d, linear, einsum, softmax


def run(unet, x, emb, ctx, hooks=None):
    outs = {}
    hs = []
    if hooks is not None:
        for name, m in unet.named_modules():
            if name.count(".") == 1 and name.split(".")[0] in ("input_blocks", "output_blocks") or name in ("middle_block", "out"):
                hs.append(m.register_forward_hook(lambda mod, i, o, name=name: outs.__setitem__(name, o.detach().clone())))
    with torch.no_grad():
        y = unet(x, emb, ctx)
    for h in hs:
        h.remove()
    return y, outs


def rel(a, b):
    return ((a.double() - b.double()).norm() / b.double().norm()).item()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    unet = L.make_unet(0)
    x = torch.randn(1, 4, 64, 64, generator=torch.Generator().manual_seed(1)).repeat(2, 1, 1, 1)
    g = torch.Generator().manual_seed(2)
    ctx = torch.cat([torch.randn(1, 77, 768, generator=g), torch.randn(1, 77, 768, generator=g)], 0)
    solver = S.OracleSolver()
    solver.prepare(20)
    with torch.no_grad():
        emb = unet.embed_time(torch.tensor(solver.table("model_ts")[:1])).repeat(2, 1)
    ref, ref_blocks = run(unet, x, emb, ctx, hooks=True)
    lines = ["eps error budget of one batch-2 UNet step (fp32 oracle vs precision-emulated oracle, host PyTorch; tools/eps_error_budget.py)",
             "product path on B200 (tests/test_gpu_model.py, bench.py parity leg): eps rel-L2 7.5e-3, tolerance 1e-2", ""]
    for tag, (w, a, p) in (("W   weights bf16", (1, 0, 0)), ("A   GEMM/conv inputs bf16", (0, 1, 0)), ("WA  both", (1, 1, 0)), ("WAP both + attention Q/K/V/P bf16", (1, 1, 1))):
        with emulate(w, a, p):
            y, blocks = run(unet, x, emb, ctx, hooks=(True if tag.startswith("WAP") else None))
        lines.append("%-40s eps rel-L2 %.3e" % (tag, rel(y, ref)))
        if blocks:
            lines.append("")
            lines.append("error of each block's output under WAP (relative L2 vs the fp32 oracle's same tensor):")
            for name in ref_blocks:
                lines.append("  %-20s %.3e" % (name, rel(blocks[name], ref_blocks[name])))
    txt = "\n".join(lines)
    print(txt)
    open(os.path.join(ROOT, "profiles", "r02_eps_error_budget.txt"), "w").write(txt + "\n")


if __name__ == "__main__":
    main()
