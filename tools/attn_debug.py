import os, sys, torch
ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod import ops
def ref(q, k, v, heads):
    B, N, Cc = q.shape; dh = Cc // heads
    qh, kh, vh = (t.float().reshape(B, t.shape[1], heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    p = torch.softmax(qh @ kh.transpose(-1, -2) * dh ** -0.5, dim=-1)
    return (p @ vh).permute(0, 2, 1, 3).reshape(B, N, Cc)
for (B, heads, dh, Nq, Nkv) in [(1, 8, 40, 128, 64), (1, 8, 40, 128, 77), (1, 8, 40, 128, 128), (1, 8, 40, 128, 100), (1, 8, 40, 128, 13), (2, 8, 40, 4096, 77), (1, 4, 64, 200, 333)]:
    torch.manual_seed(0)
    Cc = heads * dh
    q, k, v = (torch.randn(B, n, Cc, device="cuda").to(torch.bfloat16) for n in (Nq, Nkv, Nkv))
    want = ref(q, k, v, heads)
    got = torch.ops.sdod.attention(ops.pack_heads(q, heads, dh), ops.pack_heads(k, heads, dh), ops.pack_heads(v, heads, dh, True), B, heads, dh, Nkv, dh ** -0.5).float()
    d = (got - want).abs()
    print((B, heads, dh, Nq, Nkv), "max err %.4f  rel %.4f  nan %d  worst row %d col %d" % (d.max().item(), (d.max() / want.abs().max()).item(), torch.isnan(got).sum().item(), d.view(-1, Cc).max(1).values.argmax().item(), d.view(-1, Cc).max(0).values.argmax().item()))
print("---- test config (q,k scaled 1.5)")
for (B, heads, dh, Nq, Nkv) in [(2, 8, 40, 4096, 77), (1, 8, 80, 1024, 77), (2, 8, 160, 64, 77), (1, 8, 40, 256, 256), (2, 8, 40, 4096, 4096)]:
    torch.manual_seed(Nq + Nkv + dh)
    Cc = heads * dh
    q, k, v = (torch.randn(B, n, Cc).mul(s).to(torch.bfloat16).cuda() for n, s in ((Nq, 1.5), (Nkv, 1.5), (Nkv, 1.0)))
    want = ref(q, k, v, heads)
    got = torch.ops.sdod.attention(ops.pack_heads(q, heads, dh), ops.pack_heads(k, heads, dh), ops.pack_heads(v, heads, dh, True), B, heads, dh, Nkv, dh ** -0.5).float()
    d = (got - want).abs()
    r = d.view(-1, Cc).max(1).values.argmax().item()
    print((B, heads, dh, Nq, Nkv), "max err %.4f  rel %.4f  nonfinite %d  worst row %d col %d  |want|max %.3f" % (d.max().item(), (d.max() / want.abs().max()).item(), (~torch.isfinite(got)).sum().item(), r, d.view(-1, Cc).max(0).values.argmax().item(), want.abs().max().item()))
