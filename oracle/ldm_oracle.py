"""ORACLE — TEST INFRASTRUCTURE ONLY.  fp32 PyTorch restatement of the SD-v1 UNet and VAE decoder.

The reference does not vendor this arithmetic: it lives in an external, un-pinned
"modified Stable-Diffusion repo" (reference README.md:23) and reaches libsdod only
as opaque serialized graphs (csrc/libsdod/src/context.cpp:105,352,366,387).  This file
restates the PUBLIC CompVis v1 architecture, cross-checked
against the in-tree evidence the reference does hold:
  * layer names (analyze_results.py:25-87) — module names below match them, so a
    real SD-v1 state_dict loads by key;
  * mode_dim=320 / temb_dim=1280 (context.cpp:258-259), 77 tokens (tokenizer.h:24),
    latent 4x64x64, x8 upscale (api/libsdod.h:34-36);
  * exact parameter counts 859,520,964 (UNet) and 49,490,199 (decoder+post_quant_conv).
PARITY UNPINNED by the reference (it holds no test at this boundary); "parity" for
a7/a8 means: our CUDA path vs this restatement on identical random-init weights.

GroupNorm goes through the reference operator surface (sdod.EfficientGN semantics =
torch.nn.functional.group_norm, sdod/efficient_gn.py:9-12,61-70).
"""
import math

import torch
import torch.nn as nn
import torch.nn.functional as F


def group_norm_f64(x, num_groups, weight=None, bias=None, eps=1e-5):
    """float64 restatement of the GroupNorm contract (efficient_gn.py:12 -> ATen group_norm):
    biased variance over (C/G, *spatial) per (n, g); y = (x-mu)*rsqrt(var+eps)*w_c + b_c."""
    n, c = x.shape[:2]
    xd = x.detach().to(torch.float64).reshape(n, num_groups, -1)
    mu = xd.mean(dim=2, keepdim=True)
    var = ((xd - mu) ** 2).mean(dim=2, keepdim=True)
    y = ((xd - mu) / torch.sqrt(var + eps)).reshape(x.shape)
    shp = (1, c) + (1,) * (x.dim() - 2)
    if weight is not None:
        y = y * weight.detach().to(torch.float64).reshape(shp)
    if bias is not None:
        y = y + bias.detach().to(torch.float64).reshape(shp)
    return y


def timestep_embedding(t, dim=320, max_period=10000.0):
    """cos-first sinusoid, as context.cpp:257-275 (and public ldm util.timestep_embedding)."""
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


class GEGLU(nn.Module):
    def __init__(self, dim_in, dim_out):
        super().__init__()
        self.proj = nn.Linear(dim_in, dim_out * 2)

    def forward(self, x):
        x, gate = self.proj(x).chunk(2, dim=-1)
        return x * F.gelu(gate)


class FeedForward(nn.Module):
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(GEGLU(dim, dim * mult), nn.Dropout(0.0), nn.Linear(dim * mult, dim))

    def forward(self, x):
        return self.net(x)


class CrossAttention(nn.Module):
    def __init__(self, query_dim, context_dim=None, heads=8, dim_head=64):
        super().__init__()
        inner = heads * dim_head
        context_dim = context_dim or query_dim
        self.heads, self.scale = heads, dim_head ** -0.5
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_k = nn.Linear(context_dim, inner, bias=False)
        self.to_v = nn.Linear(context_dim, inner, bias=False)
        self.to_out = nn.Sequential(nn.Linear(inner, query_dim), nn.Dropout(0.0))

    def forward(self, x, context=None):
        context = x if context is None else context
        b, n, _ = x.shape
        h = self.heads
        q, k, v = self.to_q(x), self.to_k(context), self.to_v(context)
        q, k, v = (t.reshape(b, t.shape[1], h, -1).permute(0, 2, 1, 3) for t in (q, k, v))
        sim = torch.einsum("bhid,bhjd->bhij", q, k) * self.scale
        out = torch.einsum("bhij,bhjd->bhid", sim.softmax(dim=-1), v)
        return self.to_out(out.permute(0, 2, 1, 3).reshape(b, n, -1))


class BasicTransformerBlock(nn.Module):
    def __init__(self, dim, n_heads, d_head, context_dim):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.norm1, self.norm2, self.norm3 = nn.LayerNorm(dim), nn.LayerNorm(dim), nn.LayerNorm(dim)

    def forward(self, x, context):
        x = self.attn1(self.norm1(x)) + x
        x = self.attn2(self.norm2(x), context) + x
        return self.ff(self.norm3(x)) + x


class SpatialTransformer(nn.Module):
    def __init__(self, ch, n_heads, d_head, context_dim=768):
        super().__init__()
        self.norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.proj_in = nn.Conv2d(ch, n_heads * d_head, 1)
        self.transformer_blocks = nn.ModuleList([BasicTransformerBlock(n_heads * d_head, n_heads, d_head, context_dim)])
        self.proj_out = nn.Conv2d(n_heads * d_head, ch, 1)

    def forward(self, x, context):
        b, c, h, w = x.shape
        x_in = x
        x = self.proj_in(self.norm(x)).reshape(b, -1, h * w).permute(0, 2, 1)
        for blk in self.transformer_blocks:
            x = blk(x, context)
        x = x.permute(0, 2, 1).reshape(b, -1, h, w)
        return self.proj_out(x) + x_in


class ResBlock(nn.Module):
    def __init__(self, ch, emb_ch, out_ch):
        super().__init__()
        self.in_layers = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(), nn.Conv2d(ch, out_ch, 3, padding=1))
        self.emb_layers = nn.Sequential(nn.SiLU(), nn.Linear(emb_ch, out_ch))
        self.out_layers = nn.Sequential(nn.GroupNorm(32, out_ch), nn.SiLU(), nn.Dropout(0.0), nn.Conv2d(out_ch, out_ch, 3, padding=1))
        self.skip_connection = nn.Identity() if ch == out_ch else nn.Conv2d(ch, out_ch, 1)

    def forward(self, x, emb):
        h = self.in_layers(x)
        h = h + self.emb_layers(emb)[:, :, None, None]
        h = self.out_layers(h)
        return self.skip_connection(x) + h


class Downsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.op = nn.Conv2d(ch, ch, 3, stride=2, padding=1)

    def forward(self, x):
        return self.op(x)


class Upsample(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.conv = nn.Conv2d(ch, ch, 3, padding=1)

    def forward(self, x):
        return self.conv(F.interpolate(x, scale_factor=2.0, mode="nearest"))


class _Seq(nn.Sequential):
    def forward(self, x, emb, context):
        for m in self:
            if isinstance(m, ResBlock):
                x = m(x, emb)
            elif isinstance(m, SpatialTransformer):
                x = m(x, context)
            else:
                x = m(x)
        return x


class UNetModel(nn.Module):
    """SD-v1 UNet: model_channels 320, mult (1,2,4,4), 2 res blocks, attention at ds 1/2/4, 8 heads."""

    def __init__(self, in_ch=4, out_ch=4, mc=320, mult=(1, 2, 4, 4), num_res=2, attn_ds=(1, 2, 4), heads=8, context_dim=768):
        super().__init__()
        ted = mc * 4
        self.model_channels = mc
        self.time_embed = nn.Sequential(nn.Linear(mc, ted), nn.SiLU(), nn.Linear(ted, ted))
        self.input_blocks = nn.ModuleList([_Seq(nn.Conv2d(in_ch, mc, 3, padding=1))])
        chans, ch, ds = [mc], mc, 1
        for level, m in enumerate(mult):
            for _ in range(num_res):
                layers = [ResBlock(ch, ted, m * mc)]
                ch = m * mc
                if ds in attn_ds:
                    layers.append(SpatialTransformer(ch, heads, ch // heads, context_dim))
                self.input_blocks.append(_Seq(*layers))
                chans.append(ch)
            if level != len(mult) - 1:
                self.input_blocks.append(_Seq(Downsample(ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = _Seq(ResBlock(ch, ted, ch), SpatialTransformer(ch, heads, ch // heads, context_dim), ResBlock(ch, ted, ch))
        self.output_blocks = nn.ModuleList()
        for level, m in list(enumerate(mult))[::-1]:
            for i in range(num_res + 1):
                ich = chans.pop()
                layers = [ResBlock(ch + ich, ted, mc * m)]
                ch = mc * m
                if ds in attn_ds:
                    layers.append(SpatialTransformer(ch, heads, ch // heads, context_dim))
                if level and i == num_res:
                    layers.append(Upsample(ch))
                    ds //= 2
                self.output_blocks.append(_Seq(*layers))
        self.out = nn.Sequential(nn.GroupNorm(32, ch), nn.SiLU(), nn.Conv2d(ch, out_ch, 3, padding=1))

    def embed_time(self, model_t):
        """model_t: [B] float timesteps (solver model_ts) -> [B,1280]; == the reference's `temb` graph."""
        return self.time_embed(timestep_embedding(model_t, self.model_channels))

    def forward(self, x, emb, context):
        """x [B,4,H,W], emb [B,1280] (already through time_embed, as graph input 1: context.cpp:214-218), context [B,77,768]."""
        hs, h = [], x
        for m in self.input_blocks:
            h = m(h, emb, context)
            hs.append(h)
        h = self.middle_block(h, emb, context)
        for m in self.output_blocks:
            h = m(torch.cat([h, hs.pop()], dim=1), emb, context)
        return self.out(h)


# ------------------------------------------------------------------ VAE decoder
class VAEResnetBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.norm1 = nn.GroupNorm(32, cin, eps=1e-6)
        self.conv1 = nn.Conv2d(cin, cout, 3, padding=1)
        self.norm2 = nn.GroupNorm(32, cout, eps=1e-6)
        self.conv2 = nn.Conv2d(cout, cout, 3, padding=1)
        if cin != cout:
            self.nin_shortcut = nn.Conv2d(cin, cout, 1)

    def forward(self, x):
        h = self.conv1(F.silu(self.norm1(x)))
        h = self.conv2(F.silu(self.norm2(h)))
        if hasattr(self, "nin_shortcut"):
            x = self.nin_shortcut(x)
        return x + h


class VAEAttnBlock(nn.Module):
    def __init__(self, ch):
        super().__init__()
        self.norm = nn.GroupNorm(32, ch, eps=1e-6)
        self.q, self.k, self.v = nn.Conv2d(ch, ch, 1), nn.Conv2d(ch, ch, 1), nn.Conv2d(ch, ch, 1)
        self.proj_out = nn.Conv2d(ch, ch, 1)

    def forward(self, x):
        b, c, h, w = x.shape
        hn = self.norm(x)
        q = self.q(hn).reshape(b, c, h * w).permute(0, 2, 1)
        k = self.k(hn).reshape(b, c, h * w)
        v = self.v(hn).reshape(b, c, h * w)
        w_ = torch.bmm(q, k) * (c ** -0.5)
        w_ = F.softmax(w_, dim=2)
        out = torch.bmm(v, w_.permute(0, 2, 1)).reshape(b, c, h, w)
        return x + self.proj_out(out)


class _Level(nn.Module):
    pass


class VAEDecoder(nn.Module):
    """ldm Decoder (ch 128, mult (1,2,4,4), 2 res blocks (+1 in decoder), z 4) + post_quant_conv.
    forward() returns the [0,1] image the host loop quantises (context.cpp:386-395)."""

    def __init__(self, ch=128, mult=(1, 2, 4, 4), num_res=2, z_ch=4, out_ch=3, scale=0.18215):
        super().__init__()
        self.scale = scale
        self.post_quant_conv = nn.Conv2d(z_ch, z_ch, 1)
        dec = nn.Module()
        bi = ch * mult[-1]
        dec.conv_in = nn.Conv2d(z_ch, bi, 3, padding=1)
        dec.mid = nn.Module()
        dec.mid.block_1 = VAEResnetBlock(bi, bi)
        dec.mid.attn_1 = VAEAttnBlock(bi)
        dec.mid.block_2 = VAEResnetBlock(bi, bi)
        dec.up = nn.ModuleList()
        levels = []
        for i in reversed(range(len(mult))):
            lvl = _Level()
            lvl.block = nn.ModuleList()
            bo = ch * mult[i]
            for _ in range(num_res + 1):
                lvl.block.append(VAEResnetBlock(bi, bo))
                bi = bo
            if i != 0:
                lvl.upsample = Upsample(bi)
            levels.insert(0, lvl)
        for lvl in levels:
            dec.up.append(lvl)
        dec.norm_out = nn.GroupNorm(32, bi, eps=1e-6)
        dec.conv_out = nn.Conv2d(bi, out_ch, 3, padding=1)
        self.decoder = dec

    def forward(self, z):
        d = self.decoder
        h = d.conv_in(self.post_quant_conv(z / self.scale))
        h = d.mid.block_2(d.mid.attn_1(d.mid.block_1(h)))
        for i in reversed(range(len(d.up))):
            for blk in d.up[i].block:
                h = blk(h)
            if hasattr(d.up[i], "upsample"):
                h = d.up[i].upsample(h)
        h = d.conv_out(F.silu(d.norm_out(h)))
        return torch.clamp((h + 1.0) / 2.0, 0.0, 1.0)


def count_params(m):
    return sum(p.numel() for p in m.parameters())


def randomize_zero_layers(model, std=0.02):
    """Public SD zero-inits proj_out / final conv; SURVEY §8(d): re-init non-zero so errors are visible.
    (PyTorch default inits are already non-zero here; kept for checkpoints that carry zeros.)"""
    with torch.no_grad():
        for p in model.parameters():
            if p.abs().max() == 0:
                p.normal_(0, std)
    return model


def make_unet(seed=0):
    g = torch.random.get_rng_state()
    torch.manual_seed(seed)
    m = UNetModel().eval()
    torch.random.set_rng_state(g)
    return m


def make_vae(seed=0):
    g = torch.random.get_rng_state()
    torch.manual_seed(seed + 1000)
    m = VAEDecoder().eval()
    torch.random.set_rng_state(g)
    return m
