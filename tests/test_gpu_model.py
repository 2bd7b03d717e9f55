"""GPU parity of the whole UNet step / VAE decode vs the fp32 oracle restatement on identical random-init weights.

Tolerances (BASELINE.json north star): per-step eps relative L2 <= 1e-2; image PSNR >= 35 dB.
The oracle runs on the same GPU in fp32 (TF32 off) — same arithmetic as its CPU run, minutes faster.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from sdod import model as M
    from sdod import ops
from oracle import ldm_oracle as L

DEV = "cuda"


def rel_l2(got, want):
    got, want = got.double().cpu(), want.double().cpu()
    return ((got - want).norm() / want.norm()).item()


def psnr(got, want):
    mse = ((got.double().cpu() - want.double().cpu()) ** 2).mean().item()
    return 10 * math.log10(1.0 / max(mse, 1e-20))


@pytest.fixture(scope="module")
def fp32_exact():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    yield


@pytest.fixture(scope="module")
def unet_pair(fp32_exact):
    oracle = L.make_unet(seed=0).to(DEV)
    weights = M.Weights(oracle.state_dict())
    return oracle, weights


def test_time_embed_matches_oracle(unet_pair):
    oracle, weights = unet_pair
    net = M.UNet(weights, latent_hw=16, max_batch=2)
    t = torch.tensor(ops.dpm_schedule(20)["model_ts"][:20], device=DEV)
    with torch.no_grad():
        want = oracle.embed_time(t)
    got = net.time_embed(t)
    assert rel_l2(got, want) < 1e-2


@pytest.mark.parametrize("hw,B", [(16, 2), (32, 1), (64, 2)])
def test_unet_step_eps_parity(unet_pair, hw, B):
    oracle, weights = unet_pair
    net = M.UNet(weights, latent_hw=hw, max_batch=B)
    assert len(weights) == 686
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, 4, hw, hw, generator=g).to(DEV)
    ctx = torch.randn(B, 77, 768, generator=torch.Generator().manual_seed(2)).to(DEV)
    t = torch.tensor([999.0, 499.5][:B], device=DEV)
    with torch.no_grad():
        emb = oracle.embed_time(t)
        want = oracle(x, emb, ctx)
    got = net(x, emb, ctx)
    assert torch.isfinite(got).all()
    err = rel_l2(got, want)
    print("unet hw=%d B=%d eps rel-L2 = %.3e, launches/forward = %d" % (hw, B, err, net.launches_per_forward(B)))
    assert err < 1e-2
    # CUDA-graph replay gives the same bits as the eager plan
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        a = net(x, emb, None, use_graph=True)
        b = net(x, emb, None, use_graph=True)
        c = net(x, emb, None, use_graph=True)
    s.synchronize()
    assert torch.equal(a, got) and torch.equal(b, got) and torch.equal(c, got)


@pytest.mark.parametrize("B", [8, 32])
def test_unet_step_eps_parity_at_benchmarked_batches(unet_pair, B):
    """The configurations bench.py actually times: UNet batch 32 (16 images per call; selects the CTA-pair and persistent GEMM variants,
    M = 131072) and batch 8 (4 images per call).  Every sample has its own latent, prompt and timestep."""
    oracle, weights = unet_pair
    net = M.UNet(weights, latent_hw=64, max_batch=B)
    x = torch.randn(B, 4, 64, 64, generator=torch.Generator().manual_seed(11)).to(DEV)
    ctx = torch.randn(B, 77, 768, generator=torch.Generator().manual_seed(12)).to(DEV)
    t = torch.linspace(999.0, 49.95, B, device=DEV)
    with torch.no_grad():
        emb = oracle.embed_time(t)
        want = torch.cat([oracle(x[i:i + 4], emb[i:i + 4], ctx[i:i + 4]) for i in range(0, B, 4)], 0)
    got = net(x, emb, ctx)
    assert torch.isfinite(got).all()
    per = [rel_l2(got[i], want[i]) for i in range(B)]
    print("unet hw=64 B=%d eps rel-L2 = %.3e (worst sample %.3e), launches/forward = %d" % (B, rel_l2(got, want), max(per), net.launches_per_forward(B)))
    assert rel_l2(got, want) < 1e-2 and max(per) < 1e-2
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        a = net(x, emb, None, use_graph=True)
        b = net(x, emb, None, use_graph=True)
    s.synchronize()
    assert torch.equal(a, got) and torch.equal(b, got)


@pytest.mark.parametrize("hw,B", [(16, 2), (64, 1), (64, 8), (64, 16)])
def test_vae_decode_parity(fp32_exact, hw, B):
    oracle = L.make_vae(seed=0).to(DEV)
    net = M.VaeDecoder(M.Weights(oracle.state_dict()), latent_hw=hw, max_batch=B)
    z = torch.randn(B, 4, hw, hw, generator=torch.Generator().manual_seed(3)).to(DEV) * 0.18215 * 2
    with torch.no_grad():
        want = torch.cat([oracle(z[i:i + 2]) for i in range(0, B, 2)], 0).permute(0, 2, 3, 1)      # (64,8) = BASELINE config C4; (64,16) = bench.py's batch
    u8, img = net(z)
    p = psnr(img, want)
    print("vae hw=%d B=%d PSNR = %.1f dB" % (hw, B, p))
    assert p >= 35.0
    d = (u8.int() - (want * 255).clamp(0, 255).int()).abs().float()
    print("      u8 |diff|: mean %.3f  p99.9 %.0f  max %.0f" % (d.mean().item(), d.flatten().kthvalue(int(d.numel() * 0.999)).values.item(), d.max().item()))
    assert d.mean() < 1.0 and d.flatten().kthvalue(int(d.numel() * 0.999)).values <= 6
    assert torch.equal(u8, ops.image_to_u8(img))


def test_random_init_models_run():
    net = M.UNet(None, seed=7, latent_hw=16, max_batch=2)
    net.set_context(torch.randn(2, 77, 768))
    eps = net(torch.randn(2, 4, 16, 16), torch.randn(2, 1280))
    assert torch.isfinite(eps).all() and eps.abs().mean() > 0
    vae = M.VaeDecoder(None, seed=7, latent_hw=8, max_batch=1)
    u8, img = vae(torch.randn(1, 4, 8, 8))
    assert u8.shape == (1, 64, 64, 3) and img.min() >= 0 and img.max() <= 1


def test_vae_decode_more_images_than_one_launch_can_address():
    """32 images at 512x512 are 65,536 row tiles in the last level — one more than gridDim.y allows: the decoder splits the batch itself
    (found by the 4-GPU C5 sweep, where a rank decodes 32 images per call)."""
    vae = M.VaeDecoder(None, seed=1, latent_hw=64, max_batch=32)
    z = torch.randn(32, 4, 64, 64, generator=torch.Generator().manual_seed(9)).to(DEV)
    u8, _ = vae(z)
    ref = M.VaeDecoder(None, seed=1, latent_hw=64, max_batch=16)
    a, _ = ref(z[:16])
    b, _ = ref(z[16:])
    assert torch.equal(u8, torch.cat([a, b], 0))
