"""GPU: the prompt path — CLIP text encoder on the library's kernels vs HuggingFace's CLIPTextModel (fp32, identical
random-init weights; the north star's per-op bf16 tolerance 2e-2), its causal attention and QuickGELU pieces, and the libsdod entry points."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from sdod import _cabi as C
    from sdod import libsdod as A
    from sdod import model as M
    from sdod import ops
    from sdod import text as T

DEV = "cuda"
TOL_BF16 = 2e-2


from _clip_ref import clip_text_model, rel_err  # noqa: E402


def prompt_tokens(B, seed):
    g = torch.Generator().manual_seed(seed)
    toks = torch.full((B, 77), 49407, dtype=torch.int64)
    toks[:, 0] = 49406
    for b in range(B):
        n = int(torch.randint(1, 75, (1,), generator=g))
        toks[b, 1:1 + n] = torch.randint(0, 49406, (n,), generator=g)
    return toks


@pytest.mark.parametrize("B", [1, 3])
def test_text_encoder_matches_clip_text_model(B):
    from sdod import checkpoint as K
    m = clip_text_model(0)
    toks = prompt_tokens(B, 7 + B)
    with torch.no_grad():
        want = m(input_ids=toks).last_hidden_state
    enc = T.TextEncoder(M.Weights(K.text_encoder_state_dict(m.state_dict())), max_batch=4)
    got = enc(toks).cpu()
    got_bf16 = enc(toks, dtype=torch.bfloat16).cpu()
    e = rel_err(got, want)
    l2 = ((got - want).norm() / want.norm()).item()
    print("CLIP text encoder B=%d: max rel err %.3e, rel-L2 %.3e" % (B, e, l2))
    assert got.shape == (B, 77, 768) and torch.isfinite(got).all()
    assert e < TOL_BF16 and l2 < 1e-2
    assert torch.equal(got_bf16.float(), got)                       # the fp32 output is the bf16 result widened
    assert torch.equal(enc(toks).cpu(), got)                        # deterministic, plan reuse


@pytest.mark.parametrize("B,heads,dh,N", [(2, 12, 64, 77), (1, 4, 64, 300), (1, 8, 40, 256), (1, 2, 64, 129)])
def test_causal_attention(B, heads, dh, N):
    torch.manual_seed(N + dh)
    Cc = heads * dh
    q, k, v = (torch.randn(B, N, Cc).bfloat16().to(DEV) * s for s in (1.5, 1.5, 1.0))
    qh, kh, vh = (t.float().reshape(B, N, heads, dh).permute(0, 2, 1, 3) for t in (q, k, v))
    s = qh @ kh.transpose(-1, -2) * dh ** -0.5
    s = s.masked_fill(torch.triu(torch.ones(N, N, dtype=torch.bool, device=DEV), 1), float("-inf"))
    want = (torch.softmax(s, -1) @ vh).permute(0, 2, 1, 3).reshape(B, N, Cc)
    got = ops.attention_causal(ops.pack_heads(q, heads, dh), ops.pack_heads(k, heads, dh), ops.pack_heads(v, heads, dh, True), B, heads, dh, N, dh ** -0.5)
    assert torch.isfinite(got.float()).all() and rel_err(got, want) < TOL_BF16
    # the unmasked kernel on the same operands differs (the mask is really applied)
    plain = torch.ops.sdod.attention(ops.pack_heads(q, heads, dh), ops.pack_heads(k, heads, dh), ops.pack_heads(v, heads, dh, True), B, heads, dh, N, dh ** -0.5)
    assert rel_err(plain, want) > 0.1


def test_quick_gelu_epilogue():
    torch.manual_seed(3)
    a, w, b = torch.randn(231, 768).bfloat16().to(DEV), (torch.randn(3072, 768) / 768 ** 0.5).bfloat16().to(DEV), torch.randn(3072).to(DEV)
    y = a.float() @ w.float().t() + b
    want = y * torch.sigmoid(1.702 * y)
    got = torch.ops.sdod.linear(a, w, b, None, C.ACT_QUICK_GELU)
    assert rel_err(got, want) < TOL_BF16
    got32 = torch.ops.sdod.linear(a, w, b, None, C.ACT_QUICK_GELU, 1.0, True)
    assert rel_err(got32, want) < 1e-2


def test_random_init_context_encodes_prompts():
    """random-init contexts carry a byte-level tokenizer and a random-init text encoder: the prompt decides the conditioning
    (round 1 hashed the prompt into noise)."""
    with A.Context("random-init:3", latent_spatial=16, steps=2, max_images=1, device=0) as ctx:
        e1, t1 = ctx.encode_prompt("a cat", return_tokens=True)
        e2 = ctx.encode_prompt("a dog")
        e1b = ctx.encode_prompt("  A   CAT ")
        assert list(t1[:7]) == [512, 64 + 256, 66, 64, 83 + 256, 513, 513]          # byte-level ids: a</w> c a t</w>
        assert np.isfinite(e1).all() and e1.std() > 0.1
        assert np.array_equal(e1, e1b)                                # sanitised to the same prompt
        assert np.abs(e1 - e2).max() > 1e-2
        # the causal mask: the embedding of a prefix position does not depend on later tokens
        assert np.array_equal(e1[:2], e2[:2]) and not np.array_equal(e1[3:5], e2[3:5])
        # the same weights through the graph-level ABI
        enc = T.TextEncoder(None, seed=3 + 2, max_batch=1)
        assert np.array_equal(enc(torch.from_numpy(t1.astype(np.int64))[None]).cpu().numpy()[0], e1)
        ctx.set_seed(11)
        img = ctx.generate_image("a cat", 7.5)
        ctx.set_seed(11)
        imgs = ctx.generate(e1[None], ctx.encode_prompt("")[None], None, 7.5)
        assert np.array_equal(img, imgs[0])                           # generate_image == encode(prompt) / cached encode("") -> generate
        with pytest.raises(A.LibsdodError) as ei:
            ctx.generate_images(["a cat", "a dog"], 7.5)              # max_images = 1
        assert ei.value.status == A.INVALID_ARGUMENT
        with pytest.raises(A.LibsdodError) as ei:
            ctx.encode_prompt(b"bad \xff utf8")
        assert ei.value.status == A.INVALID_ARGUMENT and "Invalid UTF-8" in str(ei.value)


def test_generate_images_batches_the_prompts():
    """libsdod_b200_generate_images: n prompts -> one text-encoder batch -> one generate call; equals the per-prompt path image by image."""
    prompts = ["a cat", "a photograph of an astronaut riding a horse", ""]
    with A.Context("random-init:4", latent_spatial=16, steps=3, max_images=3, device=0) as ctx:
        embs = np.stack([ctx.encode_prompt(p) for p in prompts])
        ctx.set_seed(21)
        imgs = ctx.generate_images(prompts, 7.5)
        ctx.set_seed(21)
        want = ctx.generate(embs, np.stack([ctx.encode_prompt("")] * 3), None, 7.5)
        # (a batch of 3 may pick other split-K factors than three batches of 1: equal up to fp32 summation order)
        mse = ((imgs.astype(np.float64) - want.astype(np.float64)) ** 2).mean()
        assert imgs.shape == (3, 128, 128, 3) and 10 * np.log10(255.0 ** 2 / max(mse, 1e-12)) >= 40.0
        assert np.abs(imgs[0].astype(np.int32) - imgs[1].astype(np.int32)).mean() > 1.0
