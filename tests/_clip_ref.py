"""Test helper: HuggingFace CLIPTextModel (openai/clip-vit-large-patch14 text tower geometry) with seeded random-init weights — the
third-party oracle of the text encoder (the reference keeps this model in an opaque serialized graph)."""
import torch


def rel_err(got, want):
    return ((got.float() - want.float()).abs().max() / want.float().abs().max()).item()


def clip_text_model(seed=0):
    from transformers import CLIPTextConfig, CLIPTextModel
    torch.manual_seed(seed)
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12,
                         max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=768)
    m = CLIPTextModel(cfg).eval()
    with torch.no_grad():                        # HF initialises LayerNorm to (1, 0) and biases to 0: make every parameter matter
        for n, p in m.named_parameters():
            if n.endswith("bias"):
                p.normal_(0, 0.05)
            elif "layer_norm" in n and n.endswith("weight"):
                p.normal_(1.0, 0.1)
    return m
