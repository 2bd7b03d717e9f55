"""SD v1.x checkpoint -> the `models_dir` this library loads.

The reference's `models_dir` holds opaque serialized graphs (`unet.serialized.bin`, `vae_decoder.serialized.bin`, ...;
csrc/libsdod/src/context.cpp:105,186).  Ours holds the named fp32 tensors of the public ldm checkpoint in the flat SDODW001
format (`sdod_weights_load_file`, include/sdod_model.h): `unet.sdodw` with the keys under `model.diffusion_model.` and
`vae_decoder.sdodw` with `first_stage_model.{post_quant_conv,decoder}.*`, and `text_encoder.sdodw` with the CLIP text model under
`cond_stage_model.transformer.` (keys `text_model.*`, as HuggingFace's CLIPTextModel names them).  Packing to bf16 / NHWC happens at plan
build.  The tokenizer vocabulary `ctokenizer.txt` is the reference's own file format: `write_tokenizer_file` restates gen_tokenizer_file.py:27-42.
"""
import os

import torch

from .model import save_weight_file

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."
TEXT_PREFIX = "cond_stage_model.transformer."
UNET_TENSORS = 686      # SD v1.x UNetModel (859,520,964 parameters)
VAE_DECODER_TENSORS = 140
TEXT_TENSORS = 196      # CLIPTextModel of openai/clip-vit-large-patch14 without the position_ids buffer (123,060,480 parameters)


def split_sd_state_dict(state_dict):
    """(unet_sd, vae_decoder_sd) with the prefixes of a full SD checkpoint stripped.  A bare UNet / decoder state_dict (no
    prefixes, as the oracle modules produce) passes through unchanged."""
    sd = state_dict.get("state_dict", state_dict)
    unet, vae = {}, {}
    for k, v in sd.items():
        if not torch.is_tensor(v):
            continue
        if k.startswith(UNET_PREFIX):
            unet[k[len(UNET_PREFIX):]] = v
        elif k.startswith(VAE_PREFIX):
            kk = k[len(VAE_PREFIX):]
            if kk.startswith("decoder.") or kk.startswith("post_quant_conv."):
                vae[kk] = v
        elif k.startswith("decoder.") or k.startswith("post_quant_conv."):
            vae[k] = v
        elif k.split(".")[0] in ("time_embed", "input_blocks", "middle_block", "output_blocks", "out"):
            unet[k] = v
    return unet, vae


def text_encoder_state_dict(state_dict):
    """The CLIP text model's tensors ("text_model.*") of a full SD checkpoint or of a bare CLIPTextModel state_dict; integer buffers
    (position_ids) are dropped."""
    sd = state_dict.get("state_dict", state_dict)
    out = {}
    for k, v in sd.items():
        if not torch.is_tensor(v) or not v.is_floating_point():
            continue
        if k.startswith(TEXT_PREFIX):
            k = k[len(TEXT_PREFIX):]
        if k.startswith("text_model."):
            out[k] = v
    return out


def bytes_to_unicode():
    """CLIP's byte -> printable symbol table (the function the reference's gen_tokenizer_file.py:5-24 carries)."""
    keep = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    table, n = {b: chr(b) for b in keep}, 0
    order = keep[:]
    for b in range(256):
        if b not in table:
            table[b] = chr(256 + n)
            order.append(b)
            n += 1
    return {b: table[b] for b in order}


def write_tokenizer_file(bpe_vocab_path, out_path, n_merges=49152 - 256 - 2):
    """`ctokenizer.txt` from OpenAI's bpe_simple_vocab_16e6.txt(.gz): the 256 byte symbols, the same with "</w>", then the first n_merges
    merges, one per line — the file libsdod::Tokenizer reads (gen_tokenizer_file.py:27-42)."""
    import gzip
    opener = gzip.open if bpe_vocab_path.endswith(".gz") else open
    lines = opener(bpe_vocab_path, "rb").read().decode("utf-8").split("\n")
    merges = [tuple(l.split()) for l in lines[1:n_merges + 1]]
    if any(len(m) != 2 for m in merges):
        raise ValueError("%s: malformed merge line" % bpe_vocab_path)
    vocab = list(bytes_to_unicode().values())
    with open(out_path, "wb") as f:
        for v in vocab + [v + "</w>" for v in vocab]:
            f.write((v + "\n").encode("utf-8"))
        for a, b in merges:
            f.write((a + " " + b + "\n").encode("utf-8"))
    return len(merges)


def load_checkpoint(path):
    """state_dict of a .ckpt/.pt (torch.load, weights only) or .safetensors file."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        return load_file(path)
    return torch.load(path, map_location="cpu", weights_only=True)


def convert(src, models_dir):
    """Write `models_dir/unet.sdodw` and `models_dir/vae_decoder.sdodw` from a checkpoint path or a state_dict; returns the tensor counts."""
    sd = load_checkpoint(src) if isinstance(src, (str, os.PathLike)) else src
    unet, vae = split_sd_state_dict(sd)
    text = text_encoder_state_dict(sd)
    if text:
        os.makedirs(models_dir, exist_ok=True)
        save_weight_file(os.path.join(models_dir, "text_encoder.sdodw"), text)
    if not unet and not vae and text:
        return 0, 0
    if not unet and not vae:
        raise ValueError("no UNet ('%s*') or VAE decoder ('%s{decoder,post_quant_conv}.*') tensors found" % (UNET_PREFIX, VAE_PREFIX))
    os.makedirs(models_dir, exist_ok=True)
    if unet:
        save_weight_file(os.path.join(models_dir, "unet.sdodw"), unet)
    if vae:
        save_weight_file(os.path.join(models_dir, "vae_decoder.sdodw"), vae)
    return len(unet), len(vae)


def read_weight_file(path):
    """Parse an SDODW001 file back into {name: tensor} (tests, inspection)."""
    import struct

    import numpy as np
    out = {}
    with open(path, "rb") as f:
        if f.read(8) != b"SDODW001":
            raise ValueError("%s: not an SDODW001 file" % path)
        (n,) = struct.unpack("<I", f.read(4))
        for _ in range(n):
            (ln,) = struct.unpack("<I", f.read(4))
            name = f.read(ln).decode()
            (nd,) = struct.unpack("<I", f.read(4))
            shape = struct.unpack("<%dq" % nd, f.read(8 * nd))
            cnt = int(np.prod(shape)) if nd else 1
            out[name] = torch.from_numpy(np.frombuffer(f.read(4 * cnt), dtype=np.float32).copy()).reshape(shape)
    return out


if __name__ == "__main__":
    import sys
    if len(sys.argv) != 3:
        raise SystemExit("usage: python -m sdod.checkpoint <sd-v1-checkpoint.{ckpt,safetensors}> <models_dir>")
    print("wrote %d UNet tensors, %d VAE-decoder tensors" % convert(sys.argv[1], sys.argv[2]))
