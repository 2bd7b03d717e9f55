"""SD v1.x checkpoint -> the `models_dir` this library loads.

The reference's `models_dir` holds opaque serialized graphs (`unet.serialized.bin`, `vae_decoder.serialized.bin`, ...;
csrc/libsdod/src/context.cpp:105,186).  Ours holds the named fp32 tensors of the public ldm checkpoint in the flat SDODW001
format (`sdod_weights_load_file`, include/sdod_model.h): `unet.sdodw` with the keys under `model.diffusion_model.` and
`vae_decoder.sdodw` with `first_stage_model.{post_quant_conv,decoder}.*`.  Packing to bf16 / NHWC happens at plan build.
"""
import os

import torch

from .model import save_weight_file

UNET_PREFIX = "model.diffusion_model."
VAE_PREFIX = "first_stage_model."
UNET_TENSORS = 686      # SD v1.x UNetModel (859,520,964 parameters)
VAE_DECODER_TENSORS = 140


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
