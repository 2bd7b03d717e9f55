"""Host-side handles over the graph-level C ABI (include/sdod_model.h): weights, UNet, VAE decoder.

Tensors cross the ABI as device pointers; latents are fp32 NHWC [B,H,W,4] inside the library, and these
wrappers accept / return the reference's NCHW tensors (libsdod keeps x_host as [C,H,W]: context.cpp:228).
"""
import ctypes
import struct

import torch

from . import _cabi as C
from . import ops


def _stream():
    return torch.cuda.current_stream().cuda_stream


class Weights:
    """Named fp32 tensors on the device, keyed by ldm state_dict names."""

    def __init__(self, state_dict=None):
        h = ctypes.c_void_p()
        C.check(C.lib().sdod_weights_create(ctypes.byref(h)), "sdod_weights_create")
        self._h = h
        if state_dict is not None:
            self.update(state_dict)

    def update(self, state_dict, prefix=""):
        for k, v in state_dict.items():
            t = v.detach().to("cpu", torch.float32).contiguous()
            shape = (ctypes.c_longlong * t.dim())(*t.shape)
            C.check(C.lib().sdod_weights_set_f32(self._h, (prefix + k).encode(), t.data_ptr(), t.dim(), shape), "sdod_weights_set_f32")

    def load_file(self, path):
        C.check(C.lib().sdod_weights_load_file(self._h, path.encode()), "sdod_weights_load_file")

    def __len__(self):
        return int(C.lib().sdod_weights_count(self._h))

    def __del__(self):
        if getattr(self, "_h", None) and C is not None:
            C.lib().sdod_weights_destroy(self._h)
            self._h = None


def save_weight_file(path, state_dict):
    """Write the flat SDODW001 format read by sdod_weights_load_file (models_dir/unet.sdodw, vae_decoder.sdodw)."""
    with open(path, "wb") as f:
        f.write(b"SDODW001")
        f.write(struct.pack("<I", len(state_dict)))
        for k, v in state_dict.items():
            t = v.detach().to("cpu", torch.float32).contiguous()
            name = k.encode()
            f.write(struct.pack("<I", len(name)) + name)
            f.write(struct.pack("<I", t.dim()) + struct.pack("<%dq" % t.dim(), *t.shape))
            f.write(t.numpy().tobytes())


class UNet:
    def __init__(self, weights=None, seed=0, latent_hw=64, max_batch=2):
        h = ctypes.c_void_p()
        self._weights = weights   # keep alive
        C.check(C.lib().sdod_unet_create(ctypes.byref(h), weights._h if weights is not None else None, seed, latent_hw, max_batch), "sdod_unet_create")
        self._h, self.latent_hw, self.max_batch = h, latent_hw, max_batch

    def time_embed(self, t):
        t = t.to("cuda", torch.float32).contiguous()
        out = torch.empty(t.numel(), 1280, dtype=torch.float32, device="cuda")
        C.check(C.lib().sdod_unet_time_embed(self._h, _stream(), t.data_ptr(), t.numel(), out.data_ptr()), "sdod_unet_time_embed")
        return out

    def set_context(self, context):
        context = context.to("cuda").contiguous()
        assert context.shape[1:] == (77, 768)
        dt = C.F32 if context.dtype == torch.float32 else C.BF16
        C.check(C.lib().sdod_unet_set_context(self._h, _stream(), context.data_ptr(), dt, context.shape[0]), "sdod_unet_set_context")

    def forward_nhwc(self, x_nhwc, emb, use_graph=False):
        B = x_nhwc.shape[0]
        eps = torch.empty_like(x_nhwc)
        C.check(C.lib().sdod_unet_forward(self._h, _stream(), x_nhwc.data_ptr(), emb.data_ptr(), eps.data_ptr(), B, int(use_graph)), "sdod_unet_forward")
        return eps

    def __call__(self, x, emb, context=None, use_graph=False):
        """x [B,4,H,W] fp32 (NCHW, as the oracle), emb [B,1280] fp32 -> eps [B,4,H,W] fp32."""
        if context is not None:
            self.set_context(context)
        x_nhwc = x.to("cuda", torch.float32).permute(0, 2, 3, 1).contiguous()
        emb = emb.to("cuda", torch.float32).contiguous()
        return self.forward_nhwc(x_nhwc, emb, use_graph).permute(0, 3, 1, 2).contiguous()

    def profile(self, B, iters=5):
        """[(ms, op name)] of the forward plan, eager with CUDA events between ops (hot caches)."""
        buf = ctypes.create_string_buffer(1 << 20)
        C.check(C.lib().sdod_unet_profile(self._h, _stream(), B, iters, buf, len(buf)), "sdod_unet_profile")
        rows = [l.split("\t") for l in buf.value.decode().strip().split("\n") if l]
        return [(float(a), b) for a, b in rows]

    def launches_per_forward(self, B):
        return int(C.lib().sdod_unet_launches_per_forward(self._h, B))

    def __del__(self):
        if getattr(self, "_h", None) and C is not None:
            C.lib().sdod_unet_destroy(self._h)
            self._h = None


class VaeDecoder:
    def __init__(self, weights=None, seed=0, latent_hw=64, max_batch=1):
        h = ctypes.c_void_p()
        self._weights = weights
        C.check(C.lib().sdod_vae_create(ctypes.byref(h), weights._h if weights is not None else None, seed, latent_hw, max_batch), "sdod_vae_create")
        self._h, self.latent_hw = h, latent_hw

    def __call__(self, z, use_graph=False):
        """z [B,4,H,W] fp32 NCHW latent -> (uint8 [B,8H,8W,3], fp32 [B,8H,8W,3] in [0,1])."""
        z_nhwc = z.to("cuda", torch.float32).permute(0, 2, 3, 1).contiguous()
        B, H, W, _ = z_nhwc.shape
        u8 = torch.empty(B, 8 * H, 8 * W, 3, dtype=torch.uint8, device="cuda")
        img = torch.empty(B, 8 * H, 8 * W, 3, dtype=torch.float32, device="cuda")
        C.check(C.lib().sdod_vae_decode(self._h, _stream(), z_nhwc.data_ptr(), u8.data_ptr(), img.data_ptr(), B, int(use_graph)), "sdod_vae_decode")
        return u8, img

    def __del__(self):
        if getattr(self, "_h", None) and C is not None:
            C.lib().sdod_vae_destroy(self._h)
            self._h = None
