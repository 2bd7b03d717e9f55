"""ctypes binding of the libsdod C API (include/libsdod.h) — the binding a maintainer of the reference would add to
call libsdod_b200.so exactly as csrc/libsdod/test/simple_app.cpp calls the reference library."""
import ctypes

import numpy as np

from . import _cabi as C

NO_ERROR, INVALID_CONTEXT, INVALID_ARGUMENT, FAILED_ALLOCATION, RUNTIME_ERROR, INTERNAL_ERROR = range(6)
LOG_NOTHING, LOG_ERROR, LOG_INFO, LOG_DEBUG, LOG_ABUSIVE = range(5)

_u, _i, _vp, _f = ctypes.c_uint, ctypes.c_int, ctypes.c_void_p, ctypes.c_float
_ucp = ctypes.POINTER(ctypes.c_ubyte)
_SIGS = {
    "libsdod_setup": (_i, [ctypes.POINTER(_vp), ctypes.c_char_p, _u, _u, _u, _u, _u, _i]),
    "libsdod_set_steps": (_i, [_vp, _u]),
    "libsdod_set_log_level": (_i, [_vp, _u]),
    "libsdod_ref_context": (_i, [_vp]),
    "libsdod_release": (_i, [_vp]),
    "libsdod_generate_image": (_i, [_vp, ctypes.c_char_p, _f, ctypes.POINTER(_ucp), ctypes.POINTER(_u)]),
    "libsdod_get_error_description": (ctypes.c_char_p, [_i]),
    "libsdod_get_last_error_extra_info": (ctypes.c_char_p, [_i, _vp]),
    "libsdod_b200_set_seed": (_i, [_vp, ctypes.c_ulonglong]),
    "libsdod_b200_set_sampler": (_i, [_vp, ctypes.c_int]),
    "libsdod_b200_generate": (_i, [_vp, _u, _vp, _vp, _vp, _f, _vp, _vp]),
    "libsdod_b200_generate_device": (_i, [_vp, _u, _vp, _vp, _vp, _f, _vp]),
    "libsdod_b200_setup": (_i, [ctypes.POINTER(_vp), ctypes.c_char_p, _u, _u, _u, _u, _i]),
    "libsdod_b200_last_timings": (_i, [_vp, ctypes.POINTER(_f * 4)]),
    "libsdod_b200_encode_prompt": (_i, [_vp, ctypes.c_char_p, _vp, _vp]),
    "libsdod_b200_generate_images": (_i, [_vp, _u, ctypes.POINTER(ctypes.c_char_p), _f, _vp]),
    "libsdod_b200_pair_export": (_i, [_vp, _vp]),
    "libsdod_b200_pair_connect": (_i, [_vp, _vp, _i]),
    "libsdod_b200_generate_pair": (_i, [_vp, _u, _vp, _vp, _f, _vp, ctypes.POINTER(_u), ctypes.POINTER(_u), _vp]),
}
_bound = False


def api():
    global _bound
    lib = C.lib()
    if not _bound:
        for name, (res, args) in _SIGS.items():
            fn = getattr(lib, name)
            fn.restype, fn.argtypes = res, args
        _bound = True
    return lib


def exported_symbols():
    return sorted(_SIGS)


class LibsdodError(RuntimeError):
    def __init__(self, status, ctx=None):
        desc = api().libsdod_get_error_description(status)
        extra = api().libsdod_get_last_error_extra_info(status, ctx)
        self.status = status
        super().__init__("%s; %s" % (desc.decode() if desc else "?", extra.decode() if extra else ""))


class Context:
    """with Context("random-init:0", steps=20) as ctx: img = ctx.generate_image("a prompt", 7.5)"""

    def __init__(self, models_dir="random-init:0", latent_spatial=64, steps=20, log_level=LOG_ERROR, max_images=1, device=-1):
        self._h = _vp()
        self.latent_spatial, self.max_images = latent_spatial, max_images
        if max_images == 1 and device < 0:
            st = api().libsdod_setup(ctypes.byref(self._h), models_dir.encode(), 4, latent_spatial, 8, steps, log_level, 1)
        else:
            st = api().libsdod_b200_setup(ctypes.byref(self._h), models_dir.encode(), latent_spatial, steps, log_level, max_images, device)
        if st:
            err = LibsdodError(st, self._h)
            if self._h:
                api().libsdod_release(self._h)
            self._h = None
            raise err

    def _ok(self, st):
        if st:
            raise LibsdodError(st, self._h)

    def set_steps(self, steps):
        self._ok(api().libsdod_set_steps(self._h, steps))

    def set_seed(self, seed):
        self._ok(api().libsdod_b200_set_seed(self._h, seed))

    SAMPLERS = {"dpm": 0, "ddim": 1, "plms": 2}

    def set_sampler(self, name):
        """'dpm' = DPM-Solver++(2M), the reference's sampler (default); 'ddim' = DDIM with eta 0 (row f4, parity unpinned)."""
        self._ok(api().libsdod_b200_set_sampler(self._h, self.SAMPLERS[name] if isinstance(name, str) else int(name)))

    def generate_image(self, prompt, guidance_scale=7.5, out=None):
        side = self.latent_spatial * 8
        if out is None:
            out = np.empty((side, side, 3), dtype=np.uint8)
        buf = out.ctypes.data_as(_ucp)
        n = _u(out.nbytes)
        self._ok(api().libsdod_generate_image(self._h, prompt.encode(), guidance_scale, ctypes.byref(buf), ctypes.byref(n)))
        assert n.value == side * side * 3
        return out

    def encode_prompt(self, prompt, return_tokens=False):
        """The prompt path alone: tokenizer + CLIP text encoder -> [77,768] fp32 (and the 77 token ids)."""
        emb = np.empty((77, 768), dtype=np.float32)
        tok = np.empty(77, dtype=np.uint16)
        raw = prompt if isinstance(prompt, (bytes, bytearray)) else prompt.encode("utf-8")
        self._ok(api().libsdod_b200_encode_prompt(self._h, bytes(raw), emb.ctypes.data, tok.ctypes.data))
        return (emb, tok) if return_tokens else emb

    def generate_images(self, prompts, guidance_scale=7.5):
        """libsdod_generate_image for a list of prompts in one call -> uint8 [n, 8S, 8S, 3]."""
        n = len(prompts)
        arr = (ctypes.c_char_p * n)(*[p if isinstance(p, (bytes, bytearray)) else p.encode("utf-8") for p in prompts])
        side = self.latent_spatial * 8
        imgs = np.empty((n, side, side, 3), dtype=np.uint8)
        self._ok(api().libsdod_b200_generate_images(self._h, n, arr, guidance_scale, imgs.ctypes.data))
        return imgs

    def generate(self, cond, uncond=None, latents=None, guidance_scale=7.5, return_latents=False):
        cond = np.ascontiguousarray(cond, dtype=np.float32)
        n = cond.shape[0]
        uncond = None if uncond is None else np.ascontiguousarray(uncond, dtype=np.float32)
        latents = None if latents is None else np.ascontiguousarray(latents, dtype=np.float32)
        side = self.latent_spatial * 8
        imgs = np.empty((n, side, side, 3), dtype=np.uint8)
        lat_out = np.empty((n, 4, self.latent_spatial, self.latent_spatial), dtype=np.float32) if return_latents else None
        p = lambda a: None if a is None else a.ctypes.data
        self._ok(api().libsdod_b200_generate(self._h, n, p(cond), p(uncond), p(latents), guidance_scale, p(imgs), p(lat_out)))
        return (imgs, lat_out) if return_latents else imgs

    def generate_device(self, cond, uncond, latents_nhwc, guidance_scale, images_out):
        """torch CUDA tensors in, uint8 CUDA tensor out; no host copies (bench.py's HBM-resident leg)."""
        p = lambda a: None if a is None else a.data_ptr()
        self._ok(api().libsdod_b200_generate_device(self._h, cond.shape[0], p(cond), p(uncond), p(latents_nhwc), guidance_scale, p(images_out)))
        return images_out

    # ---- CFG split over a GPU pair (include/libsdod.h: libsdod_b200_pair_*)
    def pair_export(self):
        """64-byte CUDA IPC handle of this context's eps exchange buffer; hand it to the peer process."""
        h = (ctypes.c_ubyte * 64)()
        self._ok(api().libsdod_b200_pair_export(self._h, ctypes.addressof(h)))
        return bytes(h)

    def pair_connect(self, peer_handle, role):
        """Map the peer's exchange buffer; role 0 = this context runs the conditional half, 1 = the unconditional half."""
        h = (ctypes.c_ubyte * 64).from_buffer_copy(peer_handle)
        self._ok(api().libsdod_b200_pair_connect(self._h, ctypes.addressof(h), int(role)))

    def generate_pair(self, conditioning_half, latents=None, guidance_scale=7.5, return_latents=False):
        """This rank's half of classifier-free guidance (cond for role 0, uncond for role 1), [n,77,768]; the peer context must make the
        same call.  Returns (images [n,8S,8S,3] with this rank's decoded share filled in, first_image, n_decoded[, final latents])."""
        c = np.ascontiguousarray(conditioning_half, dtype=np.float32)
        n = c.shape[0]
        latents = None if latents is None else np.ascontiguousarray(latents, dtype=np.float32)
        side = self.latent_spatial * 8
        imgs = np.zeros((n, side, side, 3), dtype=np.uint8)
        lat_out = np.empty((n, 4, self.latent_spatial, self.latent_spatial), dtype=np.float32) if return_latents else None
        first, cnt = _u(0), _u(0)
        p = lambda a: None if a is None else a.ctypes.data
        self._ok(api().libsdod_b200_generate_pair(self._h, n, p(c), p(latents), guidance_scale, p(imgs), ctypes.byref(first), ctypes.byref(cnt), p(lat_out)))
        return (imgs, first.value, cnt.value, lat_out) if return_latents else (imgs, first.value, cnt.value)

    def last_timings(self):
        t = (_f * 4)()
        self._ok(api().libsdod_b200_last_timings(self._h, ctypes.byref(t)))
        return dict(conditioning_ms=t[0], iteration_ms=t[1], decoding_ms=t[2], total_ms=t[3])

    def release(self):
        if self._h:
            api().libsdod_release(self._h)
            self._h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.release()

    def __del__(self):
        self.release()
