"""ctypes binding of libsdod_b200.so — the C-ABI boundary (include/sdod_kernels.h, include/libsdod.h).

There is no CPU fallback: if the library is missing, import of the ops fails loudly with the build
command; if no CUDA device is present every compute entry point returns an error status.
"""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.normpath(os.path.join(_HERE, "..", "csrc"))
LIB_PATH = os.path.join(_HERE, "lib", "libsdod_b200.so")

F32, BF16 = 0, 1
NCHW, NHWC = 0, 1
ACT_NONE, ACT_SILU, ACT_GELU, ACT_GEGLU, ACT_QUICK_GELU = 0, 1, 2, 3, 4
OUT_BF16, OUT_F32, OUT_HEADS, OUT_HEADS_T, OUT_QKV = 0, 1, 2, 3, 4

c_vp, c_int, c_ll, c_f, c_sz, c_u = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float, ctypes.c_size_t, ctypes.c_uint


class SdodError(RuntimeError):
    pass


class Epilogue(ctypes.Structure):
    _fields_ = [("C", c_vp), ("C2", c_vp), ("C3", c_vp), ("ldc", c_ll), ("strideC", c_ll), ("bias", c_vp), ("row_bias", c_vp),
                ("rows_per_group", c_int), ("ld_row_bias", c_ll), ("residual", c_vp), ("ldr", c_ll), ("strideR", c_ll), ("alpha", c_f), ("act", c_int),
                ("out_mode", c_int), ("heads", c_int), ("head_dim", c_int), ("tokens", c_int), ("dpad", c_int), ("tok_pad", c_int), ("vt_rows", c_int), ("residual_f32", c_int),
                ("ln_out", c_vp), ("ld_ln", c_ll), ("ln_weight", c_vp), ("ln_bias", c_vp), ("ln_eps", c_f)]


class GemmDesc(ctypes.Structure):
    _fields_ = [("A", c_vp), ("lda", c_ll), ("strideA", c_ll), ("W", c_vp), ("ldw", c_ll), ("strideW", c_ll),
                ("M", c_int), ("N", c_int), ("K", c_int), ("batch", c_int), ("block_n", c_int), ("epi", Epilogue),
                ("A2", c_vp), ("lda2", c_ll), ("K2", c_int)]


class ConvDesc(ctypes.Structure):
    _fields_ = [("X", c_vp), ("Wt", c_vp), ("B", c_int), ("H", c_int), ("W", c_int), ("Cin", c_int), ("Cout", c_int),
                ("block_n", c_int), ("epi", Epilogue), ("X2", c_vp), ("ldx2", c_ll), ("Cin2", c_int), ("upsample2x", c_int), ("stride", c_int)]


_SIGS = {
    "sdod_last_error": (ctypes.c_char_p, []),
    "sdod_abi_version": (c_int, []),
    "sdod_launch_count": (ctypes.c_ulonglong, []),
    "sdod_group_norm_workspace": (c_sz, [c_int, c_int, c_int, c_int, c_int]),
    "sdod_group_norm": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f, c_int, c_int, c_int, c_vp, c_sz]),
    "sdod_group_norm_nhwc": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_f, c_int, c_vp, c_sz]),
    "sdod_group_norm_nhwc2": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_f, c_int, c_vp, c_sz]),
    "sdod_group_norm_nhwc2_supported": (c_int, [c_int, c_int, c_int, c_int, c_int, c_int]),
    "sdod_layer_norm": (c_int, [c_vp, c_vp, c_int, c_vp, c_vp, c_vp, c_int, c_int, c_f]),
    "sdod_cfg_dpm_step": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_sz, c_f, c_f, c_f, c_f, c_f, c_f, c_int, c_vp]),
    "sdod_cfg_dpm_step_pair": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_vp, c_u, c_int, c_sz, c_f, c_f, c_f, c_f, c_f, c_f, c_int]),
    "sdod_dpm_schedule": (c_int, [c_u, c_f, c_f, c_u] + [c_vp] * 8),
    "sdod_dpm_coeffs": (c_int, [c_u, c_f, c_f, c_u, c_u] + [c_vp] * 6),
    "sdod_ddim_schedule": (c_int, [c_u, c_f, c_f, c_u, c_vp, c_vp]),
    "sdod_cfg_lms_step": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_sz, c_f, c_vp, c_vp, c_vp, c_vp, c_vp, c_f, c_f, c_f, c_f, c_vp]),
    "sdod_timestep_sinusoid": (c_int, [c_vp, c_vp, c_int, c_int, c_f, c_vp]),
    "sdod_randn": (c_int, [c_vp, c_vp, c_sz, ctypes.c_ulonglong, ctypes.c_ulonglong]),
    "sdod_image_to_u8": (c_int, [c_vp, c_vp, c_int, c_vp, c_sz]),
    "sdod_gemm_bf16": (c_int, [c_vp, ctypes.POINTER(GemmDesc)]),
    "sdod_set_splitk_workspace": (c_int, [c_vp, c_sz, c_vp, c_int]),
    "sdod_set_gemm_timeline": (c_int, [c_vp]),
    "sdod_conv3x3_bf16": (c_int, [c_vp, ctypes.POINTER(ConvDesc)]),
    "sdod_attention_bf16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f]),
    "sdod_attention_causal_bf16": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_f]),
    "sdod_softmax_rows": (c_int, [c_vp, c_vp, c_vp, c_ll, c_int, c_ll, c_f]),
    "sdod_nchw_f32_to_nhwc_bf16": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int]),
    "sdod_nhwc_to_nchw_f32": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int]),
    "sdod_upsample2x_nhwc": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int]),
    "sdod_concat_channels": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_vp, c_ll, c_int]),
    "sdod_im2col3x3": (c_int, [c_vp, c_vp, c_int, c_vp, c_int, c_int, c_int, c_int, c_int, c_int]),
    "sdod_cast_f32_to_bf16": (c_int, [c_vp, c_vp, c_vp, c_sz]),
    "sdod_silu_bf16": (c_int, [c_vp, c_vp, c_vp, c_sz]),
    "sdod_pack_conv3x3_weight": (c_int, [c_vp, c_vp, c_vp, c_int, c_int, c_int]),
    "sdod_pack_conv3x3_up2_weight": (c_int, [c_vp, c_vp, c_vp, c_int, c_int]),
    # include/sdod_model.h
    "sdod_weights_create": (c_int, [ctypes.POINTER(c_vp)]),
    "sdod_weights_set_f32": (c_int, [c_vp, ctypes.c_char_p, c_vp, c_int, ctypes.POINTER(c_ll)]),
    "sdod_weights_load_file": (c_int, [c_vp, ctypes.c_char_p]),
    "sdod_weights_count": (c_ll, [c_vp]),
    "sdod_weights_destroy": (None, [c_vp]),
    "sdod_unet_create": (c_int, [ctypes.POINTER(c_vp), c_vp, ctypes.c_ulonglong, c_int, c_int]),
    "sdod_unet_destroy": (None, [c_vp]),
    "sdod_unet_time_embed": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp]),
    "sdod_unet_set_context": (c_int, [c_vp, c_vp, c_vp, c_int, c_int]),
    "sdod_unet_forward": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int]),
    "sdod_unet_launches_per_forward": (ctypes.c_ulonglong, [c_vp, c_int]),
    "sdod_unet_profile": (c_int, [c_vp, c_vp, c_int, c_int, ctypes.c_char_p, c_sz]),
    "sdod_vae_create": (c_int, [ctypes.POINTER(c_vp), c_vp, ctypes.c_ulonglong, c_int, c_int]),
    "sdod_vae_destroy": (None, [c_vp]),
    "sdod_vae_decode": (c_int, [c_vp, c_vp, c_vp, c_vp, c_vp, c_int, c_int]),
    "sdod_tokenizer_create": (c_int, [ctypes.POINTER(c_vp), ctypes.c_char_p]),
    "sdod_tokenizer_destroy": (None, [c_vp]),
    "sdod_tokenizer_encode": (c_int, [c_vp, ctypes.c_char_p, ctypes.POINTER(ctypes.c_ushort), c_u, ctypes.POINTER(c_int)]),
    "sdod_tokenizer_vocab_size": (c_int, [c_vp]),
    "sdod_text_encoder_create": (c_int, [ctypes.POINTER(c_vp), c_vp, ctypes.c_ulonglong, c_int]),
    "sdod_text_encoder_destroy": (None, [c_vp]),
    "sdod_text_encoder_forward": (c_int, [c_vp, c_vp, c_vp, c_int, c_vp, c_int]),
}

_lib = None


def build(verbose=False):
    """Compile every CUDA source for sm_100a in-tree (nvcc cross-compiles without a GPU)."""
    r = subprocess.run(["make", "-C", CSRC, "-j8"], capture_output=True, text=True)
    if r.returncode != 0:
        raise SdodError("building libsdod_b200.so failed:\n" + r.stdout[-4000:] + r.stderr[-4000:])
    if verbose:
        print(r.stdout[-2000:])
    return LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise SdodError("libsdod_b200.so is not built (%s). There is no CPU fallback; run `make -C %s` "
                            "or `python -c 'import __graft_entry__ as g; g.build()'`." % (LIB_PATH, CSRC))
        l = ctypes.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            if hasattr(l, name):
                fn = getattr(l, name)
                fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(status, what):
    if status != 0:
        msg = lib().sdod_last_error()
        raise SdodError("%s failed (status %d): %s" % (what, status, msg.decode() if msg else "?"))


def exported_symbols():
    return sorted(_SIGS.keys())
