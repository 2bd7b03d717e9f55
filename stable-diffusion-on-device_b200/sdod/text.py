"""Host-side handles of the prompt path (include/sdod_model.h): the CLIP byte-pair tokenizer and the CLIP text encoder.

Reference counterparts: libsdod::Tokenizer (csrc/libsdod/src/tokenizer.cpp; vocabulary file written by gen_tokenizer_file.py) and the
`cond_model` graph executed at context.cpp:237,327."""
import ctypes

import torch

from . import _cabi as C


class Tokenizer:
    """bpe_file: a ctokenizer.txt; None = byte-level vocabulary without merges (what random-init contexts use).  Needs no GPU."""

    def __init__(self, bpe_file=None):
        h = ctypes.c_void_p()
        C.check(C.lib().sdod_tokenizer_create(ctypes.byref(h), bpe_file.encode() if bpe_file is not None else None), "sdod_tokenizer_create")
        self._h = h

    @property
    def vocab_size(self):
        return int(C.lib().sdod_tokenizer_vocab_size(self._h))

    def encode(self, prompt, context_len=77, return_deviated=False):
        """-> list of context_len token ids: [start, ..., end padding] (tokenizer.cpp:258-276).  Invalid UTF-8 raises SdodError.
        return_deviated: also return whether the reference's merge loop would not have terminated on this prompt (see sdod_model.h)."""
        raw = prompt if isinstance(prompt, (bytes, bytearray)) else prompt.encode("utf-8")
        out = (ctypes.c_ushort * context_len)()
        dev = ctypes.c_int(0)
        C.check(C.lib().sdod_tokenizer_encode(self._h, bytes(raw), out, context_len, ctypes.byref(dev)), "sdod_tokenizer_encode")
        return (list(out), bool(dev.value)) if return_deviated else list(out)

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                C.lib().sdod_tokenizer_destroy(self._h)
                self._h = None
        except Exception:          # interpreter shutdown: the module globals may already be gone
            pass


class TextEncoder:
    """CLIP ViT-L/14 text encoder on the library's kernels.  weights: sdod.model.Weights holding the HuggingFace CLIPTextModel state_dict
    ("text_model...." keys) or None for random-init."""

    def __init__(self, weights=None, seed=0, max_batch=2):
        h = ctypes.c_void_p()
        self._weights = weights
        C.check(C.lib().sdod_text_encoder_create(ctypes.byref(h), weights._h if weights is not None else None, seed, max_batch), "sdod_text_encoder_create")
        self._h, self.max_batch = h, max_batch

    def __call__(self, tokens, dtype=torch.float32):
        """tokens [B,77] integer tensor -> last_hidden_state [B,77,768] on the device."""
        t = tokens.to("cuda", torch.int32).contiguous()
        assert t.dim() == 2 and t.shape[1] == 77
        out = torch.empty(t.shape[0], 77, 768, dtype=dtype, device="cuda")
        dt = C.F32 if dtype == torch.float32 else C.BF16
        C.check(C.lib().sdod_text_encoder_forward(self._h, torch.cuda.current_stream().cuda_stream, t.data_ptr(), t.shape[0], out.data_ptr(), dt),
                "sdod_text_encoder_forward")
        return out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                C.lib().sdod_text_encoder_destroy(self._h)
                self._h = None
        except Exception:
            pass
