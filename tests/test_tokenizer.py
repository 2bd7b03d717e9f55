"""CPU: the prompt tokenizer against the reference's own tokenizer.

Pins: (1) golden token ids recorded from the reference's tokenizer.cpp compiled in place (tests/golden/make_tokenizer_golden.py) on a vocabulary file
in the reference's format; (2) where oracle/_ref/libtok_ref.so is present (built from /root/reference by `make -C oracle ref`), a seeded fuzz
against it, bit-for-bit on the ids.  The library needs no GPU for this path."""
import ctypes
import json
import locale
import os
import random

import pytest

from sdod._cabi import SdodError
from sdod.text import Tokenizer

ROOT = os.path.normpath(os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
REF_LIB = os.path.join(ROOT, "oracle", "_ref", "libtok_ref.so")


@pytest.fixture(scope="module")
def golden(golden_dir):
    return json.load(open(os.path.join(golden_dir, "tokenizer_golden.json")))


@pytest.fixture(scope="module")
def tok(golden_dir):
    return Tokenizer(os.path.join(golden_dir, "ctokenizer_synth.txt"))


def test_vocabulary_file_layout(tok, golden):
    # 512 byte symbols + one token per merge + start + end (tokenizer.cpp:228-255)
    assert tok.vocab_size == 512 + golden["merges"] + 2


def test_golden_ids_from_the_compiled_reference(tok, golden):
    assert len(golden["cases"]) >= 150
    rejected = 0
    for case in golden["cases"]:
        raw = bytes.fromhex(case["utf8_hex"])
        if case["ids"] is None:                    # the reference throws INVALID_ARGUMENT "Invalid UTF-8 string" (tokenizer.cpp:77,179)
            rejected += 1
            with pytest.raises(SdodError, match="Invalid UTF-8"):
                tok.encode(raw)
            continue
        ids, deviated = tok.encode(raw, return_deviated=True)
        assert not deviated
        assert ids == case["ids"], raw
        assert len(ids) == 77 and ids[0] == tok.vocab_size - 2 and ids[-1] == tok.vocab_size - 1
    assert rejected >= 5


def test_short_context_lengths(tok, golden):
    for case in golden["short_context"]:
        assert tok.encode(bytes.fromhex(case["utf8_hex"]), case["context_len"]) == case["ids"]


def test_prompts_the_reference_never_finishes(tok, golden):
    """[a, a, b] with best-ranked pair (a, b): the reference's merge pass leaves the word unchanged and loops (tokenizer.cpp:339-356).
    The library flags these prompts and applies the textbook merge; everything else about the output format holds."""
    assert golden["reference_does_not_terminate"]
    for h in golden["reference_does_not_terminate"]:
        ids, deviated = tok.encode(bytes.fromhex(h), return_deviated=True)
        assert deviated and len(ids) == 77 and ids[0] == tok.vocab_size - 2 and ids[-1] == tok.vocab_size - 1


def test_byte_level_vocabulary_without_a_file():
    t = Tokenizer()
    assert t.vocab_size == 514
    ids = t.encode("ab A")
    # gen_tokenizer_file.py:33-34 order: '!'..'~' first -> 'a' = 64, 'b' = 65; "</w>" forms 256 later; upper case is lowered
    assert ids[:5] == [512, 64, 65 + 256, 64 + 256, 513] and ids[5:] == [513] * 72
    assert t.encode("")[:2] == [512, 513]


def test_missing_vocabulary_file_is_an_error():
    with pytest.raises(SdodError, match="does not exist"):
        Tokenizer("/nonexistent/ctokenizer.txt")


def test_process_locale_is_left_alone(tok):
    before = locale.setlocale(locale.LC_ALL)
    tok.encode("Ünïcödé 日本語 😀")
    assert locale.setlocale(locale.LC_ALL) == before


@pytest.mark.skipif(not os.path.exists(REF_LIB), reason="oracle/_ref/libtok_ref.so not built (needs /root/reference)")
@pytest.mark.timeout(300)
def test_fuzz_against_the_compiled_reference(tok, golden_dir):
    lib = ctypes.CDLL(REF_LIB)
    lib.tok_ref_create.restype = ctypes.c_void_p
    lib.tok_ref_create.argtypes = [ctypes.c_char_p]
    lib.tok_ref_tokenize.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint16), ctypes.c_uint]
    lib.tok_ref_destroy.argtypes = [ctypes.c_void_p]
    h = lib.tok_ref_create(os.path.join(golden_dir, "ctokenizer_synth.txt").encode())
    assert h
    before = locale.setlocale(locale.LC_ALL)
    rng = random.Random(20261018)
    pieces = list("abcdefghijklmnopqrstuvwxyz") * 2 + list("ABCXYZ 0123456789.,;:!?-_'\"()[]/\\@#") + [" ", " ", "  ", "\t", "\n", "'s", "'t", "'re", "'ve", "'m", "'ll", "'d",
              "é", "É", "ß", "Ω", "ω", "Ж", "ж", "١", "１", "日", "本", "😀", " ", " ", "　", "́", "​", "ing", "the", "tion", "oo", "ll", "ss"]
    compared = skipped = rejected = 0
    out = (ctypes.c_uint16 * 77)()
    try:
        for i in range(3000):
            raw = "".join(rng.choice(pieces) for _ in range(rng.randrange(0, 80))).encode("utf-8")
            if i % 50 == 0:                                   # sprinkle invalid UTF-8
                raw = raw[:len(raw) // 2] + bytes([rng.choice([0x80, 0xC0, 0xFF, 0xE2, 0xF0])]) + raw[len(raw) // 2:]
            try:
                ids, deviated = tok.encode(raw, return_deviated=True)
            except SdodError:
                ids, deviated = None, False
            if deviated:                                      # the reference would spin for ever on this one
                skipped += 1
                continue
            n = lib.tok_ref_tokenize(h, raw, out, 77)
            want = None if n < 0 else list(out)[:n]
            assert ids == want, raw
            compared += 1
            rejected += want is None
    finally:
        lib.tok_ref_destroy(h)
        locale.setlocale(locale.LC_ALL, before)              # the shim selects C.utf8 for the reference's wide-character calls
    assert compared > 2000 and rejected > 10
