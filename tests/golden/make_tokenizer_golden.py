"""Fixtures for the prompt tokenizer (run HERE, where /root/reference exists; the outputs are committed):
  tests/golden/ctokenizer_synth.txt  — a vocabulary file in the reference's format, written by the logic of the reference's
      gen_tokenizer_file.py:27-42 (512 byte symbols, then merges) with merges learnt by a few hundred rounds of textbook BPE training on a
      small prompt corpus (the real bpe_simple_vocab_16e6.txt.gz is not in the reference tree or this image);
  tests/golden/tokenizer_golden.json — prompts and the token ids the reference's OWN tokenizer (csrc/libsdod/src/tokenizer.cpp, compiled in
      place as oracle/_ref/libtok_ref.so) returns for them on that file; invalid UTF-8 cases record the reference's exception as null.
   python tests/golden/make_tokenizer_golden.py"""
import collections
import ctypes
import json
import os
import random
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.normpath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, os.path.join(ROOT, "stable-diffusion-on-device_b200"))
from sdod._cabi import SdodError  # noqa: E402
from sdod.text import Tokenizer  # noqa: E402

CORPUS = """a photograph of an astronaut riding a horse on mars, highly detailed, 8k
an oil painting of a cat wearing a top hat, in the style of van gogh
a cute corgi puppy playing in the snow, golden hour lighting, bokeh
cyberpunk city street at night, neon lights, rain, reflections, cinematic
portrait of an old fisherman, weathered face, dramatic lighting, 85mm lens
a bowl of ramen with a soft boiled egg, food photography, steam rising
the quick brown fox jumps over the lazy dog
it's a beautiful day and we're going to the park, isn't it? they've said so, i'll go, he'd go, i'm in
watercolor landscape of mountains and a lake at sunrise, pastel colors
a red sports car parked in front of a modern glass building
isometric pixel art of a tiny island with a lighthouse
a medieval castle on a cliff above a stormy sea, fantasy concept art, trending on artstation
macro photo of a dew covered spider web, shallow depth of field
a steampunk robot reading a book in a cozy library
1920s jazz club, black and white photo, film grain, 35mm
the tall cool tree stood still; all the small balls fell off the wall
good food mood wood hood book look took cook pool tool cool fool
""".strip().split("\n")


def bytes_to_unicode():                       # the mapping the reference's generator takes from the CLIP code base (gen_tokenizer_file.py:5-24)
    bs = list(range(ord("!"), ord("~") + 1)) + list(range(0xA1, 0xAC + 1)) + list(range(0xAE, 0xFF + 1))
    cs = bs[:]
    n = 0
    for b in range(256):
        if b not in bs:
            bs.append(b)
            cs.append(256 + n)
            n += 1
    return dict(zip(bs, [chr(c) for c in cs]))


def train_bpe(lines, rounds):
    b2u = bytes_to_unicode()
    words = collections.Counter()
    for line in lines:
        for w in line.lower().split():
            sym = tuple(b2u[b] for b in w.encode("utf-8"))
            words[sym[:-1] + (sym[-1] + "</w>",)] += 1
    merges = []
    for _ in range(rounds):
        pairs = collections.Counter()
        for w, c in words.items():
            for a, b in zip(w[:-1], w[1:]):
                pairs[(a, b)] += c
        if not pairs:
            break
        (a, b), c = max(pairs.items(), key=lambda kv: (kv[1], kv[0]))
        merges.append((a, b))
        new = collections.Counter()
        for w, cnt in words.items():
            out, i = [], 0
            while i < len(w):
                if i + 1 < len(w) and w[i] == a and w[i + 1] == b:
                    out.append(a + b)
                    i += 2
                else:
                    out.append(w[i])
                    i += 1
            new[tuple(out)] += cnt
        words = new
    return merges


def main():
    vocab = list(bytes_to_unicode().values())
    vocab = vocab + [v + "</w>" for v in vocab]
    merges = train_bpe(CORPUS, 700)
    path = os.path.join(HERE, "ctokenizer_synth.txt")
    with open(path, "wb") as f:                                   # gen_tokenizer_file.py:37-42
        for v in vocab:
            f.write((v + "\n").encode("utf-8"))
        for a, b in merges:
            f.write((a + " " + b + "\n").encode("utf-8"))
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libtok_ref.so"))
    lib.tok_ref_create.restype = ctypes.c_void_p
    lib.tok_ref_create.argtypes = [ctypes.c_char_p]
    lib.tok_ref_tokenize.argtypes = [ctypes.c_void_p, ctypes.c_char_p, ctypes.POINTER(ctypes.c_uint16), ctypes.c_uint]
    h = lib.tok_ref_create(path.encode())
    assert h

    def ref(raw, ctx=77):
        out = (ctypes.c_uint16 * ctx)()
        n = lib.tok_ref_tokenize(h, raw, out, ctx)
        return None if n < 0 else list(out)[:n]

    prompts = [l.encode() for l in CORPUS]
    prompts += [s.encode("utf-8") for s in [
        "", " ", "   leading and trailing   ", "A Photograph OF An ASTRONAUT", "tabs\tand  double  spaces\t\t!", "it's they're we've i'll he'd i'm can't 'tis",
        "'s 't 're 've 'm 'll 'd", "''''", "don''t", "x'y", "1234567890", "route 66 and 3.14159, 50% off!!!", "hello,world;foo:bar(baz)[qux]{quux}", "café Élan naïve ÀÉÎ",
        "Ω ω Ünïcödé", "日本語のプロンプト", "emoji 😀 party 🎉🎉", "a b nbsp", "em space　ideographic", "full１２width digits", "line\nbreak", "ǅ titlecase",
        "snake_case and kebab-case and dots...", "\"quoted\" 'single' `back`", "a" * 200, "word " * 100, "lll aba abab tool cool good", "oolong zoology",
        "!!!???...,,,", "mixed123abc456", "trailing digit 7", "7 leading digit", "a--b", "--", "́combining", "x​y zero width",
    ]]
    prompts += [b"\xff\xfe", b"ok \xc3", b"\xc0\xaf overlong", b"\xed\xa0\x80 surrogate", b"truncated \xf0\x9f\x98", b"\xf5\x80\x80\x80"]
    rng = random.Random(1234)
    alphabet = list("abcdelot aeio '.,!0123456789-") + ["é", "Ω", "日", "😀", "\t", "  ", "'s", "'ll", "ing", "the", "oo", "ll"]
    for _ in range(120):
        prompts.append("".join(rng.choice(alphabet) for _ in range(rng.randrange(1, 60))).encode("utf-8"))
    # Prompts on which the reference's merge loop does not terminate (tokenizer.cpp:339-356, see tests/test_tokenizer.py) cannot be recorded: our
    # library flags exactly those (sdod_tokenizer_encode's `deviated`), so they are filtered with it and listed separately without ids.
    ours = Tokenizer(path)
    cases, hangs = [], []
    for raw in prompts:
        try:
            _, dev = ours.encode(raw, return_deviated=True)
        except SdodError:
            dev = False
        if dev:
            hangs.append(raw.hex())
            continue
        cases.append({"utf8_hex": raw.hex(), "ids": ref(raw)})
    short = [{"utf8_hex": p.hex(), "context_len": c, "ids": ref(p, c)} for p, c in ((b"a photograph of an astronaut", 8), (b"word " * 30, 16), (b"", 4))]
    json.dump({"vocab_file": "ctokenizer_synth.txt", "merges": len(merges), "cases": cases, "short_context": short, "reference_does_not_terminate": hangs},
              open(os.path.join(HERE, "tokenizer_golden.json"), "w"), indent=0)
    print("merges %d, cases %d (%d rejected by the reference), reference-hang prompts %d" % (len(merges), len(cases), sum(c["ids"] is None for c in cases), len(hangs)))


if __name__ == "__main__":
    main()
