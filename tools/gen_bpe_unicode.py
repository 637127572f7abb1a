#!/usr/bin/env python
"""Generates arks_b200/data/bpe_unicode.bin.z: one nibble per Unicode code point for the BPE pre-tokenizer on the device
(arks_b200/csrc/bpe.cuh):  bits 0-1  class  0 other, 1 letter (\\p{L}), 2 number (\\p{N}), 3 white space (\\s)
                          bit 2     NFC-unsafe: the code point may change under NFC normalisation (Qwen2's tokenizer.json
                                    has an NFC normaliser; a text containing such a code point is reported as uncounted)

The classes are NOT taken from Python's unicodedata (its Unicode version differs from the regex engine's): they are
measured on the engine the oracle uses — HF `tokenizers`' Split pre-tokenizer with the Qwen2 pattern
(transformers/models/qwen2/tokenization_qwen2.py:33) — by probing every code point in two contexts.
tests/test_bpe.py re-derives a sample and checks the committed table is what this script produces.

    python tools/gen_bpe_unicode.py
"""
import os
import unicodedata
import zlib

import numpy as np

QWEN2_PATTERN = (r"""(?i:'s|'t|'re|'ve|'m|'ll|'d)|[^\r\n\p{L}\p{N}]?\p{L}+|\p{N}| ?[^\s\p{L}\p{N}]+[\r\n]*|\s*[\r\n]+"""
                 r"""|\s+(?!\S)|\s+""")
OTHER, LETTER, NUMBER, SPACE, NFC_UNSAFE = 0, 1, 2, 3, 4
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "arks_b200", "data", "bpe_unicode.bin.z")


def splitter():
    from tokenizers import Regex, pre_tokenizers
    return pre_tokenizers.Split(Regex(QWEN2_PATTERN), behavior="isolated", invert=False)


def classify(cp: int, split) -> int:
    """class of one code point, read off how the pattern cuts "x<c>x" and "<c><c>x" """
    if cp in (10, 13):
        return SPACE
    c = chr(cp)
    a = [sp for _, sp in split.pre_tokenize_str("x" + c + "x")]
    if a == [(0, 3)]:
        return LETTER
    b = [sp for _, sp in split.pre_tokenize_str(c + c + "x")]
    if a == [(0, 1), (1, 2), (2, 3)] and b == [(0, 1), (1, 2), (2, 3)]:
        return NUMBER
    return SPACE if b == [(0, 1), (1, 3)] else OTHER


def nfc_unsafe() -> np.ndarray:
    """code points that NFC may change: not NFC themselves, reordered (combining class != 0), or able to compose with
    what precedes them (second element of a canonical pair, Hangul vowels / trailing consonants)"""
    u = np.zeros(0x110000, bool)
    for cp in range(0x110000):
        if 0xD800 <= cp <= 0xDFFF:
            continue
        c = chr(cp)
        if unicodedata.combining(c) or unicodedata.normalize("NFC", c) != c:
            u[cp] = True
        d = unicodedata.decomposition(c)
        if d and not d.startswith("<"):
            parts = d.split()
            if len(parts) == 2:
                u[int(parts[1], 16)] = True
    u[0x1161:0x1176] = True
    u[0x11A8:0x11C3] = True
    return u


def build() -> np.ndarray:
    split = splitter()
    cls = np.zeros(0x110000, np.uint8)
    for cp in range(0x110000):
        if not 0xD800 <= cp <= 0xDFFF:
            cls[cp] = classify(cp, split)
    cls |= nfc_unsafe().astype(np.uint8) * NFC_UNSAFE
    return cls[0::2] | (cls[1::2] << 4)  # code point 2k in the low nibble of byte k


if __name__ == "__main__":
    packed = build()
    os.makedirs(os.path.dirname(OUT), exist_ok=True)
    open(OUT, "wb").write(zlib.compress(packed.tobytes(), 9))
    print("wrote", OUT, os.path.getsize(OUT), "bytes;", np.bincount(np.concatenate([packed & 3, (packed >> 4) & 3])))
