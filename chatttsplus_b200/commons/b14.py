"""base16384 ("b14") text codec — replacement for the absent ``pybase16384`` package.

The reference stores speaker embeddings, audio prompts and the DVAE ``coef`` as b14 strings
(reference: chattts_plus/models/tokenizer.py:139-148,180-222, chattts_plus/models/dvae.py:218-220,249-252,
chattts_plus/pipelines/chattts_plus_pipeline.py:56-60,310-319).  base16384 is a public format: every 7
input bytes become four 14-bit symbols, each stored as the code point ``0x4E00 + symbol``; a ragged tail of
``r`` bytes (1..6) is zero-padded to whole symbols and followed by the marker code point ``0x3D00 + r``.

Pinned by the reference's own speaker files (tests/golden/speaker_2222.txt, tests/test_b14_codec.py).
"""
from __future__ import annotations

import numpy as np

_BASE = 0x4E00
_TAIL = 0x3D00


def encode_to_string(data: bytes) -> str:
    data = bytes(data)
    n = len(data)
    r = n % 7
    nbits_sym = (n * 8 + 13) // 14
    bits = np.unpackbits(np.frombuffer(data, dtype=np.uint8))
    pad = nbits_sym * 14 - bits.size
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, dtype=np.uint8)])
    sym = bits.reshape(-1, 14).astype(np.uint32)
    weights = (1 << np.arange(13, -1, -1)).astype(np.uint32)
    vals = (sym * weights).sum(axis=1) + _BASE
    out = "".join(map(chr, vals.tolist()))
    if r:
        out += chr(_TAIL + r)
    return out


def decode_from_string(s: str) -> bytes:
    if not s:
        return b""
    r = 0
    last = ord(s[-1])
    if (last & 0xFF00) == _TAIL:
        r = last - _TAIL
        s = s[:-1]
    vals = np.fromiter((ord(c) - _BASE for c in s), dtype=np.int64, count=len(s))
    if vals.size and (vals.min() < 0 or vals.max() >= (1 << 14)):
        raise ValueError("not a base16384 string")
    bits = ((vals[:, None] >> np.arange(13, -1, -1)[None, :]) & 1).astype(np.uint8).reshape(-1)
    nbytes = bits.size // 8
    out = np.packbits(bits[: nbytes * 8]).tobytes()
    if r:
        tail_syms = (r * 8 + 13) // 14
        groups = (len(s) - tail_syms) // 4
        out = out[: groups * 7 + r]
    return out
