"""b14 + LZMA speaker codec pinned by the reference's shipped speaker string (assets/speakers/2222.pt)."""
import lzma
import os

import numpy as np

from chatttsplus_b200.commons import b14

_FILTERS = [{"id": lzma.FILTER_LZMA2, "preset": 9 | lzma.PRESET_EXTREME}]


def test_known_answer_speaker_2222(golden_dir):
    s = open(os.path.join(golden_dir, "speaker_2222.txt"), encoding="utf-8").read()
    assert len(s) == 862
    raw = b14.decode_from_string(s)
    assert len(raw) == 1506
    dec = lzma.decompress(raw, format=lzma.FORMAT_RAW, filters=_FILTERS)
    a = np.frombuffer(dec, dtype=np.float16)
    assert a.shape == (768,)
    assert np.allclose(a[:4].astype(np.float32), [0.2197, 1.773, -3.531, -1.47], atol=2e-3)
    assert abs(float(np.linalg.norm(a.astype(np.float32))) - 132.46) < 0.01
    assert b14.encode_to_string(raw) == s


def test_roundtrip_all_tail_lengths():
    rng = np.random.default_rng(0)
    for n in range(0, 64):
        d = rng.integers(0, 256, n, dtype=np.uint8).tobytes()
        assert b14.decode_from_string(b14.encode_to_string(d)) == d
