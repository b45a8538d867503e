"""Host text front-end against outputs of the reference's own commons/text_utils.py and commons/norm.py
(tests/golden/text_ref.json, produced by tests/golden/make_golden.py::gen_text)."""
import json
import os

import pytest

from chatttsplus_b200 import text as T


@pytest.fixture(scope="module")
def ref(golden_dir):
    with open(os.path.join(golden_dir, "text_ref.json"), encoding="utf-8") as f:
        return json.load(f)


def _call(fn, *a):
    try:
        return fn(*a)
    except Exception as e:   # the reference's error behaviour is part of the contract (num_to_english(10) raises IndexError)
        return "!" + type(e).__name__


def test_num_to_english_matches_reference(ref):
    for n, want in ref["num_to_english"].items():
        assert _call(T.num_to_english, int(n)) == want, n
    assert T.num_to_english(105) == "One hundred and five" and T.num_to_english(0) == ""


def test_text_utils_match_reference(ref):
    for t, want in ref["num2text"].items():
        assert _call(T.num2text, t) == want, t
    for t, want in ref["remove_brackets"].items():
        assert T.remove_brackets(t) == want, t
    for t, want in ref["get_lang"].items():
        assert T.get_lang(t) == want, t
    for t, want in ref["split_by_punct"].items():
        assert T.split_text_by_punctuation(t) == want, t


def test_normalizer_matches_reference(ref, tmp_path):
    mp = tmp_path / "homophones_map.json"
    mp.write_text(json.dumps(ref["homophones"], ensure_ascii=False), encoding="utf-8")
    nz = T.Normalizer(str(mp))
    assert nz.register("en", lambda s: s.replace("test", "TEST"))
    assert not nz.register("en", lambda s: s)                      # duplicate name refused (norm.py:161-163)
    assert not nz.register("bad", lambda s: 1)                     # wrong return type refused (norm.py:164-168)
    for c in ref["normalizer"]:
        assert nz(c["text"], c["tn"], c["hr"], c["lang"]) == c["out"], c
    assert {chr(k): chr(v) for k, v in T._SIMPLIFY.items()} == ref["tables"]["simplify"]
    assert {chr(k): chr(v) for k, v in T._HALF2FULL.items()} == ref["tables"]["half2full"]


def test_split_and_merge_flow():
    """split_text (English path: the reference falls back to num2text when nemo is unavailable, text_utils.py:147-151) and the
    short-piece merge of chattts_plus_pipeline.py:359-373."""
    out = T.split_text(["I have 3 cats [laugh] (really)", "ok"])
    assert out == ["I have Three cats  [laugh]   (really)", "ok"]   # num_to_english capitalises, the tag keeps its brackets
    long_line = ("This is a sentence, " * 15).strip()
    pieces = T.split_text([long_line])
    assert len(pieces) >= 2 and "".join(pieces) == T.num2text(T.remove_brackets(long_line))
    assert T.merge_short_texts(["a", "b"]) == ["a [uv_break] b [uv_break] "]
    assert T.merge_short_texts(["x" * 40, "tail"]) == ["x" * 40 + " [uv_break] tail [uv_break] "]
    assert T.merge_short_texts(["x" * 20, "y" * 20, "z" * 40]) == ["x" * 20 + " [uv_break] ", "y" * 20 + " [uv_break] ", "z" * 40]


def test_prompt_audio_loader_reads_pcm_wav_without_codec_backend(tmp_path):
    """speaker_audio_path loading (chattts_plus_pipeline.py:495-498): torchaudio.load needs an optional codec package; PCM WAV files
    are read with the standard library, resampled to 24 kHz and averaged over channels."""
    import wave
    import numpy as np
    import torch
    from chatttsplus_b200.pipeline import ChatTTSPlusPipeline
    t = np.arange(8000) / 8000.0
    left = (np.sin(2 * np.pi * 220 * t) * 12000).astype("<i2")
    right = (np.sin(2 * np.pi * 330 * t) * 8000).astype("<i2")
    p = str(tmp_path / "stereo8k.wav")
    with wave.open(p, "wb") as f:
        f.setnchannels(2); f.setsampwidth(2); f.setframerate(8000)
        f.writeframes(np.stack([left, right], 1).reshape(-1).tobytes())
    wav = ChatTTSPlusPipeline._load_audio_24k(p)
    assert wav.shape == (24000,) and wav.dtype == torch.float32
    assert 0.1 < float(wav.abs().max()) < 0.45                     # mean of the two channels, int16 full scale = 1.0
    p8 = str(tmp_path / "mono8bit.wav")
    with wave.open(p8, "wb") as f:
        f.setnchannels(1); f.setsampwidth(1); f.setframerate(24000)
        f.writeframes((np.sin(2 * np.pi * 100 * np.arange(2400) / 24000.0) * 100 + 128).astype(np.uint8).tobytes())
    w8 = ChatTTSPlusPipeline._load_audio_24k(p8)
    assert w8.shape == (2400,) and abs(float(w8.mean())) < 0.05
