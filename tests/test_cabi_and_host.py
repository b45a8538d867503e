"""CPU-side checks: the C-ABI library loads and exports every symbol include/ctp.h declares (no compute without a GPU),
the product path fails loudly without a device, host-side mirrors keep the reference's signatures, the config shim works."""
import ctypes
import inspect
import os
import re

import pytest
import torch

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), ".."))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "ctp.h")).read()
    return sorted(set(re.findall(r"CTP_API\s+[\w\s\*]+?\b(ctp_\w+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__
    __graft_entry__.build()
    from chatttsplus_b200 import _lib
    lib = _lib.lib()
    names = _declared_symbols()
    assert len(names) >= 20
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/ctp.h but not exported"
        assert n in _lib.SYMBOLS, f"{n} has no ctypes prototype in chatttsplus_b200/_lib.py"
    assert lib.ctp_version() >= 100


@pytest.mark.skipif(torch.cuda.is_available(), reason="CPU-only behaviour")
def test_no_cpu_fallback():
    from chatttsplus_b200 import _lib, synth
    from chatttsplus_b200.gpt import GPT
    lib = _lib.lib()
    assert lib.ctp_device_check(0) != 0
    assert b"CUDA" in lib.ctp_last_error() or b"device" in lib.ctp_last_error()
    g = GPT(dict(hidden_size=768, intermediate_size=3072, num_attention_heads=12, num_hidden_layers=1), num_text_tokens=64)
    with pytest.raises(_lib.CtpError):
        g.to("cuda")
    with pytest.raises(_lib.CtpError):
        g(torch.zeros(1, 2, 4, dtype=torch.long), torch.ones(1, 2, dtype=torch.bool))


def test_reference_signatures_are_kept():
    from chattts_plus.pipelines.chattts_plus_pipeline import ChatTTSPlusPipeline
    from chattts_plus.commons.utils import InferCodeParams, RefineTextParams, TorchSeedContext, get_inference_device  # noqa: F401
    from chattts_plus.commons import constants
    from chattts_plus import models
    sig = inspect.signature(ChatTTSPlusPipeline.infer)
    for name in ["text", "stream", "lang", "skip_refine_text", "refine_text_only", "use_decoder", "do_text_normalization",
                 "do_text_optimization", "do_homophone_replacement", "params_refine_text", "params_infer_code", "kwargs"]:
        assert name in sig.parameters, name
    gsig = inspect.signature(models.GPT.generate)
    for name in ["emb", "inputs_ids", "temperature", "eos_token", "attention_mask", "max_new_token", "min_new_token",
                 "logits_warpers", "logits_processors", "infer_text", "return_attn", "return_hidden", "stream", "show_tqdm",
                 "ensure_non_empty", "stream_batch", "context"]:
        assert name in gsig.parameters, name
    p = InferCodeParams()
    assert (p.prompt, p.temperature, p.repetition_penalty, p.max_new_token, p.top_P, p.top_K) == ("[speed_5]", 0.3, 1.05, 2048, 0.7, 20)
    assert os.path.isabs(constants.CHECKPOINT_DIR)
    assert all(hasattr(models, n) for n in ("GPT", "DVAE", "Tokenizer"))
    # the text front end and the zero-shot prompt hooks keep the reference's module paths and method names too
    from chattts_plus.commons import norm, text_utils
    assert callable(text_utils.split_text) and callable(text_utils.num2text) and callable(norm.Normalizer(None))
    assert list(inspect.signature(ChatTTSPlusPipeline.sample_audio_speaker).parameters) == ["self", "wav"]
    assert "mode" in inspect.signature(models.DVAE.__call__).parameters


def test_config_shim_matches_reference_keys():
    from omegaconf import OmegaConf
    cfg = OmegaConf.load(os.path.join(ROOT, "configs", "infer", "chattts_plus.yaml"))
    assert list(cfg.MODELS.keys()) == ["tokenizer", "dvae_encode", "dvae_decode", "vocos", "gpt"]
    assert cfg.MODELS.gpt.kwargs.gpt_config.num_hidden_layers == 20
    cfg.MODELS["dvae_decode"]["kwargs"]["coef"] = "x"      # mutation used by the pipeline (chattts_plus_pipeline.py:63-67)
    assert cfg.MODELS.dvae_decode.kwargs.coef == "x"
    assert "vocos" in cfg.MODELS


def test_gen_logits_and_flatten():
    from chatttsplus_b200.processors import flatten, gen_logits
    w, p = gen_logits(625, 0.7, 20, 1.05)
    sp = flatten(w, p)
    assert (sp.top_p, sp.top_k, sp.min_keep, sp.rep_penalty, sp.rep_window, sp.rep_max_ids) == (0.7, 20, 3, 1.05, 16, 625)
    # transformers' own warpers are accepted by duck typing
    from transformers.generation import TopKLogitsWarper, TopPLogitsWarper
    sp2 = flatten([TopPLogitsWarper(0.5, min_tokens_to_keep=3), TopKLogitsWarper(10, min_tokens_to_keep=3)], [])
    assert (sp2.top_p, sp2.top_k, sp2.rep_penalty) == (0.5, 10, 1.0)
    # no TopK warper (gen_logits(top_K=None), reference processors.py:43-47): top_k 0 = every rank may survive, TopP decides
    sp3 = flatten(*gen_logits(625, 0.7, None, 1.0))
    assert (sp3.top_p, sp3.top_k, sp3.min_keep, sp3.rep_penalty) == (0.7, 0, 3, 1.0)
    assert flatten([TopKLogitsWarper(50, min_tokens_to_keep=3)], []).top_k == 50


def test_tokenizer_layout_and_speaker_hook():
    """Left padding, [B, L, num_vq] expansion, audio-prompt tail and the speaker-embedding write (tokenizer.py:50-178)."""
    from chatttsplus_b200.tokenizer import Tokenizer, apply_spk_emb

    class FakeTok:
        vocab = {"[spk_emb]": 7, "[break_0]": 50, "[Ebreak]": 60}

        def __len__(self):
            return 100

        def convert_tokens_to_ids(self, t):
            return self.vocab.get(t, 1)

        def encode_plus(self, t, return_tensors="pt", add_special_tokens=False, padding=True):
            ids = torch.tensor([[ord(ch) % 90 + 8 for ch in t]])
            return {"input_ids": ids, "attention_mask": torch.ones_like(ids)}

        def batch_decode(self, x):
            return ["".join(chr(int(i)) for i in r) for r in x]

    tok = Tokenizer(tokenizer=FakeTok())
    prompt = Tokenizer._encode_prompt(torch.arange(12).view(4, 3))
    ids, mask, text_mask = tok.encode(["abcd", "xy"], 4, prompt_str=prompt)
    assert ids.shape == (2, 7, 4) and mask.shape == (2, 7)
    assert mask[1].tolist() == [0, 0, 1, 1, 1, 1, 1]                # left padding, prompt slots attended
    assert text_mask[0].tolist() == [True] * 4 + [False] * 3
    assert torch.equal(ids[0, 4:], torch.arange(12).view(4, 3).t())
    assert torch.equal(ids[0, :4, 0], ids[0, :4, 3])               # text ids replicated over the VQ columns
    emb = torch.zeros(1, 3, 8)
    iid = torch.tensor([[[1] * 4, [7] * 4, [2] * 4]])
    spk = torch.arange(8, dtype=torch.float32)
    apply_spk_emb(emb, spk, iid, 7)
    assert torch.allclose(emb[0, 1], torch.nn.functional.normalize(spk, dim=0)) and float(emb[0, 0].abs().sum()) == 0
    s = Tokenizer._encode_spk_emb(torch.randn(768))
    assert Tokenizer._decode_spk_emb(s).shape == (768,)


def test_omegaconf_shim_defers_to_an_installed_package(tmp_path):
    """the repo-root ``omegaconf`` must not shadow a real distribution.  A stand-in 'real' package later on sys.path is
    loaded in place of the shim (and its submodules resolve); without one the PyYAML stand-in answers."""
    import subprocess
    import sys
    real = tmp_path / "site" / "omegaconf"
    real.mkdir(parents=True)
    (real / "__init__.py").write_text("REAL = True\nclass OmegaConf:\n    @staticmethod\n    def load(p):\n        return 'real'\n")
    (real / "sub.py").write_text("X = 1\n")
    code = ("import sys; sys.path.insert(0, {root!r}); {extra}"
            "import omegaconf; print(getattr(omegaconf, 'REAL', False), type(omegaconf.OmegaConf.load({yaml!r})).__name__)")
    yaml = os.path.join(ROOT, "configs", "infer", "chattts_plus.yaml")
    lite = subprocess.run([sys.executable, "-c", code.format(root=ROOT, extra="", yaml=yaml)], capture_output=True, text=True, check=True).stdout
    assert lite.split() == ["False", "DictConfig"]
    both = subprocess.run([sys.executable, "-c", code.format(root=ROOT, extra=f"sys.path.append({str(tmp_path / 'site')!r}); ", yaml=yaml)
                           + "; import omegaconf.sub as s; print(s.X)"], capture_output=True, text=True, check=True).stdout
    assert both.split() == ["True", "str", "1"]


def test_synthetic_checkpoint_tokenizer_round_trip(tmp_path):
    """asset/tokenizer.pt as the reference stores it (a pickled BertTokenizerFast, tokenizer.py:27-31) loads through ``Tokenizer(model_path)``
    under the installed transformers (>= 5: no ``encode_plus``) and yields the left-padded [B, L, num_vq] layout of tokenizer.py:50-137."""
    import torch
    from chatttsplus_b200 import synth
    from chatttsplus_b200.tokenizer import Tokenizer
    text = "我们针对对话式任务进行了优化"
    p = tmp_path / "tokenizer.pt"
    torch.save(synth.make_bert_tokenizer([text]), p)
    tok = Tokenizer(str(p))
    assert tok.spk_emb_ids > 0 and tok.break_0_ids > tok._tokenizer.convert_tokens_to_ids("我") and tok.eos_token != tok.break_0_ids
    ids, mask, text_mask = tok.encode([f"[Stts][spk_emb][speed_3]{text} [uv_break][Ptts]", "[Stts][empty_spk]hi[Ptts]"], 4)
    assert ids.shape[0] == 2 and ids.shape[2] == 4 and mask.shape == ids.shape[:2]
    assert int(mask[1].sum()) < int(mask[0].sum()) and int(mask[1, 0]) == 0, "shorter text is LEFT padded"
    assert (ids[..., 0] == ids[..., 3]).all() and int((ids[0, :, 0] == tok.spk_emb_ids).sum()) == 1
    assert "[UNK]" not in tok.decode(ids[:1, :, 0])[0]
