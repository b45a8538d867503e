"""Generate golden vectors by running the REFERENCE's own code (from /root/reference) in this container.

    python tests/golden/make_golden.py          # writes tests/golden/*.pt (small fp32 tensors)

The reference's tests hold no golden vectors for the hot path, and /root/reference does not
exist on the GPU box, so the outputs of the reference itself are captured here once and committed:

  trunk_ref.pt        reference chattts_plus/models/llama.py  LlamaModel (SDPA + DynamicCache): prefill with
                      left padding + 3 cached decode steps        -> pins oracle.trunk_forward  (A6-A11)
  processors_ref.pt   reference chattts_plus/models/processors.py + transformers TopP/TopK warpers
                                                                  -> pins oracle.repetition_penalty/top_p/top_k (A15-A16)
  gpt_generate_ref.pt reference chattts_plus/models/gpt.py GPT.forward + GPT.generate (seeded torch.multinomial)
                                                                  -> pins oracle.gpt_embed / generate (A1,A3-A5,A13-A18)
  dvae_ref.pt         reference chattts_plus/models/dvae.py DVAE decode branch
                                                                  -> pins oracle.dvae_decode (A20-A22)
  text_ref.json       reference commons/text_utils.py + commons/norm.py on fixed samples  -> pins chatttsplus_b200/text.py (f4)
  dvae_encode_ref.pt  reference dvae.py MelSpectrogramFeatures + downsample_conv + encoder (encode branch up to the quantiser)
                                                                  -> pins oracle.mel_features / dvae_encode_features (f3)

Import shims (this script only; nothing of the reference is copied): stub parent packages so that
``models/__init__.py`` (which imports the absent pybase16384) is bypassed; stub modules ``pybase16384`` and
``vector_quantize_pytorch`` (import-time only); ``transformers.LogitsWarper`` alias (removed in transformers 5);
``DynamicCache.get_max_length / from_legacy_cache / to_legacy_cache`` re-added as trivial methods; the three
LlamaConfig attributes llama.py reads (rope_theta, rope_scaling, _attn_implementation).  Weights are the seeded
synthetic tensors of chatttsplus_b200.synth (no checkpoint exists offline).
"""
import importlib.util
import os
import sys
import types

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, "..", ".."))
sys.path.insert(0, ROOT)
REF = os.environ.get("CTP_REFERENCE_DIR", "/root/reference")
R = os.path.join(REF, "chattts_plus")

from chatttsplus_b200 import synth  # noqa: E402


def _load_reference():
    import transformers
    from transformers.cache_utils import DynamicCache

    if not hasattr(transformers, "LogitsWarper"):
        transformers.LogitsWarper = object
    if not hasattr(DynamicCache, "get_max_length"):
        DynamicCache.get_max_length = lambda self: None
    DynamicCache.from_legacy_cache = classmethod(lambda cls, past=None, *a, **k: cls())
    DynamicCache.to_legacy_cache = lambda self: self
    for name, path in [("chattts_plus", R), ("chattts_plus.models", R + "/models"),
                       ("chattts_plus.commons", R + "/commons")]:
        m = types.ModuleType(name)
        m.__path__ = [path]
        sys.modules[name] = m
    for stub in ("pybase16384",):
        sys.modules[stub] = types.ModuleType(stub)
    vq = types.ModuleType("vector_quantize_pytorch")
    vq.GroupedResidualFSQ = object
    sys.modules["vector_quantize_pytorch"] = vq

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(R, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod

    os.environ.setdefault("CHATTTS_PLUS_LOG_DIR", "/tmp/ctp_ref_logs")
    load("chattts_plus.commons.constants", "commons/constants.py")
    logger = load("chattts_plus.commons.logger", "commons/logger.py")
    sys.modules["chattts_plus.commons"].logger = logger
    llama = load("chattts_plus.models.llama", "models/llama.py")
    processors = load("chattts_plus.models.processors", "models/processors.py")
    gpt = load("chattts_plus.models.gpt", "models/gpt.py")
    dvae = load("chattts_plus.models.dvae", "models/dvae.py")
    return llama, processors, gpt, dvae


def _llama_cfg(c: synth.GPTConfig):
    from transformers import LlamaConfig
    cfg = LlamaConfig(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                      num_attention_heads=c.num_attention_heads, num_hidden_layers=c.num_hidden_layers,
                      max_position_embeddings=c.max_position_embeddings)
    cfg.__dict__["rope_theta"] = c.rope_theta
    cfg.rope_scaling = None
    cfg._attn_implementation = "sdpa"
    return cfg


SMALL = synth.GPTConfig(num_hidden_layers=2, num_text_tokens=96)


def gen_trunk(llama):
    from transformers.cache_utils import DynamicCache
    c = SMALL
    sd = synth.make_gpt_state(c, seed=11)
    trunk = llama.LlamaModel(_llama_cfg(c)).eval()
    del trunk.embed_tokens
    trunk.load_state_dict({k[len("gpt."):]: v for k, v in sd.items() if k.startswith("gpt.")}, strict=True)
    g = torch.Generator().manual_seed(5)
    B, L0, steps = 3, 7, 3
    x = torch.randn(B, L0, c.hidden_size, generator=g)
    mask = torch.ones(B, L0 + steps, dtype=torch.long)
    mask[1, :2] = 0  # left padding
    mask[2, :4] = 0
    xs = [torch.randn(B, 1, c.hidden_size, generator=g) for _ in range(steps)]
    outs = []
    with torch.no_grad():
        cache = DynamicCache()
        m = mask[:, :L0]
        pos = (m.cumsum(-1) - 1).masked_fill(m == 0, 1)
        o = trunk(inputs_embeds=x, attention_mask=m, position_ids=pos, past_key_values=cache, use_cache=True,
                  return_dict=True, cache_position=torch.arange(L0))
        outs.append(o.last_hidden_state.clone())
        for i in range(steps):
            m = mask[:, : L0 + i + 1]
            pos = (m.cumsum(-1) - 1).masked_fill(m == 0, 1)[:, -1:]
            o = trunk(inputs_embeds=xs[i], attention_mask=m, position_ids=pos, past_key_values=cache,
                      use_cache=True, return_dict=True, cache_position=torch.arange(L0 + i, L0 + i + 1))
            outs.append(o.last_hidden_state.clone())
    torch.save({"cfg": dict(num_hidden_layers=c.num_hidden_layers, num_text_tokens=c.num_text_tokens),
                "weight_seed": 11, "x": x, "xs": xs, "mask": mask, "outs": outs}, os.path.join(HERE, "trunk_ref.pt"))
    print("trunk_ref.pt", [tuple(o.shape) for o in outs])


def gen_processors(processors):
    g = torch.Generator().manual_seed(21)
    rows, A = 12, 626
    logits = torch.randn(rows, A, generator=g) * 3.0
    hist = torch.randint(0, A, (rows, 23), generator=g)
    hist[:, -6:] = hist[:, -12:-6]  # repeats inside the window
    warpers, procs = processors.gen_logits(num_code=625, top_P=0.7, top_K=20, repetition_penalty=1.05)
    out = {"logits": logits, "hist": hist}
    s = procs[0](hist, logits.clone())
    out["after_rep"] = s.clone()
    s2 = warpers[0](hist, s.clone())
    out["after_top_p"] = s2.clone()
    s3 = warpers[1](hist, s2.clone())
    out["after_top_k"] = s3.clone()
    # short history (< window) and the row-truncation quirk (rows > max_input_ids)
    proc_small = processors.CustomRepetitionPenaltyLogitsProcessorRepeat(1.2, 5, 16)
    out["short_hist"] = hist[:, :3].clone()
    out["after_rep_quirk"] = proc_small(out["short_hist"], logits.clone())
    torch.save(out, os.path.join(HERE, "processors_ref.pt"))
    print("processors_ref.pt kept:", int(torch.isfinite(s3).sum(-1).float().mean()))


def gen_gpt(gpt_mod, processors):
    c = SMALL
    sd = synth.make_gpt_state(c, seed=12)
    gcfg = dict(hidden_size=c.hidden_size, intermediate_size=c.intermediate_size,
                num_attention_heads=c.num_attention_heads, num_hidden_layers=c.num_hidden_layers,
                use_cache=False, max_position_embeddings=c.max_position_embeddings)
    # reference GPT._build_llama calls LlamaConfig(**config); patch the three attributes afterwards
    orig_build = gpt_mod.GPT._build_llama

    def build(self, config):
        from transformers import LlamaConfig
        lc = _llama_cfg(c)
        model = gpt_mod.LlamaModel(lc)
        del model.embed_tokens
        return model, lc
    gpt_mod.GPT._build_llama = build
    model = gpt_mod.GPT(gcfg, num_audio_tokens=c.num_audio_tokens, num_text_tokens=c.num_text_tokens,
                        num_vq=c.num_vq).eval()
    gpt_mod.GPT._build_llama = orig_build
    model.load_state_dict(sd, strict=True)

    g = torch.Generator().manual_seed(31)
    B, L0 = 3, 6
    ids1 = torch.randint(0, c.num_text_tokens, (B, L0, 1), generator=g)
    input_ids = ids1.expand(-1, -1, c.num_vq).clone()
    attn = torch.ones(B, L0, dtype=torch.long)
    attn[1, :2] = 0
    text_mask = attn.bool().clone()
    # last two positions of row 0..B are an "audio prompt" (text_mask False, per-vq ids)
    text_mask[:, -2:] = False
    input_ids[:, -2:] = torch.randint(0, c.num_audio_tokens - 1, (B, 2, c.num_vq), generator=g)
    with torch.no_grad():
        emb = model(input_ids, text_mask)
    res = {"weight_seed": 12, "input_ids": input_ids, "attention_mask": attn, "text_mask": text_mask, "emb": emb.clone()}
    warpers, procs = processors.gen_logits(num_code=625, top_P=0.7, top_K=20, repetition_penalty=1.05)
    for name, temp, maxn, minn in [("sampled", 0.3, 8, 3), ("neargreedy", 0.0003, 6, 0)]:
        torch.manual_seed(1234)
        out = next(model.generate(emb.clone(), input_ids.clone(), temperature=torch.tensor([temp] * c.num_vq),
                                  eos_token=625, attention_mask=attn, max_new_token=maxn, min_new_token=minn,
                                  logits_warpers=warpers, logits_processors=procs, infer_text=False,
                                  return_hidden=True, stream=False, show_tqdm=False, ensure_non_empty=True))
        res[name] = {"temperature": temp, "max_new_token": maxn, "min_new_token": minn,
                     "ids": [t.clone() for t in out.ids], "hiddens": [t.clone() for t in out.hiddens]}
        print("gpt_generate", name, [tuple(t.shape) for t in out.ids])
    # refine-text pass (infer_text=True): chattts_plus_pipeline.py:237-277 calls generate with head_text, one temperature
    warpers_t, procs_t = processors.gen_logits(num_code=c.num_text_tokens, top_P=0.7, top_K=20, repetition_penalty=1.0)
    torch.manual_seed(4321)
    out = next(model.generate(emb.clone(), input_ids.clone(), temperature=torch.tensor([0.7]), eos_token=60, attention_mask=attn,
                              max_new_token=9, min_new_token=0, logits_warpers=warpers_t, logits_processors=procs_t, infer_text=True,
                              stream=False, show_tqdm=False, ensure_non_empty=True))
    res["text"] = {"temperature": 0.7, "max_new_token": 9, "eos": 60, "ids": [t.clone() for t in out.ids]}
    print("gpt_generate text", [tuple(t.shape) for t in out.ids])
    torch.save(res, os.path.join(HERE, "gpt_generate_ref.pt"))


def gen_dvae(dvae_mod):
    cfg = synth.DVAEConfig(n_layer=3)
    sd = synth.make_dvae_state(cfg, seed=13)
    m = dvae_mod.DVAE(decoder_config=dict(idim=cfg.idim, odim=cfg.odim, hidden=cfg.hidden, n_layer=cfg.n_layer,
                                          bn_dim=cfg.bn_dim), dim=cfg.dim).eval()
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(41)
    x = torch.randn(1, 768, 9, generator=g)
    with torch.no_grad():
        mel = m(x.clone())
    torch.save({"weight_seed": 13, "n_layer": cfg.n_layer, "x": x, "mel": mel.clone()}, os.path.join(HERE, "dvae_ref.pt"))
    print("dvae_ref.pt", tuple(mel.shape))


def gen_dvae_encode(dvae_mod):
    """Zero-shot prompt encoder: the reference's own MelSpectrogramFeatures, downsample_conv and encoder
    modules chained as DVAE.forward(mode="encode") does (dvae.py:263-269).  The quantiser (vector_quantize_pytorch) is absent
    here, so the golden stops at the encoder output."""
    cfg = synth.DVAEConfig.codes_model(encoder=True, enc_layers=3)
    cfg.n_layer = 2
    sd = synth.make_dvae_state(cfg, seed=17)
    m = dvae_mod.DVAE(decoder_config=dict(idim=cfg.idim, odim=cfg.odim, hidden=cfg.hidden, n_layer=cfg.n_layer, bn_dim=cfg.bn_dim),
                      encoder_config=dict(idim=cfg.dim, odim=cfg.vq_dim, hidden=cfg.enc_hidden, n_layer=cfg.enc_layers, bn_dim=cfg.enc_bn),
                      vq_config=None, dim=cfg.dim).eval()
    missing = m.load_state_dict({k: v for k, v in sd.items() if not k.startswith("vq_layer.")}, strict=False)
    assert all(k.startswith("preprocessor_mel.") for k in missing.missing_keys), missing
    g = torch.Generator().manual_seed(43)
    audio = 0.1 * torch.randn(1, 256 * 37 + 91, generator=g)
    with torch.no_grad():
        mel = m.preprocessor_mel(audio.clone())
        x = m.downsample_conv(torch.div(mel, m.coef.view(1, 100, 1).expand(mel.shape)))
        x = m.encoder(x)
    torch.save({"weight_seed": 17, "n_layer": cfg.n_layer, "enc_layers": cfg.enc_layers, "audio": audio, "mel": mel.clone(), "x": x.clone()},
               os.path.join(HERE, "dvae_encode_ref.pt"))
    print("dvae_encode_ref.pt", tuple(mel.shape), tuple(x.shape))


TEXT_SAMPLES = [
    "Hello world, this is a test.", "I have 3 apples, 12 pears and 105 plums!", "The price is 1,234.56 dollars (50% off) [uv_break] ok",
    "2+3=5 and 7*8 = 56; 9-4", "1/2 of 3.5/7", "Call 911 now: it's urgent [laugh] really?", "x = 1000000 and y=20001 [lbreak]",
    "今天天气不错，我们去公园玩吧！", "他说: hello (world) 你好【测试】《书名》", "混合 text with 中文 and English words, 还有 numbers 42.",
    "emoji 😀 and symbols #$%^&*~ should go", "[break] leading tag and trailing tag [uv_break]", "", "   ",
    "A very long English sentence, " * 12 + "the end.", "这是一个很长的中文句子，" * 30 + "结束。",
    "Decimal 3.14159 stays, but sentence ends. Next one? Yes! Fine; good: ok) done} more… and more",
    "17 77 707 7 70", "999999999999999999999 is too long, 1234567890123456 is not", "tags [laugh][laugh] twice [uv_break][lbreak] end",
]


def gen_text():
    """Host text front-end (f4): the reference's own commons/text_utils.py and commons/norm.py functions on fixed samples.
    ``zh_normalization`` (absent) is stubbed at import time only; functions that need it (split_text on Chinese) are not recorded."""
    import json
    import tempfile
    zh = types.ModuleType("zh_normalization")
    zh.TextNormalizer = object
    sys.modules["zh_normalization"] = zh

    def load(modname, rel):
        spec = importlib.util.spec_from_file_location(modname, os.path.join(R, rel))
        mod = importlib.util.module_from_spec(spec)
        sys.modules[modname] = mod
        spec.loader.exec_module(mod)
        return mod
    tu = load("chattts_plus.commons.text_utils", "commons/text_utils.py")
    norm = load("chattts_plus.commons.norm", "commons/norm.py")
    out = {"num_to_english": {}, "num2text": {}, "remove_brackets": {}, "get_lang": {}, "split_by_punct": {}, "normalizer": []}
    nums = list(range(0, 130)) + [199, 200, 201, 210, 211, 999, 1000, 1001, 1010, 1011, 1100, 2024, 9999, 10000, 10001, 10011, 12345, 100000,
                                 100200, 1000000, 1000001, 1002003, 20001, 1234567, 999999999, 1000000000, 123456789012, 1000000000000, 1234567890123456]
    for n in nums:
        try:
            out["num_to_english"][str(n)] = tu.num_to_english(n)
        except Exception as e:  # the reference raises IndexError for ...10
            out["num_to_english"][str(n)] = "!" + type(e).__name__
    for t in TEXT_SAMPLES:
        try:
            out["num2text"][t] = tu.num2text(t)
        except Exception as e:
            out["num2text"][t] = "!" + type(e).__name__
        out["remove_brackets"][t] = tu.remove_brackets(t)
        out["get_lang"][t] = tu.get_lang(t)
        out["split_by_punct"][t] = tu.split_text_by_punctuation(t)
    hm = {"好": "郝", "天": "添", "书": "叔", "a": "b"}
    with tempfile.TemporaryDirectory() as d:
        mp = os.path.join(d, "homophones_map.json")
        with open(mp, "w", encoding="utf-8") as f:
            json.dump(hm, f, ensure_ascii=False)
        nz = norm.Normalizer(mp)
        nz.register("en", lambda s: s.replace("test", "TEST"))
        for t in TEXT_SAMPLES:
            for tn in (True, False):
                for hr in (True, False):
                    for lang in (None, "zh", "en"):
                        out["normalizer"].append({"text": t, "tn": tn, "hr": hr, "lang": lang, "out": nz(t, tn, hr, lang)})
    out["homophones"] = hm
    out["tables"] = {"simplify": {chr(k): v for k, v in nz.character_simplifier.items()},
                     "half2full": {chr(k): v for k, v in nz.halfwidth_2_fullwidth.items()}}
    with open(os.path.join(HERE, "text_ref.json"), "w", encoding="utf-8") as f:
        json.dump(out, f, ensure_ascii=False, indent=0)
    print("text_ref.json", len(out["num_to_english"]), len(out["normalizer"]))


if __name__ == "__main__":
    torch.set_num_threads(8)
    llama, processors, gpt_mod, dvae_mod = _load_reference()
    gen_trunk(llama)
    gen_processors(processors)
    gen_gpt(gpt_mod, processors)
    gen_dvae(dvae_mod)
    gen_dvae_encode(dvae_mod)
    gen_text()
