"""The oracle is pinned against outputs of the REFERENCE's own code (tests/golden/make_golden.py)."""
import os

import pytest
import torch

from chatttsplus_b200 import synth
from oracle import ctp_oracle as O


def _load(golden_dir, name):
    return torch.load(os.path.join(golden_dir, name), weights_only=False)


def test_trunk_matches_reference_llama(golden_dir):
    """oracle.trunk_forward == reference chattts_plus/models/llama.py LlamaModel (prefill w/ left pad + decode)."""
    g = _load(golden_dir, "trunk_ref.pt")
    cfg = synth.GPTConfig(**g["cfg"])
    sd = synth.make_gpt_state(cfg, seed=g["weight_seed"])
    mask = g["mask"]
    L0 = g["x"].shape[1]
    cache = O.KVCache.empty(cfg.num_hidden_layers)
    m = mask[:, :L0]
    out = O.trunk_forward(sd, g["x"], m, O.position_ids_from_mask(m), cache, cfg.num_hidden_layers,
                          cfg.num_attention_heads)
    valid = m.bool()
    assert torch.allclose(out[valid], g["outs"][0][valid], atol=2e-5, rtol=1e-5)
    for i, x in enumerate(g["xs"]):
        m = mask[:, : L0 + i + 1]
        pos = O.position_ids_from_mask(m)[:, -1:]
        out = O.trunk_forward(sd, x, m, pos, cache, cfg.num_hidden_layers, cfg.num_attention_heads)
        assert torch.allclose(out, g["outs"][i + 1], atol=2e-5, rtol=1e-5), i


def test_processors_match_reference(golden_dir):
    g = _load(golden_dir, "processors_ref.pt")
    s = O.repetition_penalty(g["hist"], g["logits"].clone(), 1.05, 625, 16)
    assert torch.equal(s, g["after_rep"])
    s2 = O.top_p_warp(s, 0.7, 3)
    assert torch.equal(s2, g["after_top_p"])
    s3 = O.top_k_warp(s2, 20, 3)
    assert torch.equal(s3, g["after_top_k"])
    q = O.repetition_penalty(g["short_hist"], g["logits"].clone(), 1.2, 5, 16)
    assert torch.equal(q, g["after_rep_quirk"])


def test_live_transformers_warpers_agree():
    """Same image on the GPU box: the installed TopP/TopK warpers equal the restatement."""
    from transformers.generation import TopKLogitsWarper, TopPLogitsWarper
    g = torch.Generator().manual_seed(3)
    s = torch.randn(16, 626, generator=g) * 4
    assert torch.equal(TopPLogitsWarper(0.7, min_tokens_to_keep=3)(None, s.clone()), O.top_p_warp(s, 0.7, 3))
    assert torch.equal(TopKLogitsWarper(20, min_tokens_to_keep=3)(None, s.clone()), O.top_k_warp(s, 20, 3))


@pytest.mark.parametrize("case", ["sampled", "neargreedy"])
def test_generate_matches_reference_gpt(golden_dir, case):
    """oracle.gpt_embed + oracle.generate(sampler='torch') reproduce reference GPT.forward/GPT.generate
    (same seed -> same torch.multinomial stream -> identical ids; hiddens to fp32 round-off)."""
    g = _load(golden_dir, "gpt_generate_ref.pt")
    cfg = synth.GPTConfig(num_hidden_layers=2, num_text_tokens=96)
    sd = synth.make_gpt_state(cfg, seed=g["weight_seed"])
    emb = O.gpt_embed(sd, g["input_ids"], g["text_mask"], cfg.num_vq)
    assert torch.allclose(emb, g["emb"], atol=1e-6)
    c = g[case]
    torch.manual_seed(1234)
    r = O.generate(sd, emb, g["input_ids"], torch.tensor([c["temperature"]] * cfg.num_vq), 625,
                   g["attention_mask"], n_layers=cfg.num_hidden_layers, n_heads=cfg.num_attention_heads,
                   max_new_token=c["max_new_token"], min_new_token=c["min_new_token"], sampler="torch")
    for a, b in zip(r.ids, c["ids"]):
        assert torch.equal(a, b)
    for a, b in zip(r.hiddens, c["hiddens"]):
        assert torch.allclose(a, b, atol=3e-5, rtol=1e-5)


def test_dvae_decode_matches_reference(golden_dir):
    g = _load(golden_dir, "dvae_ref.pt")
    cfg = synth.DVAEConfig(n_layer=g["n_layer"])
    sd = synth.make_dvae_state(cfg, seed=g["weight_seed"])
    mel = O.dvae_decode(sd, g["x"], n_layer=cfg.n_layer)
    assert mel.shape == g["mel"].shape
    assert torch.allclose(mel, g["mel"], atol=2e-5, rtol=1e-5)


def test_vocos_istft_matches_torch():
    """The ISTFT inside oracle.vocos_decode is torch.istft itself; sanity: output length hop*(T-1)."""
    cfg = synth.VocosConfig(num_layers=2)
    sd = synth.make_vocos_state(cfg, seed=5)
    mel = torch.randn(1, 100, 12, generator=torch.Generator().manual_seed(1))
    wav = O.vocos_decode(sd, mel, num_layers=2)
    assert wav.shape == (1, 256 * 11)
    assert torch.isfinite(wav).all()


def test_gfsq_embed_shapes_and_digits():
    cfg = synth.DVAEConfig.codes_model()
    sd = synth.make_dvae_state(cfg, seed=3)
    ids = torch.tensor([[[0], [624], [312], [7]]])  # [B=1, G*R=4, T=1]
    f = O.gfsq_embed(sd, ids)
    assert f.shape == (1, 1024, 1)
    # idx 0 -> all digits 0 -> code -1; idx 624 -> all digits 4 -> code +1, residual level scaled by 1/4
    w = sd["vq_layer.quantizer.rvqs.0.project_out.weight"]
    b = sd["vq_layer.quantizer.rvqs.0.project_out.bias"]
    exp = torch.nn.functional.linear(torch.full((4,), -1.0) + 0.25 * torch.full((4,), 1.0), w, b)
    assert torch.allclose(f[0, :512, 0], exp, atol=1e-6)


def test_refine_text_pass_matches_reference_gpt(golden_dir):
    """infer_text=True (scope row f1): oracle.generate reproduces the reference's text-token pass."""
    g = _load(golden_dir, "gpt_generate_ref.pt")
    cfg = synth.GPTConfig(num_hidden_layers=2, num_text_tokens=96)
    sd = synth.make_gpt_state(cfg, seed=g["weight_seed"])
    emb = O.gpt_embed(sd, g["input_ids"], g["text_mask"], cfg.num_vq)
    c = g["text"]
    torch.manual_seed(4321)
    r = O.generate(sd, emb, g["input_ids"], torch.tensor([c["temperature"]]), c["eos"], g["attention_mask"],
                   n_layers=cfg.num_hidden_layers, n_heads=cfg.num_attention_heads, max_new_token=c["max_new_token"], min_new_token=0,
                   rep_penalty=None, sampler="torch", infer_text=True)
    for a, b in zip(r.ids, c["ids"]):
        assert a.shape == b.shape and torch.equal(a, b)


def test_dvae_encode_front_end_matches_reference_golden(golden_dir):
    """Zero-shot prompt encoder (f3): mel features and the encoder output equal what the reference's own MelSpectrogramFeatures /
    downsample_conv / encoder modules produced (tests/golden/make_golden.py::gen_dvae_encode)."""
    ref = _load(golden_dir, "dvae_encode_ref.pt")
    cfg = synth.DVAEConfig.codes_model(encoder=True, enc_layers=ref["enc_layers"])
    cfg.n_layer = ref["n_layer"]
    sd = synth.make_dvae_state(cfg, seed=ref["weight_seed"])
    mel = O.mel_features(ref["audio"])
    assert mel.shape == ref["mel"].shape and float((mel - ref["mel"]).abs().max()) < 1e-4
    x = O.dvae_encode_features(sd, ref["audio"], n_layer=ref["enc_layers"])
    assert x.shape == ref["x"].shape and float((x - ref["x"]).abs().max()) < 1e-4


def test_gfsq_quantize_is_consistent_with_embed():
    """The quantiser restatement (unpinned: vector_quantize_pytorch is absent) and the pinned-by-formula embed (A23) agree:
    embed(quantize(x)) == project_out(sum_r q_r * 4^-r) and every index is a valid code."""
    cfg = synth.DVAEConfig.codes_model(encoder=True, enc_layers=1)
    sd = synth.make_dvae_state(cfg, seed=3)
    g = torch.Generator().manual_seed(5)
    x = 3.0 * torch.randn(2, 1024, 11, generator=g)
    ids = O.gfsq_quantize(sd, x)
    assert ids.shape == (2, 4, 11) and int(ids.min()) >= 0 and int(ids.max()) < 625
    feat = O.gfsq_embed(sd, ids)
    # recompute the quantised vector directly
    xt = x.transpose(1, 2)
    outs = []
    for gi in range(2):
        z = torch.nn.functional.linear(xt[..., gi * 512:(gi + 1) * 512], sd[f"vq_layer.quantizer.rvqs.{gi}.project_in.weight"],
                                       sd[f"vq_layer.quantizer.rvqs.{gi}.project_in.bias"])
        res = O.fsq_bound(z, (5, 5, 5, 5))
        q_sum = torch.zeros_like(z)
        for r in range(2):
            sc = 4.0 ** (-r)
            q = torch.round(O.fsq_bound(res / sc, (5, 5, 5, 5))) / 2
            q_sum = q_sum + q * sc
            res = res - q * sc
        outs.append(torch.nn.functional.linear(q_sum, sd[f"vq_layer.quantizer.rvqs.{gi}.project_out.weight"],
                                               sd[f"vq_layer.quantizer.rvqs.{gi}.project_out.bias"]))
    direct = torch.cat(outs, -1).transpose(1, 2)
    assert float((feat - direct).abs().max()) < 1e-4
