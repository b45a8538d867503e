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


# ---- third-party arithmetic that has no reference-produced fixture (vector_quantize_pytorch, vocos, peft are absent):
# ---- hand-derived known-answer cases against the published closed forms ------------------------------------------------------

def _identity_vq_state():
    """GFSQ projections that expose the 4-d code space: project_out = [I4; 0] (no bias), project_in = [I4 | 0]."""
    sd = {}
    for g in range(2):
        wo = torch.zeros(512, 4); wo[:4] = torch.eye(4)
        wi = torch.zeros(4, 512); wi[:, :4] = torch.eye(4)
        sd[f"vq_layer.quantizer.rvqs.{g}.project_out.weight"] = wo
        sd[f"vq_layer.quantizer.rvqs.{g}.project_out.bias"] = torch.zeros(512)
        sd[f"vq_layer.quantizer.rvqs.{g}.project_in.weight"] = wi
        sd[f"vq_layer.quantizer.rvqs.{g}.project_in.bias"] = torch.zeros(4)
    return sd


def test_gfsq_codebook_closed_form_all_625_codes():
    """FSQ (Mentzer et al. 2023, "Finite Scalar Quantization", sec. 3; vector_quantize_pytorch FSQ.indices_to_codes): with levels
    [5,5,5,5] index i has digits d_j = (i // 5^j) % 5 and code value (d_j - 2) / 2; ResidualFSQ scales residual level r by 4^-r.
    Enumerates every index at both residual levels and in both groups (dvae.py:84-94)."""
    sd = _identity_vq_state()
    idx = torch.arange(625)
    for g in range(2):
        for r in range(2):
            ids = torch.zeros(1, 4, 625, dtype=torch.long)
            ids[0, g * 2 + (1 - r)] = 312          # the other level of this group: all digits 2 -> code 0
            other = 1 - g
            ids[0, other * 2] = 312; ids[0, other * 2 + 1] = 312
            ids[0, g * 2 + r] = idx
            f = O.gfsq_embed(sd, ids)              # [1, 1024, 625]
            got = f[0, g * 512: g * 512 + 4].t()   # [625, 4]
            exp = torch.stack([((idx // 5 ** j) % 5 - 2).float() / 2 for j in range(4)], 1) * (4.0 ** -r)
            assert torch.equal(got, exp), (g, r)
            assert float(f[0, (1 - g) * 512: (1 - g) * 512 + 4].abs().max()) == 0.0


def test_fsq_quantiser_thresholds_closed_form():
    """ResidualFSQ.forward (vector_quantize_pytorch 1.17.8): residual = bound(project_in(x)); per level FSQ.quantize =
    round(bound(residual / s_r)) / (L // 2) with bound(z) = tanh(z) * (L-1)(1+eps)/2, L = 5, eps = 1e-3.  The value is therefore
    bounded TWICE before the first rounding: the level-0 decision boundaries sit at z = atanh(atanh((k + 1/2) / 2.002) / 2.002),
    k in {-2..1}; the index packs digit_j = 2 q_j + 2 with basis 5^j."""
    import math
    sd = _identity_vq_state()
    half_l = 4 * (1 + 1e-3) / 2
    for k in (-2, -1, 0, 1):
        zb = math.atanh(math.atanh((k + 0.5) / half_l) / half_l)
        for dz, digit in ((-1e-4, k + 2), (1e-4, k + 3)):
            x = torch.zeros(1, 1024, 1)
            x[0, 0, 0] = zb + dz                   # group 0, code dimension 0; everything else 0 -> digit 2
            ids = O.gfsq_quantize(sd, x)
            assert int(ids[0, 0, 0]) == digit + 2 * 5 + 2 * 25 + 2 * 125, (k, dz, int(ids[0, 0, 0]))
            assert int(ids[0, 2, 0]) == 312        # the other group saw zeros
    # second residual level: what is left after level 0, divided by 1/4, goes through the same rule
    x = torch.zeros(1, 1024, 1)
    x[0, 1, 0] = 0.3
    ids = O.gfsq_quantize(sd, x)
    res = math.tanh(0.3) * half_l                  # .5832
    q0 = round(math.tanh(res) * half_l) / 2        # tanh(.5832)*2.002 = 1.051 -> 1 -> q0 = .5
    assert q0 == 0.5 and int(ids[0, 0, 0]) == 312 + 1 * 5          # digit 3 in dimension 1
    d1 = round(math.tanh((res - q0) / 0.25) * half_l) + 2
    assert int(ids[0, 1, 0]) == 312 + (d1 - 2) * 5


def test_lora_merge_known_answer():
    """peft LoRA merge (peft/tuners/lora/layer.py Linear.get_delta_weight: B @ A * scaling, scaling = lora_alpha / r, or
    lora_alpha / sqrt(r) with use_rslora) on a hand-computable case, for an attention and an MLP target."""
    p = {"gpt.layers.0.self_attn.q_proj.weight": torch.zeros(2, 3), "gpt.layers.0.mlp.down_proj.weight": torch.ones(2, 3)}
    A = torch.tensor([[1.0, 2.0, 3.0]])            # r = 1
    B = torch.tensor([[1.0], [2.0]])
    lora = {"base_model.model.layers.0.self_attn.q_proj.lora_A.weight": A, "base_model.model.layers.0.self_attn.q_proj.lora_B.weight": B,
            "base_model.model.layers.0.mlp.down_proj.lora_A.weight": A, "base_model.model.layers.0.mlp.down_proj.lora_B.weight": B}
    m = O.lora_merge(p, lora, 1, alpha=2.0, r=1)
    assert torch.equal(m["gpt.layers.0.self_attn.q_proj.weight"], torch.tensor([[2.0, 4.0, 6.0], [4.0, 8.0, 12.0]]))
    assert torch.equal(m["gpt.layers.0.mlp.down_proj.weight"], torch.tensor([[3.0, 5.0, 7.0], [5.0, 9.0, 13.0]]))
    A4 = torch.cat([A, torch.zeros(3, 3)])         # r = 4 with three dead ranks: rsLoRA scaling = alpha / 2
    B4 = torch.cat([B, torch.zeros(2, 3)], 1)
    lora4 = {"base_model.model.layers.0.self_attn.q_proj.lora_A.weight": A4, "base_model.model.layers.0.self_attn.q_proj.lora_B.weight": B4}
    m4 = O.lora_merge(p, lora4, 1, alpha=2.0, r=4, use_rslora=True)
    assert torch.equal(m4["gpt.layers.0.self_attn.q_proj.weight"], torch.tensor([[1.0, 2.0, 3.0], [2.0, 4.0, 6.0]]))


def _tone_head_state(k0: int, log_mag: float):
    """ISTFTHead weights that emit one spectral line: |S[k0]| = exp(log_mag), phase(k0, frame f) = (pi/2) k0 f - pi k0 (what a
    stationary cosine of bin k0 has under hop = n_fft / 4 when frame f starts at sample f*hop - n_fft/2); input row f is [f, 0, ...]."""
    import math
    sd = {"head.out.weight": torch.zeros(1026, 512), "head.out.bias": torch.zeros(1026), "head.istft.window": torch.hann_window(1024, periodic=True)}
    sd["head.out.bias"][:513] = -40.0              # exp(-40): every other line is silent
    sd["head.out.bias"][k0] = log_mag
    sd["head.out.weight"][513 + k0, 0] = (math.pi / 2) * k0
    sd["head.out.bias"][513 + k0] = -math.pi * (k0 % 2)
    return sd


def test_vocos_head_and_istft_pure_tone():
    """vocos ISTFTHead + ISTFT(padding="center") on an analytic input: ONE spectral line S[k0, f] = c e^{i 2 pi k0 (hop f - N/2) / N}.
    Every frame's inverse real FFT is (2c/N) cos(2 pi k0 n / N); istft multiplies by the periodic Hann window w, overlap-adds at
    hop = N/4 and divides by the envelope sum w^2: with sum_f w = 2 and sum_f w^2 = 3/2 the output is (2c/N)(4/3) cos(2 pi k0 n / N).
    torch.istft(center=True) trims N/2 and returns hop * (T - 1) samples."""
    import math
    k0, T = 37, 20
    c = 64.0                                              # below the 1e2 clip
    amp = (2 * c / 1024) * (4.0 / 3.0)
    sd = _tone_head_state(k0, math.log(c))
    x = torch.zeros(1, T, 512)
    x[0, :, 0] = torch.arange(T, dtype=torch.float32)
    S = O.vocos_head_spec(sd, x)
    assert S.shape == (1, 513, T)
    wav = torch.istft(S, 1024, 256, 1024, sd["head.istft.window"], center=True)
    assert wav.shape == (1, 256 * (T - 1))         # centre padding length
    n = torch.arange(wav.shape[1], dtype=torch.float64)
    ref = amp * torch.cos(2 * math.pi * k0 * n / 1024)
    assert float((wav[0].double() - ref)[512:-512].abs().max()) < 1e-4   # interior: full window overlap
    # magnitude clip: exp(log 1e3) is clipped at 1e2 (ISTFTHead: torch.clip(mag, max=1e2))
    sd2 = _tone_head_state(k0, math.log(1e3))
    S2 = O.vocos_head_spec(sd2, x)
    assert abs(float(S2[0, k0, 0].abs()) - 100.0) < 1e-3
