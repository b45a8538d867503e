"""Seeded synthetic checkpoints with the reference's state-dict key layout.

No ChatTTS checkpoint exists offline (``checkpoints/`` is git-ignored and the pipeline
downloads from the hub at first run), so parity tests, ``smoke()`` and ``bench.py`` all run on weights made
here: real shapes, reference key names (SURVEY.md appendix A.3 = the module structure of
chattts_plus/models/gpt.py:43-77, dvae.py:129-168,203-241 and the vocos config in
configs/infer/chattts_plus.yaml:46-66), ``torch.Generator``-seeded so every box regenerates the same tensors.
The same dicts are bound to the CUDA kernels and to the oracle.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Dict, Optional

import torch


@dataclass
class GPTConfig:
    hidden_size: int = 768
    intermediate_size: int = 3072
    num_attention_heads: int = 12
    num_hidden_layers: int = 20
    num_audio_tokens: int = 626
    num_text_tokens: int = 21178
    num_vq: int = 4
    rms_norm_eps: float = 1e-6
    rope_theta: float = 10000.0
    max_position_embeddings: int = 4096

    @property
    def head_dim(self) -> int:
        return self.hidden_size // self.num_attention_heads


@dataclass
class DVAEConfig:
    """Decoder side of the DVAE. Defaults = ``dvae_decode`` (Decoder.pt) in configs/infer/chattts_plus.yaml."""
    dim: int = 384            # out_conv input channels == decoder odim
    idim: int = 384
    odim: int = 384
    hidden: int = 512
    n_layer: int = 12
    bn_dim: int = 128
    kernel: int = 7
    dilation: int = 2
    n_mels: int = 100
    # GFSQ (only for the codes->mel model, ``dvae_encode`` / DVAE_full.pt)
    vq: bool = False
    vq_dim: int = 1024
    vq_levels: tuple = (5, 5, 5, 5)
    vq_G: int = 2
    vq_R: int = 2
    # zero-shot prompt encoder (``encoder_config`` of ``dvae_encode``): DVAEDecoder(idim=dim, odim=vq_dim)
    encoder: bool = False
    enc_hidden: int = 256
    enc_layers: int = 12
    enc_bn: int = 128
    enc_odim: int = 1024

    @staticmethod
    def codes_model(encoder: bool = False, enc_layers: int = 12) -> "DVAEConfig":
        return DVAEConfig(dim=512, idim=512, odim=512, hidden=256, n_layer=12, bn_dim=128, vq=True, encoder=encoder, enc_layers=enc_layers)


@dataclass
class VocosConfig:
    input_channels: int = 100
    dim: int = 512
    intermediate_dim: int = 1536
    num_layers: int = 8
    n_fft: int = 1024
    hop_length: int = 256


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    return g


def _n(g, *shape, std=1.0, mean=0.0):
    return torch.randn(*shape, generator=g, dtype=torch.float32) * std + mean


def _u(g, *shape, lo=0.0, hi=1.0):
    return torch.rand(*shape, generator=g, dtype=torch.float32) * (hi - lo) + lo


def make_gpt_state(cfg: GPTConfig = GPTConfig(), seed: int = 1234, head_gain: float = 4.0) -> Dict[str, torch.Tensor]:
    """``GPT.pt``-shaped state dict (keys as loaded strictly by reference gpt.py:84-85)."""
    g = _gen(seed)
    H, I = cfg.hidden_size, cfg.intermediate_size
    sd: Dict[str, torch.Tensor] = {}
    for l in range(cfg.num_hidden_layers):
        p = f"gpt.layers.{l}."
        sd[p + "input_layernorm.weight"] = 1.0 + _n(g, H, std=0.1)
        sd[p + "post_attention_layernorm.weight"] = 1.0 + _n(g, H, std=0.1)
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            sd[p + f"self_attn.{nm}.weight"] = _n(g, H, H, std=0.02)
        # keys sharper than HF init so attention is not uniform (exercises the softmax)
        sd[p + "self_attn.q_proj.weight"] *= 3.0
        sd[p + "self_attn.k_proj.weight"] *= 3.0
        sd[p + "mlp.gate_proj.weight"] = _n(g, I, H, std=0.02)
        sd[p + "mlp.up_proj.weight"] = _n(g, I, H, std=0.02)
        sd[p + "mlp.down_proj.weight"] = _n(g, H, I, std=0.02)
    sd["gpt.norm.weight"] = 1.0 + _n(g, H, std=0.1)
    for q in range(cfg.num_vq):
        sd[f"emb_code.{q}.weight"] = _n(g, cfg.num_audio_tokens, H, std=0.5)
    sd["emb_text.weight"] = _n(g, cfg.num_text_tokens, H, std=1.0)
    v = _n(g, cfg.num_text_tokens, H, std=0.02)
    sd["head_text.parametrizations.weight.original0"] = v.norm(dim=1, keepdim=True) * head_gain
    sd["head_text.parametrizations.weight.original1"] = v
    for q in range(cfg.num_vq):
        v = _n(g, cfg.num_audio_tokens, H, std=0.02)
        sd[f"head_code.{q}.parametrizations.weight.original0"] = v.norm(dim=1, keepdim=True) * head_gain
        sd[f"head_code.{q}.parametrizations.weight.original1"] = v
    return sd


def _convnext(sd, g, prefix, dim, inter, kernel, gamma_lo, gamma_hi):
    sd[prefix + "dwconv.weight"] = _n(g, dim, 1, kernel, std=kernel ** -0.5)
    sd[prefix + "dwconv.bias"] = _n(g, dim, std=0.05)
    sd[prefix + "norm.weight"] = 1.0 + _n(g, dim, std=0.1)
    sd[prefix + "norm.bias"] = _n(g, dim, std=0.05)
    sd[prefix + "pwconv1.weight"] = _n(g, inter, dim, std=dim ** -0.5)
    sd[prefix + "pwconv1.bias"] = _n(g, inter, std=0.05)
    sd[prefix + "pwconv2.weight"] = _n(g, dim, inter, std=inter ** -0.5)
    sd[prefix + "pwconv2.bias"] = _n(g, dim, std=0.05)
    sd[prefix + "gamma"] = _u(g, dim, lo=gamma_lo, hi=gamma_hi)


def make_dvae_state(cfg: DVAEConfig = DVAEConfig(), seed: int = 4321) -> Dict[str, torch.Tensor]:
    """``Decoder.pt`` / decode-side ``DVAE_full.pt`` state dict (reference dvae.py:129-168,203-241)."""
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    sd["coef"] = _u(g, 1, cfg.n_mels, 1, lo=0.5, hi=1.5)
    sd["decoder.conv_in.0.weight"] = _n(g, cfg.bn_dim, cfg.idim, 3, std=(3 * cfg.idim) ** -0.5)
    sd["decoder.conv_in.0.bias"] = _n(g, cfg.bn_dim, std=0.05)
    sd["decoder.conv_in.2.weight"] = _n(g, cfg.hidden, cfg.bn_dim, 3, std=(3 * cfg.bn_dim) ** -0.5)
    sd["decoder.conv_in.2.bias"] = _n(g, cfg.hidden, std=0.05)
    for l in range(cfg.n_layer):
        _convnext(sd, g, f"decoder.decoder_block.{l}.", cfg.hidden, cfg.hidden * 4, cfg.kernel, 0.05, 0.3)
    sd["decoder.conv_out.weight"] = _n(g, cfg.odim, cfg.hidden, 1, std=cfg.hidden ** -0.5)
    sd["out_conv.weight"] = _n(g, cfg.n_mels, cfg.dim, 3, std=(3 * cfg.dim) ** -0.5)
    if getattr(cfg, "encoder", False):
        # zero-shot prompt encoder (dvae.py:224-233): downsample_conv + a DVAEDecoder-shaped encoder (idim dim -> odim vq_dim)
        sd["downsample_conv.0.weight"] = _n(g, cfg.dim, cfg.n_mels, 3, std=(3 * cfg.n_mels) ** -0.5 / 4)   # mel features are O(10)
        sd["downsample_conv.0.bias"] = _n(g, cfg.dim, std=0.05)
        sd["downsample_conv.2.weight"] = _n(g, cfg.dim, cfg.dim, 4, std=(4 * cfg.dim) ** -0.5)
        sd["downsample_conv.2.bias"] = _n(g, cfg.dim, std=0.05)
        eh, ebn = cfg.enc_hidden, cfg.enc_bn
        sd["encoder.conv_in.0.weight"] = _n(g, ebn, cfg.dim, 3, std=(3 * cfg.dim) ** -0.5)
        sd["encoder.conv_in.0.bias"] = _n(g, ebn, std=0.05)
        sd["encoder.conv_in.2.weight"] = _n(g, eh, ebn, 3, std=(3 * ebn) ** -0.5)
        sd["encoder.conv_in.2.bias"] = _n(g, eh, std=0.05)
        for l in range(cfg.enc_layers):
            _convnext(sd, g, f"encoder.decoder_block.{l}.", eh, eh * 4, cfg.kernel, 0.05, 0.3)
        sd["encoder.conv_out.weight"] = _n(g, cfg.vq_dim, eh, 1, std=eh ** -0.5)
    if cfg.vq:
        # GroupedResidualFSQ: one ResidualFSQ per group, each with project_in (dim/G -> len(levels)) and
        # project_out (len(levels) -> dim/G); only project_out is used by get_output_from_indices.
        gd = cfg.vq_dim // cfg.vq_G
        nl = len(cfg.vq_levels)
        for gi in range(cfg.vq_G):
            sd[f"vq_layer.quantizer.rvqs.{gi}.project_in.weight"] = _n(g, nl, gd, std=gd ** -0.5)
            sd[f"vq_layer.quantizer.rvqs.{gi}.project_in.bias"] = _n(g, nl, std=0.05)
            sd[f"vq_layer.quantizer.rvqs.{gi}.project_out.weight"] = _n(g, gd, nl, std=0.5)
            sd[f"vq_layer.quantizer.rvqs.{gi}.project_out.bias"] = _n(g, gd, std=0.05)
    return sd


def make_vocos_state(cfg: VocosConfig = VocosConfig(), seed: int = 9876) -> Dict[str, torch.Tensor]:
    """``Vocos.pt``-shaped state dict (vocos package module names: backbone.embed / norm / convnext.N /
    final_layer_norm, head.out, head.istft.window)."""
    g = _gen(seed)
    sd: Dict[str, torch.Tensor] = {}
    C, D = cfg.input_channels, cfg.dim
    sd["backbone.embed.weight"] = _n(g, D, C, 7, std=(7 * C) ** -0.5)
    sd["backbone.embed.bias"] = _n(g, D, std=0.05)
    sd["backbone.norm.weight"] = 1.0 + _n(g, D, std=0.1)
    sd["backbone.norm.bias"] = _n(g, D, std=0.05)
    for l in range(cfg.num_layers):
        _convnext(sd, g, f"backbone.convnext.{l}.", D, cfg.intermediate_dim, 7, 0.05, 0.25)
    sd["backbone.final_layer_norm.weight"] = 1.0 + _n(g, D, std=0.1)
    sd["backbone.final_layer_norm.bias"] = _n(g, D, std=0.05)
    sd["head.out.weight"] = _n(g, cfg.n_fft + 2, D, std=0.6 * D ** -0.5)
    sd["head.out.bias"] = _n(g, cfg.n_fft + 2, std=0.3)
    # log-magnitudes centred so that |S| ~ 1e-2 .. 1 (waveform amplitude O(0.1), like real speech)
    sd["head.out.bias"][: cfg.n_fft // 2 + 1] -= 2.5
    sd["head.istft.window"] = torch.hann_window(cfg.n_fft, periodic=True, dtype=torch.float32)
    return sd


def make_lora_state(cfg: GPTConfig = GPTConfig(), r: int = 8, seed: int = 777, mlp: bool = False) -> Dict[str, torch.Tensor]:
    """peft-adapter-shaped tensors (keys as written by peft ``save_pretrained`` for a LlamaModel wrapped in
    PeftModel: base_model.model.layers.N.self_attn.{q,k,v,o}_proj.lora_{A,B}.weight; ``mlp=True`` adds the
    mlp.{gate,up,down}_proj targets configs/train/train_voice_clone_lora.yaml lists as options)."""
    g = _gen(seed)
    H, I = cfg.hidden_size, cfg.intermediate_size
    sd = {}
    for l in range(cfg.num_hidden_layers):
        for nm in ("q_proj", "k_proj", "v_proj", "o_proj"):
            p = f"base_model.model.layers.{l}.self_attn.{nm}."
            sd[p + "lora_A.weight"] = _n(g, r, H, std=0.02)
            sd[p + "lora_B.weight"] = _n(g, H, r, std=0.02)
        if mlp:
            for nm, (fin, fout) in (("gate_proj", (H, I)), ("up_proj", (H, I)), ("down_proj", (I, H))):
                p = f"base_model.model.layers.{l}.mlp.{nm}."
                sd[p + "lora_A.weight"] = _n(g, r, fin, std=0.02)
                sd[p + "lora_B.weight"] = _n(g, fout, r, std=0.02)
    return sd


def make_spk_stat(seed: int = 2468, dim: int = 768) -> torch.Tensor:
    """``spk_stat.pt``: ``[2*dim]`` = (std, mean) halves (reference chattts_plus_pipeline.py:140-145)."""
    g = _gen(seed)
    return torch.cat([_u(g, dim, lo=0.5, hi=2.0), _n(g, dim, std=1.0)])


# ------------------------------------------------------------------------------------------------------------------------
# A complete synthetic ``CHATTTS_PLUS_CHECKPOINT_DIR`` (the files configs/infer/chattts_plus.yaml names), so the reference's own
# callers (tests/test_pipelines.py, webui.py) can run end to end without the hub download (chattts_plus_pipeline.py:70-92).
# ------------------------------------------------------------------------------------------------------------------------
_SPECIAL_TOKENS = ["[Stts]", "[Ptts]", "[spk_emb]", "[empty_spk]", "[Sbreak]", "[Pbreak]", "[Ebreak]", "[uv_break]", "[v_break]", "[lbreak]",
                   "[llbreak]", "[undefine]", "[laugh]", "[music]"] + [f"[speed_{i}]" for i in range(10)] + [f"[break_{i}]" for i in range(8)] + \
                  [f"[oral_{i}]" for i in range(10)] + [f"[laugh_{i}]" for i in range(3)]


def make_bert_tokenizer(texts=(), extra_chars: str = ""):
    """A small ``BertTokenizerFast`` (what asset/tokenizer.pt pickles, tokenizer.py:27-31) with ChatTTS's control tokens and one
    vocabulary entry per character of ``texts`` (CJK characters are split by the BERT normaliser, Latin text falls back to
    characters through WordPiece continuation pieces)."""
    from tokenizers import Tokenizer as _Tok, models, normalizers, pre_tokenizers
    from transformers import BertTokenizerFast
    base = ["[PAD]", "[UNK]", "[CLS]", "[SEP]", "[MASK]"]
    chars = sorted(set("".join(texts) + extra_chars + "abcdefghijklmnopqrstuvwxyz0123456789,.!?;:'\"-") - set(" \t\n[]"))
    vocab = base + chars + ["##" + c for c in chars] + _SPECIAL_TOKENS   # control tokens above the text ids, like the real vocabulary
                                                                          # (the refine pass keeps ids < [break_0], chattts_plus_pipeline.py:406)
    tok = _Tok(models.WordPiece({t: i for i, t in enumerate(dict.fromkeys(vocab))}, unk_token="[UNK]", max_input_chars_per_word=200))
    tok.normalizer = normalizers.BertNormalizer(lowercase=True, handle_chinese_chars=True)
    tok.pre_tokenizer = pre_tokenizers.BertPreTokenizer()
    fast = BertTokenizerFast(tokenizer_object=tok, unk_token="[UNK]", pad_token="[PAD]", sep_token="[SEP]", cls_token="[CLS]", mask_token="[MASK]")
    fast.add_special_tokens({"additional_special_tokens": list(_SPECIAL_TOKENS)})
    return fast


def write_checkpoint_dir(root: str, *, gpt_cfg: Optional[GPTConfig] = None, texts=(), lora_dir: Optional[str] = None, lora_r: int = 8,
                         seed: int = 1234, half: bool = True) -> Dict[str, str]:
    """Writes asset/{GPT,Decoder,DVAE_full,Vocos,spk_stat,tokenizer}.pt under ``root`` (the layout of the ChatTTS hub snapshot the
    reference downloads) from the seeded synthetic states, and optionally a peft-format LoRA adapter directory
    (adapter_config.json + adapter_model.safetensors, webui.py:48-62).  Shapes follow configs/infer/chattts_plus.yaml."""
    import json
    import os
    gpt_cfg = gpt_cfg or GPTConfig()
    asset = os.path.join(root, "asset")
    os.makedirs(asset, exist_ok=True)
    cast = (lambda sd: {k: (v.half() if v.is_floating_point() and half else v) for k, v in sd.items()})
    paths = {}

    def save(name, obj):
        paths[name] = os.path.join(asset, name)
        torch.save(obj, paths[name])

    save("GPT.pt", cast(make_gpt_state(gpt_cfg, seed=seed)))
    save("Decoder.pt", cast(make_dvae_state(DVAEConfig(), seed=seed + 1)))
    save("DVAE_full.pt", cast(make_dvae_state(DVAEConfig.codes_model(encoder=True), seed=seed + 2)))
    save("Vocos.pt", cast(make_vocos_state(VocosConfig(), seed=seed + 3)))
    save("spk_stat.pt", make_spk_stat(seed + 4))
    save("tokenizer.pt", make_bert_tokenizer(texts))
    if lora_dir:
        from safetensors.torch import save_file
        os.makedirs(lora_dir, exist_ok=True)
        save_file({k: v.contiguous() for k, v in make_lora_state(gpt_cfg, r=lora_r, seed=seed + 5).items()}, os.path.join(lora_dir, "adapter_model.safetensors"))
        with open(os.path.join(lora_dir, "adapter_config.json"), "w", encoding="utf-8") as f:
            json.dump({"peft_type": "LORA", "r": lora_r, "lora_alpha": 16, "lora_dropout": 0.05, "bias": "none", "use_rslora": False,
                       "target_modules": ["q_proj", "v_proj", "k_proj", "o_proj"], "task_type": None, "fan_in_fan_out": False}, f)
        paths["lora"] = lora_dir
    return paths
