"""CPU fp32 oracle for the ChatTTSPlus generation hot path.   *** TEST INFRASTRUCTURE ONLY ***

This file restates, in plain fp32 PyTorch on the CPU, the algorithm of the reference's hot path
(SURVEY.md §8a rows A1-A24).  It exists to CHECK the CUDA kernels; only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import it.
The product package (``chatttsplus_b200``) never imports anything from ``oracle/``.

Citations are ``path:line`` relative to the reference checkout (``/root/reference``).

Pinning status (see DESIGN.md "Oracle"):
  * trunk (A6-A11)        — pinned: checked against the reference's own ``chattts_plus/models/llama.py``
                             run in the build container (tests/golden/make_golden.py -> tests/golden/trunk_*.pt).
  * generate loop, heads, sampling processors (A1,A3-A5,A13-A18)
                           — pinned: checked against the reference's own ``gpt.py``/``processors.py`` run with
                             import shims (tests/golden/make_golden.py -> gpt_generate_*.pt, processors_*.pt).
  * DVAE decode (A20-A22)  — pinned: checked against the reference's own ``dvae.py`` (dvae_*.pt).
  * b14/LZMA speaker codec — pinned by the reference's shipped speaker strings (tests/golden/speaker_2222.txt).
  * Vocos decode (A24), GFSQ embed (A23), peft LoRA merge (A12)
                           — PARITY UNPINNED: ``vocos``, ``vector_quantize_pytorch`` and ``peft`` are third-party
                             pip dependencies (requirements.txt:6,8; unpinned ``vocos``,
                             ``vector_quantize_pytorch==1.17.8``) absent from /root/reference and from this
                             image; their published algorithms are restated below.  The ISTFT inside Vocos is
                             delegated to ``torch.istft`` exactly as the package does.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ------------------------------------------------------------------------------------------------
# Trunk: RMSNorm, RoPE, attention, MLP, decoder layer   (A6-A11)
# ------------------------------------------------------------------------------------------------

def rmsnorm(x: Tensor, w: Tensor, eps: float = 1e-6) -> Tensor:
    """llama.py:82-87 — fp32 variance, ``w * (x * rsqrt(mean(x^2)+eps))``."""
    x = x.float()
    var = x.pow(2).mean(-1, keepdim=True)
    return w * (x * torch.rsqrt(var + eps))


def rope_cos_sin(position_ids: Tensor, head_dim: int = 64, theta: float = 10000.0) -> Tuple[Tensor, Tensor]:
    """llama.py:98,106-119 — inv_freq = theta^(-2i/d); emb = cat(freqs, freqs)."""
    inv_freq = 1.0 / (theta ** (torch.arange(0, head_dim, 2, dtype=torch.int64).float() / head_dim))
    freqs = position_ids[..., None].float() * inv_freq  # [B, S, d/2]
    emb = torch.cat((freqs, freqs), dim=-1)
    return emb.cos(), emb.sin()


def rotate_half(x: Tensor) -> Tensor:
    """llama.py:151-155."""
    h = x.shape[-1] // 2
    return torch.cat((-x[..., h:], x[..., :h]), dim=-1)


def apply_rope(q: Tensor, k: Tensor, cos: Tensor, sin: Tensor) -> Tuple[Tensor, Tensor]:
    """llama.py:158-182 with unsqueeze_dim=1 (q,k are [B, heads, S, d])."""
    cos = cos.unsqueeze(1)
    sin = sin.unsqueeze(1)
    return q * cos + rotate_half(q) * sin, k * cos + rotate_half(k) * sin


@dataclass
class KVCache:
    """Growing per-layer K/V like HF DynamicCache.update (llama.py:630-633): cat along the sequence dim."""
    k: List[Optional[Tensor]]
    v: List[Optional[Tensor]]

    @staticmethod
    def empty(n_layers: int) -> "KVCache":
        return KVCache([None] * n_layers, [None] * n_layers)

    def update(self, layer: int, k: Tensor, v: Tensor) -> Tuple[Tensor, Tensor]:
        if self.k[layer] is None:
            self.k[layer], self.v[layer] = k, v
        else:
            self.k[layer] = torch.cat([self.k[layer], k], dim=2)
            self.v[layer] = torch.cat([self.v[layer], v], dim=2)
        return self.k[layer], self.v[layer]

    def seq_len(self) -> int:
        return 0 if self.k[0] is None else self.k[0].shape[2]


def build_additive_mask(attention_mask: Optional[Tensor], q_len: int, past_len: int, dtype=torch.float32) -> Optional[Tensor]:
    """llama.py:1021-1099 (_update_causal_mask, sdpa branch, CPU): additive ``finfo.min`` mask [B,1,q,L].

    key j is visible to query i (absolute position past_len+i) iff j <= past_len+i and attention_mask[b,j]==1.
    Returns None when there is no padding and the plain causal structure suffices (llama.py:1046-1053).
    """
    if attention_mask is None or bool((attention_mask != 0).all()):
        return None
    B, L = attention_mask.shape
    min_v = torch.finfo(dtype).min
    qpos = torch.arange(past_len, past_len + q_len)[:, None]
    kpos = torch.arange(L)[None, :]
    causal = torch.where(kpos > qpos, min_v, 0.0).to(dtype)  # [q, L]
    m = causal[None, None].expand(B, 1, q_len, L).clone()
    pad = (attention_mask == 0)[:, None, None, :].expand(B, 1, q_len, L)
    m = m.masked_fill(pad, min_v)
    return m


def decoder_layer(x: Tensor, p: Dict[str, Tensor], prefix: str, n_heads: int, cos: Tensor, sin: Tensor,
                  mask: Optional[Tensor], cache: KVCache, layer: int, eps: float) -> Tensor:
    """llama.py:689-749 (pre-norm residual block) with LlamaSdpaAttention llama.py:590-668 and LlamaMLP :196-216."""
    B, S, H = x.shape
    d = H // n_heads
    h = rmsnorm(x, p[prefix + "input_layernorm.weight"], eps)
    q = F.linear(h, p[prefix + "self_attn.q_proj.weight"]).view(B, S, n_heads, d).transpose(1, 2)
    k = F.linear(h, p[prefix + "self_attn.k_proj.weight"]).view(B, S, n_heads, d).transpose(1, 2)
    v = F.linear(h, p[prefix + "self_attn.v_proj.weight"]).view(B, S, n_heads, d).transpose(1, 2)
    q, k = apply_rope(q, k, cos, sin)
    k, v = cache.update(layer, k, v)
    L = k.shape[2]
    if mask is not None:
        am = mask[:, :, :, :L]
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=am, is_causal=False)
    else:
        a = F.scaled_dot_product_attention(q, k, v, attn_mask=None, is_causal=(S > 1))
    a = a.transpose(1, 2).contiguous().view(B, S, H)
    x = x + F.linear(a, p[prefix + "self_attn.o_proj.weight"])
    h = rmsnorm(x, p[prefix + "post_attention_layernorm.weight"], eps)
    g = F.linear(h, p[prefix + "mlp.gate_proj.weight"])
    u = F.linear(h, p[prefix + "mlp.up_proj.weight"])
    x = x + F.linear(F.silu(g) * u, p[prefix + "mlp.down_proj.weight"])
    return x


def trunk_forward(p: Dict[str, Tensor], inputs_embeds: Tensor, attention_mask: Optional[Tensor],
                  position_ids: Tensor, cache: KVCache, n_layers: int, n_heads: int, eps: float = 1e-6,
                  theta: float = 10000.0, prefix: str = "gpt.") -> Tensor:
    """LlamaModel.forward, llama.py:905-1019: mask build, n_layers decoder layers, final RMSNorm.

    ``attention_mask`` is the full 2-D mask [B, past+q]; ``position_ids`` [B, q].
    """
    x = inputs_embeds.float()
    B, S, H = x.shape
    past = cache.seq_len()
    mask = build_additive_mask(attention_mask, S, past)
    cos, sin = rope_cos_sin(position_ids, H // n_heads, theta)
    for l in range(n_layers):
        x = decoder_layer(x, p, f"{prefix}layers.{l}.", n_heads, cos, sin, mask, cache, l, eps)
    return rmsnorm(x, p[prefix + "norm.weight"], eps)


def position_ids_from_mask(attention_mask: Tensor) -> Tensor:
    """gpt.py:238-245 — cumsum(mask)-1 with padded slots set to 1."""
    pos = attention_mask.long().cumsum(-1) - 1
    return pos.masked_fill(attention_mask == 0, 1)


# ------------------------------------------------------------------------------------------------
# GPT wrapper: embeddings, heads, sampling processors, generate   (A1-A5, A13-A18)
# ------------------------------------------------------------------------------------------------

def gpt_embed(p: Dict[str, Tensor], input_ids: Tensor, text_mask: Tensor, num_vq: int = 4) -> Tensor:
    """GPT.forward, gpt.py:125-149 — text rows use emb_text[ids[...,0]]; the rest sum the num_vq code tables."""
    emb_text = F.embedding(input_ids[text_mask][:, 0], p["emb_text.weight"])
    inv = ~text_mask
    mids = input_ids[inv]
    emb_code = sum(F.embedding(mids[:, q], p[f"emb_code.{q}.weight"]) for q in range(num_vq)) if mids.numel() else None
    emb = torch.zeros(input_ids.shape[:-1] + (emb_text.shape[-1],), dtype=torch.float32)
    emb[text_mask] = emb_text
    if emb_code is not None:
        emb[inv] = emb_code
    return emb


def apply_spk_emb(emb: Tensor, spk: Tensor, input_ids: Tensor, spk_emb_id: int) -> Tensor:
    """tokenizer.py:150-178 — L2-normalised speaker vector written where ids[...,0]==[spk_emb]."""
    n = F.normalize(spk.float(), p=2.0, dim=0, eps=1e-12)
    cond = input_ids[..., 0:1].eq(spk_emb_id).expand(emb.shape)
    return torch.where(cond, n.view(1, 1, -1).expand(emb.shape), emb)


def weight_norm_fold(g: Tensor, v: Tensor) -> Tensor:
    """torch weight_norm(dim=0): W = g * v / ||v||_2 per output row (gpt.py:57-77)."""
    return g * v / v.norm(dim=1, keepdim=True)


def code_embed(p: Dict[str, Tensor], ids: Tensor, num_vq: int = 4) -> Tensor:
    """gpt.py:402-407 — sum_q emb_code[q][ids[..., q]]; ids [B, S, num_vq] -> [B, S, H]."""
    return sum(F.embedding(ids[..., q], p[f"emb_code.{q}.weight"]) for q in range(num_vq))


def head_code_logits(p: Dict[str, Tensor], hidden_last: Tensor, num_vq: int = 4) -> Tensor:
    """gpt.py:424-457 — logits[b, :, q] = h W_q^T; returned as [B*num_vq, num_audio] with row = b*num_vq+q."""
    outs = []
    for q in range(num_vq):
        W = weight_norm_fold(p[f"head_code.{q}.parametrizations.weight.original0"],
                             p[f"head_code.{q}.parametrizations.weight.original1"])
        outs.append(F.linear(hidden_last, W))  # [B, A]
    return torch.stack(outs, dim=1).reshape(-1, outs[0].shape[-1])


def head_text_logits(p: Dict[str, Tensor], hidden_last: Tensor) -> Tensor:
    """gpt.py:425-426."""
    W = weight_norm_fold(p["head_text.parametrizations.weight.original0"],
                         p["head_text.parametrizations.weight.original1"])
    return F.linear(hidden_last, W)


def repetition_penalty(input_ids: Tensor, scores: Tensor, penalty: float, max_input_ids: int, past_window: int) -> Tensor:
    """processors.py:18-34 (CustomRepetitionPenaltyLogitsProcessorRepeat.__call__), including the
    ``freq.narrow(0, max_input_ids, ...)`` row-truncation quirk."""
    if input_ids.size(1) > past_window:
        input_ids = input_ids.narrow(1, -past_window, past_window)
    freq = F.one_hot(input_ids, scores.size(1)).sum(1)
    if freq.size(0) > max_input_ids:
        freq.narrow(0, max_input_ids, freq.size(0) - max_input_ids).zero_()
    alpha = torch.pow(penalty, freq)
    return torch.where(scores < 0, scores * alpha, scores / alpha)


def top_p_warp(scores: Tensor, top_p: float, min_tokens_to_keep: int = 3) -> Tensor:
    """transformers TopPLogitsWarper.__call__ (generation/logits_process.py; pinned version 4.41 in
    requirements.txt:7 and unchanged in 5.5): sort ascending, drop cumulative prob <= 1-top_p, always keep the
    last ``min_tokens_to_keep``."""
    sorted_logits, sorted_indices = torch.sort(scores, descending=False)
    cumulative_probs = sorted_logits.softmax(dim=-1).cumsum(dim=-1)
    sorted_remove = cumulative_probs <= (1 - top_p)
    sorted_remove[..., -min_tokens_to_keep:] = 0
    remove = sorted_remove.scatter(1, sorted_indices, sorted_remove)
    return scores.masked_fill(remove, -float("inf"))


def top_k_warp(scores: Tensor, top_k: int, min_tokens_to_keep: int = 3) -> Tensor:
    """transformers TopKLogitsWarper.__call__: remove everything below the k-th largest."""
    k = min(max(top_k, min_tokens_to_keep), scores.size(-1))
    remove = scores < torch.topk(scores, k)[0][..., -1, None]
    return scores.masked_fill(remove, -float("inf"))


def process_logits(logits: Tensor, history: Tensor, temperature: Tensor, *, rep_penalty: Optional[float],
                   rep_max_ids: int, rep_window: int, top_p: Optional[float], top_k: Optional[int],
                   ban_eos: bool, eos_token: int) -> Tensor:
    """gpt.py:469-478: temperature -> processors -> warpers (TopP then TopK, processors.py:43-47) -> min-len EOS."""
    logits = logits / temperature
    if rep_penalty is not None and rep_penalty != 1:
        logits = repetition_penalty(history, logits, rep_penalty, rep_max_ids, rep_window)
    if top_p is not None:
        logits = top_p_warp(logits, top_p)
    if top_k is not None:
        logits = top_k_warp(logits, top_k)
    if ban_eos:
        logits = logits.clone()
        logits[:, eos_token] = -torch.inf
    return logits


def sample_inverse_cdf(scores: Tensor, u: Tensor) -> Tensor:
    """The B200 sampler's draw: inverse CDF over token id order with caller-supplied uniforms u in [0,1).

    Chooses the smallest index i with cumsum(p)[i] > u * sum(p).  (The reference calls torch.multinomial,
    gpt.py:480-481, whose stream cannot be reproduced by a custom kernel; in 'torch' mode the oracle calls
    torch.multinomial itself, in 'uniform' mode both sides consume the same uniforms.)
    """
    c = scores.double().cumsum(-1)
    t = u.double()[:, None] * c[:, -1:]
    idx = (c <= t).sum(-1).clamp_(max=scores.shape[-1] - 1)
    return idx


@dataclass
class GenerateResult:
    ids: List[Tensor]            # per sequence [n_i, num_vq]
    hiddens: List[Tensor]        # per sequence [n_i, H]
    logits: List[Tensor]         # per step raw head logits [B*num_vq, A] (oracle extra, for teacher-forced parity)
    steps: int


@torch.no_grad()
def generate(p: Dict[str, Tensor], emb: Tensor, inputs_ids: Tensor, temperature: Tensor, eos_token: int,
             attention_mask: Optional[Tensor], *, n_layers: int, n_heads: int, num_vq: int = 4,
             max_new_token: int = 2048, min_new_token: int = 0, rep_penalty: Optional[float] = 1.05,
             rep_window: int = 16, top_p: Optional[float] = 0.7, top_k: Optional[int] = 20,
             sampler: str = "torch", uniforms: Optional[Tensor] = None, forced_ids: Optional[Tensor] = None,
             eps: float = 1e-6, theta: float = 10000.0, ensure_non_empty: bool = True,
             max_steps: Optional[int] = None, infer_text: bool = False) -> GenerateResult:
    """GPT.generate, gpt.py:313-569 (stream=False).  infer_text=True is the refine-text pass: text embedding of the
    previous token (gpt.py:400-401), head_text logits (gpt.py:425-426), one sampling column per sequence, the sampled id
    written to every VQ column (gpt.py:489-494), ids returned as 1-D tensors (gpt.py:298-299).

    sampler: 'torch' (torch.multinomial, consumes the global CPU generator exactly like the reference),
             'uniform' (inverse CDF with ``uniforms`` [max_new_token, B*num_vq]),
             'forced' (teacher forcing with ``forced_ids`` [B, steps, num_vq]; still runs the processors).
    """
    B, L0, _ = inputs_ids.shape
    start_idx = L0
    end_idx = torch.zeros(B, dtype=torch.long)
    finish = torch.zeros(B, dtype=torch.bool)
    temp = temperature.float().unsqueeze(0).expand(B, -1).contiguous().view(-1, 1)  # gpt.py:346-351
    ncol = 1 if infer_text else num_vq
    mask_cache = torch.ones(B, L0 + max_new_token, dtype=torch.bool)
    if attention_mask is not None:
        mask_cache[:, : attention_mask.shape[1]] = attention_mask.bool()
    ids_buf = torch.zeros(B, L0 + max_new_token, num_vq, dtype=torch.long)
    ids_buf[:, :L0] = inputs_ids
    progress = L0
    cache = KVCache.empty(n_layers)
    hiddens: List[Tensor] = []
    all_logits: List[Tensor] = []
    steps = 0
    for i in range(max_new_token):
        if max_steps is not None and i >= max_steps:
            break
        cur_mask = mask_cache[:, :progress]
        pos_full = position_ids_from_mask(cur_mask)
        if i == 0:
            x = emb
            pos = pos_full
        else:
            if infer_text:
                x = F.embedding(ids_buf[:, progress - 1: progress, 0], p["emb_text.weight"])
            else:
                x = code_embed(p, ids_buf[:, progress - 1: progress], num_vq)
            pos = pos_full[:, -1:]
        h = trunk_forward(p, x, cur_mask, pos, cache, n_layers, n_heads, eps, theta)
        h_last = h[:, -1]
        hiddens.append(h_last)
        logits = head_text_logits(p, h_last) if infer_text else head_code_logits(p, h_last, num_vq)  # [B*ncol, V]
        all_logits.append(logits)
        history = ids_buf[:, start_idx:progress, :ncol].permute(0, 2, 1).reshape(B * ncol, -1)
        scores = process_logits(logits, history, temp, rep_penalty=rep_penalty, rep_max_ids=eos_token,
                                rep_window=rep_window, top_p=top_p, top_k=top_k,
                                ban_eos=(i < min_new_token), eos_token=eos_token)
        probs = F.softmax(scores, dim=-1)
        if sampler == "torch":
            idx_next = torch.multinomial(probs, num_samples=1).view(-1)
        elif sampler == "uniform":
            idx_next = sample_inverse_cdf(probs, uniforms[i])
        elif sampler == "forced":
            idx_next = forced_ids[:, i].reshape(-1)
        else:
            raise ValueError(sampler)
        idx_next = idx_next.view(-1, ncol)
        finish |= idx_next.eq(eos_token).any(1)
        ids_buf[:, progress] = idx_next if not infer_text else idx_next.expand(-1, num_vq)
        if i == 0 and bool(finish.any()) and ensure_non_empty and sampler == "torch":
            # gpt.py:496-525 — regenerate from scratch (consumes more RNG)
            return generate(p, emb, inputs_ids, temperature, eos_token, attention_mask, n_layers=n_layers,
                            n_heads=n_heads, num_vq=num_vq, max_new_token=max_new_token,
                            min_new_token=min_new_token, rep_penalty=rep_penalty, rep_window=rep_window,
                            top_p=top_p, top_k=top_k, sampler=sampler, uniforms=uniforms, forced_ids=forced_ids,
                            eps=eps, theta=theta, ensure_non_empty=ensure_non_empty, max_steps=max_steps, infer_text=infer_text)
        progress += 1
        end_idx += (~finish).long()
        steps += 1
        if bool(finish.all()):
            break
    ids = [ids_buf[b, start_idx: start_idx + int(end_idx[b])] for b in range(B)]
    if infer_text:
        ids = [i[:, 0] for i in ids]
    hs = torch.stack(hiddens, 1)
    hid = [hs[b, : int(end_idx[b])] for b in range(B)]
    return GenerateResult(ids=ids, hiddens=hid, logits=all_logits, steps=steps)


def lora_merge(p: Dict[str, Tensor], lora: Dict[str, Tensor], n_layers: int, alpha: float, r: int,
               prefix: str = "gpt.", use_rslora: bool = False) -> Dict[str, Tensor]:
    """peft ``merge_and_unload`` for Linear LoRA (peft/tuners/lora/layer.py ``Linear.get_delta_weight``): W' = W + scaling * B @ A
    with scaling = lora_alpha / r (lora_alpha / sqrt(r) when ``use_rslora``) on every adapted Linear — q/k/v/o
    (chattts_plus_pipeline.py:420-434; configs/train/train_voice_clone_lora.yaml:72-80) and, when the adapter carries them,
    mlp.gate/up/down.  UNPINNED against the package itself (peft is absent); pinned by the hand-derived known-answer case in
    tests/test_oracle_golden.py::test_lora_merge_known_answer."""
    out = dict(p)
    s = alpha / (r ** 0.5) if use_rslora else alpha / r
    for l in range(n_layers):
        for sub, names in (("self_attn", ("q_proj", "k_proj", "v_proj", "o_proj")), ("mlp", ("gate_proj", "up_proj", "down_proj"))):
            for nm in names:
                a = lora.get(f"base_model.model.layers.{l}.{sub}.{nm}.lora_A.weight")
                b = lora.get(f"base_model.model.layers.{l}.{sub}.{nm}.lora_B.weight")
                if a is None:
                    continue
                key = f"{prefix}layers.{l}.{sub}.{nm}.weight"
                out[key] = p[key] + s * (b.float() @ a.float())
    return out


# ------------------------------------------------------------------------------------------------
# DVAE decode   (A20-A23)
# ------------------------------------------------------------------------------------------------

def convnext_block(x: Tensor, p: Dict[str, Tensor], prefix: str, dilation: int) -> Tensor:
    """dvae.py:48-63 / vocos ConvNeXtBlock: depthwise conv -> LayerNorm(eps 1e-6) -> Linear -> exact GELU ->
    Linear -> *gamma -> + residual.  x is [B, C, T]."""
    w = p[prefix + "dwconv.weight"]
    k = w.shape[-1]
    y = F.conv1d(x, w, p[prefix + "dwconv.bias"], padding=dilation * (k // 2), dilation=dilation, groups=x.shape[1])
    y = y.transpose(1, 2)
    y = F.layer_norm(y, (y.shape[-1],), p[prefix + "norm.weight"], p[prefix + "norm.bias"], eps=1e-6)
    y = F.linear(y, p[prefix + "pwconv1.weight"], p[prefix + "pwconv1.bias"])
    y = F.gelu(y)
    y = F.linear(y, p[prefix + "pwconv2.weight"], p[prefix + "pwconv2.bias"])
    y = y * p[prefix + "gamma"]
    return y.transpose(1, 2) + x


def gfsq_embed(p: Dict[str, Tensor], ids: Tensor, levels: Sequence[int] = (5, 5, 5, 5), G: int = 2, R: int = 2) -> Tensor:
    """GFSQ._embed, dvae.py:84-94 -> GroupedResidualFSQ.get_output_from_indices (vector_quantize_pytorch 1.17.8,
    UNPINNED restatement).  ids [B, G*R, T] -> features [B, dim, T].

    Per group g: sum over residual levels r of implicit-codebook[idx] / (levels-1)^r scaled, then project_out.
    FSQ implicit codebook: digit_j = (idx // prod(levels[:j])) % levels[j]; value = (digit - half) / half with
    half = levels[j] // 2.  ResidualFSQ scales level r by (levels-1)^(-r).
    """
    B, GR, T = ids.shape
    x = ids.transpose(1, 2).reshape(B, T, G, R)
    lv = torch.tensor(list(levels), dtype=torch.long)
    basis = torch.cumprod(torch.cat([torch.ones(1, dtype=torch.long), lv[:-1]]), 0)
    half = (lv // 2).float()
    outs = []
    for g in range(G):
        acc = torch.zeros(B, T, len(levels))
        for r in range(R):
            idx = x[:, :, g, r]
            digits = (idx[..., None] // basis) % lv
            codes = (digits.float() - half) / half
            scale = (lv.float() - 1.0) ** (-r)
            acc = acc + codes * scale
        w = p[f"vq_layer.quantizer.rvqs.{g}.project_out.weight"]
        b = p[f"vq_layer.quantizer.rvqs.{g}.project_out.bias"]
        outs.append(F.linear(acc, w, b))
    feat = torch.cat(outs, dim=-1)  # [B, T, dim]
    return feat.transpose(1, 2)


def dvae_decode(p: Dict[str, Tensor], inp: Tensor, *, n_layer: int = 12, dilation: int = 2, vq: bool = False,
                levels: Sequence[int] = (5, 5, 5, 5), G: int = 2, R: int = 2) -> Tensor:
    """DVAE.forward decode branch, dvae.py:272-291: [B, C, T] -> mel [B, 100, 2T]."""
    x = gfsq_embed(p, inp, levels, G, R) if vq else inp.float()
    B, C, T = x.shape
    x = x.view(B, 2, C // 2, T).permute(0, 2, 3, 1).flatten(2)  # dvae.py:277-283
    # DVAEDecoder.forward, dvae.py:161-168
    y = F.conv1d(x, p["decoder.conv_in.0.weight"], p["decoder.conv_in.0.bias"], padding=1)
    y = F.gelu(y)
    y = F.conv1d(y, p["decoder.conv_in.2.weight"], p["decoder.conv_in.2.bias"], padding=1)
    for l in range(n_layer):
        y = convnext_block(y, p, f"decoder.decoder_block.{l}.", dilation)
    y = F.conv1d(y, p["decoder.conv_out.weight"])
    y = F.conv1d(y, p["out_conv.weight"], padding=1)  # dvae.py:285
    return y * p["coef"]  # dvae.py:291


# ------------------------------------------------------------------------------------------------
# DVAE encode = zero-shot speaker prompt: audio -> mel -> encoder -> GFSQ indices
#   mel / downsample / encoder are PINNED against the reference's own dvae.py (tests/golden/dvae_encode_ref.pt);
#   the quantiser (pip vector_quantize_pytorch==1.17.8 GroupedResidualFSQ.forward, absent here) is restated: UNPINNED
# ------------------------------------------------------------------------------------------------

def melscale_fbanks_htk(n_freqs: int = 513, f_min: float = 0.0, f_max: float = 12000.0, n_mels: int = 100,
                        sample_rate: int = 24000) -> Tensor:
    """torchaudio.functional.melscale_fbanks(norm=None, mel_scale="htk"), the filter bank of
    torchaudio.transforms.MelSpectrogram as constructed at dvae.py:183-190 -> [n_freqs, n_mels]."""
    all_freqs = torch.linspace(0, sample_rate // 2, n_freqs)
    m_min = 2595.0 * math.log10(1.0 + f_min / 700.0)
    m_max = 2595.0 * math.log10(1.0 + f_max / 700.0)
    m_pts = torch.linspace(m_min, m_max, n_mels + 2)
    f_pts = 700.0 * (10 ** (m_pts / 2595.0) - 1.0)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts.unsqueeze(0) - all_freqs.unsqueeze(1)
    down = (-1.0 * slopes[:, :-2]) / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    return torch.clamp(torch.min(down, up), min=0.0)


def mel_features(audio: Tensor, n_fft: int = 1024, hop: int = 256, n_mels: int = 100, sample_rate: int = 24000) -> Tensor:
    """MelSpectrogramFeatures.forward, dvae.py:196-199: log(clip(mel(|STFT|), 1e-5)); audio [B, N] -> [B, n_mels, N//hop + 1]
    (center=True, reflect padding, periodic Hann window, power=1)."""
    win = torch.hann_window(n_fft, periodic=True, dtype=torch.float32)
    spec = torch.stft(audio.float(), n_fft, hop_length=hop, win_length=n_fft, window=win, center=True, pad_mode="reflect",
                      normalized=False, onesided=True, return_complex=True).abs()
    fb = melscale_fbanks_htk(n_fft // 2 + 1, 0.0, sample_rate / 2.0, n_mels, sample_rate)
    mel = torch.matmul(spec.transpose(-1, -2), fb).transpose(-1, -2)
    return torch.log(torch.clip(mel, min=1e-5))


def dvae_encode_features(p: Dict[str, Tensor], audio: Tensor, *, n_layer: int = 12, dilation: int = 2) -> Tensor:
    """DVAE.forward encode branch up to the quantiser, dvae.py:263-269: audio [B, N] -> x [B, 1024, T2]."""
    mel = mel_features(audio)
    x = mel / p["coef"].view(1, -1, 1)                                   # dvae.py:266
    x = F.gelu(F.conv1d(x, p["downsample_conv.0.weight"], p["downsample_conv.0.bias"], padding=1))   # dvae.py:227-232
    x = F.gelu(F.conv1d(x, p["downsample_conv.2.weight"], p["downsample_conv.2.bias"], stride=2, padding=1))
    y = F.conv1d(x, p["encoder.conv_in.0.weight"], p["encoder.conv_in.0.bias"], padding=1)          # DVAEDecoder.forward, dvae.py:161-168
    y = F.gelu(y)
    y = F.conv1d(y, p["encoder.conv_in.2.weight"], p["encoder.conv_in.2.bias"], padding=1)
    for l in range(n_layer):
        y = convnext_block(y, p, f"encoder.decoder_block.{l}.", dilation)
    return F.conv1d(y, p["encoder.conv_out.weight"])


def fsq_bound(z: Tensor, levels: Sequence[int], eps: float = 1e-3) -> Tensor:
    """vector_quantize_pytorch FSQ.bound: tanh squashing to (levels-1)/2 * (1+eps); offset/shift are 0 for odd levels."""
    lv = torch.tensor(list(levels), dtype=z.dtype)
    half_l = (lv - 1) * (1 + eps) / 2
    offset = torch.where(lv % 2 == 0, torch.tensor(0.5, dtype=z.dtype), torch.tensor(0.0, dtype=z.dtype))
    shift = torch.atanh(offset / half_l)
    return torch.tanh(z + shift) * half_l - offset


def gfsq_quantize(p: Dict[str, Tensor], x: Tensor, levels: Sequence[int] = (5, 5, 5, 5), G: int = 2, R: int = 2) -> Tensor:
    """GFSQ.forward, dvae.py:98-126 -> GroupedResidualFSQ.forward (UNPINNED restatement of vector_quantize_pytorch 1.17.8):
    x [B, dim, T] -> ind [B, G*R, T].  Per group: z = project_in(x_g) (Linear dim/G -> 4); residual = bound(z);
    for r < R: q = round(bound(residual / s_r)) / (levels // 2), ind_r = sum_j (q_j * hw_j + hw_j) * prod(levels[:j]),
    residual -= q * s_r, with s_r = (levels - 1) ** -r."""
    B, D, T = x.shape
    xt = x.transpose(1, 2).float()                     # dvae.py:99-100
    gd = D // G
    lv = torch.tensor(list(levels), dtype=torch.float32)
    half_width = torch.div(lv, 2, rounding_mode="floor")
    basis = torch.cumprod(torch.tensor([1] + list(levels[:-1]), dtype=torch.float32), 0)
    out = torch.zeros(B, T, G, R, dtype=torch.long)
    for g in range(G):
        z = F.linear(xt[..., g * gd:(g + 1) * gd], p[f"vq_layer.quantizer.rvqs.{g}.project_in.weight"],
                     p[f"vq_layer.quantizer.rvqs.{g}.project_in.bias"])
        residual = fsq_bound(z, levels)
        for r in range(R):
            scale = (lv - 1) ** (-r)
            q = torch.round(fsq_bound(residual / scale, levels)) / half_width
            out[:, :, g, r] = ((q * half_width + half_width) * basis).sum(-1).round().long()
            residual = residual - q * scale
    return out.view(B, T, G * R).transpose(1, 2)       # dvae.py:109-126: [B, T, (g r)] -> [B, G*R, T]


def gfsq_pre_round(p: Dict[str, Tensor], x: Tensor, levels: Sequence[int] = (5, 5, 5, 5), G: int = 2, R: int = 2) -> Tensor:
    """The values gfsq_quantize rounds, in float64: bound(residual / s_r) per (frame, group, residual level, code dimension),
    [B, T, G, R, len(levels)].  Test helper: an index may legitimately depend on fp32 summation order only where one of these
    sits within rounding error of k + 1/2 (a numerical tie)."""
    B, D, T = x.shape
    xt = x.transpose(1, 2).double()
    gd = D // G
    lv = torch.tensor(list(levels), dtype=torch.float64)
    out = torch.zeros(B, T, G, R, len(levels), dtype=torch.float64)
    for g in range(G):
        z = F.linear(xt[..., g * gd:(g + 1) * gd], p[f"vq_layer.quantizer.rvqs.{g}.project_in.weight"].double(),
                     p[f"vq_layer.quantizer.rvqs.{g}.project_in.bias"].double())
        residual = fsq_bound(z, levels)
        for r in range(R):
            scale = (lv - 1) ** (-r)
            v = fsq_bound(residual / scale, levels)
            out[:, :, g, r] = v
            residual = residual - torch.round(v) / torch.div(lv, 2, rounding_mode="floor") * scale
    return out


def dvae_encode(p: Dict[str, Tensor], audio: Tensor, **kw) -> Tensor:
    return gfsq_quantize(p, dvae_encode_features(p, audio, **kw))


# ------------------------------------------------------------------------------------------------
# Vocos decode   (A24; third-party `vocos` package restated — UNPINNED)
# ------------------------------------------------------------------------------------------------

def vocos_backbone(p: Dict[str, Tensor], mel: Tensor, num_layers: int = 8) -> Tensor:
    """vocos.models.VocosBackbone.forward: embed Conv1d(k7,p3) -> LayerNorm -> ConvNeXt x N -> final LayerNorm."""
    x = F.conv1d(mel.float(), p["backbone.embed.weight"], p["backbone.embed.bias"], padding=3)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), p["backbone.norm.weight"], p["backbone.norm.bias"], eps=1e-6)
    x = x.transpose(1, 2)
    for l in range(num_layers):
        x = convnext_block(x, p, f"backbone.convnext.{l}.", 1)
    x = F.layer_norm(x.transpose(1, 2), (x.shape[1],), p["backbone.final_layer_norm.weight"],
                     p["backbone.final_layer_norm.bias"], eps=1e-6)
    return x  # [B, T, C]


def vocos_head_spec(p: Dict[str, Tensor], x: Tensor) -> Tensor:
    """vocos.heads.ISTFTHead.forward up to the complex spectrogram: Linear -> chunk(mag, phase) ->
    exp -> clip(max=1e2) -> mag * (cos p + i sin p).  Returns complex [B, n_fft/2+1, T]."""
    y = F.linear(x, p["head.out.weight"], p["head.out.bias"]).transpose(1, 2)
    mag, ph = y.chunk(2, dim=1)
    mag = torch.exp(mag).clip(max=1e2)
    return torch.complex(mag * torch.cos(ph), mag * torch.sin(ph))


def vocos_decode(p: Dict[str, Tensor], mel: Tensor, num_layers: int = 8, n_fft: int = 1024, hop: int = 256) -> Tensor:
    """vocos.Vocos.decode = head(backbone(mel)); ISTFT(padding="center") = torch.istft(center=True) with the
    periodic Hann window buffer.  mel [B, 100, T] -> wav [B, hop*(T-1)]."""
    S = vocos_head_spec(p, vocos_backbone(p, mel, num_layers))
    return torch.istft(S, n_fft, hop, n_fft, p["head.istft.window"], center=True)


def decode_to_wav(dvae_p: Dict[str, Tensor], vocos_p: Dict[str, Tensor], hiddens: Tensor, **dvae_kw) -> Tensor:
    """ChatTTSPlusPipeline._decode_to_wavs for one utterance, chattts_plus_pipeline.py:298-304:
    hiddens [n, 768] -> permute -> DVAE -> vocos.decode -> wav [256*(2n-1)]."""
    src = hiddens.permute(1, 0)[None]
    mel = dvae_decode(dvae_p, src, **dvae_kw)
    return vocos_decode(vocos_p, mel)[0]
