"""Helpers shared by the -m gpu parity tests (all calls go through the C ABI via chatttsplus_b200._lib)."""
import ctypes as C

import torch

from chatttsplus_b200 import _lib, synth
from chatttsplus_b200.gpt import GPT


def rel_rms(a: torch.Tensor, b: torch.Tensor) -> float:
    a = a.double().cpu()
    b = b.double().cpu()
    return float(((a - b).pow(2).mean().sqrt()) / (b.pow(2).mean().sqrt() + 1e-30))


def max_abs(a, b) -> float:
    return float((a.double().cpu() - b.double().cpu()).abs().max())


def gemm(A, B, *, out_f16=False, gelu=False, atomic=False, swap=False, bias=None, block_n=128, split_k=1, out=None):
    """C-ABI ctp_gemm_f16: out[m, n] = sum_k A[m,k] * B[n,k]  (swap: out[n, m])."""
    M, K = A.shape
    N = B.shape[0]
    assert A.dtype == torch.float16 and B.dtype == torch.float16
    if out is None:
        shape = (N, M) if swap else (M, N)
        out = torch.zeros(*shape, device=A.device, dtype=torch.float16 if out_f16 else torch.float32)
    flags = (1 if out_f16 else 0) | (2 if gelu else 0) | (4 if atomic else 0) | (8 if swap else 0)
    st = _lib.lib().ctp_gemm_f16(M, N, K, _lib.ptr(A), A.stride(0), _lib.ptr(B), B.stride(0), _lib.ptr(out), out.stride(0),
                                 _lib.ptr(bias), flags, block_n, split_k, _lib.stream_ptr())
    _lib.check(st, "ctp_gemm_f16")
    return out


def make_gpt(cfg: synth.GPTConfig, seed: int, max_batch=32, half_round_oracle=True):
    """Returns (GPT on cuda, fp32 state dict for the oracle).  With half_round_oracle the oracle sees the same
    fp16-rounded matrices the kernels stream (isolates kernel error from weight quantisation)."""
    sd = synth.make_gpt_state(cfg, seed=seed)
    g = GPT(dict(hidden_size=cfg.hidden_size, intermediate_size=cfg.intermediate_size,
                 num_attention_heads=cfg.num_attention_heads, num_hidden_layers=cfg.num_hidden_layers),
            num_audio_tokens=cfg.num_audio_tokens, num_text_tokens=cfg.num_text_tokens, num_vq=cfg.num_vq, max_batch=max_batch)
    g.load_state_dict(sd)
    g.to("cuda")
    osd = dict(sd)
    if half_round_oracle:
        for k, v in sd.items():
            if "layernorm" in k or k == "gpt.norm.weight":
                continue
            if "parametrizations" in k:
                continue  # folded below
            osd[k] = v.half().float()
        # heads are folded before rounding in the product path: give the oracle g=||W||, v=W (fold is identity)
        for q in range(cfg.num_vq):
            W = (sd[f"head_code.{q}.parametrizations.weight.original0"] * sd[f"head_code.{q}.parametrizations.weight.original1"]
                 / sd[f"head_code.{q}.parametrizations.weight.original1"].norm(dim=1, keepdim=True)).half().float()
            osd[f"head_code.{q}.parametrizations.weight.original0"] = W.norm(dim=1, keepdim=True)
            osd[f"head_code.{q}.parametrizations.weight.original1"] = W
    return g, osd


def sample_cfg(temperature=0.3, top_p=0.7, top_k=20, rep=1.05, eos=625, min_new=0, num_vq=4, rep_max_ids=625, window=16, min_keep=3):
    c = _lib.SampleCfg()
    for i in range(num_vq):
        c.temperature[i] = temperature if not isinstance(temperature, (list, tuple)) else temperature[i]
    c.rep_penalty = rep
    c.rep_window = window
    c.rep_max_ids = rep_max_ids
    c.top_p = top_p
    c.top_k = top_k
    c.min_keep = min_keep
    c.eos = eos
    c.min_new = min_new
    c.seed = 1234
    return c
