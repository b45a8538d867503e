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


def make_gpt(cfg: synth.GPTConfig, seed: int, max_batch=32, half_round_oracle=True, mutate=None):
    """Returns (GPT on cuda, fp32 state dict for the oracle).  With half_round_oracle the oracle sees the same
    fp16-rounded matrices the kernels stream (isolates kernel error from weight quantisation).  ``mutate(sd)`` edits the
    synthetic checkpoint before either side sees it (e.g. to rig the EOS row of a head)."""
    sd = synth.make_gpt_state(cfg, seed=seed)
    if mutate is not None:
        mutate(sd)
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


def check_generate_against_oracle(gpt, osd, cfg, ids, mask, text_mask, temp, u, *, max_new, min_new, eos=625, tol=2e-2, rep=1.05):
    """Free-running CUDA generate (prefill + graph-replayed decode + fused sampler, shared uniforms) judged by the oracle run
    TEACHER-FORCED on the CUDA path's own tokens: every draw must be the oracle's choice or lie within ``tol`` (raw-logit units, after
    the repetition penalty) of the oracle's best token — fp16 operands may flip an argmax only at such near-ties — and the
    finish / end_idx / length bookkeeping must follow gpt.py:483-494,527-532 from those tokens.  Returns a report dict."""
    from chatttsplus_b200.processors import gen_logits
    from oracle import ctp_oracle as O
    warpers, procs = gen_logits(num_code=eos, top_P=0.7, top_K=20, repetition_penalty=rep)
    B = ids.shape[0]
    nq = cfg.num_vq
    emb = gpt(ids.cuda(), text_mask.cuda())
    out = list(gpt.generate(emb, ids.cuda(), temp.cuda(), eos, mask.cuda(), max_new_token=max_new, min_new_token=min_new,
                            logits_warpers=warpers, logits_processors=procs, return_hidden=True, show_tqdm=False, uniforms=u))[-1]
    last = gpt._last_run
    steps = int(last["steps"])
    raw = last["ids_buf"][:, :steps].cpu().long()                    # [B, steps, nq]: finished rows keep decoding
    hid = last["hid_buf"][:, :steps].cpu()
    # bookkeeping implied by the tokens themselves
    is_eos = (raw == eos).any(-1)                                    # [B, steps]
    first = torch.where(is_eos.any(1), is_eos.float().argmax(1), torch.full((B,), steps))
    assert torch.equal(last["end_idx"].cpu().long(), first), (last["end_idx"].tolist(), first.tolist())
    assert torch.equal(last["finish"].cpu().bool(), is_eos.any(1))
    for b in range(B):
        n = int(first[b])
        assert out.ids[b].shape == (n, nq) and out.hiddens[b].shape[0] == n
        assert torch.equal(out.ids[b].cpu().long(), raw[b, :n]) and not bool((out.ids[b] == eos).any())
    # early exit: the loop stops within one poll interval (16 steps) of the step at which the last sequence finished
    if bool(is_eos.any(1).all()):
        assert steps <= int(first.max()) + 1 + 2 * 16
    emb_ref = O.gpt_embed(osd, ids, text_mask, nq)
    ref = O.generate(osd, emb_ref, ids, temp, eos, mask, n_layers=cfg.num_hidden_layers, n_heads=cfg.num_attention_heads,
                     max_new_token=steps, min_new_token=min_new, sampler="forced", forced_ids=raw, rep_penalty=rep, ensure_non_empty=False)
    exact = near = 0
    worst = 0.0
    t1 = torch.ones(B * nq, 1)
    tq = temp.float().reshape(-1).repeat(B).view(-1, 1)
    for i in range(len(ref.logits)):   # (the oracle stops at the step the last sequence finishes; the CUDA loop polls every 16 steps)
        hist = raw[:, :i].permute(0, 2, 1).reshape(B * nq, -1)
        sc = O.process_logits(ref.logits[i], hist, t1, rep_penalty=rep, rep_max_ids=eos, rep_window=16, top_p=None, top_k=None,
                              ban_eos=(i < min_new), eos_token=eos)
        full = O.process_logits(ref.logits[i], hist, tq, rep_penalty=rep, rep_max_ids=eos, rep_window=16, top_p=0.7, top_k=20,
                                ban_eos=(i < min_new), eos_token=eos)
        choice = O.sample_inverse_cdf(torch.softmax(full, -1), u[i])
        got = raw[:, i].reshape(-1)
        same = got == choice
        gap = sc.max(-1).values - sc.gather(1, got[:, None])[:, 0]
        ok = same | (gap <= tol)
        assert bool(ok.all()), f"step {i}: {int((~ok).sum())} draws differ from the oracle by more than a near-tie (max gap {float(gap[~ok].max()):.3e})"
        exact += int(same.sum())
        near += int((~same).sum())
        if bool((~same).any()):
            worst = max(worst, float(gap[~same].max()))
    hid_err = max([rel_rms(out.hiddens[b], ref.hiddens[b]) for b in range(B) if ref.hiddens[b].numel()] or [0.0])
    return {"steps": steps, "oracle_steps": len(ref.logits), "exact": exact, "near_tie": near, "worst_gap": worst, "end_idx": first.tolist(),
            "hidden_rel_rms": hid_err, "out": out, "ref": ref}
