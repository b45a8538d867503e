"""Multi-GPU plumbing for the hot path: utterances are independent (gpt.py:483-494, chattts_plus_pipeline.py:298-304), so
the batch shards across ranks with NO data-path collective.  torch.distributed (NCCL over NVLink on GPUs, gloo in the CPU
tests) is used only for: the one-off broadcast of the speaker embedding / LoRA-merged q,k,v,o at setup, the gather of
per-utterance lengths at the end, and the max-over-ranks timing in bench.py.
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous split of n_items over `world` ranks (config 4: 256 utterances -> 8 x 32)."""
    lo = (n_items * rank) // world
    hi = (n_items * (rank + 1)) // world
    return lo, hi


def broadcast_setup(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    """Speaker embedding (768 x 2 B) or merged LoRA weights (94 MB) from rank `src` to every replica."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(t, src)
    return t


def gather_lengths(local_lengths: Sequence[int], device=None) -> List[int]:
    """All ranks learn every utterance length (to lay out / order the waveforms); variable shard sizes allowed."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return list(local_lengths)
    world = dist.get_world_size()
    counts = [torch.zeros(1, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(counts, torch.tensor([len(local_lengths)], dtype=torch.int64, device=device))
    m = int(max(int(c) for c in counts))
    buf = torch.full((m,), -1, dtype=torch.int64, device=device)
    buf[: len(local_lengths)] = torch.tensor(list(local_lengths), dtype=torch.int64, device=device)
    outs = [torch.empty(m, dtype=torch.int64, device=device) for _ in range(world)]
    dist.all_gather(outs, buf)
    res: List[int] = []
    for c, o in zip(counts, outs):
        res.extend(int(v) for v in o[: int(c)].tolist())
    return res


def max_over_ranks(value: float, device=None) -> float:
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return value
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def rank_seed(seed: int, global_index: int) -> int:
    """Per-utterance RNG stream = (seed, global utterance index): results do not depend on the GPU count."""
    return (int(seed) * 1000003 + int(global_index)) & 0x7FFFFFFF


def utterance_uniforms(seed: int, g0: int, g1: int, max_new: int, num_vq: int) -> torch.Tensor:
    """Uniforms for the inverse-CDF draws of global utterances [g0, g1): fp32 [max_new, (g1 - g0) * num_vq], column (g - g0) * num_vq + q.
    Utterance g always consumes the stream seeded with rank_seed(seed, g), whatever the rank count or slice size."""
    cols = []
    for g in range(g0, g1):
        gen = torch.Generator().manual_seed(rank_seed(seed, g))
        cols.append(torch.rand(max_new, num_vq, generator=gen))
    return torch.cat(cols, dim=1) if cols else torch.zeros(max_new, 0)


def infer_ids_sharded(pipe, input_ids: torch.Tensor, attention_mask: torch.Tensor, text_mask: torch.Tensor, params, *, seed: int,
                      slice_size: int = 32, spk_emb_ids: Optional[int] = None, use_decoder: bool = True, return_codes: bool = False):
    """The reference's slice loop (chattts_plus_pipeline.py:391-397) over a job that is sharded by utterance across the ranks of the
    current process group: rank r takes the contiguous range shard_range(n, world, r), runs it in slices of
    ``slice_size`` through ``pipe.infer_ids`` (per-utterance RNG streams, so tokens do not depend on the rank count), and every rank
    learns every utterance's waveform length.  No collective touches the data path; the speaker embedding is broadcast once.

    Returns (lo, hi, local_wavs, all_lengths[, local_results])."""
    world = dist.get_world_size() if (dist.is_available() and dist.is_initialized()) else 1
    rank = dist.get_rank() if world > 1 else 0
    n = int(input_ids.shape[0])
    lo, hi = shard_range(n, world, rank)
    nq = int(input_ids.shape[2])
    if params.spk_emb is not None and isinstance(params.spk_emb, torch.Tensor) and params.spk_emb.is_cuda:
        broadcast_setup(params.spk_emb, 0)
    wavs, results = [], []
    for a in range(lo, hi, slice_size):
        b = min(hi, a + slice_size)
        u = utterance_uniforms(seed, a, b, int(params.max_new_token), nq)
        out = None
        for out in pipe.infer_ids(input_ids[a:b], attention_mask[a:b], text_mask[a:b], params, use_decoder=use_decoder,
                                  spk_emb_ids=spk_emb_ids, uniforms=u, return_codes=True):
            pass
        wavs.extend(out[0])
        results.append(out[1])
    dev = wavs[0].device if wavs and wavs[0].is_cuda else None
    all_lengths = gather_lengths([int(w.numel()) for w in wavs], device=dev)
    return (lo, hi, wavs, all_lengths, results) if return_codes else (lo, hi, wavs, all_lengths)
