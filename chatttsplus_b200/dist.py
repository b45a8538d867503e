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
