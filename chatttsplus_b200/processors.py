"""Host-side description of the sampling processors (mirror of reference chattts_plus/models/processors.py).

In the reference these objects *execute* the logit transforms in ~25 PyTorch kernels per step
(processors.py:18-34, transformers TopP/TopK warpers).  Here they only carry parameters: the transforms run
inside the fused CUDA sampler (csrc/gpt_kernels.cuh: k_sample).  ``gen_logits`` keeps the reference's
signature and return shape (``(logits_warpers, logits_processors)``, processors.py:37-57), and GPT.generate
also accepts transformers' own TopP/TopK instances by duck-typing their attributes.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple


class CustomRepetitionPenaltyLogitsProcessorRepeat:
    def __init__(self, penalty: float, max_input_ids: int, past_window: int):
        if not isinstance(penalty, float) or not (penalty > 0):
            raise ValueError(f"`penalty` has to be a strictly positive float, but is {penalty}")
        self.penalty = penalty
        self.max_input_ids = max_input_ids
        self.past_window = past_window


@dataclass
class TopPLogitsWarper:
    top_p: float
    min_tokens_to_keep: int = 1

    def __post_init__(self):
        self.top_p = float(self.top_p)
        if self.top_p < 0 or self.top_p > 1.0:
            raise ValueError(f"`top_p` has to be a float > 0 and < 1, but is {self.top_p}")


@dataclass
class TopKLogitsWarper:
    top_k: int
    min_tokens_to_keep: int = 1

    def __post_init__(self):
        if not isinstance(self.top_k, int) or self.top_k <= 0:
            raise ValueError(f"`top_k` has to be a strictly positive integer, but is {self.top_k}")
        self.top_k = max(self.top_k, self.min_tokens_to_keep)


def gen_logits(num_code: int, top_P=0.7, top_K=20, repetition_penalty=1.0):
    logits_warpers = []
    if top_P is not None:
        logits_warpers.append(TopPLogitsWarper(top_P, min_tokens_to_keep=3))
    if top_K is not None:
        logits_warpers.append(TopKLogitsWarper(top_K, min_tokens_to_keep=3))
    logits_processors = []
    if repetition_penalty is not None and repetition_penalty != 1:
        logits_processors.append(CustomRepetitionPenaltyLogitsProcessorRepeat(repetition_penalty, num_code, 16))
    return logits_warpers, logits_processors


@dataclass
class SamplerParams:
    """Flattened parameters for the fused sampler (ctp_sample_cfg)."""
    top_p: float = 0.0       # 0 disables
    top_k: int = 20
    min_keep: int = 3
    rep_penalty: float = 1.0
    rep_window: int = 16
    rep_max_ids: int = 1 << 30


def flatten(logits_warpers, logits_processors) -> SamplerParams:
    p = SamplerParams()
    saw_k = False
    for w in logits_warpers or []:
        if hasattr(w, "top_p"):
            p.top_p = float(w.top_p)
            p.min_keep = int(getattr(w, "min_tokens_to_keep", p.min_keep))
        elif hasattr(w, "top_k"):
            p.top_k = int(w.top_k)
            saw_k = True
            p.min_keep = max(p.min_keep, int(getattr(w, "min_tokens_to_keep", 1)))
        else:
            raise TypeError(f"unsupported logits warper {type(w).__name__}: the fused sampler implements TopP and TopK")
    if not saw_k:
        p.top_k = 0   # no TopK warper (gen_logits(top_K=None), processors.py:43-47): only TopP / min_keep limit the survivors
    for q in logits_processors or []:
        if hasattr(q, "penalty") and hasattr(q, "past_window"):
            p.rep_penalty = float(q.penalty)
            p.rep_window = int(q.past_window)
            p.rep_max_ids = int(q.max_input_ids)
        else:
            raise TypeError(f"unsupported logits processor {type(q).__name__}")
    return p
