"""Parameter dataclasses / seed context the callers of the hot path pass in.

Mirrors the public names of reference ``chattts_plus/commons/utils.py:12-58`` (same field names and defaults,
so ``webui.py`` / ``tests/test_pipelines.py`` construct them unchanged).
"""
from dataclasses import dataclass
from typing import Optional

import torch


@dataclass(repr=False, eq=False)
class RefineTextParams:
    prompt: str = ""
    top_P: float = 0.7
    top_K: int = 20
    temperature: float = 0.7
    repetition_penalty: float = 1.0
    max_new_token: int = 384
    min_new_token: int = 0
    show_tqdm: bool = True
    ensure_non_empty: bool = True


@dataclass(repr=False, eq=False)
class InferCodeParams(RefineTextParams):
    prompt: str = "[speed_5]"
    spk_emb: Optional[str] = None
    spk_smp: Optional[str] = None
    txt_smp: Optional[str] = None
    temperature: float = 0.3
    repetition_penalty: float = 1.05
    max_new_token: int = 2048
    stream_batch: int = 24
    stream_speed: int = 12000
    pass_first_n_batches: int = 2


def get_inference_device():
    if torch.cuda.is_available():
        return torch.device("cuda")
    return torch.device("cpu")


class TorchSeedContext:
    """Seeds the global torch generators for the ``with`` body (reference utils.py:48-58).

    The B200 sampler draws its uniforms from torch's CUDA generator (one ``torch.rand`` per generate call),
    so the CUDA generator state is saved / seeded / restored as well as the CPU one.
    """

    def __init__(self, seed):
        self.seed = seed
        self.state = None
        self.cuda_state = None

    def __enter__(self):
        self.state = torch.random.get_rng_state()
        if torch.cuda.is_available():
            self.cuda_state = torch.cuda.get_rng_state()
        torch.manual_seed(self.seed)

    def __exit__(self, type, value, traceback):
        torch.random.set_rng_state(self.state)
        if self.cuda_state is not None:
            torch.cuda.set_rng_state(self.cuda_state)
