"""``Tokenizer`` — host-side mirror of reference ``chattts_plus/models/tokenizer.py`` (boundary row A2/A7).

Wraps the pickled ``BertTokenizerFast`` of ``asset/tokenizer.pt`` (tokenizer.py:20-48), builds left-padded
``[B, L, num_vq]`` ids + masks (tokenizer.py:50-137), applies the speaker embedding (tokenizer.py:150-178) and
provides the b14+LZMA codecs for speaker embeddings / audio prompts (tokenizer.py:139-148,180-222) on top of
``commons.b14`` (``pybase16384`` is not installed in this image).
"""
from __future__ import annotations

import lzma
import os
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

from .commons import b14
from .commons import logger as _logger

os.environ.setdefault("TOKENIZERS_PARALLELISM", "false")
_LZMA_FILTERS = [{"id": lzma.FILTER_LZMA2, "preset": 9 | lzma.PRESET_EXTREME}]


class Tokenizer:
    def __init__(self, model_path=None, tokenizer=None, **kwargs):
        self.logger = _logger.get_logger(self.__class__.__name__)
        if tokenizer is None:
            self.logger.info(f"loading Tokenizer pretrained model: {model_path}")
            tokenizer = torch.load(model_path, map_location="cpu", mmap=True, weights_only=False)
        self._tokenizer = tokenizer
        try:
            self._tokenizer.eos_token = "[SEP]"
            self._tokenizer.pad_token = "[PAD]"
        except Exception:
            pass
        self.len = len(tokenizer)
        self.spk_emb_ids = tokenizer.convert_tokens_to_ids("[spk_emb]")
        self.break_0_ids = tokenizer.convert_tokens_to_ids("[break_0]")
        self.eos_token = tokenizer.convert_tokens_to_ids("[Ebreak]")
        self.decode = self._tokenizer.batch_decode

    @torch.inference_mode()
    def encode(self, text: List[str], num_vq: int, prompt_str: Optional[str] = None, device="cpu"
               ) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
        ids_lst, mask_lst = [], []
        prompt = self._decode_prompt(prompt_str) if prompt_str is not None else None
        prompt_size = 0
        if prompt is not None:
            assert prompt.size(0) == num_vq, "prompt dim 0 must equal to num_vq"
            prompt_size = prompt.size(1)
        for t in text:
            # tokenizer.py:69-71 calls encode_plus; transformers >= 5 dropped that alias of __call__
            enc = getattr(self._tokenizer, "encode_plus", None) or self._tokenizer
            x = enc(t, return_tensors="pt", add_special_tokens=False, padding=True)
            ids_lst.append(x["input_ids"].squeeze(0))
            mask_lst.append(x["attention_mask"].squeeze(0))
        L = max(i.size(0) for i in ids_lst) + prompt_size
        B = len(ids_lst)
        input_ids = torch.zeros(B, L, dtype=ids_lst[0].dtype)
        attention_mask = torch.zeros(B, L, dtype=mask_lst[0].dtype)
        for i in range(B):  # left padding; the audio prompt (if any) occupies the last prompt_size slots
            n = ids_lst[i].size(0)
            input_ids[i, L - prompt_size - n: L - prompt_size] = ids_lst[i]
            attention_mask[i, L - prompt_size - n: L - prompt_size] = mask_lst[i]
            if prompt_size:
                attention_mask[i, L - prompt_size:] = 1
        text_mask = attention_mask.bool()
        new_input_ids = input_ids.unsqueeze(-1).expand(-1, -1, num_vq).clone()
        if prompt_size:
            text_mask[:, L - prompt_size:] = False
            new_input_ids[:, L - prompt_size:] = prompt.t().unsqueeze(0).expand(B, -1, -1)
        return new_input_ids.to(device), attention_mask.to(device), text_mask.to(device)

    @staticmethod
    def _decode_spk_emb(spk_emb: str) -> np.ndarray:
        return np.frombuffer(lzma.decompress(b14.decode_from_string(spk_emb), format=lzma.FORMAT_RAW, filters=_LZMA_FILTERS),
                             dtype=np.float16).copy()

    @torch.no_grad()
    def apply_spk_emb(self, emb: torch.Tensor, spk_emb, input_ids: torch.Tensor, device=None):
        return apply_spk_emb(emb, spk_emb, input_ids, self.spk_emb_ids)

    @staticmethod
    @torch.no_grad()
    def _decode_prompt(prompt: str) -> torch.Tensor:
        dec = b14.decode_from_string(prompt)
        shp = np.frombuffer(dec[:4], dtype="<u2")
        p = np.frombuffer(lzma.decompress(dec[4:], format=lzma.FORMAT_RAW, filters=_LZMA_FILTERS), dtype="<u2").copy()
        return torch.from_numpy(p.astype(np.int64)).view(*[int(s) for s in shp])

    @staticmethod
    @torch.no_grad()
    def _encode_prompt(prompt: torch.Tensor) -> str:
        arr = prompt.to(device="cpu").numpy().astype("<u2")
        assert arr.ndim == 2, "prompt must be a 2D tensor"
        return b14.encode_to_string(np.array(arr.shape, dtype="<u2").tobytes()
                                    + lzma.compress(arr.tobytes(), format=lzma.FORMAT_RAW, filters=_LZMA_FILTERS))

    @staticmethod
    @torch.no_grad()
    def _encode_spk_emb(spk_emb: torch.Tensor) -> str:
        arr = spk_emb.to(dtype=torch.float16, device="cpu").numpy()
        return b14.encode_to_string(lzma.compress(arr.tobytes(), format=lzma.FORMAT_RAW, filters=_LZMA_FILTERS))


@torch.no_grad()
def apply_spk_emb(emb: torch.Tensor, spk_emb, input_ids: torch.Tensor, spk_emb_ids: int) -> torch.Tensor:
    """tokenizer.py:150-178: write the L2-normalised speaker vector (str-encoded or a 1-D tensor; a ``[1, dim]``
    tensor is accepted too, the evident intent of chattts_plus_pipeline.py:533-534) where ids[..., 0] == [spk_emb]."""
    if isinstance(spk_emb, str):
        t = torch.from_numpy(Tokenizer._decode_spk_emb(spk_emb))
    else:
        t = spk_emb
    t = t.reshape(-1).float()
    n = F.normalize(t, p=2.0, dim=0, eps=1e-12).to(emb.device, dtype=emb.dtype)
    cond = input_ids[..., 0:1].to(emb.device).eq(spk_emb_ids).expand(emb.shape)
    emb.copy_(torch.where(cond, n.view(1, 1, -1).expand(emb.shape), emb))
    return emb
