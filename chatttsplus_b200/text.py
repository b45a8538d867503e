"""Minimal host text front-end (reference commons/norm.py + text_utils.py are CPU string processing, out of the
hot-path scope — SURVEY.md §2 row 11; their heavy deps zh_normalization / nemo are absent).  Keeps the call shapes
the pipeline uses: ``Normalizer(map_path)(text, do_text_normalization, do_homophone_replacement, lang)`` and
``split_text(list[str]) -> list[str]``."""
from __future__ import annotations

import json
import os
import re
from typing import List, Optional

_SENT_END = re.compile(r"(?<=[。！？!?；;\.])\s*")


def split_text(text_list: List[str], max_len: int = 200) -> List[str]:
    """Sentence-level split with a soft length cap (text_utils.py:127-189 intent)."""
    out: List[str] = []
    for t in text_list:
        t = t.strip()
        if not t:
            continue
        cur = ""
        for piece in (p for p in _SENT_END.split(t) if p):
            if len(cur) + len(piece) > max_len and cur:
                out.append(cur)
                cur = ""
            cur += piece
        if cur:
            out.append(cur)
    return out


class Normalizer:
    def __init__(self, map_file_path: Optional[str] = None, logger=None):
        self.homophones = {}
        if map_file_path and os.path.exists(map_file_path):
            with open(map_file_path, "r", encoding="utf-8") as f:
                self.homophones = {ord(k): v for k, v in json.load(f).items()}
        self.normalizers = {}
        self._reject = re.compile(r"[^一-鿿A-Za-z，。、,\. \[\]_0-9？?！!：:；;'\"-]")

    def register(self, name, fn) -> bool:
        self.normalizers[name] = fn
        return True

    def __call__(self, text: str, do_text_normalization=True, do_homophone_replacement=True, lang=None) -> str:
        if do_text_normalization:
            zh = len(re.findall(r"[一-鿿]", text)) > len(re.findall(r"\b[A-Za-z]+\b", text))
            key = lang or ("zh" if zh else "en")
            if key in self.normalizers:
                text = self.normalizers[key](text)
        if do_homophone_replacement and self.homophones:
            text = text.translate(self.homophones)
        return text
