"""Host text front-end of the pipeline: CPU string processing that runs before tokenisation.

Restates, with the reference's exact observable behaviour (pinned by ``tests/golden/text_ref.json``, produced by running the
reference's own functions):

  * ``chattts_plus/commons/text_utils.py`` — ``num_to_english`` (:8-67), ``get_lang`` (:70-76), ``num2text`` (:87-114),
    ``remove_brackets`` (:117-123), ``split_text`` (:127-157), ``split_text_by_punctuation`` (:160-189)
  * ``chattts_plus/commons/norm.py`` — ``Normalizer`` (:36-209): language detection, registered normalisers, half-width ->
    full-width map for Chinese, invalid-character handling, homophone replacement (a numba loop over UTF-16 code units in the
    reference; a ``str.translate`` table here — same result, the table is the same ``homophones_map.json``).

Quirks of the reference are kept because callers see them: ``num_to_english`` raises ``IndexError`` for a group whose last two
digits are ``10``; ``num2text`` replaces a leftover ``7`` by ``"seven"`` without the spaces the other digits get;
``remove_brackets`` passes regex flags in the ``count`` position (at most 26 control tags are rewritten).

Optional third-party normalisers: Chinese text uses ``zh_normalization.TextNormalizer`` when that package is importable and
English text ``nemo_text_processing`` (the reference falls back to ``num2text`` when nemo fails; so does this module).  Neither
package is in this image: Chinese input then passes through un-normalised with a warning instead of an ImportError.
"""
from __future__ import annotations

import json
import os
import re
import sys
from typing import Callable, Dict, List, Optional, Set

from .commons import logger as _logger

_DIGIT_WORDS = ("zero", "one", "two", "three", "four", "five", "six", "seven", "eight", "nine")
_TEEN_WORDS = {11: "eleven", 12: "twelve", 13: "thirteen", 14: "fourteen", 15: "fifteen", 16: "sixteen", 17: "seventeen",
               18: "eighteen", 19: "nineteen"}
_TENS_WORDS = ("", "", "twenty", "thirty", "forty", "fifty", "sixty", "seventy", "eighty", "ninety")
_GROUP_WORDS = ("", "thousand", "million", "billion", "trillion")


def _group_to_words(group: str) -> str:
    """One group of up to three decimal digits (text_utils.py:31-56)."""
    value = int(group)
    rem = value % 100 if len(group) >= 2 else value
    words = ""
    if len(group) == 3 and value // 100:
        words = _DIGIT_WORDS[value // 100] + " hundred"
        if rem:
            words += " and "
    if 10 < rem < 20:
        return words + _TEEN_WORDS[rem]
    tens, ones = divmod(rem, 10)
    if tens >= 2:
        return words + _TENS_WORDS[tens] + (" " + _DIGIT_WORDS[ones] if ones else "")
    if rem:
        return words + _DIGIT_WORDS[rem]          # rem == 10 -> IndexError, exactly like the reference
    return words


def num_to_english(num) -> str:
    """text_utils.py:8-67: British-style reading of a non-negative integer ("One hundred and five"); 0 reads as ""."""
    digits = str(num)
    groups = [digits[max(0, end - 3):end] for end in range(len(digits), 0, -3)][::-1]
    out = ""
    seen_group = False      # a group has been emitted or inspected as non-skipped
    seen_nonzero = False
    last = len(groups) - 1
    for i, grp in enumerate(groups):
        if int(grp) == 0 and i < last:
            continue
        words = _group_to_words(grp)
        if words and seen_group and seen_nonzero:
            out += " and "
        out += words
        if i < last and int(grp) != 0:
            out += " " + _GROUP_WORDS[last - i] + ", "
        seen_group = True
        seen_nonzero = seen_nonzero or int(grp) != 0
    return out.capitalize()


_ZH_PUNCT = re.compile("[。？！，、；：‘’“”（）《》【】…—　]")
_ZH_CHAR = re.compile("[一-鿿]")


def get_lang(text: str) -> str:
    """text_utils.py:70-76"""
    return "zh" if _ZH_CHAR.search(_ZH_PUNCT.sub("", text)) is not None else "en"


def num2text(text: str) -> str:
    """text_utils.py:87-114: digits and simple arithmetic signs to English words."""
    spoken = [f" {w} " for w in _DIGIT_WORDS]
    text = re.sub(r"(\d)\,(\d)", r"\1\2", text)
    text = re.sub(r"(\d+)\s*\+", r"\1 plus ", text)
    text = re.sub(r"(\d+)\s*\-", r"\1 minus ", text)
    text = re.sub(r"(\d+)\s*[\*x]", r"\1 times ", text)
    text = re.sub(r"((?:\d+\.)?\d+)\s*/\s*(\d+)", lambda m: m.group(1) + " over " + m.group(2), text)
    for whole, int_part, frac_part in re.findall(r"((\d+)(?:\.(\d+))?%?)", text):
        if len(int_part) > 16:
            continue
        words = num_to_english(int_part)
        if frac_part:
            words += " point " + "".join(spoken[int(d)] for d in frac_part)
        if whole[-1] == "%":
            words = f" the pronunciation of  {words}"
        text = text.replace(whole, words)
    for d in "123456":
        text = text.replace(d, spoken[int(d)])
    text = text.replace("7", "seven")               # (sic: no spaces in the reference)
    for d in "890":
        text = text.replace(d, spoken[int(d)])
    return text.replace("=", " equals ")


_TAG_COUNT = int(re.I | re.S | re.M)                # the reference passes the flags as `count`


def remove_brackets(text: str) -> str:
    """text_utils.py:117-123: control tags survive, every other bracket (and a few full-width marks) is dropped."""
    text = re.sub(r"\[(uv_break|laugh|lbreak|break)\]", r" \1 ", text, _TAG_COUNT)
    text = re.sub(r"\[|\]|！|：|｛|｝", "", text)
    return re.sub(r"\s(uv_break|laugh|lbreak|break)(?=\s|$)", r" [\1] ", text)


_SPLIT_MARKS = frozenset("。？！，、；：”’》」』）】…—" ".?!,:;)}…")


def split_text_by_punctuation(text: str, min_length: int = 150) -> List[str]:
    """text_utils.py:160-189: cut after a punctuation mark once the running piece is longer than 150 characters."""
    pieces: List[str] = []
    start = 0
    n = len(text)
    for i, ch in enumerate(text):
        if ch not in _SPLIT_MARKS:
            continue
        if ch == "." and i < n - 1 and re.match(r"\d", text[i + 1]):
            continue                                 # decimal point
        if i - start > min_length:
            pieces.append(text[start:i + 1])
            start = i + 1
    if start < n:
        pieces.append(text[start:])
    return pieces


_warned_zh = False


def _normalize_zh(text: str) -> str:
    global _warned_zh
    try:
        from zh_normalization import TextNormalizer   # type: ignore
    except Exception:
        if not _warned_zh:
            _logger.get_logger("text").warning("zh_normalization is not installed: Chinese text is passed on without number/date normalisation")
            _warned_zh = True
        return text
    return "".join(TextNormalizer().normalize(text))


def _normalize_en(text: str, state: dict) -> str:
    if not state.get("nemo_failed"):
        try:
            from nemo_text_processing.text_normalization.normalize import Normalizer as _Nemo   # type: ignore
            return _Nemo(input_case="cased", lang="en").normalize(text, verbose=False, punct_post_process=True)
        except Exception:
            state["nemo_failed"] = True               # text_utils.py:147-151: fall back for the rest of the call
    return num2text(text)


def split_text(text_list: List[str]) -> List[str]:
    """text_utils.py:127-157: per line — strip brackets, normalise numbers (by language), split lines longer than 200."""
    state: dict = {}
    out: List[str] = []
    for text in text_list:
        text = remove_brackets(text)
        norm = _normalize_zh(text) if get_lang(text) == "zh" else _normalize_en(text, state)
        if len(norm) > 200:
            out.extend(split_text_by_punctuation(norm))
        else:
            out.append(norm)
    return out


def merge_short_texts(pieces: List[str]) -> List[str]:
    """chattts_plus_pipeline.py:359-373: pieces shorter than 30 characters are glued together with [uv_break]."""
    merged: List[str] = []
    pending = ""
    for it in pieces:
        if len(it) < 30:
            pending += f"{it} [uv_break] "
            if len(pending) > 30:
                merged.append(pending)
                pending = ""
        else:
            merged.append(pending + it)
            pending = ""
    if len(pending) > 30 or len(merged) < 1:
        merged.append(pending)
    elif pending:
        merged[-1] += f" [uv_break] {pending}"
    return merged


# norm.py:66-127 — the two character tables, written as aligned strings (source -> target)
_SIMPLIFY = str.maketrans("：；！（）【】『』「」《》－:;!()><-",
                          "，，。，，，，，，，，，，，,,.,,,,,")
_HALF2FULL = str.maketrans("!\"'#$%&(),-*+./:;<=>?@\\^`{|}~",
                           "！“‘＃＄％＆（），－＊＋。／：；＜＝＞？＠＼＾｀｛｜｝～")


class Normalizer:
    """norm.py:36-209.  ``Normalizer(map_path)(text, do_text_normalization, do_homophone_replacement, lang)``."""

    def __init__(self, map_file_path: Optional[str] = None, logger=None):
        self.logger = logger or _logger.get_logger(self.__class__.__name__)
        self.normalizers: Dict[str, Callable[[str], str]] = {}
        self.homophones_map = self._load_homophones_map(map_file_path) if map_file_path else {}
        self.coding = "utf-16-le" if sys.byteorder == "little" else "utf-16-be"
        self.reject_pattern = re.compile(r"[^一-鿿A-Za-z，。、,\. ]")
        self.sub_pattern = re.compile(r"\[uv_break\]|\[laugh\]|\[lbreak\]")
        self.chinese_char_pattern = _ZH_CHAR
        self.english_word_pattern = re.compile(r"\b[A-Za-z]+\b")

    @staticmethod
    def _load_homophones_map(map_file_path: str) -> Dict[int, int]:
        if not os.path.exists(map_file_path):
            return {}
        with open(map_file_path, "r", encoding="utf-8") as f:
            table = json.load(f)
        # the reference replaces UTF-16 code units (norm.py:21-33): only BMP characters can match, so can str.translate here
        return {ord(k): ord(v) for k, v in table.items() if len(k) == 1 and len(v) == 1 and ord(k) < 0x10000}

    def __call__(self, text: str, do_text_normalization=True, do_homophone_replacement=True, lang: Optional[str] = None) -> str:
        if do_text_normalization:
            _lang = self._detect_language(text) if lang is None else lang
            if _lang in self.normalizers:
                text = self.normalizers[_lang](text)
            if _lang == "zh":
                text = text.translate(_HALF2FULL)
        invalid = self._count_invalid_characters(text)
        if invalid:
            self.logger.debug(f"found invalid characters: {invalid}")
            text = text.translate(_SIMPLIFY)
        if do_homophone_replacement and self.homophones_map:
            replaced = text.translate(self.homophones_map)
            if replaced != text:
                self.logger.debug("replace homophones: " + ", ".join(f"{a}->{b}" for a, b in zip(text, replaced) if a != b))
                text = replaced
        if invalid:
            text = self.reject_pattern.sub("", text)
        return text

    def register(self, name: str, normalizer: Callable[[str], str]) -> bool:
        if name in self.normalizers:
            self.logger.warning(f"name {name} has been registered")
            return False
        try:
            if not isinstance(normalizer("test string 测试字符串"), str):
                self.logger.warning("normalizer must have caller type (str) -> str")
                return False
        except Exception as e:  # noqa: BLE001 - the reference logs and refuses any failing normaliser
            self.logger.warning(e)
            return False
        self.normalizers[name] = normalizer
        return True

    def unregister(self, name: str):
        self.normalizers.pop(name, None)

    def destroy(self):
        self.homophones_map = {}

    def _count_invalid_characters(self, s: str) -> Set[str]:
        return set(self.reject_pattern.findall(self.sub_pattern.sub("", s)))

    def _detect_language(self, sentence: str) -> str:
        return "zh" if len(self.chinese_char_pattern.findall(sentence)) > len(self.english_word_pattern.findall(sentence)) else "en"
