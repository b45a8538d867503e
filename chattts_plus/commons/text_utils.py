"""Import shim: ``chattts_plus.commons.text_utils`` (reference commons/text_utils.py) -> chatttsplus_b200.text."""
from chatttsplus_b200.text import (get_lang, num2text, num_to_english, remove_brackets, split_text,  # noqa: F401
                                   split_text_by_punctuation)
