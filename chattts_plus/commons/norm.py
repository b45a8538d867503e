"""Import shim: ``chattts_plus.commons.norm`` (reference commons/norm.py) -> chatttsplus_b200.text."""
from chatttsplus_b200.text import Normalizer  # noqa: F401
