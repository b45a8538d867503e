from chatttsplus_b200.gpt import GPT  # noqa: F401
from chatttsplus_b200.tokenizer import Tokenizer  # noqa: F401
from chatttsplus_b200.vocoder import DVAE, Vocos  # noqa: F401
from chatttsplus_b200 import processors  # noqa: F401
