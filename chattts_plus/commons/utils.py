from chatttsplus_b200.commons.utils import *  # noqa: F401,F403
from chatttsplus_b200.commons.utils import InferCodeParams, RefineTextParams, TorchSeedContext, get_inference_device  # noqa: F401
