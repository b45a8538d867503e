from chatttsplus_b200.commons.constants import *  # noqa: F401,F403
from chatttsplus_b200.commons.constants import CHECKPOINT_DIR, LOG_DIR, PROJECT_DIR  # noqa: F401
