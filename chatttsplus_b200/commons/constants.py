"""Directory constants (reference ``chattts_plus/commons/constants.py:9-16``; same env overrides)."""
import os

CURRENT_DIR = os.path.dirname(os.path.abspath(__file__))
PROJECT_DIR = os.path.abspath(os.environ.get("CHATTTS_PLUS_PROJECT_DIR", os.path.join(CURRENT_DIR, "..", "..")))
CHECKPOINT_DIR = os.path.abspath(
    os.environ.get("CHATTTS_PLUS_CHECKPOINT_DIR", os.path.join(PROJECT_DIR, "checkpoints"))
)
LOG_DIR = os.path.abspath(os.environ.get("CHATTTS_PLUS_LOG_DIR", os.path.join(PROJECT_DIR, "logs")))
