"""Drop-in import surface: the reference's callers import ``chattts_plus.pipelines.chattts_plus_pipeline``,
``chattts_plus.commons.{utils,constants}`` and ``chattts_plus.models`` (webui.py:17-18,49; tests/test_pipelines.py:14-15).
Everything here re-exports the B200-native implementation in ``chatttsplus_b200``."""
