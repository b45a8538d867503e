from .chattts_plus_pipeline import ChatTTSPlusPipeline  # noqa: F401
