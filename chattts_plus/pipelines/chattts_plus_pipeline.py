from chatttsplus_b200.pipeline import ChatTTSPlusPipeline  # noqa: F401
