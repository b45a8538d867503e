from chatttsplus_b200.commons import constants, logger, utils  # noqa: F401
