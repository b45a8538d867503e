from chatttsplus_b200.commons.logger import get_logger  # noqa: F401
