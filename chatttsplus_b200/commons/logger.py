"""Minimal ``get_logger`` (the reference's coloured/file logger, commons/logger.py, is out of scope)."""
import logging
import os

_FMT = "%(asctime)s %(name)s %(levelname)s: %(message)s"


def get_logger(name: str = "chattts_plus", level=None) -> logging.Logger:
    logger = logging.getLogger(name)
    if not logger.handlers:
        h = logging.StreamHandler()
        h.setFormatter(logging.Formatter(_FMT, "%H:%M:%S"))
        logger.addHandler(h)
        logger.propagate = False
    lvl = level or os.environ.get("CHATTTS_PLUS_LOG_LEVEL", "INFO")
    logger.setLevel(lvl)
    return logger
