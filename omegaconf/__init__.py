"""Stand-in used only when the real ``omegaconf`` package is not installed (it is absent from this image): the
reference's callers do ``from omegaconf import OmegaConf; OmegaConf.load(path)`` (webui.py:461,
tests/test_pipelines.py:16-19).  See chatttsplus_b200/commons/omegaconf_lite.py."""
from chatttsplus_b200.commons.omegaconf_lite import DictConfig, OmegaConf  # noqa: F401
