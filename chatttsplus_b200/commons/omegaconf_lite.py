"""A tiny ``OmegaConf.load`` stand-in on top of PyYAML.

The reference's callers load ``configs/infer/*.yaml`` with ``omegaconf.OmegaConf.load`` (webui.py:461,
tests/test_pipelines.py:16-19) and the pipeline uses attribute access, item access, ``in`` and in-place
mutation on the result (chattts_plus_pipeline.py:63-67,70,78).  ``omegaconf`` is not installed in this image,
so the repo-root ``omegaconf`` shim re-exports this class when the real package is missing.
"""
from __future__ import annotations

import yaml


class DictConfig(dict):
    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = _wrap(v)

    def __setitem__(self, k, v):
        super().__setitem__(k, _wrap(v))

    def get(self, k, default=None):
        return self[k] if k in self else default


def _wrap(v):
    if isinstance(v, DictConfig):
        return v
    if isinstance(v, dict):
        d = DictConfig()
        for k, x in v.items():
            dict.__setitem__(d, k, _wrap(x))
        return d
    if isinstance(v, (list, tuple)):
        return [_wrap(x) for x in v]
    return v


def _unwrap(v):
    if isinstance(v, dict):
        return {k: _unwrap(x) for k, x in v.items()}
    if isinstance(v, list):
        return [_unwrap(x) for x in v]
    return v


class OmegaConf:
    @staticmethod
    def load(path) -> DictConfig:
        with open(path, "r", encoding="utf-8") as f:
            return _wrap(yaml.safe_load(f))

    @staticmethod
    def create(obj=None) -> DictConfig:
        return _wrap(obj or {})

    @staticmethod
    def to_container(cfg, resolve=True):
        return _unwrap(cfg)
