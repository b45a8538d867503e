"""``from omegaconf import OmegaConf`` for the reference's callers (webui.py:461, tests/test_pipelines.py:16-19).

The repository root comes first on ``sys.path`` when those scripts run from it, so this package would shadow an installed
``omegaconf``.  It therefore looks for the real distribution on the rest of ``sys.path`` first and, when there is one, loads it
IN PLACE of itself (interpolation, ``merge``, ``MISSING`` ... all stay available to third-party importers); only when the
package is missing — as in this image — does it fall back to the PyYAML-based stand-in
(chatttsplus_b200/commons/omegaconf_lite.py), which covers what the pipeline uses: ``OmegaConf.load``, attribute / item access,
``in`` and in-place mutation (chattts_plus_pipeline.py:63-67,70,78).
"""
import importlib.machinery as _machinery
import importlib.util as _util
import os as _os
import sys as _sys

_here = _os.path.dirname(_os.path.abspath(__file__))
_root = _os.path.dirname(_here)
_paths = [p for p in _sys.path if _os.path.abspath(p or ".") != _root]
_spec = _machinery.PathFinder.find_spec("omegaconf", _paths)
if _spec is not None and _spec.origin and _os.path.dirname(_os.path.abspath(_spec.origin)) != _here:
    _real = _util.module_from_spec(_spec)
    _sys.modules[__name__] = _real          # importers (and submodule imports) now see the real package
    _spec.loader.exec_module(_real)
else:
    from chatttsplus_b200.commons.omegaconf_lite import DictConfig, OmegaConf  # noqa: F401
