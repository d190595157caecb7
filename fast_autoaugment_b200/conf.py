"""The configuration singleton the reference reads through ``theconf`` (``from theconf import Config as C``,
reference data.py:16, train.py:20): ``C.get()['aug']``, ``['cutout']``, ``['model']['type']``, ``['mixup']`` ...

If the real ``theconf`` package is importable its ``Config`` is used, so a caller that configured the
reference (``ConfigArgumentParser``) configures this package as well.  Otherwise a minimal stand-in with the
same ``Config.get()`` accessor is provided; fill it with ``Config.get().update({...})``.
"""
from __future__ import annotations

try:                                     # pragma: no cover - theconf is not installed in the build image
    from theconf import Config           # noqa: F401
except Exception:
    class _Conf(dict):
        @property
        def conf(self):                  # reference train.py:51 reads C.get().conf.get('mixup', 0.0)
            return self

    class Config:                        # noqa: D401 - mirrors theconf.Config's accessor
        _instance = None

        @classmethod
        def get(cls):
            if cls._instance is None:
                cls._instance = _Conf()
            return cls._instance
