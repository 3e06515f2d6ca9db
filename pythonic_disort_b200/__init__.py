"""pythonic_disort_b200 -- B200-native, column-batched PythonicDISORT solver hot path.

Public surface mirrors the reference package (src/PythonicDISORT/__init__.py:1-2):
``pydisort`` and ``subroutines``.
"""
from . import subroutines  # noqa: F401


def pydisort(*args, **kwargs):
    from .api import pydisort as _impl
    return _impl(*args, **kwargs)


def __getattr__(name):  # HenyeyGreenstein, LevelSource: inputs expanded on the device (inputs.py)
    if name in ("HenyeyGreenstein", "LevelSource"):
        from . import inputs
        return getattr(inputs, name)
    raise AttributeError(name)


__all__ = ["pydisort", "subroutines", "HenyeyGreenstein", "LevelSource"]
