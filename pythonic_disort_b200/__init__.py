"""pythonic_disort_b200 -- B200-native, column-batched PythonicDISORT solver hot path.

Public surface mirrors the reference package (src/PythonicDISORT/__init__.py:1-2):
``pydisort`` and ``subroutines``.
"""
from . import subroutines  # noqa: F401


def pydisort(*args, **kwargs):
    from .api import pydisort as _impl
    return _impl(*args, **kwargs)


__all__ = ["pydisort", "subroutines"]
