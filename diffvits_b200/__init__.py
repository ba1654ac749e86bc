"""Import shim: the package sources live in `diff-vits_b200/` (the project's directory name is not a valid Python
identifier), this makes them importable as `diffvits_b200`."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))), "diff-vits_b200")
if not _os.path.isdir(_real):  # pragma: no cover
    raise ImportError(f"diffvits_b200: package directory {_real} is missing")
__path__.insert(0, _real)

from ._pkg import *  # noqa: F401,F403,E402
from ._pkg import LAZY as _LAZY, __all__ as _eager, lazy as _lazy  # noqa: E402

__all__ = list(_eager) + list(_LAZY)


def __getattr__(name):
    if name in _LAZY:
        return _lazy(name)
    raise AttributeError(f"module 'diffvits_b200' has no attribute {name!r}")
