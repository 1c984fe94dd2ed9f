"""Importable alias of the `ted-spad_b200/` package directory.

The product package lives in `ted-spad_b200/` (the name the build contract fixes); a hyphen is
not a legal Python identifier, so this shim makes the same files importable as `tedspad_b200`
by pointing the package search path at that directory.
"""
import os as _os

_here = _os.path.dirname(_os.path.abspath(__file__))
__path__.insert(0, _os.path.join(_os.path.dirname(_here), "ted-spad_b200"))

from ._api import *  # noqa: E402,F401,F403
from ._api import __all__  # noqa: E402,F401
