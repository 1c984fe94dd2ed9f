"""Public surface of the package (re-exported by the `tedspad_b200` import shim)."""
from . import _lib, ops  # noqa: F401

__all__ = ["_lib", "ops"]
