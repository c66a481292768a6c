"""Stub for matplotlib.collections (see package docstring)."""
from . import _Anything


def __getattr__(name):
    return _Anything()
