"""bwt-algorithm_b200: B200-native (sm_100a) index-and-scan hot path behind the
reference's ``bwt`` module surface.  ``from bwt_algorithm_b200 import bwt``."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
