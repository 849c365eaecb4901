"""jat_b200 -- B200-native (sm_100a) implementation of the JaT-AudioSR DiT denoiser hot path.

Import name: ``jat_b200`` (the on-disk directory carries the reference's full name and is aliased by
the top-level ``jat_b200.py`` shim, because hyphens are not importable).
"""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
__version__ = "0.1.0"
