"""mvae_b200 - B200-native MVAE training step."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
