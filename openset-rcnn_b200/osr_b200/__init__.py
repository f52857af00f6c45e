"""osr_b200 - B200-native (sm_100a) RoI hot path of Openset-RCNN behind the reference's operator surface.

Host code is Python (like the reference); all compute goes through ``libosr_sm100a.so``
(C ABI in ``include/osr.h``) - there is no CPU / eager fallback.
"""
from . import _lib  # noqa: F401
from .structures import Boxes, Instances  # noqa: F401

__all__ = ["Boxes", "Instances"]
