"""B200-native (sm_100a) implementation of the SZN pixel-embedding hot path.

``models.FCN32s`` and ``utils.*`` mirror the reference's ``models.py`` / ``utils.py`` surface
(RohanDoshi2018/ZeroshotSemanticSegmentation) so its trainers' call sequence runs unchanged; all
device work is done by hand-written CUDA kernels in ``libszn.so`` (C ABI in ``include/szn.h``).
"""
from . import models, utils  # noqa: F401
from .models import FCN32s  # noqa: F401

__all__ = ["models", "utils", "FCN32s"]
