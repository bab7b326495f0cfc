"""ssv_b200 — B200-native (sm_100a) drop-in for the loss layer of NightShade99/Self-Supervised-Vision.

Same class names, constructor kwargs and forward signatures as the reference's `utils/losses.py`
and the ring buffers of `models/moco.py` / `models/swav.py`; the arithmetic runs in hand-written
CUDA kernels behind the C ABI of include/ssv_b200.h.  No CPU fallback.
"""
from . import _cabi  # noqa: F401
from .losses import (BarlowLoss, DinoLoss, MocoLoss, MSELoss, PirlLoss, RelicLoss, SimclrLoss, SimSiamLoss, SwavLoss,  # noqa: F401
                     update_teacher_center)
from .ema import EmaUpdater, momentum_update  # noqa: F401
from .banks import FeatureBank, MemoryBank, PirlMemoryBank, Prototypes  # noqa: F401
from .fused import NormalizedMSELoss, NormalizedSimSiamLoss, SelaLabeler, graphed  # noqa: F401

__all__ = ["SimclrLoss", "MocoLoss", "BarlowLoss", "SimSiamLoss", "RelicLoss", "SwavLoss", "MSELoss",
           "MemoryBank", "FeatureBank", "Prototypes", "DinoLoss", "PirlLoss", "PirlMemoryBank", "update_teacher_center", "EmaUpdater", "momentum_update",
           "NormalizedMSELoss", "NormalizedSimSiamLoss", "SelaLabeler", "graphed"]
