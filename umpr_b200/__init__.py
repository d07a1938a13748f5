"""umpr_b200 — B200-native (sm_100a) implementation of the UMPR review-network hot path.

Drop-in for the reference's ``src/model.py`` classes (see ``umpr_b200.model``); the kernels live in
``umpr_b200/libumpr_b200.so`` behind the C-ABI declared in ``include/umpr_b200.h``.
"""
from . import _lib
from .config import Config
from .model import (CNet, ControlNet, ImprovedRnn, PackedReviews, ReviewNet, RNet, SNet, SSNet, UMPR, VisualNet)
from .plan import PackPlan
from .eval import evaluate_mse

__all__ = ["Config", "ImprovedRnn", "RNet", "SNet", "CNet", "SSNet", "ReviewNet", "ControlNet", "VisualNet", "UMPR",
           "PackedReviews", "PackPlan", "evaluate_mse", "require_lib"]


def require_lib():
    """Load the CUDA library now; raises RuntimeError if it has not been built (no fallback exists)."""
    return _lib.load()
