"""text2human_b200 — B200-native (sm_100a) implementation of the Text2Human
VQGAN encode/quantize/decode and index-prediction transformer hot path.

Host code is Python over a C-ABI CUDA library (libt2h.so, include/t2h.h).
"""
from .ops import set_precision, get_terms  # noqa: F401

__all__ = ["set_precision", "get_terms"]
