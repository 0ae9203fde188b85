"""rsis_b200 -- B200-native (sm_100a) implementation of the RSIS encoder -> ConvLSTM-decoder hot path.

Public surface mirrors the reference's `src/modules` (imatge-upc/rsis):

    from rsis_b200.modules.model import FeatureExtractor, RSIS     # model.py:15,72
    from rsis_b200.modules.clstm import ConvLSTMCell               # clstm.py:7
    from rsis_b200.inference import test                           # test.py:16

All arithmetic runs in hand-written CUDA behind the C ABI of include/rsis_b200.h
(librsis_b200.so, bound with ctypes).  There is no PyTorch/CPU fallback: modules raise if
their tensors are not on a CUDA device or the library cannot be loaded.
"""
from . import _lib, ops  # noqa: F401
from .modules.clstm import ConvLSTMCell  # noqa: F401
from .modules.model import RSIS, FeatureExtractor  # noqa: F401
from .inference import test  # noqa: F401

__all__ = ["FeatureExtractor", "RSIS", "ConvLSTMCell", "test", "ops"]
