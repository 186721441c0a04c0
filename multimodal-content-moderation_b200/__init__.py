"""B200-native scoring path for amirhossein-yousefi/multimodal-content-moderation.

Public surface (mirrors `src.models` of the reference, R/src/models/__init__.py:3-6):

    MultiModalFusionClassifier, MultiTaskClassifier      drop-in nn.Modules (modules.py)
    Engine                                               thin owner of the C handle (engine.py)
    lib                                                  ctypes binding of include/mmcm.h (lib.py)

Importing this package never touches CUDA; the extension is loaded (and, if stale, rebuilt) on first use.
"""
from . import arch, build, checkpoint, lib, prepost, sharding, synthetic  # noqa: F401
from .engine import Engine  # noqa: F401
from .checkpoint import PackedScorer, load_checkpoint  # noqa: F401
from .pipeline import BatchedScorer  # noqa: F401
from .tokenizer import ClipTokenizer  # noqa: F401
from .modules import FocalWithLogitsLoss, MultiModalFusionClassifier, MultiTaskClassifier  # noqa: F401

__all__ = ["MultiModalFusionClassifier", "MultiTaskClassifier", "FocalWithLogitsLoss", "Engine", "BatchedScorer", "ClipTokenizer", "PackedScorer", "load_checkpoint", "arch", "build",
           "checkpoint", "lib", "prepost", "sharding", "synthetic"]
