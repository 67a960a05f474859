"""B200-native inference engine for UniMedVL's unified understanding + generation forward path.

Public surface (mirrors the reference's objects): ``Engine`` (handle on the CUDA library),
``Bagel`` (codes/modeling/unimedvl/bagel.py:Bagel), ``AutoEncoder`` (codes/modeling/autoencoder.py),
``NaiveCache`` (codes/modeling/unimedvl/qwen2_navit.py), ``InterleaveInferencer`` (codes/inferencer.py),
``ImageTransform`` (codes/data/transforms.py); ``BatchedInferencer`` (the same workflows packed over a batch of requests);
``ContinuousBatcher`` (request scheduler over the paged KV, no reference
counterpart).  Importing works without a GPU; creating an ``Engine`` does not.
"""
from . import checkpoint, config  # noqa: F401
from .autoencoder import AutoEncoder  # noqa: F401
from .bagel import Bagel  # noqa: F401
from .batched import BatchedInferencer  # noqa: F401
from .cache import NaiveCache  # noqa: F401
from .engine import Engine  # noqa: F401
from .inferencer import InterleaveInferencer  # noqa: F401
from .packing import ImageTransform  # noqa: F401
from .scheduler import ContinuousBatcher  # noqa: F401

__all__ = ["checkpoint", "config", "Engine", "Bagel", "AutoEncoder", "NaiveCache", "InterleaveInferencer", "BatchedInferencer", "ImageTransform",
           "ContinuousBatcher"]
