from .._core import GEGLU, AdaLayerNorm, Attention, FeedForward  # noqa: F401
