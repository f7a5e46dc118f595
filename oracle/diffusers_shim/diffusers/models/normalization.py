from .._core import AdaLayerNormSingle  # noqa: F401
