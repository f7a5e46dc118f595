from .._core import DualTransformer2DModel  # noqa: F401
