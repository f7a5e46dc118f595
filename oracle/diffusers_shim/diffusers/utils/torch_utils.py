from .._core import apply_freeu, randn_tensor  # noqa: F401
