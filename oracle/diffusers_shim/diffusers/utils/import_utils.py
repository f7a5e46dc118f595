from .._core import is_accelerate_available, is_xformers_available  # noqa: F401
