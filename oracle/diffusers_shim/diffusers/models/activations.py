from .._core import get_activation  # noqa: F401
