from .._core import ModelMixin  # noqa: F401
