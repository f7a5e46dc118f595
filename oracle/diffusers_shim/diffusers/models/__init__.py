from .._core import ModelMixin  # noqa: F401
from .._pipeline import AutoencoderKL  # noqa: F401
