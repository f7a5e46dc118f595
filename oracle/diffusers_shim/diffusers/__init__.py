"""oracle/diffusers_shim: see _core.py. Import path mirrors diffusers==0.24.0 for the symbols the reference touches."""
from ._core import ConfigMixin, ModelMixin  # noqa: F401
from ._pipeline import AutoencoderKL, DDIMScheduler, DiffusionPipeline  # noqa: F401

__version__ = "0.24.0+oracle-shim"
