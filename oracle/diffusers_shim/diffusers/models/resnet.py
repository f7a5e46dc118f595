from .._core import Downsample2D, ResnetBlock2D, Upsample2D  # noqa: F401
