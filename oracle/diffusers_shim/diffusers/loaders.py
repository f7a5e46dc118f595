from ._core import UNet2DConditionLoadersMixin  # noqa: F401
