from .._core import LoRACompatibleConv, LoRACompatibleLinear  # noqa: F401
