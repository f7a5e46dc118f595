"""reference: src/models/unet_2d_condition.py (UNet2DConditionModel) -> engine-backed reference ("write") network."""
from mimo_b200.host.modules import UNet2DConditionModel  # noqa: F401
