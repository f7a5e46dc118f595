"""reference: src/models/unet_3d_edit_bkfill.py (UNet3DConditionModel :30-682) -> engine-backed implementation."""
from mimo_b200.host.modules import UNet3DConditionModel, UNet3DConditionOutput  # noqa: F401
