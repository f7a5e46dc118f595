"""reference: src/models/pose_guider.py (PoseGuider :12-57) -> engine-backed implementation."""
from mimo_b200.host.modules import PoseGuider  # noqa: F401
