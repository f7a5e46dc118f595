"""reference: src/pipelines/pipeline_pose2vid_long_edit_bkfill_roiclip.py (Pose2VideoPipeline :36-578)."""
from mimo_b200.host.pipeline import Pose2VideoPipeline, Pose2VideoPipelineOutput  # noqa: F401
