"""reference: src/models/mutual_self_attention.py (ReferenceAttentionControl :19-374) -> engine state handles."""
from mimo_b200.host.modules import ReferenceAttentionControl  # noqa: F401
