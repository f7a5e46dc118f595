"""reference: src/pipelines/context.py (uniform / ordered_halving / get_context_scheduler)."""
from mimo_b200.host.context import get_context_scheduler, ordered_halving, uniform  # noqa: F401
