"""reference: src/pipelines/context.py (uniform / ordered_halving / get_context_scheduler / get_total_steps)."""
from mimo_b200.host.context import get_context_scheduler, get_total_steps, ordered_halving, uniform  # noqa: F401
