"""Sliding context windows over the frame axis — same generator as src/pipelines/context.py:7-49 (pure integer
logic, must be bit-exact; pinned by tests/golden/integer_tables.json, which the reference itself produced)."""
from __future__ import annotations

from typing import Callable, Iterator, List, Optional

import numpy as np


def ordered_halving(val: int) -> float:
    """Bit-reversal of a 64-bit integer, as a fraction in [0, 1)."""
    rev = 0
    for _ in range(64):
        rev = (rev << 1) | (val & 1)
        val >>= 1
    return rev / (1 << 64)


def uniform(step: int = ..., num_steps: Optional[int] = None, num_frames: int = ..., context_size: Optional[int] = None,
            context_stride: int = 3, context_overlap: int = 4, closed_loop: bool = True) -> Iterator[List[int]]:
    if num_frames <= context_size:
        yield list(range(num_frames))
        return
    context_stride = min(context_stride, int(np.ceil(np.log2(num_frames / context_size))) + 1)
    frac = ordered_halving(step)
    for k in range(context_stride):
        context_step = 1 << k
        pad = int(round(num_frames * frac))
        start = int(frac * context_step) + pad
        stop = num_frames + pad + (0 if closed_loop else -context_overlap)
        stride = context_size * context_step - context_overlap
        for j in range(start, stop, stride):
            yield [e % num_frames for e in range(j, j + context_size * context_step, context_step)]


def get_context_scheduler(name: str) -> Callable:
    if name == "uniform":
        return uniform
    raise ValueError(f"Unknown context_overlap policy {name}")


def get_total_steps(scheduler, timesteps: List[int], num_steps: Optional[int] = None, num_frames: int = ...,
                    context_size: Optional[int] = None, context_stride: int = 3, context_overlap: int = 4,
                    closed_loop: bool = True) -> int:
    """Number of windows over a whole sampling run (src/pipelines/context.py:52-75): the scheduler is evaluated at step
    index i for every timestep; like the reference, `closed_loop` is accepted but not forwarded."""
    total = 0
    for i in range(len(timesteps)):
        total += sum(1 for _ in scheduler(i, num_steps, num_frames, context_size, context_stride, context_overlap))
    return total
