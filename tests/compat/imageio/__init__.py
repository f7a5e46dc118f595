"""Test-only stand-in for `imageio` (absent from this image): tools/util.py of the reference imports it at module level
for video I/O helpers that the compositing test never calls."""


def __getattr__(name):
    raise NotImplementedError(f"imageio.{name}: video I/O is outside the hot path and not available in this image")
