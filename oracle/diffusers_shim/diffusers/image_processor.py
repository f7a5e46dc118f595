from ._pipeline import VaeImageProcessor  # noqa: F401
