from ._core import ConfigMixin, FrozenDict, register_to_config  # noqa: F401
