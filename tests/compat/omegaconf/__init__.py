"""Test-only stand-in for `omegaconf` (absent from this image): the two calls run_animate.py:63-98 makes —
OmegaConf.load(path) with attribute access and OmegaConf.to_container(node)."""
import yaml


class _Node(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return _wrap(v)


def _wrap(v):
    if isinstance(v, dict):
        return _Node(v)
    if isinstance(v, list):
        return [_wrap(x) for x in v]
    return v


def _plain(v):
    if isinstance(v, dict):
        return {k: _plain(x) for k, x in v.items()}
    if isinstance(v, (list, tuple)):
        return [_plain(x) for x in v]
    return v


class OmegaConf:
    @staticmethod
    def load(path):
        with open(path) as f:
            return _Node(yaml.safe_load(f))

    @staticmethod
    def to_container(node, resolve=True):
        return _plain(node)
