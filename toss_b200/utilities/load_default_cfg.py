"""load_default_cfg -- mirrors toss/utilities/load_default_cfg.py:5-14 (TOML -> DotMap(_dynamic=False))."""
import os

from .dotmap import DotMap

try:
    import toml as _toml

    def _load(path):
        return _toml.load(path)
except ImportError:   # Python >= 3.11
    import tomllib

    def _load(path):
        with open(path, "rb") as f:
            return tomllib.load(f)


def load_default_cfg(path: str = None):
    """Loads the default toml config file from the resources folder (or `path`)."""
    if path is None:
        path = os.path.join(os.path.dirname(__file__), "../resources/default_cfg.toml")
    return DotMap(_load(path), _dynamic=False)
