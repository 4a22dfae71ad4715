from .dotmap import DotMap
from .load_default_cfg import load_default_cfg

__all__ = ["DotMap", "load_default_cfg"]
