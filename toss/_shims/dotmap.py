"""`dotmap` stand-in, put on sys.path ONLY where the real package is absent (this image has no dotmap):
the reference and its tests do `from dotmap import DotMap`."""
from toss_b200.utilities.dotmap import DotMap   # noqa: F401
