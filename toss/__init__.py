"""`toss` -- the reference's import name (toss/__init__.py:1-49), served by the B200 path.

`from toss import compute_trajectory, setup_parameters, ...` resolves to toss_b200's implementations, so code and
tests written against the reference package run unmodified on the sm_100a kernels (tests/test_gpu_reference_tests_unmodified.py
executes the reference's own six test files this way).  The sub-packages `toss.fitness`, `toss.trajectory`,
`toss.mesh`, `toss.optimization`, `toss.utilities` alias the toss_b200 ones.  The five plotting helpers of
toss/visualization/plotting_utility.py are out of scope: the names exist and raise.
"""
import importlib as _importlib
import sys as _sys

import toss_b200 as _impl
from toss_b200 import *            # noqa: F401,F403  (the names of toss/__init__.py:16-49)
from toss_b200 import _compute_squared_distance   # noqa: F401  (underscore name: not covered by *)

for _sub in ("fitness", "trajectory", "mesh", "optimization", "utilities"):
    globals()[_sub] = _importlib.import_module(f"toss_b200.{_sub}")
# every module of the implementation answers to its reference name too (toss.trajectory.compute_trajectory, ...):
# ONE module object under two names, so there is one device-context cache whichever name the caller used
for _name, _m in list(_sys.modules.items()):
    if _name.startswith("toss_b200.") and _m is not None:
        _sys.modules.setdefault(__name__ + _name[len("toss_b200"):], _m)


def _no_plotting(name):
    def _stub(*args, **kwargs):
        raise NotImplementedError(f"toss.{name}: plotting (toss/visualization/plotting_utility.py) is outside the "
                                  "B200 population-fitness path")
    _stub.__name__ = name
    return _stub


_PLOTTING = ("plot_UDP_3D", "plot_UDP_2D", "fitness_over_generations", "fitness_over_time",
             "distance_deviation_over_time")
for _n in _PLOTTING:
    globals()[_n] = _no_plotting(_n)

__all__ = list(_impl.__all__) + list(_PLOTTING)
__version__ = _impl.__version__
