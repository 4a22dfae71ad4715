"""Trajectory post-processing -- mirrors toss/trajectory/trajectory_tools.py."""
from typing import Union

import numpy as np


def get_trajectory_adaptive_step(list_of_ode_objects: list) -> Union[np.ndarray, np.ndarray]:
    """Concatenated knots of all segments: states (6, sum n_k), timesteps (sum n_k)  (:6-25)."""
    states, timesteps = None, None
    for ode_object in list_of_ode_objects:
        y = np.transpose(ode_object.y)
        states = y if states is None else np.hstack((states, y))
        timesteps = ode_object.t if timesteps is None else np.hstack((timesteps, ode_object.t))
    return states, timesteps


def get_trajectory_fixed_step(args, list_of_ode_objects: list) -> Union[np.ndarray, np.ndarray]:
    """positions (3,N), velocities (3,N), timesteps (N) at the fixed measurement period (:28-76).

    One device call: the knots of all segments go to the GPU as a one-trajectory workspace and `toss_resample`
    (csrc/fitness.cu: resample_kernel) assigns every sample to its segment -- a sample at a manoeuvre time belongs
    to the earlier segment, a segment may serve no sample -- and evaluates the cubic-Hermite
    dense output there.  Samples past the last knot (which the reference leaves as np.empty garbage) come out as
    the last segment's extrapolation."""
    from .._runtime import context_for_args, params_from_args
    segments = []
    for seg in list_of_ode_objects:
        if not hasattr(seg, "f"):
            raise TypeError("get_trajectory_fixed_step needs the segment objects compute_trajectory returns "
                            "(knot slopes .f are part of the dense output)")
        segments.append((seg.t, seg.y, seg.f))
    ctx = next((seg._ctx for seg in list_of_ode_objects if getattr(seg, "_ctx", None) is not None), None)
    if ctx is None:
        ctx = context_for_args(args)
    return ctx.resample_segments(params_from_args(args), segments)
