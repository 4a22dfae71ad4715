"""compute_trajectory -- mirrors toss/trajectory/compute_trajectory.py:13-162 on the GPU stepper.

The reference returns desolver OdeSystem objects; callers read `.t`, `.y`, `.events[i].y` and the
name-mangled dense-output callable `._OdeSystem__sol(t)` (trajectory_tools.py:19-20,55,67;
compute_trajectory.py:138,141-143).  `TrajectorySegment` carries exactly those.
"""
from typing import Callable, Union

import numpy as np

from .._runtime import context_for_args, params_from_args
from .Integrator import IntegrationScheme
from .equations_of_motion import compute_motion


class _Event:
    __slots__ = ("t", "y")

    def __init__(self, t, y):
        self.t, self.y = t, y


class TrajectorySegment:
    """One integration interval: knots of the adaptive stepper + cubic-Hermite dense output."""

    def __init__(self, ctx, t, y, f, risk_radius, activate_event):
        self._ctx = ctx
        self.t = t                  # (n,)
        self.y = y                  # (n, 6)
        self.f = f                  # (n, 6)  RHS at the knots
        r2 = np.einsum("ij,ij->i", y[1:, :3], y[1:, :3])
        idx = np.nonzero(r2 <= risk_radius ** 2)[0] + 1 if activate_event else np.zeros(0, dtype=int)
        self.events = [_Event(t[i], y[i]) for i in idx]   # accepted states inside the risk sphere
        self.success = True

    def _OdeSystem__sol(self, times):
        """dense output at `times` -> (len(times), 6)  (desolver's interpolant, evaluated on the GPU)"""
        times = np.atleast_1d(np.asarray(times, dtype=np.float64))
        if times.size == 0:
            return np.zeros((0, 6))
        return self._ctx.dense_eval(self.t, self.y, self.f, times).cpu().numpy()

    sol = _OdeSystem__sol

    def __getstate__(self):
        d = dict(self.__dict__)
        d["_ctx"] = None
        return d


def compute_trajectory(x: np.ndarray, args, func: Callable) -> Union[bool, list, np.ndarray]:
    """x = [r(3), v_mag, v_dir(3), (t_m, dv_mag, dv_dir(3)) x number_of_maneuvers]
    -> (collision_detected, list_of_ode_objects, integration_intervals)."""
    if func is not compute_motion and getattr(func, "__name__", "") != "compute_motion":
        raise NotImplementedError("the B200 path integrates the built-in compute_motion only (no Python RHS callbacks)")
    if not IntegrationScheme(args.integrator.algorithm).supported:
        raise NotImplementedError("only IntegrationScheme.RK8713MSolver (algorithm = 3) is implemented on the B200 path")
    x = np.asarray(x, dtype=np.float64)
    n_man = int(args.problem.number_of_maneuvers)
    if len(x) != 7 + 5 * n_man:
        raise ValueError(f"state vector has {len(x)} entries, expected 7 + 5*number_of_maneuvers = {7 + 5 * n_man}")
    ctx = context_for_args(args)
    p = params_from_args(args)
    knot_cap = int(args.integrator.get("knot_capacity", 8192)) if hasattr(args.integrator, "get") else 8192
    h = ctx.propagate(p, x[None, :], np.zeros(0), n_man, knot_cap).to_host()
    status = int(h["counters"]["status"][0])
    if status == 3:
        raise RuntimeError("trajectory needs more knots than args.integrator.knot_capacity")
    if status != 0:
        raise FloatingPointError(f"integration failed (status {status}: 1 = step cap, 2 = non-finite state)")
    n_seg = int(h["n_seg"][0])
    seg = h["seg_start"][0]
    segments = []
    for s in range(n_seg):
        a, b = seg[s], seg[s + 1]
        segments.append(TrajectorySegment(ctx, h["knots_t"][0, a:b].copy(), h["knots_y"][0, a:b].copy(),
                                          h["knots_f"][0, a:b].copy(), p.radius_inner, bool(p.activate_event)))
    intervals = h["intervals"][0, :n_seg + 1].copy()
    # the reference mutates the shared args on every segment (compute_trajectory.py:131-132)
    args.integrator.t0 = intervals[-2]
    args.integrator.tf = intervals[-1]
    return bool(h["collision"][0]), segments, intervals
