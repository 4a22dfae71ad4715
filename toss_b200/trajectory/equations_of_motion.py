"""Equations of motion -- mirrors toss/trajectory/equations_of_motion.py.

compute_motion :62-95, compute_acceleration :41-59 and rotate_point :98-117 evaluate on the GPU
(toss_compute_motion / toss_gravity_eval / toss_rotate_points).  They exist for API parity (tests,
notebooks, single evaluations); the population path never calls back into Python per RK stage -- the
stepper kernel evaluates the same device functions in place.
"""
from math import radians

import numpy as np

from .._runtime import bare_context, context_for_args, params_from_args


def setup_spin_axis(args):
    """Unit spin axis from declination / right ascension in degrees (equations_of_motion.py:12-38)."""
    dec = radians(args.body.declination)
    ra = radians(args.body.right_ascension)
    return np.array([np.cos(dec) * np.cos(ra), np.cos(dec) * np.sin(ra), np.sin(dec)])


def compute_acceleration(x: np.ndarray, args) -> np.ndarray:
    """-grad V of the polyhedral model at x (attraction), equations_of_motion.py:41-59."""
    ctx = context_for_args(args)
    acc = ctx.gravity(np.asarray(x, dtype=np.float64).reshape(1, 3)).cpu().numpy()[0]
    return -acc


def compute_motion(t: float, x: np.ndarray, args) -> np.ndarray:
    """RHS [v, a]: the position is rotated by -w t about the spin axis when args.problem.activate_rotation,
    gravity is evaluated there and -- as in the reference -- NOT rotated back (:84-95)."""
    ctx = context_for_args(args)
    p = params_from_args(args)
    out = ctx.compute_motion(p, np.array([float(t)]), np.asarray(x, dtype=np.float64).reshape(1, 6))
    return out.cpu().numpy()[0]


def rotate_point(t: float, x: np.ndarray, spin_axis: np.ndarray, spin_velocity: float) -> np.ndarray:
    """Quaternion(axis=spin_axis, angle=-(spin_velocity*t)).rotate(x) (:98-117)."""
    ctx = bare_context()
    out = ctx.rotate_points(np.asarray(spin_axis, dtype=np.float64), float(spin_velocity), np.array([float(t)]),
                            np.asarray(x, dtype=np.float64).reshape(1, 3))
    return out.cpu().numpy()[0]
