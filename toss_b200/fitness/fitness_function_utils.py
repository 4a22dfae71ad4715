"""Fitness helpers -- mirrors toss/fitness/fitness_function_utils.py.

compute_space_coverage :54-162 and update_spherical_tensor_grid :165-228 run on the GPU (K4: voxel
bitmap in shared memory, exact integer counts).  create_spherical_tensor_grid :231-279 and the
coordinate transforms are set-up / analysis helpers on the host, as in the reference.
"""
from typing import Union

import numpy as np

from .._runtime import bare_context, default_params, set_grid


def _positions_3xN(positions):
    pos = np.asarray(positions, dtype=np.float64)
    if pos.ndim == 1:
        pos = pos.reshape(3, 1)
    return np.ascontiguousarray(pos)


def _coverage_params(number_of_spacecrafts, spin_axis, spin_velocity, radius_min, radius_max, single):
    return default_params(number_of_spacecrafts=1 if single else int(number_of_spacecrafts),
                          spin_axis=np.asarray(spin_axis, dtype=np.float64), spin_velocity=float(spin_velocity),
                          radius_inner=float(radius_min), radius_outer=float(radius_max))


def compute_space_coverage(number_of_spacecrafts: int, spin_axis: np.ndarray, spin_velocity: float,
                           positions: np.ndarray, velocities: np.ndarray, timesteps: np.ndarray, radius_min: float,
                           radius_max: float, r: np.ndarray, theta: np.ndarray, phi: np.ndarray,
                           bool_tensor: np.ndarray) -> float:
    """ratio of visited grid points + sum of 1/r over them (:54-162).  `positions` is (3,N), or (3,) for a
    single position rotated with timesteps[0] (:84-93).  `velocities` is unused, as in the reference."""
    ctx = bare_context()
    set_grid(ctx, r, theta, phi, bool_tensor)
    single = np.asarray(positions).ndim == 1
    pos = _positions_3xN(positions)
    ts = np.ascontiguousarray(np.atleast_1d(np.asarray(timesteps, dtype=np.float64)))
    p = _coverage_params(number_of_spacecrafts, spin_axis, spin_velocity, radius_min, radius_max, single)
    return float(ctx.fitness(8, p, pos, ts).cpu()[0])


def update_spherical_tensor_grid(number_of_spacecrafts: int, spin_axis: np.ndarray, spin_velocity: float,
                                 positions: np.ndarray, velocities: np.ndarray, timesteps: np.ndarray,
                                 radius_min: float, radius_max: float, r: np.ndarray, theta: np.ndarray,
                                 phi: np.ndarray, bool_tensor: np.ndarray) -> np.ndarray:
    """bool_tensor OR the voxels visited by `positions` (:165-228); returns the new tensor."""
    ctx = bare_context()
    set_grid(ctx, r, theta, phi, bool_tensor)
    pos = _positions_3xN(positions)
    ts = np.ascontiguousarray(np.atleast_1d(np.asarray(timesteps, dtype=np.float64)))
    p = _coverage_params(number_of_spacecrafts, spin_axis, spin_velocity, radius_min, radius_max, False)
    out = ctx.update_tensor(p, pos, ts)
    ctx._grid_key = None   # the context now holds the updated tensor
    return out


def create_spherical_tensor_grid(time_step: int, radius_min: float, radius_max: float,
                                 max_velocity_scaling_factor: float,
                                 fixed_velocity: np.ndarray) -> Union[np.ndarray, np.ndarray, np.ndarray, np.ndarray]:
    """Grid spacing from the distance travelled in one time step at the scaled sample velocity (:231-279); the
    reference's own input assertions (:251-254)."""
    assert (time_step > 0)
    assert (radius_min > 0)
    assert (radius_max > radius_min)
    assert (max_velocity_scaling_factor > 0)
    max_velocity = np.max(np.linalg.norm(fixed_velocity)) * max_velocity_scaling_factor
    max_distance_traveled = max_velocity * time_step
    r_steps = np.floor((radius_max - radius_min) / max_distance_traveled)
    theta_steps = np.floor(np.pi * radius_min / max_distance_traveled)
    phi_steps = np.floor(2 * np.pi * radius_min / max_distance_traveled)
    # (a spacing larger than the shell gives an empty axis, as in the reference; uploading such a grid is refused)
    r = np.linspace(radius_min, radius_max, int(r_steps))
    theta = np.linspace(-np.pi / 2, np.pi / 2, int(theta_steps))
    phi = np.linspace(-np.pi, np.pi, int(phi_steps))
    bool_tensor = np.full((len(r), len(theta), len(phi)), False)
    return r, theta, phi, bool_tensor


def cart2sphere(x, y, z) -> tuple:
    """(r, theta in [-pi/2, pi/2], phi in [-pi, pi])  (:282-314)"""
    r = np.sqrt(x ** 2 + y ** 2 + z ** 2)
    theta = np.arctan2(z, np.sqrt(x ** 2 + y ** 2))
    phi = np.arctan2(y, x)
    return r, theta, phi


def sphere2cart(r, theta, phi) -> tuple:
    """inverse of cart2sphere (:316-342)"""
    x = r * np.cos(theta) * np.cos(phi)
    y = r * np.cos(theta) * np.sin(phi)
    z = r * np.sin(theta)
    return x, y, z


def estimate_covered_volume(positions: np.ndarray) -> float:
    """Sum of the measurement-sphere volumes along the trajectory (:8-38); the GPU evaluates the same sum
    inside fitness ids 4-7 -- this helper returns it alone."""
    pos = _positions_3xN(positions)
    if pos.shape[1] < 2:      # the reference fails on `sphere_radii[0] = sphere_radii[1]` (:30-31)
        raise IndexError("index 1 is out of bounds for axis 0 with size 1")
    ctx = bare_context()
    p = default_params(total_measurable_volume=1.0)
    if ctx.grid_shape is None:
        ctx.upload_grid([1.0], [0.0], [0.0], np.zeros((1, 1, 1), dtype=np.uint8))
    return -float(ctx.fitness(5, p, pos, np.zeros(pos.shape[1])).cpu()[0])


def _compute_squared_distance(positions: np.ndarray, constant: float) -> np.ndarray:
    """|p_n|^2 - constant^2 per sample (:41-51)"""
    return positions[0, :] ** 2 + positions[1, :] ** 2 + positions[2, :] ** 2 - constant ** 2
