"""Fitness functions -- mirrors toss/fitness/fitness_functions.py (same names and argument order);
every value is computed by the K4 kernel through toss_fitness_eval."""
import numpy as np

from .._runtime import bare_context, default_params, params_from_args, set_grid
from .fitness_function_enums import FitnessFunctions
from .fitness_function_utils import _positions_3xN, compute_space_coverage


def _eval(fn: int, p, positions, timesteps=None) -> float:
    pos = _positions_3xN(positions)
    if fn in (4, 5, 6, 7) and pos.shape[1] < 2:
        # estimate_covered_volume needs two samples: the reference fails on `sphere_radii[0] = sphere_radii[1]`
        # (fitness_function_utils.py:30-31) -- same exception type here
        raise IndexError("index 1 is out of bounds for axis 0 with size 1")
    ctx = bare_context()
    if ctx.grid_shape is None:   # ids 1..7 never touch the grid, but the kernel signature carries one
        set_grid(ctx, [1.0], [0.0], [0.0], np.zeros((1, 1, 1), dtype=np.uint8))
    ts = np.zeros(pos.shape[1]) if timesteps is None else np.ascontiguousarray(timesteps, dtype=np.float64)
    return float(ctx.fitness(fn, p, pos, ts).cpu()[0])


def get_fitness(chosen_fitness_function: FitnessFunctions, args, positions: np.ndarray, velocities: np.ndarray,
                timesteps: np.ndarray):
    """Dispatcher (:7-45).  Note the reference passes args.problem.number_of_spacecrafts (plural, from the
    cfg) to the coverage functions -- reproduced by params_from_args."""
    fn = FitnessFunctions(chosen_fitness_function).value
    p = params_from_args(args)
    ctx = bare_context()
    if fn in (8, 9):
        set_grid(ctx, args.problem.tensor_grid_r, args.problem.tensor_grid_theta, args.problem.tensor_grid_phi,
                 args.problem.bool_tensor)
    return _eval(fn, p, positions, timesteps)


def target_altitude_distance(target_squared_altitude: float, positions: np.ndarray) -> float:
    return _eval(1, default_params(target_squared_altitude=float(target_squared_altitude)), positions)


def close_distance_penalty(radius_inner_bounding_sphere: float, positions: np.ndarray,
                           penalty_scaling_factor: float) -> float:
    return _eval(2, default_params(radius_inner=float(radius_inner_bounding_sphere),
                                   penalty_scaling_factor=float(penalty_scaling_factor)), positions)


def far_distance_penalty(radius_outer_bounding_sphere: float, positions: np.ndarray,
                         penalty_scaling_factor: float) -> float:
    return _eval(3, default_params(radius_outer=float(radius_outer_bounding_sphere),
                                   penalty_scaling_factor=float(penalty_scaling_factor)), positions)


def covered_volume(maximal_measurement_sphere_volume: float, positions: np.ndarray) -> float:
    return _eval(4, default_params(maximal_measurement_sphere_volume=float(maximal_measurement_sphere_volume)),
                 positions)


def total_covered_volume(total_measurable_volume: float, positions: np.ndarray) -> float:
    return _eval(5, default_params(total_measurable_volume=float(total_measurable_volume)), positions)


def covered_volume_far_distance_penalty(maximal_measurement_sphere_volume: float, radius_outer_bounding_sphere: float,
                                        positions: np.ndarray, penalty_scaling_factor: float) -> float:
    return _eval(6, default_params(maximal_measurement_sphere_volume=float(maximal_measurement_sphere_volume),
                                   radius_outer=float(radius_outer_bounding_sphere),
                                   penalty_scaling_factor=float(penalty_scaling_factor)), positions)


def covered_volume_close_distance_penalty_far_distance_penalty(maximal_measurement_sphere_volume: float,
                                                               radius_inner_bounding_sphere: float,
                                                               radius_outer_bounding_sphere: float,
                                                               positions: np.ndarray, penalty_scaling_factor):
    return _eval(7, default_params(maximal_measurement_sphere_volume=float(maximal_measurement_sphere_volume),
                                   radius_inner=float(radius_inner_bounding_sphere),
                                   radius_outer=float(radius_outer_bounding_sphere),
                                   penalty_scaling_factor=float(penalty_scaling_factor)), positions)


def covered_space(number_of_spacecrafts: int, spin_axis: np.ndarray, spin_velocity: float,
                  radius_inner_bounding_sphere: float, radius_outer_bounding_sphere: float, positions: np.ndarray,
                  velocities: np.ndarray, timesteps: np.ndarray, r: np.ndarray, theta: np.ndarray, phi: np.ndarray,
                  bool_tensor: np.ndarray) -> float:
    return compute_space_coverage(number_of_spacecrafts, spin_axis, spin_velocity, positions, velocities, timesteps,
                                  radius_inner_bounding_sphere, radius_outer_bounding_sphere, r, theta, phi,
                                  bool_tensor)


def covered_space_close_distance_penalty_far_distance_penalty(number_of_spacecrafts: int, spin_axis: np.ndarray,
                                                              spin_velocity: float,
                                                              radius_inner_bounding_sphere: float,
                                                              radius_outer_bounding_sphere: float,
                                                              positions: np.ndarray, velocities: np.ndarray,
                                                              timesteps: np.ndarray, penalty_scaling_factor: float,
                                                              r: np.ndarray, theta: np.ndarray, phi: np.ndarray,
                                                              bool_tensor: np.ndarray) -> float:
    """close + far - coverage (:198-227)"""
    ctx = bare_context()
    set_grid(ctx, r, theta, phi, bool_tensor)
    p = default_params(number_of_spacecrafts=int(number_of_spacecrafts),
                       spin_axis=np.asarray(spin_axis, dtype=np.float64), spin_velocity=float(spin_velocity),
                       radius_inner=float(radius_inner_bounding_sphere),
                       radius_outer=float(radius_outer_bounding_sphere),
                       penalty_scaling_factor=float(penalty_scaling_factor))
    return _eval(9, p, positions, timesteps)
