"""toss_b200 -- B200-native (sm_100a) population-fitness path of TOSS behind the reference's Python API.

Flat re-export of the same public names as the reference's toss/__init__.py:1-49 (plotting excluded:
out of scope).  The arithmetic lives in hand-written CUDA kernels (toss_b200/csrc) reached through the C
ABI of include/toss_b200.h; `toss_b200._native` is the ctypes binding.  There is no CPU fallback.
"""
from .fitness.fitness_function_enums import FitnessFunctions
from .fitness.fitness_function_utils import (estimate_covered_volume, _compute_squared_distance,
                                             compute_space_coverage, update_spherical_tensor_grid,
                                             create_spherical_tensor_grid, cart2sphere, sphere2cart)
from .fitness.fitness_functions import (get_fitness, target_altitude_distance, close_distance_penalty,
                                        far_distance_penalty, covered_volume, total_covered_volume, covered_space)
from .trajectory.compute_trajectory import compute_trajectory
from .trajectory.equations_of_motion import compute_motion, setup_spin_axis, rotate_point
from .trajectory.trajectory_tools import get_trajectory_fixed_step, get_trajectory_adaptive_step
from .mesh.mesh_utility import create_mesh, is_outside
from .trajectory.Integrator import IntegrationScheme
from .utilities.load_default_cfg import load_default_cfg
from .optimization.udp_initial_condition import udp_initial_condition
from .optimization.setup_parameters import setup_parameters
from .optimization.setup_state import setup_initial_state_domain

__version__ = "0.1"

__all__ = [
    "FitnessFunctions", "estimate_covered_volume", "_compute_squared_distance", "compute_space_coverage",
    "update_spherical_tensor_grid", "create_spherical_tensor_grid", "cart2sphere", "sphere2cart", "get_fitness",
    "compute_trajectory", "compute_motion", "setup_spin_axis", "rotate_point", "get_trajectory_fixed_step",
    "get_trajectory_adaptive_step", "target_altitude_distance", "close_distance_penalty", "far_distance_penalty",
    "covered_volume", "total_covered_volume", "covered_space", "create_mesh", "is_outside", "IntegrationScheme",
    "load_default_cfg", "udp_initial_condition", "setup_parameters", "setup_initial_state_domain",
]
