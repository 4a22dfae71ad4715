from .fitness_function_enums import FitnessFunctions
from .fitness_function_utils import (_compute_squared_distance, cart2sphere, compute_space_coverage,
                                     create_spherical_tensor_grid, estimate_covered_volume, sphere2cart,
                                     update_spherical_tensor_grid)
from .fitness_functions import (close_distance_penalty, covered_space, covered_volume, far_distance_penalty,
                                get_fitness, target_altitude_distance, total_covered_volume)

__all__ = ["FitnessFunctions", "_compute_squared_distance", "cart2sphere", "compute_space_coverage",
           "create_spherical_tensor_grid", "estimate_covered_volume", "sphere2cart", "update_spherical_tensor_grid",
           "close_distance_penalty", "covered_space", "covered_volume", "far_distance_penalty", "get_fitness",
           "target_altitude_distance", "total_covered_volume"]
