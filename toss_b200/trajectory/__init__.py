from .Integrator import IntegrationScheme
from .compute_trajectory import compute_trajectory
from .equations_of_motion import compute_motion, rotate_point, setup_spin_axis
from .trajectory_tools import get_trajectory_adaptive_step, get_trajectory_fixed_step

__all__ = ["IntegrationScheme", "compute_trajectory", "compute_motion", "rotate_point", "setup_spin_axis",
           "get_trajectory_adaptive_step", "get_trajectory_fixed_step"]
