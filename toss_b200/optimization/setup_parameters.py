"""setup_parameters -- mirrors toss/optimization/setup_parameters.py:12-69: load the TOML, derive the
volumes / spin state, load the mesh, build the gravity handle and the voxel grid."""
from math import pi

import numpy as np

from ..fitness.fitness_function_utils import create_spherical_tensor_grid
from ..mesh.mesh_utility import create_mesh
from ..trajectory.equations_of_motion import setup_spin_axis
from ..utilities.load_default_cfg import load_default_cfg


class GravityEvaluable:
    """Stand-in for polyhedral_gravity.GravityEvaluable((vertices, faces), density)
    (setup_parameters.py:63).  Holds only host arrays (picklable); evaluation goes to the device context
    of the mesh.  Callable as evaluable(x, parallel) -> (potential, acceleration, tensor) in the
    package's sign convention; the second-derivative tensor is never used by TOSS and is returned as None."""

    def __init__(self, polyhedral_source, density):
        self.vertices = np.asarray(polyhedral_source[0], dtype=np.float64)
        self.faces = np.asarray(polyhedral_source[1])
        self.density = float(density)

    def __call__(self, computation_points, parallel=True):
        from .._runtime import context_for_mesh
        ctx = context_for_mesh(self.vertices, self.faces, self.density)
        pts = np.asarray(computation_points, dtype=np.float64)
        single = pts.ndim == 1
        acc, pot = ctx.gravity(pts.reshape(-1, 3), potential=True)
        acc, pot = acc.cpu().numpy(), pot.cpu().numpy()
        if single:
            return float(pot[0]), acc[0], None
        return [(float(pot[i]), acc[i], None) for i in range(len(pot))]


def setup_parameters(cfg_path: str = None):
    """Returns the nested `args` DotMap with the same sections and derived fields as the reference."""
    args = load_default_cfg(cfg_path)

    args.problem.squared_volume_inner_bounding_sphere = (4 / 3) * pi * (args.problem.radius_inner_bounding_sphere ** 3)
    args.problem.squared_volume_outer_bounding_sphere = (4 / 3) * pi * (args.problem.radius_outer_bounding_sphere ** 3)
    args.problem.total_measurable_volume = (args.problem.squared_volume_outer_bounding_sphere
                                            - args.problem.squared_volume_inner_bounding_sphere)
    args.problem.maximal_measurement_sphere_volume = (4 / 3) * pi * (args.problem.maximal_measurement_sphere_radius ** 3)

    args.body.spin_velocity = (2 * pi) / args.body.spin_period
    if (args.body.spin_axis_x is not None) and (args.body.spin_axis_y is not None) and (args.body.spin_axis_z is not None):
        args.body.spin_axis = np.array([args.body.spin_axis_x, args.body.spin_axis_y, args.body.spin_axis_z])
    else:
        args.body.spin_axis = setup_spin_axis(args)

    (args.mesh.body, args.mesh.vertices, args.mesh.faces,
     args.mesh.largest_body_protuberant) = create_mesh(args.mesh.mesh_path)
    args.mesh.evaluable = GravityEvaluable((args.mesh.vertices, args.mesh.faces), args.body.density)

    args.problem.fixed_velocity = np.array([args.problem.sample_vx, args.problem.sample_vy, args.problem.sample_vz])
    (args.problem.tensor_grid_r, args.problem.tensor_grid_theta, args.problem.tensor_grid_phi,
     args.problem.bool_tensor) = create_spherical_tensor_grid(
        args.problem.measurement_period, args.problem.radius_inner_bounding_sphere,
        args.problem.radius_outer_bounding_sphere, args.problem.max_velocity_scaling_factor,
        args.problem.fixed_velocity)
    return args
