"""Use-Case-1 driver -- mirrors the reference's toss.py:19-191 on top of the B200 fitness path.

Spacecraft are optimised one after the other; after each one the champion's trajectory is recomputed and its
voxels are OR-ed into args.problem.bool_tensor, which changes the coverage term of the next spacecraft
(toss.py:104-145).  Where the reference spreads `udp.fitness` over a `pg.mp_bfe` process pool, every generation
here is one `udp.batch_fitness` call.  With pygmo installed the optimiser is pygmo's GACO driven through
`pg.bfe(pg.member_bfe())`; without it `toss_b200.optimization.gaco.Gaco`, which honours the same eleven parameters and
draws every generation's candidates on the device (csrc/gaco.cu), is used.
"""
from __future__ import annotations

import time

import numpy as np

from .fitness.fitness_function_utils import update_spherical_tensor_grid
from .optimization.gaco import Gaco
from .optimization.setup_parameters import setup_parameters
from .optimization.setup_state import setup_initial_state_domain
from .optimization.udp_initial_condition import udp_initial_condition
from .trajectory.compute_trajectory import compute_trajectory
from .trajectory.equations_of_motion import compute_motion
from .trajectory.trajectory_tools import get_trajectory_fixed_step

try:                                   # pragma: no cover - pygmo is not in this image
    import pygmo as pg
except ImportError:
    pg = None


def load_udp(args, initial_state, lower_bounds, upper_bounds):
    """toss.py:19-36 -- the UDP (wrapped in pg.problem when pygmo is present)."""
    udp = udp_initial_condition(args, initial_state, lower_bounds, upper_bounds)
    return pg.problem(udp) if pg is not None else udp


def load_uda(args, bfe=None, seed: int = 0):
    """toss.py:39-68 -- GACO with the [algorithm] parameters of the cfg."""
    a, o = args.algorithm, args.optimization
    params = (o.number_of_generations, a.kernel_size, a.convergence_speed_parameter, a.oracle_parameter,
              a.accuracy_parameter, a.threshold_parameter, a.std_convergence_speed_parameter,
              a.improvement_stopping_criterion, a.evaluation_stopping_criterion, a.focus_parameter, a.memory_parameter)
    if pg is not None:                 # pragma: no cover
        uda = pg.gaco(*params)
        uda.set_bfe(bfe)
        return pg.algorithm(uda)
    return Gaco(*params, seed=seed)


def run_optimization(args, initial_state, lower_bounds, upper_bounds, seed: int = 0):
    """toss.py:71-154.  Returns (run_time, champion_f_list, champion_x_list, fitness_array)."""
    n_sc, n_gen, size = args.problem.number_of_spacecrafts, args.optimization.number_of_generations, \
        args.optimization.population_size
    champion_f_list, champion_x_list = [], []
    fitness_array = np.empty((n_sc, n_gen), dtype=np.float64)
    lb, ub = np.asarray(lower_bounds, dtype=np.float64), np.asarray(upper_bounds, dtype=np.float64)
    rng = np.random.default_rng(seed)
    timer_start = time.time()
    for sc in range(n_sc):
        prob = load_udp(args, initial_state[sc], lower_bounds, upper_bounds)
        if pg is not None:             # pragma: no cover
            bfe = pg.bfe(pg.member_bfe())
            pop = pg.population(prob=prob, size=size, b=bfe)
            algo = load_uda(args, bfe)
            pop = algo.evolve(pop)
            log = algo.extract(pg.gaco).get_log()
            champion_f, champion_x = pop.champion_f, pop.champion_x
            fitness_list = [info[2] for info in log]
        else:
            udp = prob
            x0 = lb + (ub - lb) * rng.random((size, len(lb)))
            f0 = np.asarray(udp.batch_fitness(x0.ravel()))
            algo = load_uda(args, seed=seed + sc)
            _, _, champion_x, cf = algo.evolve(udp, x0, f0)
            champion_f = np.array([cf])
            fitness_list = [info[2] for info in algo.get_log()]
        fitness_list = list(fitness_list) + [fitness_list[-1]] * (n_gen - len(fitness_list))
        # champion trajectory -> voxel feedback for the next spacecraft (toss.py:126-140)
        state_vector = np.hstack((initial_state[sc], champion_x)) if len(initial_state[sc]) > 0 else champion_x
        _, objs, _ = compute_trajectory(state_vector, args, compute_motion)
        positions, velocities, timesteps = get_trajectory_fixed_step(args, objs)
        pr = args.problem
        pr.bool_tensor = update_spherical_tensor_grid(pr.number_of_spacecrafts, args.body.spin_axis,
                                                      args.body.spin_velocity, positions, velocities, timesteps,
                                                      pr.radius_inner_bounding_sphere, pr.radius_outer_bounding_sphere,
                                                      pr.tensor_grid_r, pr.tensor_grid_theta, pr.tensor_grid_phi,
                                                      pr.bool_tensor)
        champion_f_list.append(champion_f)
        champion_x_list.append(champion_x)
        fitness_array[sc, :] = fitness_list[:n_gen]
    return time.time() - timer_start, champion_f_list, champion_x_list, fitness_array


def main(out_prefix: str = "test_", cfg_path: str = None):
    """toss.py:157-191: default cfg, position-fixed chromosomes, four CSV files."""
    args = setup_parameters(cfg_path)
    pr = args.problem
    initial_state = np.array_split([pr.initial_x, pr.initial_y, pr.initial_z] * pr.number_of_spacecrafts,
                                   pr.number_of_spacecrafts)
    lower_bounds, upper_bounds = setup_initial_state_domain(initial_state[0], pr.start_time, pr.final_time,
                                                            pr.number_of_maneuvers, pr.number_of_spacecrafts,
                                                            args.chromosome)
    run_time, champion_f, champion_x, fitness_list = run_optimization(args, initial_state, lower_bounds, upper_bounds)
    np.savetxt(out_prefix + "run_time.csv", np.asarray([float(run_time)]), delimiter=",")
    np.savetxt(out_prefix + "champion_f.csv", np.asarray(champion_f), delimiter=",")
    np.savetxt(out_prefix + "champion_x.csv", np.asarray(champion_x), delimiter=",")
    np.savetxt(out_prefix + "fitness_list.csv", np.asarray(fitness_list), delimiter=",")
    return run_time, champion_f, champion_x, fitness_list


if __name__ == "__main__":
    main()
