"""Chromosome bounds -- mirrors toss/optimization/setup_state.py:5-49.

Full state vector: [x, y, z, v_mag, vx, vy, vz, (t_m, dv_mag, dvx, dvy, dvz) x number_of_maneuvers];
`initial_condition` (0, 3 or 7 leading values) is fixed and prepended by the UDP, the rest is optimised.
"""
import numpy as np


def setup_initial_state_domain(initial_condition, start_time, final_time, number_of_maneuvers,
                               number_of_spacecrafts, state):
    tm = [(start_time + 1), (final_time - 1)]
    man_lo = [tm[0], state.dv_min, state.dvx_min, state.dvy_min, state.dvz_min] * number_of_maneuvers
    man_hi = [tm[1], state.dv_max, state.dvx_max, state.dvy_max, state.dvz_max] * number_of_maneuvers
    if len(initial_condition) == 0:
        lower_bounds = np.concatenate(([state.x_min, state.y_min, state.z_min, state.v_min, state.vx_min,
                                        state.vy_min, state.vz_min], man_lo), axis=None)
        upper_bounds = np.concatenate(([state.x_max, state.y_max, state.z_max, state.v_max, state.vx_max,
                                        state.vy_max, state.vz_max], man_hi), axis=None)
    elif len(initial_condition) == 3:
        lower_bounds = np.concatenate(([state.v_min, state.vx_min, state.vy_min, state.vz_min], man_lo), axis=None)
        upper_bounds = np.concatenate(([state.v_max, state.vx_max, state.vy_max, state.vz_max], man_hi), axis=None)
    else:
        lower_bounds = man_lo
        upper_bounds = man_hi
    return lower_bounds, upper_bounds
