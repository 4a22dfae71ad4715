"""Trajectory post-processing -- mirrors toss/trajectory/trajectory_tools.py."""
from typing import Union

import numpy as np


def get_trajectory_adaptive_step(list_of_ode_objects: list) -> Union[np.ndarray, np.ndarray]:
    """Concatenated knots of all segments: states (6, sum n_k), timesteps (sum n_k)  (:6-25)."""
    states, timesteps = None, None
    for ode_object in list_of_ode_objects:
        y = np.transpose(ode_object.y)
        states = y if states is None else np.hstack((states, y))
        timesteps = ode_object.t if timesteps is None else np.hstack((timesteps, ode_object.t))
    return states, timesteps


def get_trajectory_fixed_step(args, list_of_ode_objects: list) -> Union[np.ndarray, np.ndarray]:
    """positions (3,N), velocities (3,N), timesteps (N) at the fixed measurement period (:28-76).
    The sample -> segment bookkeeping is the reference's; the dense output of every segment is
    evaluated on the GPU (TrajectorySegment._OdeSystem__sol -> toss_dense_eval)."""
    timesteps = np.arange(args.problem.start_time, args.problem.final_time, args.problem.measurement_period)
    positions = np.empty((3, len(timesteps)), dtype=np.float64)
    velocities = np.empty((3, len(timesteps)), dtype=np.float64)
    start_idx = 0
    for ode_object in list_of_ode_objects:
        ode_end_time = ode_object.t[-1]
        end_time_idx = (np.abs(timesteps - ode_end_time)).argmin()
        if ode_end_time < timesteps[end_time_idx]:
            end_time_idx -= 1
        if end_time_idx >= start_idx:
            dense = np.transpose(ode_object._OdeSystem__sol(timesteps[start_idx:end_time_idx + 1]))
            positions[:, start_idx:end_time_idx + 1] = dense[0:3, :]
            velocities[:, start_idx:end_time_idx + 1] = dense[3:6, :]
        start_idx = end_time_idx + 1
        if len(timesteps) - 1 < start_idx:
            break
    return positions, velocities, timesteps
