"""The reference's own test files (rasmusmarak/TOSS tests/*.py), restated against the drop-in package: same
calls, same arguments, same tolerances (np.isclose rtol = atol = 1e-5) -- `from toss import ...` becomes
`from toss_b200 import ...`.  Plus the UDP surface the reference never tests.  GPU only."""
import math

import numpy as np
import pytest

from tests.golden import reference_goldens as G

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-5, atol=1e-5)


def _args(**over):
    import toss_b200 as T
    args = T.setup_parameters()
    args.problem.start_time = 0
    args.problem.final_time = 20 * 3600.0
    args.problem.initial_time_step = 600
    args.problem.activate_event = True
    args.problem.number_of_maneuvers = 0
    args.problem.activate_rotation = True
    args.problem.radius_inner_bounding_sphere = 4000
    args.problem.measurement_period = 2500
    args.mesh.mesh_path = "3dmeshes/churyumov-gerasimenko_llp.pk"
    (args.mesh.body, args.mesh.vertices, args.mesh.faces,
     args.mesh.largest_body_protuberant) = T.create_mesh(args.mesh.mesh_path)
    for k, v in over.items():
        setattr(args.problem, k, v)
    return args


def test_integration():
    """tests/integration_test.py:13-46"""
    import toss_b200 as T
    args = _args()
    _, objs, _ = T.compute_trajectory(G.X0, args, T.compute_motion)
    states, _ = T.get_trajectory_adaptive_step(objs)
    assert all(np.isclose(G.FINAL_STATE, states[0:6, -1], **TOL))


def test_fixed_step_trajectory():
    """tests/fixed_step_trajectory_test.py:14-58"""
    import toss_b200 as T
    args = _args()
    _, objs, _ = T.compute_trajectory(G.X0, args, T.compute_motion)
    positions, _, timesteps = T.get_trajectory_fixed_step(args, objs)
    for c in range(3):
        assert all(np.isclose(G.FIXED_STEP_POSITIONS[c, :], positions[c, :], **TOL))
    assert all(np.isclose(G.FIXED_STEP_TIMESTEPS, timesteps, **TOL))
    d = np.diff(timesteps)
    assert all(x == d[0] for x in d)


def test_multiple_impulsive_maneuvers():
    """tests/multiple_impulsive_maneuvers_test.py:13-68"""
    import toss_b200 as T
    args = _args()
    for n_man in range(0, 9):
        chromosome = G.X0
        args.problem.number_of_maneuvers = n_man
        t_m = 0
        for _ in range(1, n_man + 1):
            t_m += 5000
            chromosome = np.concatenate((chromosome, [t_m, 1, 0.1, 0.1, 0.1]), axis=None)
        _, objs, intervals = T.compute_trajectory(chromosome, args, T.compute_motion)
        positions, _ = T.get_trajectory_adaptive_step(objs)
        assert all(np.isclose(G.MANEUVER_FINAL_POSITIONS[:, n_man], positions[0:3, -1], **TOL))
        assert len(objs) == n_man + 1 and len(intervals) == n_man + 2


def test_fitness_functions():
    """tests/fitness_functions_test.py:66-157"""
    import toss_b200 as T
    args = _args(target_squared_altitude=8000 ** 2, penalty_scaling_factor=1, measurement_period=100,
                 radius_outer_bounding_sphere=10000)
    _, objs, _ = T.compute_trajectory(G.X0, args, T.compute_motion)
    positions, _, _ = T.get_trajectory_fixed_step(args, objs)
    assert np.isclose(T.covered_volume(args.problem.maximal_measurement_sphere_volume, positions),
                      G.FIT_COVERED_VOLUME, **TOL)
    assert np.isclose(T.target_altitude_distance(args.problem.target_squared_altitude, positions),
                      G.FIT_TARGET_ALTITUDE, **TOL)
    assert np.isclose(T.total_covered_volume(args.problem.total_measurable_volume, positions),
                      G.FIT_TOTAL_COVERED_VOLUME, **TOL)      # stale total_measurable_volume, as in the reference
    pts = np.arange(args.problem.radius_inner_bounding_sphere - 1000, args.problem.radius_inner_bounding_sphere + 1000, 100)
    pos = np.zeros((3, len(pts)))
    pos[0, :] = pts
    assert np.isclose(T.close_distance_penalty(args.problem.radius_inner_bounding_sphere, pos,
                                               args.problem.penalty_scaling_factor), G.FIT_CLOSE_PENALTY, **TOL)
    pts = np.arange(args.problem.radius_outer_bounding_sphere - 1000, args.problem.radius_outer_bounding_sphere + 1000, 100)
    pos[0, :] = pts
    assert np.isclose(T.far_distance_penalty(args.problem.radius_outer_bounding_sphere, pos,
                                             args.problem.penalty_scaling_factor), G.FIT_FAR_PENALTY, **TOL)


def test_covered_space_half_hemisphere():
    """tests/covered_space_test.py:118-146"""
    import toss_b200 as T
    args = T.setup_parameters()
    args.problem.number_of_spacecrafts = 1
    pr = args.problem
    idx = np.arange(int((pr.tensor_grid_phi.shape[0] - 1) / 2), pr.tensor_grid_phi.shape[0], 1)
    pr.bool_tensor[:, :, idx] = True
    for k, gold in ((9, G.COVERAGE_NO_GAIN), (6, G.COVERAGE_WITH_GAIN)):
        position = np.asarray(T.sphere2cart(pr.tensor_grid_r[0], pr.tensor_grid_theta[0], pr.tensor_grid_phi[k]))
        fit = T.compute_space_coverage(pr.number_of_spacecrafts, args.body.spin_axis, args.body.spin_velocity,
                                       position, np.array([0, 0, 0]), [0], pr.radius_inner_bounding_sphere,
                                       pr.radius_outer_bounding_sphere, pr.tensor_grid_r, pr.tensor_grid_theta,
                                       pr.tensor_grid_phi, pr.bool_tensor)
        assert np.isclose(fit, gold, **TOL)
        assert abs(fit - gold) < 1e-14


def test_covered_space_perfect_ratio_and_random_sample():
    """tests/covered_space_test.py:4-115 (sample counts reduced: the reference loops 1e7 rotations in Python)"""
    import toss_b200 as T
    args = T.setup_parameters()
    n = 200000
    rng = np.random.default_rng(0)
    positions = 10 * rng.random((3, n)) + 2
    cov = T.compute_space_coverage(1, args.body.spin_axis, args.body.spin_velocity, positions, None,
                                   np.arange(0, n + 1, 1), 4, 10, args.problem.tensor_grid_r,
                                   args.problem.tensor_grid_theta, args.problem.tensor_grid_phi, args.problem.bool_tensor)
    assert cov >= 0
    r, th, ph, bt = T.create_spherical_tensor_grid(1, 2, 8, 1, args.problem.fixed_velocity)
    for g in (r, th, ph):
        g[0] += 1e-4
        g[-1] -= 1e-4
    R, TH, PH = np.meshgrid(r, th, ph, indexing="ij")
    x, y, z = T.sphere2cart(R.ravel(), TH.ravel(), PH.ravel())
    pos = np.vstack((x, y, z))
    ts = np.arange(pos.shape[1], dtype=np.float64)
    from toss_b200._runtime import bare_context
    rot = bare_context().rotate_points(args.body.spin_axis, args.body.spin_velocity, -ts, pos.T.copy()).cpu().numpy().T
    cov = T.compute_space_coverage(1, args.body.spin_axis, args.body.spin_velocity, np.ascontiguousarray(rot),
                                   args.problem.fixed_velocity, ts, 2, 8, r, th, ph, bt)
    assert cov >= 1


def test_rotation_of_point():
    """tests/point_rotation_test.py:15-76"""
    import toss_b200 as T
    from toss_b200.utilities.dotmap import DotMap
    args = DotMap(body=DotMap(declination=64, right_ascension=69, spin_period=12.06 * 3600, _dynamic=False),
                  _dynamic=False)
    args.body.spin_velocity = (2 * math.pi) / args.body.spin_period
    args.body.spin_axis = T.setup_spin_axis(args)
    rng = np.random.default_rng(2)
    for _ in range(9):
        x = rng.integers(0, 20000, 3)
        t = rng.random()
        got = T.rotate_point(t, x, args.body.spin_axis, args.body.spin_velocity)
        axis = np.asarray(args.body.spin_axis)
        axis = axis / math.sqrt(np.dot(axis, axis))
        a = math.cos((-(args.body.spin_velocity * t)) / 2.0)
        b, c, d = -axis * math.sin((-(args.body.spin_velocity * t)) / 2.0)
        aa, bb, cc, dd = a * a, b * b, c * c, d * d
        bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
        rot = np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                        [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                        [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])
        assert all(np.isclose(np.dot(rot, x), got, **TOL))


# ------------------------------------------------------------------- the surface the reference never tests
def _udp(n_man=2, final_time=86400.0):
    import toss_b200 as T
    args = T.setup_parameters()
    args.problem.number_of_maneuvers = n_man
    args.problem.final_time = final_time
    fixed = np.array([args.problem.initial_x, args.problem.initial_y, args.problem.initial_z])
    lb, ub = T.setup_initial_state_domain(fixed, args.problem.start_time, args.problem.final_time, n_man,
                                          args.problem.number_of_spacecrafts, args.chromosome)
    return T.udp_initial_condition(args, fixed, lb, ub), np.asarray(lb, dtype=float), np.asarray(ub, dtype=float)


def test_udp_fitness_equals_batch_fitness_and_the_api_chain():
    """udp.fitness(x) (udp_initial_condition.py:106-137) == row of batch_fitness == the explicit chain
    compute_trajectory -> get_trajectory_fixed_step -> get_fitness(enum 9), bit for bit."""
    import toss_b200 as T
    udp, lb, ub = _udp()
    rng = np.random.default_rng(11)
    xs = lb + (ub - lb) * rng.random((12, len(lb)))
    batch = udp.batch_fitness(xs.ravel())
    assert batch.shape == (12,)
    for i in (0, 5, 11):
        assert udp.fitness(xs[i]) == [batch[i]]
        info = np.hstack((udp.initial_condition, xs[i]))
        coll, objs, _ = T.compute_trajectory(info, udp.args, T.compute_motion)
        if coll:
            assert batch[i] == 1e30
            continue
        pos, vel, ts = T.get_trajectory_fixed_step(udp.args, objs)
        f = T.get_fitness(T.FitnessFunctions.CoveredSpaceCloseDistancePenaltyFarDistancePenalty, udp.args, pos, vel, ts)
        assert f == batch[i]


def test_udp_collision_is_a_value_not_an_error():
    """a chromosome aimed straight at the body: fitness 1e30 (udp_initial_condition.py:126-128)"""
    udp, lb, ub = _udp(n_man=0, final_time=40000.0)
    r = udp.initial_condition
    x = np.array([1.0, *(-r / np.linalg.norm(r))])      # 1 m/s toward the centre: hits the surface after ~1 h
    assert udp.fitness(x) == [1e30]
    assert udp.last_batch["collision"][0] == 1 and udp.last_batch["counters"]["n_event_points"][0] > 0


def test_bool_tensor_feedback_between_spacecraft():
    """toss.py:136-140: the champion's voxels are OR-ed into bool_tensor and change the next spacecraft's fitness."""
    import toss_b200 as T
    udp, lb, ub = _udp()
    rng = np.random.default_rng(3)
    xs = lb + (ub - lb) * rng.random((8, len(lb)))
    f0 = udp.batch_fitness(xs.ravel())
    best = int(np.argmin(f0))
    a = udp.args
    coll, objs, _ = T.compute_trajectory(np.hstack((udp.initial_condition, xs[best])), a, T.compute_motion)
    pos, vel, ts = T.get_trajectory_fixed_step(a, objs)
    a.problem.bool_tensor = T.update_spherical_tensor_grid(
        a.problem.number_of_spacecrafts, a.body.spin_axis, a.body.spin_velocity, pos, vel, ts,
        a.problem.radius_inner_bounding_sphere, a.problem.radius_outer_bounding_sphere, a.problem.tensor_grid_r,
        a.problem.tensor_grid_theta, a.problem.tensor_grid_phi, a.problem.bool_tensor)
    assert a.problem.bool_tensor.any()
    f1 = udp.batch_fitness(xs.ravel())
    assert np.all(f1 <= f0 + 1e-15) and f1[best] == f0[best]      # previous visits only ever add coverage


def test_is_outside_and_evaluable_api():
    import toss_b200 as T
    args = T.setup_parameters()
    pts = np.array([[0.0, 0.0, 0.0], [0.0, 0.0, 9000.0], [100.0, 200.0, -300.0]])
    out = T.is_outside(pts, args.mesh.vertices, args.mesh.faces)
    assert out.tolist() == [False, True, False]
    pot, acc, tensor = args.mesh.evaluable(np.array([0.0, 0.0, 9000.0]), False)
    assert pot > 0 and acc[2] > 0 and tensor is None          # package convention: +grad V, TOSS negates it
    rhs = T.compute_motion(0.0, np.array([0.0, 0.0, 9000.0, 1.0, 2.0, 3.0]), args)
    assert np.array_equal(rhs[:3], [1.0, 2.0, 3.0]) and np.allclose(rhs[3:], -acc, rtol=1e-14)
    with pytest.raises(NotImplementedError):
        T.compute_trajectory(G.X0, _args(), lambda t, y, args: y)


def test_use_case_driver_sequential_spacecraft(tmp_path):
    """toss.py:71-191 at BASELINE configs[0] size (lllp mesh, 1 manoeuvre, pop 16, 2 generations), 2 spacecraft:
    one batch_fitness call per generation, champions improve monotonically, the first champion's voxels reach the
    second spacecraft through bool_tensor, four CSV files are written."""
    import toss_b200 as T
    from toss_b200 import driver
    args = T.setup_parameters()
    args.mesh.mesh_path = "3dmeshes/churyumov-gerasimenko_lllp.pk"
    (args.mesh.body, args.mesh.vertices, args.mesh.faces,
     args.mesh.largest_body_protuberant) = T.create_mesh(args.mesh.mesh_path)
    args.problem.number_of_maneuvers = 1
    args.problem.number_of_spacecrafts = 2
    args.problem.final_time = 86400
    args.optimization.population_size = 16
    args.optimization.number_of_generations = 2
    pr = args.problem
    init = np.array_split([pr.initial_x, pr.initial_y, pr.initial_z] * 2, 2)
    lb, ub = T.setup_initial_state_domain(init[0], pr.start_time, pr.final_time, 1, 2, args.chromosome)
    rt, cf, cx, fl = driver.run_optimization(args, init, lb, ub, seed=1)
    assert fl.shape == (2, 2) and np.all(np.diff(fl, axis=1) <= 0)
    assert len(cx) == 2 and len(cx[0]) == len(lb) and cf[0][0] == fl[0, -1]
    assert args.problem.bool_tensor.any()
    # the champion of spacecraft 0, re-evaluated against the UPDATED tensor, gains nothing new
    udp = T.udp_initial_condition(args, init[0], lb, ub)
    assert udp.fitness(cx[1])[0] == cf[1][0]
