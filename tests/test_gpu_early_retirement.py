"""Collision verdict computed inside the stepper (a ray-cast sweep per event point) and early retirement of
collided trajectories (toss_params.stop_on_collision): same fitness vector as the reference behaviour, bit-identical
to the oracle in either mode, and consistent with the stand-alone K5 pass over the stored knots."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_parity import FIXED_POS, both_params, ctx_for, random_population

pytestmark = pytest.mark.gpu


def _cnt(res):
    from toss_b200._native import COUNTERS_DTYPE
    return res["counters"].cpu().numpy().view(COUNTERS_DTYPE)


@pytest.mark.parametrize("mesh_name,pop,ft", [("llp", 256, 604800.0), ("lp", 48, 172800.0)])
def test_early_retirement_changes_work_not_results(mesh_name, pop, ft):
    c, m = ctx_for(mesh_name)
    xs = random_population(pop, 2, 31, ft)
    k = pop // 4                                               # a quarter is aimed at the body, manoeuvres late
    xs[:k, 0] = 1.0
    xs[:k, 1:4] = -FIXED_POS / np.linalg.norm(FIXED_POS) + 0.05 * np.random.default_rng(1).normal(size=(k, 3))
    xs[:k, 4::5] = ft - 10.0 - np.arange(2)[None, :]
    out = {}
    for stop in (0, 1):
        p, po = both_params(final_time=ft, stop_on_collision=stop)
        res = c.population_fitness(p, xs, FIXED_POS, 2)
        ref = O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, 2)
        cnt = _cnt(res)
        for k in ("n_rhs", "n_accepted", "n_rejected", "n_event_points", "status"):
            assert np.array_equal(cnt[k], ref["counters"][k]), (stop, k)
        assert np.array_equal(res["collision"].cpu().numpy(), ref["collision"])
        assert np.array_equal(res["fitness"].cpu().numpy(), ref["fitness"])
        assert np.array_equal(res["n_visited"].cpu().numpy(), ref["n_visited"])
        out[stop] = (res["fitness"].cpu().numpy(), res["collision"].cpu().numpy(), cnt["n_rhs"].copy())
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1])   # same answers
    coll = out[0][1] == 1
    assert coll.sum() >= pop // 8
    assert np.all(out[1][2][coll] < out[0][2][coll]) and np.array_equal(out[1][2][~coll], out[0][2][~coll])


def test_inline_verdict_equals_the_pass_over_stored_knots():
    c, m = ctx_for("llp")
    p, _ = both_params(final_time=200000.0)
    xs = random_population(128, 1, 5, 200000.0)
    xs[:40, 0] = 1.0
    xs[:40, 1:4] = -FIXED_POS / np.linalg.norm(FIXED_POS) + 0.05 * np.random.default_rng(2).normal(size=(40, 3))
    xs[:40, 4] = 199990.0
    prop = c.propagate(p, xs, FIXED_POS, 1)
    a = prop.to_host()
    assert a["collision"].sum() > 5
    b = c.collision_verdict(p, prop).to_host()
    assert np.array_equal(a["collision"], b["collision"])
    assert np.array_equal(a["counters"]["n_event_points"], b["counters"]["n_event_points"])


def test_udp_batch_uses_early_retirement_and_matches_fitness():
    """udp.batch_fitness (stop_on_collision = 1) == the explicit reference-order chain (integrate to the end)."""
    import toss_b200 as T
    args = T.setup_parameters()
    args.problem.number_of_maneuvers = 1
    args.problem.final_time = 100000
    fixed = np.array([args.problem.initial_x, args.problem.initial_y, args.problem.initial_z])
    lb, ub = T.setup_initial_state_domain(fixed, 0, 100000, 1, 4, args.chromosome)
    udp = T.udp_initial_condition(args, fixed, lb, ub)
    xs = random_population(24, 1, 77, 100000.0)
    xs[:8, 0] = 1.0
    xs[:8, 1:4] = -fixed / np.linalg.norm(fixed) + 0.05 * np.random.default_rng(3).normal(size=(8, 3))
    xs[:8, 4] = 99990.0
    f = udp.batch_fitness(xs.ravel())
    assert (f == 1e30).sum() >= 4
    for i in (0, 3, 12, 20):
        coll, objs, _ = T.compute_trajectory(np.hstack((fixed, xs[i])), args, T.compute_motion)
        if coll:
            assert f[i] == 1e30
        else:
            pos, vel, ts = T.get_trajectory_fixed_step(args, objs)
            assert f[i] == T.get_fitness(T.FitnessFunctions.CoveredSpaceCloseDistancePenaltyFarDistancePenalty, args, pos, vel, ts)
