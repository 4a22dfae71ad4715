"""Edge cases of the path on the GPU, each checked against the oracle (bit-identical) or against the
documented error behaviour of the C ABI: degenerate populations, simultaneous / unsorted manoeuvres, failure
statuses (step cap, non-finite chromosome, knot capacity), events off, rotation off, very fine voxel grids,
sample/manoeuvre coincidences, and BASELINE's other configs at test size."""
import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_parity import FIXED_POS, both_params, bound_population, ctx_for, random_population

pytestmark = pytest.mark.gpu


def _counters(res):
    from toss_b200._native import COUNTERS_DTYPE
    return res["counters"].cpu().numpy().view(COUNTERS_DTYPE)


def _same(res, ref):
    cnt = _counters(res)
    for k in ("n_rhs", "n_accepted", "n_rejected", "n_event_points", "status"):
        assert np.array_equal(cnt[k], ref["counters"][k]), k
    assert np.array_equal(res["collision"].cpu().numpy(), ref["collision"])
    assert np.array_equal(res["n_visited"].cpu().numpy(), ref["n_visited"])
    assert np.array_equal(res["fitness"].cpu().numpy(), ref["fitness"])


def test_population_of_one_and_odd_sizes():
    c, m = ctx_for("llp")
    p, po = both_params(final_time=86400.0)
    for pop in (1, 7, 9, 33):          # fewer trajectories than warps in a CTA, not a multiple of anything
        xs = random_population(pop, 2, 60 + pop, 86400.0)
        _same(c.population_fitness(p, xs, FIXED_POS, 2), O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, 2))


def test_empty_population_is_refused_loudly():
    from toss_b200._native import TossError
    c, m = ctx_for("llp")
    p, _ = both_params()
    with pytest.raises(TossError):
        c.population_fitness_host(p, np.zeros((0, 14)), FIXED_POS, 2)
    with pytest.raises(TossError):                      # ragged: dim does not match 7 + 5 M
        c.population_fitness_host(p, np.zeros((4, 13)), FIXED_POS, 2)
    with pytest.raises(TossError):
        c.population_fitness_host(p, np.zeros((4, 14)), FIXED_POS, 2, knot_cap=2)


def test_all_three_chromosome_layouts():
    """setup_state.py:35-47: nothing fixed (dim 7+5M), position fixed (4+5M), state fixed (5M)."""
    c, m = ctx_for("llp")
    p, po = both_params(final_time=50000.0)
    rng = np.random.default_rng(8)
    full = np.hstack([np.tile(FIXED_POS, (10, 1)) + rng.normal(size=(10, 3)) * 300, random_population(10, 1, 9, 50000.0)])
    ref = O.population_fitness(m, po, O.default_grid(), full, np.zeros(0), 1)
    _same(c.population_fitness(p, full, np.zeros(0), 1), ref)
    # same trajectories with the leading 3 / 7 values moved into the fixed part
    one = full[3]
    a = c.population_fitness_host(p, one[None, 3:], one[:3], 1)
    b = c.population_fitness_host(p, one[None, 7:], one[:7], 1)
    assert a["fitness"][0] == ref["fitness"][3] == b["fitness"][0]


def test_simultaneous_unsorted_and_boundary_maneuvers():
    """compute_trajectory.py:87-121: int32 truncation, argsort, equal times merged with the dv SUMMED; samples
    that coincide with a manoeuvre time belong to the earlier segment (trajectory_tools.py:55-58)."""
    c, m = ctx_for("llp")
    p, po = both_params(final_time=40000.0, measurement_period=100.0)
    base = [0.3, 0.5, -0.4, 0.2]
    rows = [
        base + [20000.9, 0.4, 1, 0, 0] + [20000.2, 0.3, 0, 1, 0] + [5000.0, 0.2, 0, 0, 1],       # two truncate to 20000, unsorted
        base + [30000.0, 0.1, 1, 1, 0] + [10000.0, 0.2, -1, 0, 0] + [20000.0, 0.05, 0, 0, 1],    # all on sample times
        base + [1.0, 0.5, 0.3, 0.3, 0.3] + [39999.0, 0.5, 0, -1, 0] + [1.9, 0.1, 1, 0, 0],        # at the bounds, 1 and 1.9 merge
        base + [12345.6, 0.0, 0, 0, 0] + [12345.1, 0.0, 0, 0, 0] + [12345.9, 0.3, 0.1, 0.1, 0.1],  # triple merge
    ]
    xs = np.array(rows, dtype=float)
    res = c.population_fitness(p, xs, FIXED_POS, 3)
    ref = O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, 3)
    _same(res, ref)
    h = c.propagate(p, xs, FIXED_POS, 3).to_host()
    assert h["n_seg"].tolist() == [3, 4, 3, 2]
    assert h["intervals"][0, :4].tolist() == [0.0, 5000.0, 20000.0, 40000.0]
    assert h["intervals"][3, :3].tolist() == [0.0, 12345.0, 40000.0]


def test_failure_statuses_score_1e30():
    """status 1 (step cap), 2 (non-finite state), 3 (knot capacity): fitness 1e30, n_visited -1, same as the oracle."""
    c, m = ctx_for("llp")
    xs = random_population(6, 1, 3, 86400.0)
    p, po = both_params(final_time=86400.0, max_steps=20)
    res, ref = c.population_fitness(p, xs, FIXED_POS, 1), O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, 1)
    assert (ref["counters"]["status"] == 1).all()
    _same(res, ref)
    assert (res["fitness"].cpu().numpy() == 1e30).all()
    p, po = both_params(final_time=86400.0)
    bad = xs.copy()
    bad[2, 0] = np.nan
    bad[4, 1] = np.inf
    res, ref = c.population_fitness(p, bad, FIXED_POS, 1), O.population_fitness(m, po, O.default_grid(), bad, FIXED_POS, 1)
    assert ref["counters"]["status"].tolist() == [0, 0, 2, 0, 2, 0]
    _same(res, ref)
    res = c.population_fitness(p, xs, FIXED_POS, 1, knot_cap=16)     # too few knots for a day
    cnt = _counters(res)
    assert (cnt["status"] == 3).all() and (res["fitness"].cpu().numpy() == 1e30).all()


def test_events_off_rotation_off_and_other_spin_axis():
    c, m = ctx_for("llp")
    xs = bound_population(12, 1, 21, 86400.0)
    xs[:, 0] *= 0.1                                   # slow: falls toward the body
    for kw in (dict(activate_event=0), dict(activate_rotation=0),
               dict(spin_axis=(0.15709817, 0.40925472, 0.89879405), spin_velocity=3e-4)):
        p, po = both_params(final_time=86400.0, **kw)
        res, ref = c.population_fitness(p, xs, FIXED_POS, 1), O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, 1)
        _same(res, ref)
        if "activate_event" in kw:
            assert ref["collision"].sum() == 0 and (ref["counters"]["n_event_points"] == 0).all()


def test_collisions_detected_and_scored():
    """plunging trajectories: event points recorded, ray-cast verdict identical, fitness 1e30"""
    c, m = ctx_for("llp")
    p, po = both_params(final_time=60000.0)
    rng = np.random.default_rng(2)
    d = -FIXED_POS / np.linalg.norm(FIXED_POS)
    xs = np.array([[rng.uniform(0.5, 1.5), *(d + 0.1 * rng.normal(size=3))] for _ in range(32)])
    res, ref = c.population_fitness(p, xs, FIXED_POS, 0), O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, 0)
    assert 4 < ref["collision"].sum() <= 32 and (ref["counters"]["n_event_points"] > 0).all()
    _same(res, ref)


def test_fine_grid_and_single_sample():
    c, _ = ctx_for("llp")
    grid = O.default_grid(scale=4.0)                        # 60 x 90 x 181 ~ 1 M voxels: the largest class that fits shared memory
    c, m = ctx_for("llp", grid)
    p, po = both_params(final_time=43200.0, number_of_spacecrafts=1)
    xs = bound_population(4, 0, 5, 43200.0)
    res = c.population_fitness(p, xs, FIXED_POS, 0)
    ref = O.population_fitness(m, po, grid, xs, FIXED_POS, 0)
    _same(res, ref)
    assert ref["n_visited"].max() > 50
    p1, po1 = both_params(final_time=100.0, measurement_period=100.0)   # exactly one sample
    grid = O.default_grid()
    c, m = ctx_for("llp", grid)
    _same(c.population_fitness(p1, xs, FIXED_POS, 0), O.population_fitness(m, po1, grid, xs, FIXED_POS, 0))


def test_grid_too_fine_is_refused():
    from toss_b200._native import TossError
    c, _ = ctx_for("llp", O.default_grid(scale=1.5))       # ~19 M voxels > shared-memory bitmap
    p, _ = both_params(final_time=1000.0)
    with pytest.raises(TossError, match="voxel grid too fine"):
        c.population_fitness_host(p, bound_population(2, 0, 1, 1000.0), FIXED_POS, 0)
    ctx_for("llp")                                         # restore the default grid for later tests


def test_baseline_config_shapes_at_test_size():
    """configs[0] (1 manoeuvre, lllp, pop 16), configs[2] (llp, 2 manoeuvres, pop 256 trimmed to 64, the
    4-spacecraft chunking quirk on), configs[3] (full mesh, 1 day, pop 8)."""
    for mesh_name, pop, n_man, ft, nsc in (("lllp", 16, 1, 604800.0, 4), ("llp", 64, 2, 604800.0, 4), ("full", 8, 2, 86400.0, 1)):
        c, m = ctx_for(mesh_name)
        p, po = both_params(final_time=ft, number_of_spacecrafts=nsc)
        xs = random_population(pop, n_man, 900 + pop, ft)
        _same(c.population_fitness(p, xs, FIXED_POS, n_man), O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, n_man))


@pytest.mark.parametrize("mesh_name,pop,n_spot", [("lp", 32768, 6), ("full", 4096, 2)])
def test_full_size_populations_through_size_independent_properties(mesh_name, pop, n_spot):
    """BASELINE configs[4] (lp, pop 32768) and configs[3] (full mesh, pop 4096) at FULL size, 7-day window, where the
    oracle cannot follow the whole population: (1) a trajectory does not depend on its population -- a random subset
    evaluated on its own gives the same bits (other team size, other queue order, other CTA neighbours); (2) permuting
    the population permutes the result; (3) no trajectory fails, collided ones score 1e30 and only they; (4) a few
    trajectories against the oracle, bit for bit (fitness, counts, collision flag)."""
    c, m = ctx_for(mesh_name)
    p, po = both_params(stop_on_collision=1)
    xs = random_population(pop, 2, 4242)
    whole = c.population_fitness(p, xs, FIXED_POS, 2)
    fit, col, cnt = whole["fitness"].cpu().numpy(), whole["collision"].cpu().numpy(), _counters(whole)
    assert fit.shape == (pop,) and np.all(cnt["status"] == 0) and np.all(np.isfinite(fit))
    assert np.array_equal(fit == 1e30, col != 0) and 0 < col.sum() < pop
    rng = np.random.default_rng(9)
    sub = np.sort(rng.choice(pop, 96 if mesh_name == "lp" else 24, replace=False))
    alone = c.population_fitness(p, xs[sub], FIXED_POS, 2)
    assert np.array_equal(alone["fitness"].cpu().numpy(), fit[sub])
    assert np.array_equal(_counters(alone)["n_rhs"], cnt["n_rhs"][sub])
    if mesh_name == "lp":
        perm = rng.permutation(pop)
        assert np.array_equal(c.population_fitness(p, xs[perm], FIXED_POS, 2)["fitness"].cpu().numpy(), fit[perm])
    order = np.argsort(cnt["n_rhs"][sub])
    spot = sub[order[len(sub) // 2 - n_spot // 2:len(sub) // 2 - n_spot // 2 + n_spot]]   # median cost: the oracle is a CPU
    ref = O.population_fitness(m, po, O.default_grid(), xs[spot], FIXED_POS, 2)
    assert np.array_equal(ref["fitness"], fit[spot]) and np.array_equal(ref["collision"], col[spot])
    assert np.array_equal(ref["counters"]["n_rhs"], cnt["n_rhs"][spot])
    assert np.array_equal(ref["n_visited"], whole["n_visited"].cpu().numpy()[spot])
