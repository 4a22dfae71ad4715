"""Scheduling never changes a result bit: warps per trajectory (1 / 2 / 4), the longest-first work queue built from
the mascon-proxy run, and the population a trajectory is evaluated in are pure scheduling -- knots, counters,
collision flags and fitness must be identical for every setting and equal to the oracle's."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_parity import FIXED_POS, both_params, ctx_for, random_population

pytestmark = pytest.mark.gpu


def _run(c, p, xs, n_man, **env):
    from toss_b200._native import COUNTERS_DTYPE
    old = {k: os.environ.get(k) for k in env}
    os.environ.update({k: str(v) for k, v in env.items()})
    try:
        res = c.population_fitness(p, xs, FIXED_POS, n_man)
        h = c.propagate(p, xs, FIXED_POS, n_man).to_host()
    finally:
        for k, v in old.items():
            if v is None:
                os.environ.pop(k, None)
            else:
                os.environ[k] = v
    return (res["fitness"].cpu().numpy(), res["collision"].cpu().numpy(), res["n_visited"].cpu().numpy(),
            res["counters"].cpu().numpy().view(COUNTERS_DTYPE).copy(), h["knots_t"], h["knots_y"], h["knots_f"])


@pytest.mark.parametrize("mesh_name,pop,ft", [("llp", 96, 172800.0), ("lp", 80, 86400.0), ("full", 12, 21600.0)])
def test_team_size_and_queue_order_do_not_change_a_bit(mesh_name, pop, ft):
    c, m = ctx_for(mesh_name)
    p, po = both_params(final_time=ft)
    xs = random_population(pop, 2, 700 + pop, ft)
    ref = O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, 2)
    base = None
    for team in (1, 2, 4):
        for lpt in (0, 1):
            got = _run(c, p, xs, 2, TOSS_B200_TEAM=team, TOSS_B200_LPT=lpt)
            assert np.array_equal(got[0], ref["fitness"]), (team, lpt)
            assert np.array_equal(got[1], ref["collision"]) and np.array_equal(got[2], ref["n_visited"])
            for k in ("n_rhs", "n_accepted", "n_rejected", "n_event_points", "status"):
                assert np.array_equal(got[3][k], ref["counters"][k]), (team, lpt, k)
            if base is None:
                base = got
            else:
                # the knots themselves (the workspace beyond a trajectory's last knot is not defined: compare the
                # first n_accepted of each, a lower bound of its knot count)
                for i in range(pop):
                    n = int(got[3]["n_accepted"][i])
                    for a, b in zip(got[4:], base[4:]):
                        assert np.array_equal(a[i, :n], b[i, :n]), (team, lpt, i)


def test_a_trajectory_does_not_depend_on_its_population():
    """fitness(x) == the corresponding entry of batch_fitness, whatever else is in the batch."""
    c, m = ctx_for("lp")
    p, _ = both_params(final_time=43200.0)
    xs = random_population(40, 2, 77, 43200.0)
    whole = c.population_fitness(p, xs, FIXED_POS, 2)["fitness"].cpu().numpy()
    for i in (0, 13, 39):
        alone = c.population_fitness(p, xs[i:i + 1], FIXED_POS, 2)["fitness"].cpu().numpy()
        assert alone[0] == whole[i]
    perm = np.random.default_rng(1).permutation(40)
    assert np.array_equal(c.population_fitness(p, xs[perm], FIXED_POS, 2)["fitness"].cpu().numpy(), whole[perm])
