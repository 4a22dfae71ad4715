"""CPU tests of the generational optimiser (toss_b200/optimization/gaco.py) on closed-form problems: the eleven
parameters of toss.py:51-66 / default_cfg.toml:72-82 each change what pagmo's documentation says they change."""
import math

import numpy as np
import pytest

from toss_b200.optimization.gaco import Gaco, sample_host


class Sphere:
    def __init__(self, dim=4, centre=1.5):
        self.dim, self.centre, self.calls = dim, centre, 0

    def get_bounds(self):
        return ([-5.0] * self.dim, [5.0] * self.dim)

    def batch_fitness(self, dvs):
        self.calls += 1
        x = np.asarray(dvs).reshape(-1, self.dim)
        return ((x - self.centre) ** 2).sum(axis=1)


def _start(u, n=32, seed=0):
    x = np.random.default_rng(seed).uniform(-5, 5, (n, u.dim))
    f = u.batch_fitness(x.ravel())
    u.calls = 0
    return x, f


def test_converges_with_one_batch_call_per_generation_and_is_deterministic():
    u = Sphere()
    x, f = _start(u)
    g = Gaco(60, 13, 1.0, 1e9, 0.0, 30, 7, 100000, 100000, 0.0, False, seed=3)
    px, pf, cx, cf = g.evolve(u, x, f)
    log = g.get_log()
    assert u.calls == 60 == len(log) and log[-1][1] == 60 * 32
    best = [row[2] for row in log]
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:])) and cf == best[-1] < 1e-3
    assert np.all(px >= -5) and np.all(px <= 5) and np.allclose(cx, 1.5, atol=0.05)
    assert px.shape == x.shape and pf.shape == f.shape
    g2 = Gaco(60, 13, 1.0, 1e9, 0.0, 30, 7, 100000, 100000, 0.0, False, seed=3)
    assert g2.evolve(u, x, f)[3] == cf
    g3 = Gaco(60, 13, 1.0, 1e9, 0.0, 30, 7, 100000, 100000, 0.0, False, seed=4)
    assert g3.evolve(u, x, f)[3] != cf


def test_reference_default_parameters_run():
    """default_cfg.toml:72-82 with its population of 15 (>= kernel 13) and 15 generations (toml :44-45)"""
    u = Sphere(dim=6)
    x, f = _start(u, n=15)
    g = Gaco(15, 13, 1.0, 1e9, 0.0, 1, 7, 100000, 100000, 0.0, False, seed=0)
    _, _, cx, cf = g.evolve(u, x, f)
    assert len(g.get_log()) == 15 and cf <= f.min()
    with pytest.raises(ValueError):
        Gaco(2, 13).evolve(u, x[:10], f[:10])               # population smaller than the kernel


def test_oracle_penalty_formula_and_ranking():
    g = Gaco(1, 13, oracle_parameter=2.0)
    f = np.array([-1.0, 2.0, 2.5, 10.0])
    p = g.penalty(f)
    alpha = (6 * math.sqrt(3) - 2) / (6 * math.sqrt(3))
    assert np.allclose(p, [-3.0, 0.0, alpha * 0.5, alpha * 8.0], rtol=0, atol=1e-15)
    assert np.all(np.diff(p) > 0)                            # monotone: without constraints the ranking is the fitness's
    # the collision constant of the UDP (1e30) and the default oracle (1e9) stay finite and ordered
    g = Gaco(1, 13, oracle_parameter=1e9)
    assert np.all(np.diff(g.penalty(np.array([-0.9, 0.0, 3.0, 1e30]))) > 0)


def test_threshold_collapses_the_weights_and_q_spreads_them():
    g = Gaco(10, 13, q=1.0, threshold=5)
    w = g.weights()
    assert abs(w.sum() - 1) < 1e-15 and np.all(np.diff(w) < 0) and w[0] < 0.2
    g.q = 0.01                                               # what the loop sets from generation `threshold` on
    w = g.weights()
    assert w[0] > 1 - 1e-12
    u = Sphere()
    x, f = _start(u)
    g = Gaco(6, 13, 1.0, 1e9, 0.0, 4, 7, 0, 0, 0.0, False, seed=1)
    g.evolve(u, x, f)
    assert g.q == 0.01 and g.counter == 6


def test_sigma_uses_the_gen_mark_counter_and_focus_caps_it():
    u = Sphere()
    x, f = _start(u)
    g = Gaco(0, 13, n_gen_mark=3, oracle_parameter=1e9, acc=0.0)
    g.evolve(u, x, f)                                        # gen = 0: archive only
    lb, ub = (np.asarray(b) for b in u.get_bounds())
    spread = g.arch_x.max(axis=0) - g.arch_x.min(axis=0)
    for mark in (1, 2, 3):
        g.gen_mark = mark
        assert np.array_equal(g.sigmas(lb, ub), spread / mark)
    g.gen_mark, g.focus = 1, 100.0
    assert np.array_equal(g.sigmas(lb, ub), np.minimum(spread, (ub - lb) / 100.0))
    seen = []
    g = Gaco(7, 13, 1.0, 1e9, 0.0, 7, 3, 0, 0, 0.0, False, seed=1)
    orig = g.sigmas
    g.sigmas = lambda lb, ub: (seen.append(g.gen_mark), orig(lb, ub))[1]
    g.evolve(u, x, f)
    assert seen == [1, 2, 3, 1, 2, 3, 1]


def test_accuracy_keeps_archived_penalties_apart():
    u = Sphere()
    x, f = _start(u, n=64)
    g = Gaco(25, 8, 1.0, 1e9, 0.05, 25, 7, 0, 0, 0.0, False, seed=2)
    g.evolve(u, x, f)
    gaps = np.diff(np.sort(g.arch_p))
    assert len(g.arch_p) == 8 and np.all(gaps >= 0.05 - 1e-15)
    g0 = Gaco(25, 8, 1.0, 1e9, 0.0, 25, 7, 0, 0, 0.0, False, seed=2)
    g0.evolve(u, x, f)
    assert np.min(np.diff(np.sort(g0.arch_p))) < 0.05        # without it the archive crowds around the optimum


def test_memory_keeps_the_archive_between_calls_and_stopping_criteria_stop():
    u = Sphere()
    x, f = _start(u)
    g = Gaco(5, 13, 1.0, 1e9, 0.0, 5, 7, 0, 0, 0.0, True, seed=5)
    px, pf, _, cf1 = g.evolve(u, x, f)
    px, pf, _, cf2 = g.evolve(u, px, pf)
    assert g.counter == 10 and len(g.get_log()) == 10 and cf2 <= cf1
    assert [r[0] for r in g.get_log()] == list(range(1, 11))
    g = Gaco(5, 13, 1.0, 1e9, 0.0, 5, 7, 0, 0, 0.0, False, seed=5)
    g.evolve(u, x, f)
    g.evolve(u, x, f)
    assert g.counter == 5 and len(g.get_log()) == 5          # memory off: every call starts again

    class Flat(Sphere):
        def batch_fitness(self, dvs):
            self.calls += 1
            return np.full(len(dvs) // self.dim, 7.0)

    fl = Flat()
    xf, ff = _start(fl)
    g = Gaco(50, 13, 1.0, 1e9, 0.0, 50, 7, 3, 0, 0.0, False, seed=5)      # impstop = 3 generations
    ff = ff + np.arange(len(ff)) * 1e-3                      # a spread-out start, then nothing ever improves
    g.evolve(fl, xf, ff)
    assert len(g.get_log()) == 3
    g = Gaco(50, 13, 1.0, 1e9, 0.0, 50, 7, 0, 32 * 4, 0.0, False, seed=5)  # evalstop = 4 generations' evaluations
    g.evolve(fl, xf, ff)
    assert len(g.get_log()) == 4


def test_host_sampler_statistics_bounds_and_counter_based_streams():
    arch = np.array([[0.0, 100.0], [1.0, 200.0], [2.0, 300.0]])
    cum = np.array([0.5, 0.8, 1.0])
    sigma, lb, ub = np.array([0.1, 2.0]), np.array([-0.05, 0.0]), np.array([5.0, 1000.0])
    x = sample_host(arch, cum, sigma, lb, ub, seed=7, generation=3, first_ant=0, n=40000)
    assert np.all(x >= lb) and np.all(x <= ub)
    k = np.rint((x[:, 1] / 100.0)).clip(1, 3).astype(int) - 1
    share = np.bincount(k, minlength=3) / len(k)
    assert np.allclose(share, [0.5, 0.3, 0.2], atol=0.01)
    sel = k == 1
    assert abs(x[sel, 0].mean() - 1.0) < 0.005 and abs(x[sel, 0].std() - 0.1) < 0.005
    assert x[k == 0, 0].min() >= -0.05                       # truncated at the bound by re-drawing, not piled up on it
    assert (x[k == 0, 0] == -0.05).sum() == 0
    # an ant's genes depend on (seed, generation, ant) only: any slice of the population gives the same rows
    part = sample_host(arch, cum, sigma, lb, ub, seed=7, generation=3, first_ant=1000, n=50)
    assert np.array_equal(part, x[1000:1050])
    assert not np.array_equal(sample_host(arch, cum, sigma, lb, ub, 7, 4, 0, 50), x[:50])
    assert not np.array_equal(sample_host(arch, cum, sigma, lb, ub, 8, 3, 0, 50), x[:50])
