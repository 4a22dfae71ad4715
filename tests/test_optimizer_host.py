"""CPU tests of the host-side generational optimiser (toss_b200/optimization/gaco_lite.py) on an analytic problem."""
import numpy as np

from toss_b200.optimization.gaco_lite import GacoLite


class Sphere:
    calls = 0

    def get_bounds(self):
        return ([-5.0] * 4, [5.0] * 4)

    def batch_fitness(self, dvs):
        Sphere.calls += 1
        x = np.asarray(dvs).reshape(-1, 4)
        return ((x - 1.5) ** 2).sum(axis=1)


def test_gaco_lite_converges_with_one_batch_call_per_generation():
    u = Sphere()
    rng = np.random.default_rng(0)
    x = rng.uniform(-5, 5, (32, 4))
    f = u.batch_fitness(x.ravel())
    Sphere.calls = 0
    g = GacoLite(40, seed=3)
    px, pf, cx, cf = g.evolve(u, x, f)
    log = g.get_log()
    assert Sphere.calls == 40 == len(log)
    best = [row[2] for row in log]
    assert all(b1 <= b0 for b0, b1 in zip(best, best[1:])) and cf == best[-1] < 1e-2
    assert np.all(px >= -5) and np.all(px <= 5) and np.allclose(cx, 1.5, atol=0.1)
    g2 = GacoLite(40, seed=3)
    assert g2.evolve(u, x, f)[3] == cf          # deterministic for a given seed
