"""A small generational optimiser standing in for pygmo's GACO when pygmo is not installed.

pygmo's `gaco` (Extended Ant Colony Optimisation, Schlueter et al. 2009) keeps an archive of the `ker` best
solutions and samples every new generation from a Gaussian kernel mixture centred on the archive members:
rank weights w_k ~ N(k; 1, q * ker), per-dimension standard deviations from the mean distance between
archive members divided by a convergence-speed factor that grows with the generation count.  This is a
restatement of that published scheme with the parameters of default_cfg.toml [algorithm]; it is NOT pygmo's
code and makes no claim of sample-for-sample equivalence.  It lives on the host: every generation is ONE
`udp.batch_fitness` call, which is the whole point of the B200 path.
"""
from __future__ import annotations

import numpy as np


class GacoLite:
    def __init__(self, gen: int, ker: int = 13, q: float = 1.0, oracle_parameter: float = 1e9, acc: float = 0.0,
                 threshold: int = 1, n_gen_mark: int = 7, impstop: int = 100000, evalstop: int = 100000,
                 focus: float = 0.0, memory: bool = False, seed: int = 0):
        self.gen, self.ker, self.q, self.n_gen_mark = int(gen), int(ker), float(q), int(n_gen_mark)
        self.impstop, self.evalstop = int(impstop), int(evalstop)
        self.rng = np.random.default_rng(seed)
        self.log = []          # (generation, evaluations, best fitness) per generation, like gaco.get_log()

    def evolve(self, udp, pop_x: np.ndarray, pop_f: np.ndarray):
        """pop_x (n, dim), pop_f (n,) evaluated already.  Returns the final (pop_x, pop_f, champion_x, champion_f)."""
        lb, ub = (np.asarray(b, dtype=np.float64) for b in udp.get_bounds())
        n, dim = pop_x.shape
        ker = min(self.ker, n)
        order = np.argsort(pop_f, kind="stable")
        arch_x, arch_f = pop_x[order[:ker]].copy(), pop_f[order[:ker]].copy()
        ranks = np.arange(1, ker + 1)
        w = np.exp(-((ranks - 1) ** 2) / (2.0 * (self.q * ker) ** 2))
        w /= w.sum()
        evals, stall = n, 0
        for g in range(1, self.gen + 1):
            # std per dimension: mean distance to the other archive members, shrinking with the generation count
            speed = max(1.0, g / max(self.n_gen_mark, 1))
            dist = np.abs(arch_x[:, None, :] - arch_x[None, :, :]).sum(axis=1) / max(ker - 1, 1)      # (ker, dim)
            sigma = np.maximum(dist / speed, 1e-12 * (ub - lb))
            pick = self.rng.choice(ker, size=n, p=w)
            cand = arch_x[pick] + self.rng.normal(size=(n, dim)) * sigma[pick]
            cand = np.clip(cand, lb, ub)
            f = np.asarray(udp.batch_fitness(cand.ravel()), dtype=np.float64)
            evals += n
            allx, allf = np.vstack((arch_x, cand)), np.concatenate((arch_f, f))
            order = np.argsort(allf, kind="stable")[:ker]
            improved = allf[order[0]] < arch_f[0]
            arch_x, arch_f = allx[order].copy(), allf[order].copy()
            pop_x, pop_f = cand, f
            self.log.append((g, evals, float(arch_f[0])))
            stall = 0 if improved else stall + 1
            if stall >= self.impstop or evals >= self.evalstop * max(n, 1):
                break
        return pop_x, pop_f, arch_x[0].copy(), float(arch_f[0])

    def get_log(self):
        return list(self.log)
