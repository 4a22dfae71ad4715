"""Extended Ant Colony Optimisation (GACO) with the reference's eleven parameters -- SURVEY.md 8 f2.

The reference evolves every spacecraft's population with pygmo's `gaco` (toss.py:39-68: gen, ker, q, oracle_parameter, acc,
threshold, n_gen_mark, impstop, evalstop, focus, memory; values in toss/resources/default_cfg.toml:72-82).  pygmo /
pagmo is an un-vendored, un-pinned dependency (environment.yml) that is absent from this image, so this module restates
the algorithm as pagmo 2 documents it (Schlueter, Egea, Banga: "Extended ant colony optimization for non-convex mixed
integer nonlinear programming", 2009; pagmo `gaco` class documentation).  It is not pagmo's code and does not share its
random-number stream: parity here means the same parameters with the same meaning, not the same samples.

What every parameter does here:

* `gen`        generations per `evolve` call; every generation is ONE batch evaluation of `pop_size` candidates.
* `ker`        size of the solution archive (the kernels of the Gaussian mixture); pop_size >= ker is required.
* `q`          convergence speed: rank weights w_k ~ exp(-k^2 / (2 q^2 ker^2)) / (q ker sqrt(2 pi)), k = 0 .. ker-1.
* `oracle_parameter`  pagmo's penalty reference value Omega: solutions are ranked by p(f) = alpha |f - Omega|
               (f > Omega), -|f - Omega| otherwise,
               alpha = (6 sqrt 3 - 2) / (6 sqrt 3) for a problem without constraints (residual 0) -- the reference's UDP has none.
* `acc`        a candidate only enters the archive when its penalty differs by at least `acc` from every archived one.
* `threshold`  from this generation count on q is set to 0.01 (the search collapses onto the best kernels).
* `n_gen_mark` the standard deviations are (max - min of the archive per gene) divided by a counter that runs
               1 .. n_gen_mark and then starts again.
* `impstop`    stop after this many consecutive generations without an improvement of the best penalty (0: never).
* `evalstop`   stop after this many fitness evaluations without an improvement (0: never).
* `focus`      if non-zero, sigma_h is capped at (ub_h - lb_h) / focus.
* `memory`     keep the archive and the counters between `evolve` calls.

Where the candidates are made.  With a UDP that offers `batch_fitness_device` (toss_b200's does) the ants of a
generation are drawn by `toss_gaco_sample` (csrc/gaco.cu) straight into the device matrix the stepper reads; per
generation only the archive tables go up (ker x dim doubles) and the fitness vector plus the rows that entered the
archive come down.  Any other UDP gets `sample_host`, the numpy mirror of the same counter-based generator (same
integers, same formula; the device's log / cos differ from numpy's in the last bits, tests compare at 1e-12).
"""
from __future__ import annotations

import math

import numpy as np

_M64 = (1 << 64) - 1
_ALPHA_UNCONSTRAINED = (6.0 * math.sqrt(3.0) - 2.0) / (6.0 * math.sqrt(3.0))


def _mix64(z):
    """splitmix64 finaliser on uint64 arrays (csrc/gaco.cu: gaco_mix64)."""
    with np.errstate(over="ignore"):
        z = (z + np.uint64(0x9E3779B97F4A7C15)).astype(np.uint64)
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        return z ^ (z >> np.uint64(31))


def _u01(h):
    return ((h >> np.uint64(11)).astype(np.float64) + 0.5) * (1.0 / 9007199254740992.0)


def sample_host(archive_x, cum_weights, sigma, lb, ub, seed: int, generation: int, first_ant: int, n: int):
    """numpy mirror of gaco_sample_kernel: ants first_ant .. first_ant+n-1 of a generation, (n, dim)."""
    archive_x = np.asarray(archive_x, dtype=np.float64)
    ker, dim = archive_x.shape
    cum, sigma, lb, ub = (np.asarray(a, dtype=np.float64) for a in (cum_weights, sigma, lb, ub))
    key = _mix64(np.uint64(seed & _M64) ^ _mix64(np.uint64(generation)))
    with np.errstate(over="ignore"):
        ka = _mix64(key + np.arange(first_ant, first_ant + n, dtype=np.uint64))
        u = _u01(_mix64(ka))
        k = np.minimum((u[:, None] > cum[None, :]).sum(axis=1), ker - 1)     # first i with u <= cum[i]
        out = np.empty((n, dim))
        for h in range(dim):
            centre = archive_x[k, h]
            g = np.empty(n)
            todo = np.arange(n)
            for t in range(64):
                c = np.uint64((h * 64 + t) * 2)
                u1 = _u01(_mix64(ka[todo] + np.uint64(1) + c))
                u2 = _u01(_mix64(ka[todo] + np.uint64(2) + c))
                z = np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)
                g[todo] = centre[todo] + sigma[h] * z
                todo = todo[(g[todo] < lb[h]) | (g[todo] > ub[h])]
                if todo.size == 0:
                    break
            g[todo] = np.clip(g[todo], lb[h], ub[h])
            out[:, h] = g
    return out


class Gaco:
    """pygmo-shaped GACO: `Gaco(*the eleven cfg values, seed=...)`, `.evolve(udp, pop_x, pop_f)`, `.get_log()`."""

    def __init__(self, gen: int = 1, ker: int = 63, q: float = 1.0, oracle_parameter: float = 0.0, acc: float = 0.01,
                 threshold: int = 1, n_gen_mark: int = 7, impstop: int = 100000, evalstop: int = 100000,
                 focus: float = 0.0, memory: bool = False, seed: int = 0):
        if ker < 2:
            raise ValueError("gaco: the kernel size must be at least 2")
        if q < 0 or acc < 0 or focus < 0:
            raise ValueError("gaco: q, acc and focus must not be negative")
        if not (1 <= threshold <= max(gen, 1)) and gen != 0 and not memory:
            raise ValueError("gaco: the threshold parameter must be in [1, gen]")
        self.gen, self.ker, self.q0, self.oracle_parameter, self.acc = int(gen), int(ker), float(q), float(oracle_parameter), float(acc)
        self.threshold, self.n_gen_mark = int(threshold), max(int(n_gen_mark), 1)
        self.impstop, self.evalstop, self.focus, self.memory = int(impstop), int(evalstop), float(focus), bool(memory)
        self.seed = int(seed)
        self.log = []                      # (gen, fevals, best fitness, ker, oracle_parameter, dx, dp) per generation
        self._reset()

    def _reset(self):
        self.arch_x = self.arch_f = self.arch_p = None
        self.counter = 0                   # generations seen (across calls when memory is on)
        self.gen_mark = 1
        self.q = self.q0
        self.n_impstop = self.n_evalstop = 0
        self.fevals = 0

    # ---- the pieces (each is tested on its own) ---------------------------------------------------------------
    def penalty(self, f):
        """pagmo's penalty function for an unconstrained problem (residual 0): monotone in f, zero at f = oracle_parameter"""
        f = np.asarray(f, dtype=np.float64)
        d = np.abs(f - self.oracle_parameter)
        return np.where(f > self.oracle_parameter, _ALPHA_UNCONSTRAINED * d, -d)

    def weights(self):
        k = np.arange(self.ker, dtype=np.float64)
        s = self.q * self.ker
        w = np.exp(-(k * k) / (2.0 * s * s)) / (s * math.sqrt(2.0 * math.pi))
        tot = w.sum()
        if not (tot > 0.0):                # q so small that only the best kernel survives in double precision
            w = np.zeros(self.ker)
            w[0] = tot = 1.0
        return w / tot

    def sigmas(self, lb, ub):
        spread = (self.arch_x.max(axis=0) - self.arch_x.min(axis=0)) / self.gen_mark
        if self.focus != 0.0:
            spread = np.minimum(spread, (ub - lb) / self.focus)
        return spread

    def _update_archive(self, x_of, f, p):
        """merge a generation (fitness f, penalties p, rows fetched through x_of(indices)) into the sorted archive"""
        order = np.argsort(p, kind="stable")
        entering = []
        arch_p = list(self.arch_p) if self.arch_p is not None else []
        worst = arch_p[-1] if len(arch_p) == self.ker else math.inf
        for i in order:
            if len(arch_p) == self.ker and not (p[i] < worst):
                break
            if self.acc != 0.0 and arch_p and np.min(np.abs(np.asarray(arch_p) - p[i])) < self.acc:
                continue
            entering.append(int(i))
            arch_p.append(float(p[i]))
            arch_p.sort()
            del arch_p[self.ker:]
            worst = arch_p[-1] if len(arch_p) == self.ker else math.inf
        if not entering:
            return False
        rows = x_of(np.asarray(entering, dtype=np.int64))
        if self.arch_x is None:
            ax, af, ap = rows, f[entering], p[entering]
        else:
            ax = np.vstack((self.arch_x, rows))
            af = np.concatenate((self.arch_f, f[entering]))
            ap = np.concatenate((self.arch_p, p[entering]))
        keep = np.argsort(ap, kind="stable")[:self.ker]
        best_before = self.arch_p[0] if self.arch_p is not None else math.inf
        self.arch_x, self.arch_f, self.arch_p = ax[keep].copy(), af[keep].copy(), ap[keep].copy()
        return self.arch_p[0] < best_before

    # ---- the loop --------------------------------------------------------------------------------------------
    def evolve(self, udp, pop_x, pop_f):
        """pop_x (n, dim) with its fitness pop_f (n,).  Returns (pop_x, pop_f, champion_x, champion_f) after `gen`
        generations: the population is the last generation's ants (like pagmo's, which overwrites every individual),
        the champion the best solution seen."""
        lb, ub = (np.asarray(b, dtype=np.float64) for b in udp.get_bounds())
        pop_x = np.ascontiguousarray(pop_x, dtype=np.float64)
        pop_f = np.asarray(pop_f, dtype=np.float64).reshape(-1)
        n, dim = pop_x.shape
        if n < self.ker:
            raise ValueError(f"gaco: the population ({n}) must be at least as large as the kernel ({self.ker})")
        if not self.memory:
            self._reset()
            self.log = []
        on_device = hasattr(udp, "batch_fitness_device")
        ctx = udp.device_context() if on_device else None
        self._update_archive(lambda idx: pop_x[idx], pop_f, self.penalty(pop_f))
        if self.arch_x is None or len(self.arch_x) < 2:
            raise ValueError("gaco: fewer than two distinct solutions to build the archive from (accuracy too coarse?)")
        champion_x, champion_f = self.arch_x[0].copy(), float(self.arch_f[0])
        d_x = None
        for g in range(1, self.gen + 1):
            self.counter += 1
            if self.counter >= self.threshold:
                self.q = 0.01
            cum = np.cumsum(self.weights())
            cum[-1] = 1.0
            sigma = self.sigmas(lb, ub)
            kerx = np.zeros((self.ker, dim))
            kerx[:len(self.arch_x)] = self.arch_x
            kerx[len(self.arch_x):] = self.arch_x[-1]           # an archive still filling up repeats its last member
            if on_device:
                d_x = ctx.gaco_sample(kerx, cum, sigma, lb, ub, self.seed, self.counter, 0, n, out=d_x)
                f = np.asarray(udp.batch_fitness_device(d_x), dtype=np.float64)
                fetch = lambda idx, d=d_x: d[ctx.torch.from_numpy(idx).to(d.device)].cpu().numpy()
            else:
                x = sample_host(kerx, cum, sigma, lb, ub, self.seed, self.counter, 0, n)
                f = np.asarray(udp.batch_fitness(x.ravel()), dtype=np.float64).reshape(-1)
                fetch = lambda idx, x=x: x[idx]
            self.fevals += n
            best_p_before = float(self.arch_p[0])
            best_x_before = self.arch_x[0].copy()
            improved = self._update_archive(fetch, f, self.penalty(f))
            self.gen_mark = self.gen_mark + 1 if self.gen_mark < self.n_gen_mark else 1
            if improved:
                self.n_impstop = self.n_evalstop = 0
            else:
                self.n_impstop += 1
                self.n_evalstop += n
            if self.arch_f[0] < champion_f or self.penalty(self.arch_f[0]) < self.penalty(champion_f):
                champion_x, champion_f = self.arch_x[0].copy(), float(self.arch_f[0])
            self.log.append((self.counter, self.fevals, float(self.arch_f[0]), self.ker, self.oracle_parameter,
                             float(np.abs(self.arch_x[0] - best_x_before).sum()),
                             float(abs(self.arch_p[0] - best_p_before))))
            pop_f = f
            last = (d_x, None) if on_device else (None, x)
            if (self.impstop and self.n_impstop >= self.impstop) or (self.evalstop and self.n_evalstop >= self.evalstop):
                break
        if self.gen >= 1:
            pop_x = last[0].cpu().numpy() if on_device else last[1]
        return pop_x, pop_f, champion_x, champion_f

    def get_log(self):
        return list(self.log)
