"""udp_initial_condition -- the pygmo user-defined problem of toss/optimization/udp_initial_condition.py:13-147,
with the new `batch_fitness`.

`fitness(x)` keeps the reference's contract (one chromosome -> [float]); `batch_fitness(dvs)` follows pygmo's
batch protocol (1-D array of n*dim concatenated decision vectors -> 1-D array of n fitness values), so
`pg.bfe(pg.member_bfe())` / `pg.default_bfe()` route a whole generation to ONE call of the B200 path:
propagate -> collision verdict -> fused resample + fitness, population sharded over ranks when a
torch.distributed process group is active.  The object holds no device handles: pygmo may deep-copy and pickle it.
"""
from typing import Union

import numpy as np

from .. import parallel as P
from .._runtime import context_for_args, params_from_args


def ctx_counters_dtype():
    from .._native import COUNTERS_DTYPE
    return COUNTERS_DTYPE


class udp_initial_condition:

    def __init__(self, args, initial_condition, lower_bounds, upper_bounds):
        """Same assertions as the reference (:80-96), same attributes (:99-103)."""
        self.target_sq_alt = args.problem.target_squared_altitude
        assert self.target_sq_alt > 0
        assert all(np.greater(upper_bounds, lower_bounds))
        assert args.body.mu > 0
        assert args.problem.final_time > args.problem.start_time
        assert args.problem.initial_time_step <= args.problem.final_time - args.problem.start_time
        assert args.problem.radius_inner_bounding_sphere > args.mesh.largest_body_protuberant
        assert args.body.density > 0
        assert args.body.declination >= 0
        assert args.body.right_ascension >= 0
        assert args.body.spin_period >= 0
        assert args.problem.number_of_maneuvers >= 0
        assert isinstance(args.problem.number_of_maneuvers, int)
        assert isinstance(args.problem.activate_event, bool)
        assert args.problem.radius_outer_bounding_sphere > args.problem.radius_inner_bounding_sphere
        assert args.problem.squared_volume_outer_bounding_sphere > args.problem.squared_volume_inner_bounding_sphere
        assert args.problem.total_measurable_volume > 0
        assert args.problem.measurement_period > 0

        self.args = args
        self.initial_condition = initial_condition
        self.lower_bounds = lower_bounds
        self.upper_bounds = upper_bounds
        self.args.problem.number_of_spacecraft = 1   # (sic) the reference sets this unused singular key (:103)
        self.knot_capacity = 2048                    # knots per trajectory in the device workspace (status 3 if exceeded)
        self.last_batch = None                       # counters / collision flags of the latest batch
        # Integrator failures (counters.status 1 = step cap / dt <= 0, 2 = non-finite state) are NOT collisions.  Policy
        # (SURVEY.md 8b "Errors"): the trajectory scores 1e30 so that the optimiser moves on, the batch counts them in
        # last_batch["n_failed"] (summed over ranks), a RuntimeWarning names the count -- and with
        # raise_on_integration_failure = True every rank raises FloatingPointError instead, as compute_trajectory does.
        self.raise_on_integration_failure = False

    # ------------------------------------------------------------------------------------------
    def get_bounds(self) -> Union[np.ndarray, np.ndarray]:
        return (self.lower_bounds, self.upper_bounds)

    def get_nobj(self) -> int:
        return 1

    def fitness(self, chromosome: np.ndarray) -> list:
        """One chromosome -> [fitness] (1e30 when a collision is detected, :126-128): a batch of one."""
        x = np.asarray(chromosome, dtype=np.float64).reshape(1, -1)
        return [float(self._evaluate(x)[0])]

    def batch_fitness(self, dvs: np.ndarray) -> np.ndarray:
        """pygmo batch protocol: dvs holds n decision vectors back to back; returns n fitness values."""
        dvs = np.asarray(dvs, dtype=np.float64)
        dim = len(self.lower_bounds)
        if dvs.ndim == 1:
            if dvs.size % dim != 0:
                raise ValueError(f"batch_fitness: {dvs.size} values is not a multiple of the problem dimension {dim}")
            dvs = dvs.reshape(-1, dim)
        return self._evaluate(dvs)

    def device_context(self):
        """The device context this UDP evaluates on (mesh + voxel grid of self.args uploaded)."""
        return context_for_args(self.args, need_grid=True)

    def batch_fitness_device(self, d_xs) -> np.ndarray:
        """`batch_fitness` for decision vectors that already live on the GPU: d_xs is a CUDA float64 tensor (n, dim),
        e.g. the candidates `toss_gaco_sample` drew (optimization/gaco.py).  Returns the host fitness vector (n)."""
        if d_xs.dim() != 2 or d_xs.shape[1] != len(self.lower_bounds):
            raise ValueError(f"batch_fitness_device: expected (n, {len(self.lower_bounds)}), got {tuple(d_xs.shape)}")
        return self._evaluate(d_xs)

    # ------------------------------------------------------------------------------------------
    def _evaluate(self, xs) -> np.ndarray:
        """xs: host array (n, dim) or CUDA tensor (n, dim)"""
        args = self.args
        on_device = not isinstance(xs, np.ndarray)
        n_man = int(args.problem.number_of_maneuvers)
        fixed = np.asarray(self.initial_condition, dtype=np.float64).reshape(-1)
        if fixed.size + xs.shape[1] != 7 + 5 * n_man:
            raise ValueError(f"len(initial_condition) + len(chromosome) = {fixed.size + xs.shape[1]}, "
                             f"expected 7 + 5*number_of_maneuvers = {7 + 5 * n_man}")
        ctx = context_for_args(args, need_grid=True)
        # a collided trajectory scores the constant 1e30 whatever happens afterwards, so the batch path retires it at
        # the first event point found inside the mesh (same fitness vector, fewer gravity evaluations); the
        # compute_trajectory API keeps the reference behaviour and integrates to final_time
        p = params_from_args(args, stop_on_collision=1)
        n = len(xs)
        rank, w = P.world()
        if w == 1 and on_device:
            r = ctx.population_fitness(p, xs, fixed, n_man, self.knot_capacity)
            local = {k: v.cpu().numpy() for k, v in r.items()}
            local["counters"] = local["counters"].view(ctx_counters_dtype())
            fit = local["fitness"]
            status = local["counters"]["status"]
            short, failed = int((status == 3).sum()), int(((status == 1) | (status == 2)).sum())
        elif w == 1:
            local = ctx.population_fitness_host(p, xs, fixed, n_man, self.knot_capacity)
            fit = local["fitness"]
            status = local["counters"]["status"]
            short, failed = int((status == 3).sum()), int(((status == 1) | (status == 2)).sum())
        elif P._dist().get_backend() == "nccl":
            # rows dealt by predicted cost, everything device-side between the H2D copy of the chromosomes and the
            # D2H copy of the gathered fitness vector (parallel.py)
            r = P.sharded_population_fitness(ctx, p, xs if on_device else ctx.to_device_pinned(xs), fixed, n_man,
                                             self.knot_capacity)
            out = ctx.torch.cat((r["fitness"], r["flags"])).cpu().numpy()
            fit, short, failed = out[:n], int(out[n]), int(out[n + 1])
            local = r["local"]
            if local is not None:
                local = {k: v.cpu().numpy() for k, v in local.items()}
                local["counters"] = local["counters"].view(ctx_counters_dtype())
                local["rows"] = r["order"].cpu().numpy()[rank::w]
        else:
            # any other backend (gloo): the same dealing with the collectives on host arrays
            if on_device:
                xs = xs.cpu().numpy()
            lo, hi = P.shard_bounds(n, rank, w)
            cost = ctx.predict_cost(p, ctx._dev(xs[lo:hi]), fixed, n_man).cpu().numpy() if hi > lo else np.zeros(0, np.int32)
            order = P.cost_order_host(P.gather_costs_host(cost, n))
            rows = P.deal_host(order, rank, w)
            local = ctx.population_fitness_host(p, xs[rows], fixed, n_man, self.knot_capacity) if len(rows) else None
            fit, (short, failed) = P.exchange_dealt(None if local is None else local["fitness"],
                                                    None if local is None else local["counters"]["status"], order, n)
            if local is not None:
                local["rows"] = rows
        self.last_batch = local
        if local is not None:
            local["n_failed"] = failed
        if short:
            raise RuntimeError("a trajectory needs more knots than udp.knot_capacity")
        if failed:
            msg = (f"{failed} of {len(xs)} trajectories failed to integrate (step cap or non-finite state); "
                   "they score 1e30 like collisions -- see udp.last_batch['counters']['status']")
            if self.raise_on_integration_failure:
                raise FloatingPointError(msg)
            import warnings
            warnings.warn(msg, RuntimeWarning, stacklevel=3)
        return fit

    # pygmo deep-copies / pickles the UDP: nothing device-side lives on the object
    def __getstate__(self):
        d = dict(self.__dict__)
        d["last_batch"] = None
        return d
