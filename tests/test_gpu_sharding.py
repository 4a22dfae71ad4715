"""The multi-GPU deal on ONE GPU: the ranks of a world are emulated as successive calls (no collective involved --
the all-gathers are concatenations here), so the kernels of csrc/sharding.cu and the pre-ordered work queue are
checked against their numpy mirrors (toss_b200/parallel.py) and against the single-shot evaluation, bit for bit."""
import numpy as np
import pytest

from toss_b200 import parallel as P
from tests.test_gpu_parity import FIXED_POS, both_params, bound_population, ctx_for, random_population

pytestmark = pytest.mark.gpu


def test_cost_order_kernel_equals_its_numpy_mirror():
    import torch
    c, _ = ctx_for("llp")
    rng = np.random.default_rng(11)
    for n in (1, 31, 1024, 1025, 5000, 32768):
        cost = rng.integers(0, 20000, size=n).astype(np.int32)
        cost[rng.integers(0, n, size=max(1, n // 7))] = 2500      # many ties: the order among them is by index
        order = c.cost_order(torch.from_numpy(cost).cuda()).cpu().numpy()
        assert np.array_equal(order, P.cost_order_host(cost)), n


@pytest.mark.parametrize("world", [1, 2, 3, 8])
def test_dealt_shards_reassemble_to_the_single_shot_result(world):
    import torch
    from toss_b200._native import COUNTERS_DTYPE
    c, _ = ctx_for("lp")
    ft = 43200.0
    p, _ = both_params(final_time=ft)
    n = 61
    xs_h = random_population(n, 2, 4242, ft)
    xs = torch.from_numpy(xs_h).cuda()
    whole = c.population_fitness(p, xs, FIXED_POS, 2)
    whole_fit = whole["fitness"].cpu().numpy()
    whole_cnt = whole["counters"].cpu().numpy().view(COUNTERS_DTYPE).copy()
    # (1) + (2): cost slices, concatenated
    cap = -(-n // world)
    cost = torch.zeros(cap * world, dtype=torch.int32, device="cuda")
    for r in range(world):
        lo, hi = P.shard_bounds(n, r, world)
        if hi > lo:
            c.predict_cost(p, xs[lo:hi], FIXED_POS, 2, out=cost[r * cap:r * cap + hi - lo])
    cost = cost[:n].contiguous()
    cost_h = cost.cpu().numpy()
    # the proxy ranks trajectories by length: on this mesh it must correlate with the real RHS counts
    assert np.corrcoef(cost_h, whole_cnt["n_rhs"])[0, 1] > 0.8
    order = c.cost_order(cost)
    order_h = order.cpu().numpy()
    assert np.array_equal(order_h, P.cost_order_host(cost_h))
    # (3)-(5) per emulated rank
    recv = []
    for r in range(world):
        x_local = c.deal_rows(xs, order, r, world)
        rows = P.deal_host(order_h, r, world)
        assert np.array_equal(x_local.cpu().numpy(), xs_h[rows])
        c.set_queue_order(True)
        try:
            loc = c.population_fitness(p, x_local, FIXED_POS, 2) if len(rows) else None
        finally:
            c.set_queue_order(False)
        if loc is not None:
            assert np.array_equal(loc["fitness"].cpu().numpy(), whole_fit[rows])
            lc = loc["counters"].cpu().numpy().view(COUNTERS_DTYPE)
            assert np.array_equal(lc["n_rhs"], whole_cnt["n_rhs"][rows])
        send = c.pack_result(loc["fitness"] if loc else None, loc["counters"] if loc else None, len(rows), cap)
        assert np.array_equal(send.cpu().numpy(), P.pack_host(whole_fit[rows], whole_cnt["status"][rows], cap),
                              equal_nan=True)
        recv.append(send)
    fit, flags = c.unpack_result(torch.cat(recv), order, n, world, cap)
    assert np.array_equal(fit.cpu().numpy(), whole_fit)
    assert flags.cpu().numpy().tolist() == [0.0, 0.0]
    f2, fl2 = P.unpack_host(torch.cat(recv).cpu().numpy(), order_h, n, world, cap)
    assert np.array_equal(f2, whole_fit) and fl2 == (0, 0)


def test_failure_counts_ride_with_the_fitness():
    import torch
    c, _ = ctx_for("llp")
    ft = 86400.0
    p, _ = both_params(final_time=ft)
    xs = torch.from_numpy(random_population(8, 2, 5, ft)).cuda()
    loc = c.population_fitness(p, xs, FIXED_POS, 2, 4)        # 4 knots: every trajectory runs out of capacity
    send = c.pack_result(loc["fitness"], loc["counters"], 8, 10).cpu().numpy()
    assert send[10] == 8.0 and send[11] == 0.0 and np.isnan(send[8:10]).all() and (send[:8] == 1e30).all()


def test_the_proxy_predicts_colliders_short_and_never_a_survivor():
    """The cost predictor retires a proxy trajectory deep inside the body when the real run would retire at its first
    event point inside the mesh (stop_on_collision): colliders are predicted short, and -- what matters for the
    schedule's tail -- no trajectory that survives is (the occupancy grid is eroded by one cell)."""
    import torch
    from toss_b200._native import COUNTERS_DTYPE
    c, _ = ctx_for("lp")
    n_man = 0
    r = FIXED_POS
    toward = -r / np.linalg.norm(r)
    rng = np.random.default_rng(12)
    # 24 chromosomes aimed at the body (1 m/s, +- 5 degrees), 40 from the converged-like family (long bound orbits)
    aimed = np.stack([np.concatenate(([1.0], toward + 0.08 * rng.normal(size=3))) for _ in range(24)])
    bound = bound_population(40, n_man, 77)
    xs_h = np.vstack((aimed, bound))
    xs = torch.from_numpy(xs_h).cuda()
    p1, _ = both_params(stop_on_collision=1)
    p0, _ = both_params(stop_on_collision=0)
    cost1 = c.predict_cost(p1, xs, FIXED_POS, n_man).cpu().numpy()
    cost0 = c.predict_cost(p0, xs, FIXED_POS, n_man).cpu().numpy()
    res = c.population_fitness(p1, xs, FIXED_POS, n_man, 2048)
    coll = res["collision"].cpu().numpy().astype(bool)
    n_rhs = res["counters"].cpu().numpy().view(COUNTERS_DTYPE)["n_rhs"]
    assert coll[:24].all() and not coll[24:].any()
    assert np.array_equal(cost1[~coll], cost0[~coll])             # a survivor's prediction is untouched
    assert (cost1[coll] < cost0[coll]).mean() > 0.8               # most colliders are caught ...
    assert np.median(cost1[coll]) < 0.5 * np.median(cost1[~coll])  # ... and predicted short, like their real cost
    assert np.median(n_rhs[coll]) < 0.5 * np.median(n_rhs[~coll])
