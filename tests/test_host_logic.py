"""CPU tests of the host side: config contract, chromosome bounds, sharding + gather over a 2-rank gloo group,
the C-ABI library loads and exports every symbol include/toss_b200.h declares (no compute without a GPU)."""
import copy
import os
import pickle
import re
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def test_library_builds_and_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from toss_b200 import _native as N
    lib = N.lib()
    header = (ROOT / "include" / "toss_b200.h").read_text()
    declared = set(re.findall(r"\b(toss_[a-z0-9_]+)\s*\(", header))
    assert declared, "no prototypes found"
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in include/toss_b200.h but not exported"
    assert declared == set(N.EXPORTS)
    assert lib.toss_version().decode().startswith("toss_b200")


def test_struct_layouts_match_the_header():
    import ctypes as C
    from toss_b200 import _native as N
    from oracle import oracle as O
    assert C.sizeof(N.TossParams) == C.sizeof(O.OrcParams) == 160
    assert N.COUNTERS_DTYPE.itemsize == 24
    L = N.WsLayout()
    assert N.lib().toss_workspace_layout(10, 2, 64, C.byref(L)) == 0
    assert L.knots_y - L.knots_t >= 10 * 64 * 8 and L.total_bytes > L.queue
    assert N.lib().toss_workspace_layout(0, 2, 64, C.byref(L)) != 0
    assert b"bad argument" in N.lib().toss_last_error()


def test_no_gpu_means_a_loud_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from toss_b200 import _native as N
    with pytest.raises(N.TossError):
        N.Context(0)
    import toss_b200 as T
    args = T.setup_parameters()
    with pytest.raises(N.TossError):
        T.compute_motion(0.0, np.zeros(6) + 5000.0, args)


def test_product_never_imports_the_oracle():
    for f in (ROOT / "toss_b200").rglob("*"):
        if f.suffix in (".py", ".cu", ".cuh", ".h") and f.is_file():
            txt = f.read_text()
            assert "oracle" not in txt.replace("oracle_parameter", ""), f"{f} mentions the oracle"


def test_default_cfg_contract():
    import toss_b200 as T
    a = T.load_default_cfg()
    for sec in ("body", "integrator", "problem", "optimization", "chromosome", "algorithm", "mesh"):
        assert sec in a
    assert a.integrator.algorithm == 3 and a.integrator.rtol == 1e-12
    assert a.problem.final_time == 604800 and a.problem.number_of_spacecrafts == 4
    with pytest.raises(AttributeError):
        a.problem.no_such_key            # DotMap(_dynamic=False) semantics
    a.problem.new_key = 3
    assert a.problem.new_key == 3
    b = copy.deepcopy(a)
    b.problem.final_time = 1
    assert a.problem.final_time == 604800
    assert pickle.loads(pickle.dumps(a)).body.density == 533


def test_setup_parameters_derived_fields():
    import toss_b200 as T
    from math import pi
    a = T.setup_parameters()
    assert a.mesh.vertices.shape == (93, 3) and a.mesh.faces.shape == (182, 3)
    assert np.isclose(a.body.spin_velocity, 2 * pi / 43416)
    assert np.array_equal(a.body.spin_axis, [0, 0, 1])
    assert np.isclose(a.problem.total_measurable_volume, 4 / 3 * pi * (12500 ** 3 - 4000 ** 3))
    assert a.problem.bool_tensor.shape == (6, 8, 17) and not a.problem.bool_tensor.any()
    assert np.isclose(a.problem.tensor_grid_r[1] - a.problem.tensor_grid_r[0], 1700.0)
    assert 3000 < a.mesh.largest_body_protuberant < 4000


def test_spin_axis_goldens():
    """reference tests/point_rotation_test.py:78-103"""
    import toss_b200 as T
    from toss_b200.utilities.dotmap import DotMap
    from tests.golden import reference_goldens as G
    for (dec, ra), axis in G.SPIN_AXIS_CASES:
        a = DotMap(body=DotMap(declination=dec, right_ascension=ra, _dynamic=False), _dynamic=False)
        assert np.allclose(T.setup_spin_axis(a), axis, rtol=1e-5, atol=1e-5)


def test_chromosome_bounds_layouts():
    """setup_state.py:35-47: dim = 7+5M, 4+5M, 5M"""
    import toss_b200 as T
    a = T.load_default_cfg()
    for fixed, dim in ((np.zeros(0), 17), (np.zeros(3), 14), (np.zeros(7), 10)):
        lb, ub = T.setup_initial_state_domain(fixed, 0, 604800, 2, 4, a.chromosome)
        assert len(lb) == len(ub) == dim
        assert lb[-5] == 1 and ub[-5] == 604799 and ub[-4] == 2.5


def test_udp_is_picklable_and_validates_shapes():
    import toss_b200 as T
    a = T.setup_parameters()
    fixed = np.array([a.problem.initial_x, a.problem.initial_y, a.problem.initial_z])
    lb, ub = T.setup_initial_state_domain(fixed, a.problem.start_time, a.problem.final_time,
                                          a.problem.number_of_maneuvers, a.problem.number_of_spacecrafts, a.chromosome)
    udp = T.udp_initial_condition(a, fixed, lb, ub)
    assert udp.args.problem.number_of_spacecraft == 1 and udp.get_nobj() == 1
    assert np.array_equal(udp.get_bounds()[0], lb)
    u2 = pickle.loads(pickle.dumps(copy.deepcopy(udp)))
    assert np.array_equal(u2.get_bounds()[1], ub) and u2.args.mesh.faces.shape == (182, 3)
    with pytest.raises(ValueError):
        udp.batch_fitness(np.zeros(15))          # not a multiple of dim = 14
    a.problem.radius_inner_bounding_sphere = 100.0
    with pytest.raises(AssertionError):
        T.udp_initial_condition(a, fixed, lb, ub)


def test_cost_slices_and_dealt_rows_cover_the_population():
    """parallel.py: contiguous slices for the cost prediction (their concatenation IS the cost array), rows dealt
    round-robin over the longest-first order: every row exactly once, shard sizes within one, predicted work balanced."""
    from toss_b200.parallel import shard_bounds, champion, cost_order_host, deal_host, dealt_count
    rng = np.random.default_rng(3)
    for n in (0, 1, 7, 16, 4099, 32768):
        for w in (1, 2, 3, 8):
            cuts = [shard_bounds(n, r, w) for r in range(w)]
            cap = -(-n // w)
            assert cuts[0][0] == 0 and cuts[-1][1] == n
            assert all(cuts[i][1] == cuts[i + 1][0] for i in range(w - 1))
            assert all(a == min(n, r * cap) for r, (a, b) in enumerate(cuts)) and all(b - a <= cap for a, b in cuts)
            cost = rng.integers(200, 7000, size=n).astype(np.int32)
            order = cost_order_host(cost)
            assert sorted(order.tolist()) == list(range(n))
            key = np.clip(cost >> 2, 0, 4095)
            assert all(key[order[i]] > key[order[i + 1]] or (key[order[i]] == key[order[i + 1]] and order[i] < order[i + 1])
                       for i in range(n - 1))
            shards = [deal_host(order, r, w) for r in range(w)]
            assert [len(s) for s in shards] == [dealt_count(n, r, w) for r in range(w)]
            assert sorted(np.concatenate(shards).tolist()) == list(range(n))
            if n >= 4096:   # the deal balances the predicted work to a fraction of one trajectory
                work = [cost[s].sum() for s in shards]
                assert max(work) - min(work) <= cost.max()
    assert champion(np.array([3.0, 1.0, 1.0, 2.0])) == (1, 1.0)


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from toss_b200.parallel import (shard_bounds, gather_costs_host, cost_order_host, deal_host, exchange_dealt, champion,
                                world)
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:" + sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
assert world() == (int(sys.argv[3]), 2)
n = 11
rng = np.random.default_rng(4)
truth, cost = rng.normal(size=n), rng.integers(900, 6000, size=n).astype(np.int32)
# (1) every rank predicts the cost of its contiguous slice, (2) all-gather, (3) same order everywhere
lo, hi = shard_bounds(n)
costs = gather_costs_host(cost[lo:hi], n)
assert np.array_equal(costs, cost)
order = cost_order_host(costs)
rows = deal_host(order, rank, 2)
# (4) "evaluate" the dealt rows, (5) exchange + un-deal: the full vector in population order on every rank
full, (n_short, n_failed) = exchange_dealt(truth[rows], np.zeros(len(rows), np.int32), order, n)
assert np.array_equal(full, truth), (full, truth)
assert (n_short, n_failed) == (0, 0)
assert champion(full) == (int(np.argmin(truth)), float(truth.min()))
# a rank with an empty shard (n < world) still takes part in the collectives
o1 = cost_order_host(gather_costs_host(cost[:1] if shard_bounds(1)[1] > shard_bounds(1)[0] else cost[:0], 1))
r1 = deal_host(o1, rank, 2)
one, _ = exchange_dealt(truth[r1], None, o1, 1)
assert one.shape == (1,) and one[0] == truth[0] and len(r1) == (1 if rank == 0 else 0)
# failures seen by ONE rank (knot capacity: status 3; integrator: status 1 / 2) reach every rank as counts
st = np.zeros(len(rows), np.int32)
if rank == 1:
    st[0], st[1] = 3, 2
full, flags = exchange_dealt(truth[rows], st, order, n)
assert flags == (1, 1) and np.array_equal(full, truth)
dist.barrier(); dist.destroy_process_group()
print("ok", lo, hi, len(rows))
'''


def test_deal_and_gather_over_two_gloo_ranks(tmp_path):
    """SURVEY.md 8e at world size 2 (gloo stands in for NCCL): cost slices -> all-gather -> common longest-first order ->
    dealt rows -> all-gather of the fitness shards (+ failure counts) -> population order on every rank."""
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), str(ROOT), port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(2)]
    outs = [p.communicate(timeout=180)[0] for p in procs]
    assert all(p.returncode == 0 for p in procs), outs
    assert outs[0].strip().endswith("ok 0 6 6") and outs[1].strip().endswith("ok 6 11 5")


def test_grid_builder_input_assertions_are_the_reference_s():
    """fitness_function_utils.py:251-254 asserts on its four inputs (not on the resulting step counts): a spacing larger
    than the shell gives an EMPTY axis, which toss_grid_upload then refuses."""
    import toss_b200 as T
    v = np.array([0.1, 0.2, -0.3])
    for bad in ((0, 4000, 12500, 40), (100, 0, 12500, 40), (100, 4000, 4000, 40), (100, 4000, 12500, 0)):
        with pytest.raises(AssertionError):
            T.create_spherical_tensor_grid(*bad, v)
    r, th, ph, bt = T.create_spherical_tensor_grid(100, 4000, 12500, 400, v)     # 15 km per step > 8.5 km of shell
    assert len(r) == 0 and bt.shape == (0, len(th), len(ph)) and bt.dtype == bool


def test_covered_volume_of_a_single_sample_fails_like_the_reference():
    """estimate_covered_volume copies the second radius into the first (fitness_function_utils.py:30-31): with one
    sample the reference raises IndexError; so do the mirrors, before anything reaches the device."""
    import toss_b200 as T
    one = np.array([[5000.0], [100.0], [200.0]])
    with pytest.raises(IndexError):
        T.estimate_covered_volume(one)
    with pytest.raises(IndexError):
        T.covered_volume(1.0e5, one)
    with pytest.raises(IndexError):
        T.total_covered_volume(1.0e12, one)
