"""Bit-for-bit parity of the GPU path with the CPU oracle.

Both sides implement ONE arithmetic specification: every value is a fixed sequence of correctly rounded
IEEE-754 operations (+ - * / sqrt fma rint; oracle/orc_math.h <-> toss_b200/csrc/tmath.cuh), the sum over face
pairs (same deterministic pairing on both sides) runs in warp order (lane-strided partial sums + xor butterfly)
and the sample reductions in 256-thread-block order.
So the comparison is `array_equal`, not a tolerance: trajectories, knots, step counts, collision flags,
voxel counts and fitness values are IDENTICAL.  (The thread-per-point gravity kernel K1a sums the pairs
sequentially instead and agrees to ~1e-14.)"""
import numpy as np
import pytest

from oracle import oracle as O
from tests.test_gpu_parity import FIXED_POS, both_params, bound_population, ctx_for, points, random_population

pytestmark = pytest.mark.gpu


def test_guard_free_div_and_sqrt_are_ieee():
    c, _ = ctx_for("lllp")
    assert c.selftest_div_sqrt(400_000_000) == (0, 0)


def _fn_inputs():
    rng = np.random.default_rng(0)
    n = 300000
    return {
        "two_atanh": (np.concatenate([rng.uniform(0, 0.2, n), rng.uniform(0.2, 0.9999, n), 10.0 ** rng.uniform(-9, -1, n)]), None),
        "log": (np.concatenate([rng.uniform(1, 4, n), 10.0 ** rng.uniform(-200, 200, n)]), None),
        "atan2": (rng.normal(size=2 * n) * 10.0 ** rng.uniform(-3, 6, 2 * n), rng.normal(size=2 * n) * 10.0 ** rng.uniform(-3, 6, 2 * n)),
        "sincos": (np.concatenate([rng.uniform(-100, 100, n), rng.uniform(-1e4, 1e4, n)]), None),
        "root8": (10.0 ** rng.uniform(-30, 30, n), None),
        "root4": (10.0 ** rng.uniform(-30, 30, n), None),
        "atan_series8": (rng.uniform(-0.1, 0.1, n), None),
        "div": (10.0 ** rng.uniform(-20, 20, n) * rng.choice([-1, 1], n), 10.0 ** rng.uniform(-20, 60, n)),
        "sqrt": (10.0 ** rng.uniform(-20, 60, n), None),
        "two_atanh_mm": (np.concatenate([rng.uniform(-0.2, 0.2, n), 10.0 ** rng.uniform(-9, -1, n)]), None),
        "two_atan_mm": (np.concatenate([rng.uniform(-0.1, 0.1, n), 10.0 ** rng.uniform(-9, -1, n)]), None),
        "two_atanh_far": (np.concatenate([rng.uniform(-0.05, 0.05, n), 10.0 ** rng.uniform(-9, -2, n)]), None),
        "two_atan_far": (np.concatenate([rng.uniform(-0.05, 0.05, n), 10.0 ** rng.uniform(-9, -2, n)]), None),
        "two_atanh_mid": (np.concatenate([rng.uniform(-0.1, 0.1, n), 10.0 ** rng.uniform(-9, -1, n)]), None),
        "two_atanh_wide": (np.concatenate([rng.uniform(-0.4, 0.4, n), 10.0 ** rng.uniform(-9, -1, n)]), None),
    }


@pytest.mark.parametrize("mesh_name", ["lllp", "llp", "lp", "full"])
def test_face_pairing_identical(mesh_name):
    """the library's matching of faces into edge-sharing pairs == the oracle's (integer work, exact)"""
    c, m = ctx_for(mesh_name)
    got, ref = c.mesh_pairs(), m.pairs()
    assert np.array_equal(got, ref)
    single = ref[:, 2] < 0
    faces = np.concatenate([ref[:, 0], ref[~single, 2]])
    assert np.array_equal(np.sort(faces), np.arange(m.n_faces))     # every face exactly once
    assert single.sum() <= 0.02 * m.n_faces + 2


@pytest.mark.parametrize("fn", list(O.MATH_FN))
def test_elementary_functions_bit_identical(fn):
    c, _ = ctx_for("lllp")
    a, b = _fn_inputs()[fn]
    ref = O.math_probe(fn, a, b)
    got = c.math_probe(O.MATH_FN[fn], a, b)
    if fn == "sincos":
        assert np.array_equal(got[0], ref[0]) and np.array_equal(got[1], ref[1])
    else:
        assert np.array_equal(got, ref)


@pytest.mark.parametrize("mesh_name", ["lllp", "llp", "lp", "full"])
def test_rhs_bit_identical(mesh_name):
    """compute_motion (rotation + warp-order gravity): identical bits, far field, near field and inside."""
    c, m = ctx_for(mesh_name)
    p, po = both_params()
    rng = np.random.default_rng(5)
    n = 256 if mesh_name != "full" else 48
    pos = np.vstack([points(n // 2, 9), points(n // 4, 10, 300.0, 4000.0), points(n // 4, 11, 12500.0, 6.0e5)])
    ys = np.hstack([pos, rng.normal(size=(n, 3))])
    ts = rng.uniform(0, 604800, n)
    out = c.compute_motion(p, ts, ys).cpu().numpy()
    ref = np.array([O.compute_motion(m, po, ts[i], ys[i]) for i in range(n)])
    assert np.array_equal(out, ref)
    acc, pot = c.gravity(pos[:100], potential=True)          # warp-per-point kernel
    racc, rpot = O.gravity(m, pos[:100], potential=True)
    assert np.array_equal(acc.cpu().numpy(), racc) and np.array_equal(pot.cpu().numpy(), rpot)


@pytest.mark.parametrize("mesh_name,pop,n_man,kind,ft", [
    ("llp", 128, 2, "init", 604800.0), ("llp", 64, 2, "bound", 604800.0), ("lllp", 16, 1, "init", 604800.0),
    ("lp", 32, 2, "init", 172800.0), ("full", 8, 1, "bound", 43200.0), ("llp", 24, 8, "init", 604800.0)])
def test_population_bit_identical(mesh_name, pop, n_man, kind, ft):
    """knots, step counts, event counts, collision flags, voxel counts and fitness: all identical."""
    from toss_b200._native import COUNTERS_DTYPE
    c, m = ctx_for(mesh_name)
    p, po = both_params(final_time=ft)
    xs = (random_population if kind == "init" else bound_population)(pop, n_man, 4000 + pop, ft)
    res = c.population_fitness(p, xs, FIXED_POS, n_man)
    ref = O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, n_man)
    cnt = res["counters"].cpu().numpy().view(COUNTERS_DTYPE)
    for k in ("n_rhs", "n_accepted", "n_rejected", "n_event_points", "status"):
        assert np.array_equal(cnt[k], ref["counters"][k]), k
    assert np.array_equal(res["collision"].cpu().numpy(), ref["collision"])
    assert np.array_equal(res["n_visited"].cpu().numpy(), ref["n_visited"])
    assert np.array_equal(res["fitness"].cpu().numpy(), ref["fitness"])
    h = c.propagate(p, xs, FIXED_POS, n_man).to_host()
    for i in range(min(pop, 12)):
        tr = O.compute_trajectory(m, po, np.concatenate((FIXED_POS, xs[i])), n_man)
        n = len(tr.t)
        assert h["n_seg"][i] == tr.n_seg and np.array_equal(h["seg_start"][i, :tr.n_seg + 1], tr.seg_start)
        assert np.array_equal(h["knots_t"][i, :n], tr.t)
        assert np.array_equal(h["knots_y"][i, :n], tr.y)
        assert np.array_equal(h["knots_f"][i, :n], tr.f)
        assert np.array_equal(h["intervals"][i, :tr.n_seg + 1], tr.intervals)


def test_every_fitness_function_bit_identical(fixtures_npz):
    fx = fixtures_npz
    for ci in range(6):
        N, nsc, fine = (int(a) for a in fx[f"cov{ci}_meta"])
        s = "grid2" if fine else "grid"
        grid = (fx[f"{s}_r"], fx[f"{s}_theta"], fx[f"{s}_phi"], fx[f"cov{ci}_bt"])
        c, m = ctx_for("llp", grid)
        p, po = both_params(number_of_spacecrafts=nsc, spin_axis=fx["axis"], spin_velocity=float(fx["w"]),
                            target_squared_altitude=8000.0 ** 2)
        pos, ts = fx[f"cov{ci}_pos"], fx[f"cov{ci}_ts"]
        for e in range(1, 10):
            got = float(c.fitness(e, p, pos, ts).cpu()[0])
            assert got == O.get_fitness(e, po, pos, ts, grid), (ci, e)


def test_thread_per_point_kernel_close_to_warp_order():
    """K1a accumulates the faces sequentially (one thread owns a point): same terms, different summation order."""
    c, m = ctx_for("lp")
    pts = points(8192, 3)
    a = c.gravity(pts).cpu().numpy()
    r = O.gravity(m, pts, threads=8)
    assert (np.linalg.norm(a - r, axis=1) / np.linalg.norm(r, axis=1)).max() < 1e-13
