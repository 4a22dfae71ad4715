"""Pins the CPU oracle (oracle/toss_oracle.c) to every golden vector the reference's tests hold
for the hot path (tests/golden/reference_goldens.py cites file:line) and to outputs of the
reference's own numpy code run in the build container (tests/golden/ref_numpy_fixtures.npz,
made by tools/gen_golden_fixtures.py).  CPU only."""
import math

import numpy as np
import pytest

from oracle import oracle as O
from tests.golden import reference_goldens as G

TOL = dict(rtol=1e-5, atol=1e-5)   # the reference's own np.isclose tolerance


@pytest.fixture(scope="module")
def llp():
    v, f = O.load_mesh("llp")
    return O.Mesh(v, f)


def test_params(**over):
    kw = dict(final_time=G.TEST_FINAL_TIME, initial_time_step=G.TEST_DT0, measurement_period=2500.0)
    kw.update(over)
    return O.default_params(**kw)


test_params.__test__ = False


def test_mesh_sizes():
    for name, (nv, nf) in dict(lllp=(11, 18), llp=(93, 182), lp=(916, 1828), full=(9149, 18294)).items():
        v, f = O.load_mesh(name)
        assert v.shape == (nv, 3) and f.shape == (nf, 3)
        # closed, consistently oriented: every directed edge appears once and its reverse once
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        fwd = set(map(tuple, e))
        assert len(fwd) == len(e) and all((b, a) in fwd for a, b in fwd)


def test_integration_final_state(llp):
    """reference tests/integration_test.py:42-46"""
    tr = O.compute_trajectory(llp, test_params(), G.X0, 0)
    assert np.allclose(tr.y[-1], G.FINAL_STATE, **TOL)
    # far tighter than the reference asks: limited by the 9 printed digits of the golden
    assert np.max(np.abs(tr.y[-1] - G.FINAL_STATE) / np.abs(G.FINAL_STATE)) < 5e-9
    assert tr.counters["status"] == 0 and not tr.collision_detected


def test_fixed_step_positions(llp):
    """reference tests/fixed_step_trajectory_test.py:44-58 -- all 87 values, including the one
    (z at t = 45000 s) that only desolver's step sequence + cubic-Hermite interpolant reproduces."""
    p = test_params()
    tr = O.compute_trajectory(llp, p, G.X0, 0)
    pos, vel, ts = O.fixed_step(p, tr)
    assert pos.shape == (3, 29)
    assert np.allclose(ts, G.FIXED_STEP_TIMESTEPS, **TOL)
    assert np.all(np.isclose(pos, G.FIXED_STEP_POSITIONS, **TOL))
    assert len(set(np.diff(ts))) == 1


@pytest.mark.parametrize("n_man", range(9))
def test_multiple_impulsive_maneuvers(llp, n_man):
    """reference tests/multiple_impulsive_maneuvers_test.py:44-68"""
    x = G.X0.copy()
    for k in range(1, n_man + 1):
        x = np.concatenate((x, [5000.0 * k, 1.0, 0.1, 0.1, 0.1]))
    tr = O.compute_trajectory(llp, test_params(), x, n_man)
    assert np.allclose(tr.y[-1, 0:3], G.MANEUVER_FINAL_POSITIONS[:, n_man], **TOL)
    assert tr.n_seg == n_man + 1


def _fitness_case(llp):
    p = test_params(measurement_period=100.0, penalty_scaling_factor=1.0, radius_outer=10000.0,
                    target_squared_altitude=8000.0 ** 2)
    tr = O.compute_trajectory(llp, p, G.X0, 0)
    pos, vel, ts = O.fixed_step(p, tr)
    return p, pos, ts


def test_fitness_goldens(llp):
    """reference tests/fitness_functions_test.py:66-157"""
    p, pos, ts = _fitness_case(llp)
    grid = O.default_grid()
    assert pos.shape == (3, 720)
    assert np.isclose(O.get_fitness(4, p, pos, ts, grid), G.FIT_COVERED_VOLUME, **TOL)
    assert np.isclose(O.get_fitness(1, p, pos, ts, grid), G.FIT_TARGET_ALTITUDE, **TOL)
    assert np.isclose(O.get_fitness(5, p, pos, ts, grid), G.FIT_TOTAL_COVERED_VOLUME, **TOL)
    rin, rout = 4000.0, 10000.0
    line = np.zeros((3, 20))
    line[0] = np.arange(rin - 1000, rin + 1000, 100)
    assert np.isclose(O.get_fitness(2, p, line, np.zeros(20), grid), G.FIT_CLOSE_PENALTY, **TOL)
    line[0] = np.arange(rout - 1000, rout + 1000, 100)
    assert np.isclose(O.get_fitness(3, p, line, np.zeros(20), grid), G.FIT_FAR_PENALTY, **TOL)


def _sphere2cart(r, th, ph):
    return np.array([r * np.cos(th) * np.cos(ph), r * np.cos(th) * np.sin(ph), r * np.sin(th)])


def test_half_hemisphere_coverage():
    """reference tests/covered_space_test.py:118-146"""
    r, th, ph, bt = O.default_grid()
    assert (len(r), len(th), len(ph)) == (6, 8, 17)
    bt[:, :, int((len(ph) - 1) / 2):] = True
    axis, w = (0.0, 0.0, 1.0), 2 * math.pi / 43416
    for k, gold in ((9, G.COVERAGE_NO_GAIN), (6, G.COVERAGE_WITH_GAIN)):
        pos = _sphere2cart(r[0], th[0], ph[k])
        cov = O.space_coverage(1, axis, w, pos, [0.0], 4000.0, 12500.0, (r, th, ph, bt))
        assert abs(cov - gold) < 1e-14


def test_perfect_ratio_coverage():
    """reference tests/covered_space_test.py:35-115 (visit every grid point -> coverage >= 1)"""
    r, th, ph, bt = O.default_grid(period=1.0, rmin=2.0, rmax=8.0, scale=1.0)
    r, th, ph = r.copy(), th.copy(), ph.copy()
    for g in (r, th, ph):
        g[0] += 1e-4
        g[-1] -= 1e-4
    R, T, P = np.meshgrid(r, th, ph, indexing="ij")
    pos = np.vstack([c.ravel() for c in _sphere2cart(R, T, P)])
    axis, w = np.array([0.0, 0.0, 1.0]), 2 * math.pi / 43416
    ts = np.arange(pos.shape[1], dtype=np.float64)
    rot = np.array([O.rotate_point(-ts[i], pos[:, i], axis, w) for i in range(pos.shape[1])]).T
    cov, nvis = O.space_coverage(1, axis, w, rot, ts, 2.0, 8.0, (r, th, ph, bt), return_visited=True)
    assert nvis == bt.size and cov >= 1.0


def test_rotation_euler_rodrigues():
    """reference tests/point_rotation_test.py:47-76"""
    rng = np.random.default_rng(3)
    axis = np.array([0.15709817, 0.40925472, 0.89879405])
    w = 2 * math.pi / (12.06 * 3600)
    for _ in range(20):
        x = rng.integers(0, 20000, 3).astype(float)
        t = rng.random() * 1e5
        ax = axis / np.linalg.norm(axis)
        a = math.cos(-(w * t) / 2.0)
        b, c, d = -ax * math.sin(-(w * t) / 2.0)
        aa, bb, cc, dd = a * a, b * b, c * c, d * d
        bc, ad, ac, ab, bd, cd = b * c, a * d, a * c, a * b, b * d, c * d
        rot = np.array([[aa + bb - cc - dd, 2 * (bc + ad), 2 * (bd - ac)],
                        [2 * (bc - ad), aa + cc - bb - dd, 2 * (cd + ab)],
                        [2 * (bd + ac), 2 * (cd - ab), aa + dd - bb - cc]])
        assert np.allclose(rot @ x, O.rotate_point(t, x, axis, w), rtol=1e-12, atol=1e-9)


# ------------------------- outputs of the reference's own numpy code ------------------------------

@pytest.mark.parametrize("mesh_name", ["llp", "lp"])
def test_is_outside_matches_reference(fixtures_npz, mesh_name):
    """toss/mesh/mesh_utility.py:70-166 run verbatim (tools/gen_golden_fixtures.py): exact flags."""
    v, f = O.load_mesh(mesh_name)
    m = O.Mesh(v, f)
    pts = fixtures_npz[f"outside_{mesh_name}_pts"]
    ref = fixtures_npz[f"outside_{mesh_name}_flags"]
    assert 0.05 < ref.mean() < 0.95
    keep = np.ones(len(pts), dtype=bool)
    keep[fixtures_npz[f"outside_{mesh_name}_degenerate"]] = False   # rays exactly through a vertex/edge: BLAS-dependent
    assert np.array_equal(O.is_outside(m, pts)[keep], ref[keep])


def _grid(fx, fine):
    s = "grid2" if fine else "grid"
    return fx[f"{s}_r"], fx[f"{s}_theta"], fx[f"{s}_phi"]


def test_grid_matches_reference(fixtures_npz):
    """fitness_function_utils.py:231-279"""
    r, th, ph, bt = O.default_grid()
    assert np.array_equal(r, fixtures_npz["grid_r"]) and np.array_equal(th, fixtures_npz["grid_theta"])
    assert np.array_equal(ph, fixtures_npz["grid_phi"])
    r, th, ph, bt = O.default_grid(scale=8.0)
    assert np.array_equal(r, fixtures_npz["grid2_r"]) and np.array_equal(ph, fixtures_npz["grid2_phi"])


@pytest.mark.parametrize("ci", range(6))
def test_coverage_and_fitness_match_reference(fixtures_npz, ci):
    """compute_space_coverage / update_spherical_tensor_grid / get_fitness(1..9) of the reference,
    incl. the number_of_spacecrafts chunked rotation clock (fitness_function_utils.py:96-102):
    voxel tensors exact, values to 1e-13 relative."""
    fx = fixtures_npz
    N, nsc, fine = (int(a) for a in fx[f"cov{ci}_meta"])
    pos, ts, bt = fx[f"cov{ci}_pos"], fx[f"cov{ci}_ts"], fx[f"cov{ci}_bt"]
    g = _grid(fx, fine)
    axis, w = fx["axis"], float(fx["w"])
    cov, nvis, newt = O.space_coverage(nsc, axis, w, pos, ts, 4000.0, 12500.0, (*g, bt),
                                       return_visited=True, return_tensor=True)
    assert np.array_equal(newt, fx[f"cov{ci}_new_bt"])
    assert nvis == int(fx[f"cov{ci}_new_bt"].sum())
    assert abs(cov - float(fx[f"cov{ci}_coverage"])) <= 1e-13 * abs(float(fx[f"cov{ci}_coverage"]))
    assert np.array_equal(O.update_tensor(nsc, axis, w, pos, ts, 4000.0, 12500.0, (*g, bt)), fx[f"cov{ci}_new_bt"])
    p = O.default_params(number_of_spacecrafts=nsc, spin_axis=axis, spin_velocity=w,
                         target_squared_altitude=8000.0 ** 2)
    ref = fx[f"cov{ci}_fitness_1_9"]
    for e in range(1, 10):
        got = O.get_fitness(e, p, pos, ts, (*g, bt))
        assert abs(got - ref[e - 1]) <= 1e-12 * max(abs(ref[e - 1]), 1e-300), (e, got, ref[e - 1])


@pytest.mark.parametrize("si", range(5))
def test_fixed_step_index_logic_matches_reference(fixtures_npz, si):
    """trajectory_tools.py:46-74 run verbatim on stand-in segments whose dense output is a known
    quadratic: which segment serves which sample (manoeuvre-time samples -> earlier segment,
    sample-free segments) must agree.  A cubic Hermite reproduces a quadratic exactly up to rounding."""
    fx = fixtures_npz
    t0, tf, per = fx[f"seg{si}_params"]
    iv = fx[f"seg{si}_intervals"]
    p = O.default_params(start_time=t0, final_time=tf, measurement_period=per)
    c = np.arange(1, 7, dtype=np.float64)
    kt, ky, kf, seg = [], [], [], [0]
    for k in range(len(iv) - 1):
        t = np.linspace(iv[k], iv[k + 1], 5)
        kt.append(t)
        ky.append((1000.0 * k + c)[None, :] + 1e-3 * t[:, None] * c[None, :] + 1e-9 * t[:, None] ** 2)
        kf.append(1e-3 * c[None, :] + 2e-9 * t[:, None] * np.ones(6)[None, :])
        seg.append(seg[-1] + 5)
    tr = O.Trajectory(np.concatenate(kt), np.vstack(ky), np.vstack(kf), np.array(seg, dtype=np.int32), iv, 0,
                      np.zeros((0, 3)), {})
    pos, vel, ts = O.fixed_step(p, tr)
    assert np.array_equal(ts, fx[f"seg{si}_ts"])
    assert np.allclose(pos, fx[f"seg{si}_pos"], rtol=1e-13, atol=1e-9)
    assert np.allclose(vel, fx[f"seg{si}_vel"], rtol=1e-13, atol=1e-9)
    # the segment id is encoded in the thousands digit: exact agreement of the assignment
    assert np.array_equal(np.floor(pos[0] / 1000.0), np.floor(fx[f"seg{si}_pos"][0] / 1000.0))
