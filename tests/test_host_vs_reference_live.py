"""The host-side mirrors of the reference's setup code against the reference ITSELF, run live from /root/reference in a
subprocess (build container only: the GPU box has no such path, the test skips there).  Everything compared here is
plain numpy on both sides and is compared for equality: the derived `args` fields of setup_parameters (volumes, spin
state, mesh vertices / faces after the reference's scaling, the voxel grid), the chromosome bounds of every layout,
the spherical-coordinate helpers.  CPU only -- nothing here touches the CUDA library."""
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "toss" / "__init__.py").exists(), reason="reference checkout not present")


@pytest.fixture(scope="module")
def ref(tmp_path_factory):
    out = tmp_path_factory.mktemp("ref_live") / "ref.npz"
    r = subprocess.run([sys.executable, str(ROOT / "tests" / "_reference_live_worker.py"), str(REF), str(out)],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    return np.load(out, allow_pickle=False)


def test_setup_parameters_fields_equal_the_reference(ref):
    import toss_b200 as T
    args = T.setup_parameters()
    keys = [k for k in ref.files if k.startswith("args.")]
    assert len(keys) > 60
    for k in keys:
        _, sec, name = k.split(".", 2)
        want = ref[k]
        assert name in args[sec], f"missing args.{sec}.{name}"
        got = args[sec][name]
        if want.dtype.kind in "US":
            if str(want) == "None":
                assert got is None, k
            elif name == "mesh_path":
                assert Path(str(got)).name == Path(str(want)).name, k
            else:
                assert str(got) == str(want), k
        else:
            got = np.asarray(got)
            assert got.shape == want.shape and np.array_equal(got, want), k
    # nothing extra that the reference does not have, except the documented stand-ins
    for sec in ("problem", "body", "integrator", "optimization", "algorithm", "chromosome"):
        extra = [n for n in args[sec] if f"args.{sec}.{n}" not in ref.files]
        assert extra == [], (sec, extra)


def test_coordinate_helpers_and_grid_builder_equal_the_reference(ref):
    import toss_b200 as T
    xyz = ref["xyz"]
    assert np.array_equal(np.vstack(T.cart2sphere(xyz[0], xyz[1], xyz[2])), ref["cart2sphere"])
    r, th, ph = ref["rtp"]
    assert np.array_equal(np.vstack(T.sphere2cart(r, th, ph)), ref["sphere2cart"])
    assert np.array_equal(T._compute_squared_distance(xyz, 4000.0), ref["sqdist"])
    for i, (dt, rmin, rmax, scale) in enumerate([(100, 4000, 12500, 40), (100, 4000, 12500, 8), (50, 3000, 9000, 15.5)]):
        g = T.create_spherical_tensor_grid(dt, rmin, rmax, scale, ref[f"grid{i}_v"])
        for nme, arr in zip("r theta phi bool".split(), g):
            assert np.array_equal(np.asarray(arr), ref[f"grid{i}_{nme}"]), (i, nme)


def test_chromosome_bounds_and_spin_axis_equal_the_reference(ref):
    import toss_b200 as T
    from toss_b200.utilities.dotmap import DotMap
    args = T.load_default_cfg()
    for n_fixed in (0, 3, 7):
        for n_man in (0, 1, 3):
            lb, ub = T.setup_initial_state_domain(np.zeros(n_fixed), 0, 604800, n_man, 4, args.chromosome)
            assert np.array_equal(np.asarray(lb, float), ref[f"bounds_{n_fixed}_{n_man}_lo"])
            assert np.array_equal(np.asarray(ub, float), ref[f"bounds_{n_fixed}_{n_man}_hi"])
    a2 = DotMap({"body": DotMap({"declination": 64.0, "right_ascension": 69.0})})
    assert np.array_equal(np.asarray(T.setup_spin_axis(a2)), ref["spin_axis_64_69"])


def test_manoeuvre_decode_and_segment_loop_equal_the_reference(ref):
    """The reference's OWN compute_trajectory (compute_trajectory.py:53-136: v0 = v_mag * v_dir, int32 truncation of the
    manoeuvre times, stable sort, merge of simultaneous manoeuvres, velocity REPLACED at a manoeuvre) run around a
    stand-in integrator, against the oracle's trajectory on the same chromosomes: same integration intervals, same
    number of segments, the same velocity at the first knot of every segment, the same initial state."""
    from oracle import oracle as O
    v, f = O.load_mesh("lllp")
    m = O.Mesh(v, f)
    t_end = float(ref["traj_T_END"])
    n_merged = 0
    for i in range(int(ref["traj_n"])):
        x, n_man = ref[f"traj{i}_x"], int(ref[f"traj{i}_nman"])
        iv, calls = ref[f"traj{i}_intervals"], ref[f"traj{i}_calls"]
        p = O.default_params(final_time=t_end, activate_event=0, rtol=1e-6, atol=1e-6)
        tr = O.compute_trajectory(m, p, x, n_man)
        assert np.array_equal(tr.intervals, iv), (i, tr.intervals, iv)
        assert tr.n_seg == len(calls) == len(iv) - 1
        n_merged += int(len(iv) - 1 < n_man + 1)
        for k in range(tr.n_seg):
            t_k, y_k, _ = tr.segment(k)
            assert t_k[0] == calls[k, 0] and t_k[-1] == calls[k, 1]
            assert np.array_equal(y_k[0, 3:6], calls[k, 5:8]), (i, k)      # v0, then the (merged) dv of the manoeuvre
        assert np.array_equal(tr.y[0, 0:3], calls[0, 2:5])
    assert n_merged >= 4          # the simultaneous-manoeuvre cases did merge


def test_coverage_voxels_and_fitness_fuzz_equal_the_reference(ref):
    """24 random walks (1 .. 400 samples, 1 .. 7 spacecraft chunks, empty / sparse / dense previous tensors, samples
    inside, outside and across the shell) plus constructed TIES (no rotation, samples exactly midway between two r
    grid points and on the equator, where theta = 0 is midway between the two middle points of the even theta grid:
    argmin takes the first minimum) through the reference's compute_space_coverage / update_spherical_tensor_grid /
    get_fitness(1..9), against the oracle: voxel tensors exact, values to 1e-12."""
    from oracle import oracle as O
    g = (ref["cov_grid_r"], ref["cov_grid_theta"], ref["cov_grid_phi"])
    axis = np.array([0.15709817, 0.40925472, 0.89879405])
    n_ties = 0
    for ci in range(int(ref["cv_n"])):
        N, nsc, w = ref[f"cv{ci}_meta"]
        N, nsc = int(N), int(nsc)
        pos, ts, bt = ref[f"cv{ci}_pos"], ref[f"cv{ci}_ts"], ref[f"cv{ci}_bt"]
        want_bt, want_cov, want_fit = ref[f"cv{ci}_new_bt"], float(ref[f"cv{ci}_coverage"]), ref[f"cv{ci}_fitness"]
        n_ties += int(w == 0.0)
        cov, nvis, newt = O.space_coverage(nsc, axis, float(w), pos, ts, 4000.0, 12500.0, (*g, bt),
                                           return_visited=True, return_tensor=True)
        assert np.array_equal(newt, want_bt), ci
        assert nvis == int(want_bt.sum()), ci
        assert abs(cov - want_cov) <= 1e-12 * max(abs(want_cov), 1e-300), (ci, cov, want_cov)
        assert np.array_equal(O.update_tensor(nsc, axis, float(w), pos, ts, 4000.0, 12500.0, (*g, bt)), want_bt), ci
        p = O.default_params(number_of_spacecrafts=nsc, spin_axis=axis, spin_velocity=float(w),
                             target_squared_altitude=8000.0 ** 2)
        for e in range(1, 10):
            if np.isnan(want_fit[e - 1]):
                continue
            got = O.get_fitness(e, p, pos, ts, (*g, bt))
            assert abs(got - want_fit[e - 1]) <= 1e-12 * max(abs(want_fit[e - 1]), 1e-300), (ci, e, got, want_fit[e - 1])
    assert n_ties >= 3


@pytest.mark.parametrize("mesh_name", ["lllp", "llp", "lp"])
def test_is_outside_fuzz_equals_the_reference(ref, mesh_name):
    """mesh_utility.is_outside (Moeller-Trumbore, +z ray, hit parity) on thousands of points within +-40 % of the
    surface, against the oracle's ray casting: every flag equal, and both classes well represented."""
    from oracle import oracle as O
    v, f = O.load_mesh(mesh_name)
    pts, want = ref[f"io_{mesh_name}_pts"], ref[f"io_{mesh_name}_flags"]
    got = O.is_outside(O.Mesh(v, f), pts)
    assert np.array_equal(np.asarray(got, dtype=bool), want)
    assert 0.2 < want.mean() < 0.8


def test_resampling_plumbing_on_real_trajectories_equals_the_reference(ref):
    """The reference's get_trajectory_fixed_step / get_trajectory_adaptive_step run on segment objects built from the
    oracle's knots and dense output (a merged manoeuvre pair, manoeuvres exactly on sample times, a 7 s period), against
    the oracle's own resampler on the same trajectory: identical samples, times and concatenated knots."""
    from oracle import oracle as O
    v, f = O.load_mesh("lllp")
    m = O.Mesh(v, f)
    for i in range(int(ref["rs_n"])):
        n_man, t_end, period = ref[f"rs{i}_meta"]
        p = O.default_params(final_time=float(t_end), measurement_period=float(period), activate_event=0,
                             rtol=1e-8, atol=1e-8)
        tr = O.compute_trajectory(m, p, ref[f"rs{i}_x"], int(n_man))
        pos, vel, ts = O.fixed_step(p, tr)
        assert np.array_equal(ts, ref[f"rs{i}_ts"]), i
        assert np.array_equal(pos, ref[f"rs{i}_pos"]) and np.array_equal(vel, ref[f"rs{i}_vel"]), i
        assert np.array_equal(tr.y.T, ref[f"rs{i}_states"]) and np.array_equal(tr.t, ref[f"rs{i}_times"]), i
