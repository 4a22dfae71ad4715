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
