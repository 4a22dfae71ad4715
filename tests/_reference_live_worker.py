"""Run by tests/test_host_vs_reference_live.py in a SUBPROCESS, in the build container only (needs /root/reference):
imports the reference's own modules from where they lie, with inert stand-ins for the third-party packages that are
not installed here (the same ones tools/gen_golden_fixtures.py uses; none of them does arithmetic on this path), runs
the reference's HOST-side functions on seeded inputs and writes inputs and outputs to an .npz.  Test infrastructure."""
import sys
import types

import numpy as np

ref_root, out_path = sys.argv[1], sys.argv[2]
sys.path.insert(0, ref_root)


class _Inert(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Inert(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return self

    def __iter__(self):          # `_, _ = tgen.tetrahedralize()` (mesh_utility.py:65: the result feeds plotting only)
        return iter((self, self))


for name in ["tetgen", "tetgen.pytetgen", "polyhedral_gravity", "polyhedral_gravity.model", "desolver",
             "desolver.backend", "pykep", "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits",
             "mpl_toolkits.mplot3d", "pyvista", "pyquaternion", "pygmo"]:
    sys.modules[name] = _Inert(name)


class DotMap(dict):
    def __init__(self, *a, _dynamic=True, **k):
        super().__init__(*a, **k)
        for key, v in list(self.items()):
            if isinstance(v, dict) and not isinstance(v, DotMap):
                self[key] = DotMap(v)

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v


dm = types.ModuleType("dotmap")
dm.DotMap = DotMap
sys.modules["dotmap"] = dm
try:
    import toml  # noqa: F401
except ImportError:
    import tomllib
    tm = types.ModuleType("toml")
    tm.load = lambda f: tomllib.loads(f.read()) if hasattr(f, "read") else tomllib.load(open(f, "rb"))
    sys.modules["toml"] = tm

import os  # noqa: E402
os.chdir(ref_root)                       # the cfg's mesh_path is relative to the repository root
from toss.optimization.setup_parameters import setup_parameters  # noqa: E402
from toss.optimization.setup_state import setup_initial_state_domain  # noqa: E402
from toss.trajectory.equations_of_motion import setup_spin_axis  # noqa: E402
from toss.fitness.fitness_function_utils import (create_spherical_tensor_grid, cart2sphere, sphere2cart,  # noqa: E402
                                                 _compute_squared_distance)

out = {}
args = setup_parameters()
for sec in ("problem", "body", "integrator", "optimization", "algorithm", "chromosome", "mesh"):
    for k, v in args[sec].items():
        if k in ("body", "evaluable"):
            continue
        if v is None:
            out[f"args.{sec}.{k}"] = np.array("None")
        elif isinstance(v, (bool, int, float, str, np.ndarray, list)):
            out[f"args.{sec}.{k}"] = np.asarray(v)
rng = np.random.default_rng(31)
xyz = rng.normal(size=(3, 500)) * 8000.0
out["xyz"] = xyz
out["cart2sphere"] = np.vstack(cart2sphere(xyz[0], xyz[1], xyz[2]))
r, th, ph = rng.uniform(1, 2e4, 500), rng.uniform(-np.pi / 2, np.pi / 2, 500), rng.uniform(-np.pi, np.pi, 500)
out["rtp"] = np.vstack([r, th, ph])
out["sphere2cart"] = np.vstack(sphere2cart(r, th, ph))
out["sqdist"] = _compute_squared_distance(xyz, 4000.0)
for i, (dt, rmin, rmax, scale) in enumerate([(100, 4000, 12500, 40), (100, 4000, 12500, 8), (50, 3000, 9000, 15.5)]):
    v = rng.normal(size=3)
    g = create_spherical_tensor_grid(dt, rmin, rmax, scale, v)
    out[f"grid{i}_v"] = v
    for nme, arr in zip("r theta phi bool".split(), g):
        out[f"grid{i}_{nme}"] = np.asarray(arr)
for n_fixed in (0, 3, 7):
    for n_man in (0, 1, 3):
        lb, ub = setup_initial_state_domain(np.zeros(n_fixed), 0, 604800, n_man, 4, args.chromosome)
        out[f"bounds_{n_fixed}_{n_man}_lo"], out[f"bounds_{n_fixed}_{n_man}_hi"] = np.asarray(lb, float), np.asarray(ub, float)
# ---- the reference's own compute_trajectory (decode, int32 truncation, sort, merge of simultaneous manoeuvres, the
# segment loop that REPLACES the velocity) around a stand-in integrator that returns its initial state: what is
# recorded is every (t0, tf, initial state) the reference hands to desolver, plus the intervals it returns
import toss.trajectory.compute_trajectory as ct  # noqa: E402


class _Traj:
    def __init__(self, x):
        self.y = np.array([x], dtype=np.float64)
        self.events = []


_calls = []


def _fake_integrate(func, x, a):
    _calls.append(np.concatenate(([float(a.integrator.t0), float(a.integrator.tf)], np.asarray(x, dtype=np.float64))))
    return _Traj(np.array(x, dtype=np.float64))


ct.integrate_system = _fake_integrate
ct.point_is_outside_mesh = lambda pts, v, f: np.ones(len(pts), dtype=bool)
T_END = 20000.0
args.problem.start_time, args.problem.final_time = 0.0, T_END
cases = []
for n_man in (0, 1, 2, 3, 5):
    for rep in range(6):
        x = np.concatenate((rng.normal(size=3) * 6000.0, [rng.uniform(0, 1)], rng.uniform(-1, 1, 3)))
        tm = rng.uniform(1, T_END - 1, n_man)
        if n_man >= 2 and rep % 3 == 1:
            tm[1] = np.floor(tm[0]) + 0.75          # equal after the int32 truncation
            tm[0] = np.floor(tm[0]) + 0.25
        if n_man >= 3 and rep % 3 == 2:
            tm[:] = np.floor(tm[0]) + rng.uniform(0, 0.99, n_man)   # all simultaneous
        if n_man >= 3 and rep == 3:
            tm[2] = np.floor(tm[0]) + 0.5           # first and third equal, second in between or not
        for k in range(n_man):
            x = np.concatenate((x, [tm[k], rng.uniform(0, 2.5)], rng.uniform(-1, 1, 3)))
        cases.append((n_man, x))
for i, (n_man, x) in enumerate(cases):
    args.problem.number_of_maneuvers = n_man
    _calls.clear()
    collided, objs, iv = ct.compute_trajectory(x.copy(), args, None)
    out[f"traj{i}_x"] = x
    out[f"traj{i}_nman"] = np.array(n_man)
    out[f"traj{i}_intervals"] = np.asarray(iv, dtype=np.float64)
    out[f"traj{i}_calls"] = np.array(_calls)
    assert collided is False and len(objs) == len(_calls)
out["traj_n"] = np.array(len(cases))
out["traj_T_END"] = np.array(T_END)

# ---- coverage / voxel update / all nine fitness functions on random walks and on constructed TIES --------------------
class Quaternion:
    """axis-angle unit quaternion; rotate(v) = q v q* (pyquaternion semantics, axis normalised) -- the one stand-in
    that does arithmetic (the same one tools/gen_golden_fixtures.py uses)"""

    def __init__(self, axis, angle):
        axis = np.asarray(axis, dtype=np.float64)
        axis = axis / np.linalg.norm(axis)
        self.w = np.cos(angle / 2.0)
        self.v = axis * np.sin(angle / 2.0)

    def rotate(self, x):
        x = np.asarray(x, dtype=np.float64)
        t = 2.0 * np.cross(self.v, x)
        return x + self.w * t + np.cross(self.v, t)


import toss.fitness.fitness_function_utils as ffu  # noqa: E402
import toss.trajectory.equations_of_motion as eom  # noqa: E402
for mod in (ffu, eom):
    if hasattr(mod, "Quaternion"):
        mod.Quaternion = Quaternion
from toss.fitness.fitness_functions import get_fitness  # noqa: E402
from toss.fitness.fitness_function_enums import FitnessFunctions  # noqa: E402

axis = np.array([0.15709817, 0.40925472, 0.89879405])
wspin = 2 * np.pi / 43416.0
gr, gth, gph, _ = create_spherical_tensor_grid(100, 4000, 12500, 40, np.array([-0.02826052, 0.1784372, -0.29885126]))
out["cov_grid_r"], out["cov_grid_theta"], out["cov_grid_phi"] = gr, gth, gph
cov_cases = []
for ci in range(24):
    N = int(rng.choice([1, 2, 3, 5, 29, 120, 400]))
    nsc = int(rng.choice([1, 2, 3, 4, 7]))
    step = rng.normal(size=(3, N)) * rng.choice([30.0, 300.0, 1500.0])
    pos = (rng.normal(size=(3, 1)) * rng.choice([2500.0, 5000.0, 9000.0])) + np.cumsum(step, axis=1)
    ts = np.arange(N) * 100.0
    w_case = wspin
    if ci % 6 == 4:        # ties: no rotation (angle exactly 0), samples exactly midway between two r grid points,
        w_case = 0.0       # on the equator (theta = 0 is midway between the two middle points of an even theta grid)
        mids = 0.5 * (gr[:-1] + gr[1:])
        k = rng.integers(0, len(mids), N)
        sgn = rng.choice([-1.0, 1.0], N)
        pos = np.vstack((sgn * mids[k], np.zeros(N), np.zeros(N)))
    if ci % 6 == 5:        # nothing inside the shell
        pos = pos / np.linalg.norm(pos, axis=0) * rng.uniform(13000.0, 30000.0, N)
    bt = rng.random((len(gr), len(gth), len(gph))) < rng.choice([0.0, 0.15, 0.9])
    cov_cases.append((N, nsc, w_case, pos, ts, bt))
for ci, (N, nsc, w_case, pos, ts, bt) in enumerate(cov_cases):
    vel = np.zeros_like(pos)
    cov = ffu.compute_space_coverage(nsc, axis, w_case, pos, vel, ts, 4000.0, 12500.0, gr, gth, gph, bt.copy())
    new_bt = ffu.update_spherical_tensor_grid(nsc, axis, w_case, pos, vel, ts, 4000.0, 12500.0, gr, gth, gph, bt.copy())
    fa = DotMap({"body": DotMap({"spin_axis": axis, "spin_velocity": w_case}),
                 "problem": DotMap({"number_of_spacecrafts": nsc, "target_squared_altitude": 8000.0 ** 2,
                                    "radius_inner_bounding_sphere": 4000.0, "radius_outer_bounding_sphere": 12500.0,
                                    "penalty_scaling_factor": 0.1,
                                    "maximal_measurement_sphere_volume": (4 / 3) * np.pi * 35.95398913 ** 3,
                                    "total_measurable_volume": (4 / 3) * np.pi * (12500.0 ** 3 - 4000.0 ** 3),
                                    "tensor_grid_r": gr, "tensor_grid_theta": gth, "tensor_grid_phi": gph,
                                    "bool_tensor": bt.copy()})})
    fits = []
    for e in range(1, 10):
        try:
            fits.append(float(get_fitness(FitnessFunctions(e), fa, pos, vel, ts)))
        except Exception:      # e.g. the covered-volume functions on a single sample
            fits.append(np.nan)
    out[f"cv{ci}_pos"], out[f"cv{ci}_ts"], out[f"cv{ci}_bt"] = pos, ts, bt
    out[f"cv{ci}_meta"] = np.array([N, nsc, w_case])
    out[f"cv{ci}_coverage"] = np.float64(cov)
    out[f"cv{ci}_new_bt"] = np.asarray(new_bt, dtype=bool)
    out[f"cv{ci}_fitness"] = np.array(fits)
out["cv_n"] = np.array(len(cov_cases))

# ---- is_outside on many points hugging the surface (where the +z ray parity is decided by few faces) -----------------
from toss.mesh.mesh_utility import is_outside, create_mesh  # noqa: E402
for mesh_name, n in (("lllp", 3000), ("llp", 3000), ("lp", 600)):
    _, mv, mf, _ = create_mesh(f"3dmeshes/churyumov-gerasimenko_{mesh_name}.pk")
    cen = mv[mf].mean(axis=1)                                   # face centroids
    pick = rng.integers(0, len(cen), n)
    pts = cen[pick] * rng.uniform(0.6, 1.4, (n, 1)) + rng.normal(size=(n, 3)) * 30.0
    out[f"io_{mesh_name}_pts"] = pts
    out[f"io_{mesh_name}_flags"] = np.asarray(is_outside(pts, mv, mf), dtype=bool)

# ---- the reference's resampling code fed with segment objects built from REAL trajectories (the oracle's knots and its
# dense output behind the name-mangled callable trajectory_tools.py:67 uses): pins the sample -> segment plumbing end to end
from pathlib import Path as _P  # noqa: E402
sys.path.insert(0, str(_P(__file__).resolve().parent.parent))
from oracle import oracle as _O  # noqa: E402
from toss.trajectory.trajectory_tools import get_trajectory_fixed_step, get_trajectory_adaptive_step  # noqa: E402


class _Seg:
    def __init__(self, t, y, f):
        self.t, self.y, self._f = t, y, f

    def _OdeSystem__sol(self, tq):
        return _O.dense_eval(self.t, self.y, self._f, np.atleast_1d(np.asarray(tq, dtype=np.float64)))


_mv, _mf = _O.load_mesh("lllp")
_mesh = _O.Mesh(_mv, _mf)
rs_cases = [(0, 86400.0, 100.0), (1, 86400.0, 100.0), (2, 50000.0, 250.0), (3, 20000.0, 7.0), (3, 20000.0, 100.0)]
for i, (n_man, t_end, period) in enumerate(rs_cases):
    x = np.concatenate(([-135.13402075, -4089.53592604, 6050.17636635], [rng.uniform(0.2, 0.6)], rng.uniform(-1, 1, 3)))
    tm = rng.uniform(1, t_end - 1, n_man)
    if n_man == 3 and period == 100.0:
        tm = np.array([5000.0, 5000.4, 7300.0])          # a merged pair, and manoeuvres exactly ON sample times
    for k in range(n_man):
        x = np.concatenate((x, [tm[k], rng.uniform(0, 0.3)], rng.uniform(-1, 1, 3)))
    p_ = _O.default_params(final_time=t_end, measurement_period=period, activate_event=0, rtol=1e-8, atol=1e-8)
    tr = _O.compute_trajectory(_mesh, p_, x, n_man)
    objs = [_Seg(*tr.segment(k)) for k in range(tr.n_seg)]
    ra = DotMap({"problem": DotMap({"start_time": 0.0, "final_time": t_end, "measurement_period": period})})
    pos, vel, ts = get_trajectory_fixed_step(ra, objs)
    st, tt = get_trajectory_adaptive_step(objs)
    out[f"rs{i}_x"], out[f"rs{i}_meta"] = x, np.array([n_man, t_end, period])
    out[f"rs{i}_pos"], out[f"rs{i}_vel"], out[f"rs{i}_ts"] = pos, vel, ts
    out[f"rs{i}_states"], out[f"rs{i}_times"] = st, tt
out["rs_n"] = np.array(len(rs_cases))

a2 = DotMap({"body": {"declination": 64.0, "right_ascension": 69.0}})
out["spin_axis_64_69"] = np.asarray(setup_spin_axis(a2))
np.savez(out_path, **out)
print("ok", len(out))
