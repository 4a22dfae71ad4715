"""Generate tests/golden/ref_numpy_fixtures.npz by RUNNING the reference's own pure-numpy code.

Run ONCE in the build container (needs /root/reference; the GPU box has no such path):

    python tools/gen_golden_fixtures.py

The reference's fitness / ray-casting / resampling-index code is plain numpy, but `import toss`
pulls in third-party packages that are not installed here.  They are replaced by inert stand-ins
in sys.modules; the only stand-in that does arithmetic is `pyquaternion.Quaternion` (unit
quaternion sandwich product with a normalised axis, which is what pyquaternion documents).  No
reference source is copied: the reference modules are imported from where they lie.

What is pinned (inputs AND the reference's outputs are stored):
  * toss.mesh.mesh_utility.is_outside                 -> collision flags (integer parity, exact)
  * toss.fitness.fitness_function_utils.compute_space_coverage / update_spherical_tensor_grid /
    create_spherical_tensor_grid                       -> voxel tensors (exact) + coverage value
  * toss.fitness.fitness_functions.get_fitness for every enum 1..9
  * toss.trajectory.trajectory_tools.get_trajectory_fixed_step index logic (segment / sample
    assignment) with stand-in OdeSystem objects whose dense output is a known polynomial
"""
import sys
import types
from pathlib import Path

import numpy as np

REF = Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference")
OUT = Path(__file__).resolve().parent.parent / "tests" / "golden" / "ref_numpy_fixtures.npz"
sys.path.insert(0, str(REF))
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))


class _Inert(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        m = _Inert(self.__name__ + "." + name)
        setattr(self, name, m)
        return m

    def __call__(self, *a, **k):
        return self


for name in ["tetgen", "tetgen.pytetgen", "polyhedral_gravity", "polyhedral_gravity.model", "desolver",
             "desolver.backend", "pykep", "matplotlib", "matplotlib.pyplot", "matplotlib.patches",
             "mpl_toolkits", "mpl_toolkits.mplot3d", "pyvista"]:
    sys.modules[name] = _Inert(name)


class Quaternion:
    """axis-angle unit quaternion; rotate(v) = q v q* (pyquaternion semantics, axis normalised)."""

    def __init__(self, axis, angle):
        axis = np.asarray(axis, dtype=np.float64)
        axis = axis / np.linalg.norm(axis)
        self.w = np.cos(angle / 2.0)
        self.v = axis * np.sin(angle / 2.0)

    def rotate(self, x):
        x = np.asarray(x, dtype=np.float64)
        t = 2.0 * np.cross(self.v, x)
        return x + self.w * t + np.cross(self.v, t)


pq = types.ModuleType("pyquaternion")
pq.Quaternion = Quaternion
sys.modules["pyquaternion"] = pq


class DotMap(dict):
    def __init__(self, *a, _dynamic=True, **k):
        super().__init__(*a, **k)

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
except ImportError:  # the reference only calls toml.load(path)
    import tomllib
    tm = types.ModuleType("toml")
    tm.load = lambda p: tomllib.load(open(p, "rb"))
    sys.modules["toml"] = tm

from toss.fitness.fitness_function_utils import (compute_space_coverage, update_spherical_tensor_grid,  # noqa: E402
                                                 create_spherical_tensor_grid)
from toss.fitness.fitness_functions import get_fitness  # noqa: E402
from toss.fitness.fitness_function_enums import FitnessFunctions  # noqa: E402
from toss.mesh.mesh_utility import is_outside  # noqa: E402
from toss.trajectory.trajectory_tools import get_trajectory_fixed_step  # noqa: E402

from oracle import oracle as O  # noqa: E402  (mesh loader only: the bundled .npz + the reference scaling)

rng = np.random.default_rng(20261019)
out = {}

# ---- is_outside on the llp and lp meshes: points from deep inside to well outside --------------
for mesh_name, n in (("llp", 400), ("lp", 120)):
    v, f = O.load_mesh(mesh_name)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    pts = d * rng.uniform(100.0, 4000.0, size=(n, 1))
    # a few points exactly above vertices / on edge projections (ray through an edge or a vertex)
    pts[:8, 0:2] = v[rng.integers(0, len(v), 8), 0:2]
    pts[:8, 2] = -5000.0
    mid = 0.5 * (v[f[:8, 0]] + v[f[:8, 1]])
    pts[8:16, 0:2] = mid[:, 0:2]
    pts[8:16, 2] = -5000.0
    out[f"outside_{mesh_name}_pts"] = pts
    # the first 16 rays pass EXACTLY through a vertex / an edge: there the barycentric tests sit on a
    # rounding knife-edge and the reference's answer depends on how its BLAS (np.dot -> ddot/dgemv)
    # orders and fuses three products -- not reproducible across machines, excluded from exact parity
    out[f"outside_{mesh_name}_degenerate"] = np.arange(16)
    out[f"outside_{mesh_name}_flags"] = np.asarray(is_outside(pts, v, f), dtype=bool)

# ---- grid -----------------------------------------------------------------------------------------
fixed_velocity = np.array([-0.02826052, 0.1784372, -0.29885126])
gr, gth, gph, bt0 = create_spherical_tensor_grid(100, 4000, 12500, 40, fixed_velocity)
out["grid_r"], out["grid_theta"], out["grid_phi"] = gr, gth, gph
gr2, gth2, gph2, bt2 = create_spherical_tensor_grid(100, 4000, 12500, 8, fixed_velocity)   # finer grid
out["grid2_r"], out["grid2_theta"], out["grid2_phi"] = gr2, gth2, gph2

axis = np.array([0.15709817, 0.40925472, 0.89879405])
w = 2 * np.pi / 43416.0
out["axis"], out["w"] = axis, w


# ---- coverage / fitness on synthetic orbit-like trajectories ---------------------------------------
def make_positions(N, seed, rlo, rhi):
    r = np.random.default_rng(seed)
    ph = np.cumsum(r.uniform(0.0, 0.02, N)) + r.uniform(0, 6)
    inc = r.uniform(0, np.pi)
    rad = rlo + (rhi - rlo) * (0.5 + 0.5 * np.sin(np.linspace(0, r.uniform(2, 9), N) + r.uniform(0, 6)))
    x = rad * np.cos(ph)
    y = rad * np.sin(ph) * np.cos(inc)
    z = rad * np.sin(ph) * np.sin(inc)
    return np.vstack((x, y, z))


cases = []
for ci, (N, nsc, rlo, rhi, fine, prev) in enumerate([
        (6048, 4, 3500.0, 14000.0, False, False),
        (6048, 1, 3000.0, 13000.0, False, True),
        (6048, 4, 4500.0, 12000.0, True, True),
        (721, 3, 2000.0, 9000.0, False, False),     # N not divisible by nsc: array_split remainder
        (50, 7, 5000.0, 6000.0, True, False),
        (300, 1, 13000.0, 20000.0, False, True),    # nothing inside the shell: "empty" branch
]):
    pos = make_positions(N, 100 + ci, rlo, rhi)
    ts = np.arange(0, N * 100.0, 100.0)[:N]
    g = (gr2, gth2, gph2) if fine else (gr, gth, gph)
    bt = np.full((len(g[0]), len(g[1]), len(g[2])), False)
    if prev:
        bt |= np.random.default_rng(200 + ci).random(bt.shape) < 0.2
    vel = np.zeros_like(pos)
    cov = compute_space_coverage(nsc, axis, w, pos, vel, ts, 4000.0, 12500.0, g[0], g[1], g[2], bt.copy())
    new_bt = update_spherical_tensor_grid(nsc, axis, w, pos, vel, ts, 4000.0, 12500.0, g[0], g[1], g[2], bt.copy())
    args = DotMap(body=DotMap(spin_axis=axis, spin_velocity=w),
                  problem=DotMap(number_of_spacecrafts=nsc, target_squared_altitude=8000.0 ** 2,
                                 radius_inner_bounding_sphere=4000.0, radius_outer_bounding_sphere=12500.0,
                                 penalty_scaling_factor=0.1,
                                 maximal_measurement_sphere_volume=(4 / 3) * np.pi * 35.95398913 ** 3,
                                 total_measurable_volume=(4 / 3) * np.pi * (12500.0 ** 3 - 4000.0 ** 3),
                                 tensor_grid_r=g[0], tensor_grid_theta=g[1], tensor_grid_phi=g[2], bool_tensor=bt.copy()))
    fits = np.array([get_fitness(FitnessFunctions(e), args, pos, vel, ts) for e in range(1, 10)], dtype=np.float64)
    out[f"cov{ci}_pos"], out[f"cov{ci}_ts"], out[f"cov{ci}_bt"] = pos, ts, bt
    out[f"cov{ci}_meta"] = np.array([N, nsc, 1 if fine else 0])
    out[f"cov{ci}_coverage"] = np.float64(cov)
    out[f"cov{ci}_new_bt"] = np.asarray(new_bt, dtype=bool)
    out[f"cov{ci}_fitness_1_9"] = fits
out["n_cov_cases"] = np.int64(6)


# ---- get_trajectory_fixed_step index logic with stand-in segment objects ---------------------------
class FakeOde:
    """Dense output = a fixed cubic in t whose coefficients depend on the segment id."""

    def __init__(self, t_knots, sid):
        self.t = np.asarray(t_knots, dtype=np.float64)
        self.sid = sid

    def _OdeSystem__sol(self, t):
        t = np.atleast_1d(np.asarray(t, dtype=np.float64))
        c = np.arange(1, 7, dtype=np.float64)[None, :]
        return (1000.0 * self.sid + c) + 1e-3 * t[:, None] * c + 1e-9 * (t[:, None] ** 2)


seg_cases = [
    (0.0, 72000.0, 2500.0, [0, 72000]),
    (0.0, 72000.0, 2500.0, [0, 5000, 10000, 72000]),                 # samples exactly at manoeuvre times
    (0.0, 604800.0, 100.0, [0, 1234, 1250, 1299, 300000, 604800]),   # a segment with no sample (1250..1299)
    (0.0, 1000.0, 300.0, [0, 301, 1000]),
    (10.0, 2000.0, 7.0, [10, 11, 12, 500, 2000]),
]
for si, (t0, tf, per, iv) in enumerate(seg_cases):
    args = DotMap(problem=DotMap(start_time=t0, final_time=tf, measurement_period=per))
    objs = [FakeOde(np.linspace(iv[k], iv[k + 1], 5), k) for k in range(len(iv) - 1)]
    pos, vel, ts = get_trajectory_fixed_step(args, objs)
    out[f"seg{si}_params"] = np.array([t0, tf, per])
    out[f"seg{si}_intervals"] = np.array(iv, dtype=np.float64)
    out[f"seg{si}_pos"], out[f"seg{si}_vel"], out[f"seg{si}_ts"] = pos, vel, ts
out["n_seg_cases"] = np.int64(len(seg_cases))

np.savez_compressed(OUT, **out)
print("wrote", OUT, OUT.stat().st_size, "bytes;", len(out), "arrays")
