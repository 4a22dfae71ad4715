"""ctypes front-end of the CPU oracle (oracle/toss_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under toss_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

_HERE = Path(__file__).resolve().parent
_LIB_PATH = _HERE / "libtoss_oracle.so"
_MESH_DIR = _HERE.parent / "toss_b200" / "resources" / "meshes"

c_double_p = C.POINTER(C.c_double)
c_int_p = C.POINTER(C.c_int)


class OrcParams(C.Structure):
    _fields_ = [
        ("density", C.c_double),
        ("spin_axis", C.c_double * 3),
        ("spin_velocity", C.c_double),
        ("activate_rotation", C.c_int32),
        ("rtol", C.c_double),
        ("atol", C.c_double),
        ("start_time", C.c_double),
        ("final_time", C.c_double),
        ("initial_time_step", C.c_double),
        ("activate_event", C.c_int32),
        ("number_of_spacecrafts", C.c_int32),
        ("radius_inner", C.c_double),
        ("radius_outer", C.c_double),
        ("measurement_period", C.c_double),
        ("penalty_scaling_factor", C.c_double),
        ("target_squared_altitude", C.c_double),
        ("maximal_measurement_sphere_volume", C.c_double),
        ("total_measurable_volume", C.c_double),
        ("max_steps", C.c_int32),
        ("stop_on_collision", C.c_int32),
    ]


class OrcCounters(C.Structure):
    _fields_ = [
        ("n_rhs", C.c_int64),
        ("n_accepted", C.c_int32),
        ("n_rejected", C.c_int32),
        ("n_event_points", C.c_int32),
        ("status", C.c_int32),
    ]


COUNTERS_DTYPE = np.dtype(
    [("n_rhs", "<i8"), ("n_accepted", "<i4"), ("n_rejected", "<i4"), ("n_event_points", "<i4"), ("status", "<i4")]
)


def build(force: bool = False) -> Path:
    """Compile the oracle with the committed Makefile (gcc only)."""
    srcs = [_HERE / n for n in ("toss_oracle.c", "toss_oracle.h", "orc_math.h", "orc_math_constants.h",
                                "rk8713m_tableau.h", "Makefile")]
    if force or not _LIB_PATH.exists() or any(_LIB_PATH.stat().st_mtime < f.stat().st_mtime for f in srcs):
        subprocess.run(["make", "-C", str(_HERE)], check=True, capture_output=True)
    return _LIB_PATH


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(_LIB_PATH))
        L.orc_mesh_create.restype = C.c_void_p
        L.orc_mesh_create.argtypes = [c_double_p, C.c_int, C.POINTER(C.c_int32), C.c_int, C.c_double]
        L.orc_mesh_free.argtypes = [C.c_void_p]
        L.orc_mesh_pairs.restype = C.c_int
        L.orc_mesh_pairs.argtypes = [C.c_void_p, C.POINTER(C.c_int32)]
        L.orc_gravity.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p, C.c_int]
        L.orc_gravity_direct.argtypes = [C.c_void_p, c_double_p, C.c_int64, c_double_p, c_double_p]
        L.orc_gravity_tiers.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.POINTER(C.c_int64)]
        L.orc_math_probe.argtypes = [C.c_int, c_double_p, c_double_p, C.c_int64, c_double_p, c_double_p]
        L.orc_rotate_point.argtypes = [C.c_double, c_double_p, c_double_p, C.c_double, c_double_p]
        L.orc_compute_motion.argtypes = [C.c_void_p, C.POINTER(OrcParams), C.c_double, c_double_p, c_double_p]
        L.orc_compute_trajectory.restype = C.c_int
        L.orc_compute_trajectory.argtypes = [
            C.c_void_p, C.POINTER(OrcParams), c_double_p, C.c_int,
            c_double_p, c_double_p, c_double_p, C.c_int,
            c_int_p, c_double_p, c_int_p, c_int_p, c_int_p, c_double_p, C.c_int, C.POINTER(OrcCounters)]
        L.orc_num_samples.restype = C.c_int64
        L.orc_num_samples.argtypes = [C.POINTER(OrcParams)]
        L.orc_fixed_step.argtypes = [C.POINTER(OrcParams), c_double_p, c_double_p, c_double_p, c_int_p, C.c_int,
                                     c_double_p, c_double_p, c_double_p]
        L.orc_dense_eval.argtypes = [c_double_p, c_double_p, c_double_p, C.c_int, c_double_p, C.c_int64, c_double_p]
        L.orc_is_outside.argtypes = [C.c_void_p, c_double_p, C.c_int64, C.POINTER(C.c_uint8)]
        L.orc_get_fitness.restype = C.c_double
        L.orc_get_fitness.argtypes = [C.c_int, C.POINTER(OrcParams), c_double_p, C.c_int64, c_double_p,
                                      c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int,
                                      C.POINTER(C.c_uint8), C.POINTER(C.c_int64)]
        L.orc_space_coverage.restype = C.c_double
        L.orc_space_coverage.argtypes = [C.c_int, c_double_p, C.c_double, c_double_p, C.c_int64, c_double_p,
                                         C.c_double, C.c_double, c_double_p, C.c_int, c_double_p, C.c_int,
                                         c_double_p, C.c_int, C.POINTER(C.c_uint8), C.c_int,
                                         C.POINTER(C.c_int64), C.POINTER(C.c_uint8)]
        L.orc_update_tensor.argtypes = [C.c_int, c_double_p, C.c_double, c_double_p, C.c_int64, c_double_p,
                                        C.c_double, C.c_double, c_double_p, C.c_int, c_double_p, C.c_int,
                                        c_double_p, C.c_int, C.POINTER(C.c_uint8)]
        L.orc_population_fitness.argtypes = [
            C.c_void_p, C.POINTER(OrcParams), c_double_p, C.c_int, c_double_p, C.c_int, c_double_p, C.c_int,
            C.POINTER(C.c_uint8), c_double_p, C.c_int, C.c_int, c_double_p, C.c_int, C.c_int,
            c_double_p, C.POINTER(C.c_int32), C.c_void_p, C.POINTER(C.c_int64), C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


# --------------------------------------------------------------------------------------
# mesh fixtures (bundled .npz written by tools/import_meshes.py)
# --------------------------------------------------------------------------------------
MESH_SCALE = 3126.6064453124995  # mesh_utility.py:60


def load_mesh(name: str):
    """name in {lllp, llp, lp, full} -> (vertices[m] (V,3) f64, faces (F,3) i32).
    Scaling as read_pk_file + create_mesh (mesh_utility.py:29-33,60)."""
    z = np.load(_MESH_DIR / f"67p_{name}.npz")
    pts = np.array(z["points"], dtype=np.float64)
    L = max(pts[:, 0]) - min(pts[:, 0])
    pts = pts / L * 2 * 0.8
    pts = pts * float(MESH_SCALE)
    return pts, np.ascontiguousarray(z["triangles"], dtype=np.int32)


def default_params(**over) -> OrcParams:
    """Values of toss/resources/default_cfg.toml + the derived fields of setup_parameters.py:49-56."""
    from math import pi
    p = OrcParams()
    p.density = 533.0
    p.spin_axis[:] = (0.0, 0.0, 1.0)
    p.spin_velocity = (2 * pi) / 43416
    p.activate_rotation = 1
    p.rtol = 1e-12
    p.atol = 1e-12
    p.start_time = 0.0
    p.final_time = 604800.0
    p.initial_time_step = 1.0
    p.activate_event = 1
    p.number_of_spacecrafts = 4
    p.radius_inner = 4000.0
    p.radius_outer = 12500.0
    p.measurement_period = 100.0
    p.penalty_scaling_factor = 0.1
    p.target_squared_altitude = 16000000.0
    p.maximal_measurement_sphere_volume = (4 / 3) * pi * (35.95398913 ** 3)
    p.total_measurable_volume = (4 / 3) * pi * (12500.0 ** 3) - (4 / 3) * pi * (4000.0 ** 3)
    p.max_steps = 200000
    p.stop_on_collision = 0
    for k, v in over.items():
        if k == "spin_axis":
            p.spin_axis[:] = tuple(float(a) for a in v)
        else:
            setattr(p, k, v)
    return p


def default_grid(period=100.0, rmin=4000.0, rmax=12500.0, scale=40.0,
                 fixed_velocity=(-0.02826052, 0.1784372, -0.29885126)):
    """create_spherical_tensor_grid (fitness_function_utils.py:231-279)."""
    max_velocity = np.max(np.linalg.norm(np.asarray(fixed_velocity))) * scale
    d = max_velocity * period
    r_steps = np.floor((rmax - rmin) / d)
    th_steps = np.floor(np.pi * rmin / d)
    ph_steps = np.floor(2 * np.pi * rmin / d)
    r = np.linspace(rmin, rmax, int(r_steps))
    th = np.linspace(-np.pi / 2, np.pi / 2, int(th_steps))
    ph = np.linspace(-np.pi, np.pi, int(ph_steps))
    return r, th, ph, np.full((len(r), len(th), len(ph)), False)


class Mesh:
    def __init__(self, vertices, faces, density=533.0):
        self.vertices = _f64(vertices)
        self.faces = np.ascontiguousarray(faces, dtype=np.int32)
        self.handle = lib().orc_mesh_create(_dp(self.vertices), len(self.vertices),
                                            self.faces.ctypes.data_as(C.POINTER(C.c_int32)), len(self.faces),
                                            float(density))

    def __del__(self):
        try:
            lib().orc_mesh_free(self.handle)
        except Exception:
            pass

    @property
    def n_faces(self):
        return len(self.faces)

    def pairs(self):
        """The face pairing of the gravity sum: int32 [n][3] rows (fA, qA, fB), fB = -1 for a face left single."""
        out = np.empty((len(self.faces), 3), np.int32)
        n = lib().orc_mesh_pairs(self.handle, out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out[:n].copy()


def gravity(mesh: Mesh, pts, threads=1, potential=False):
    """Package-convention acceleration (TOSS uses the negative)."""
    pts = _f64(pts).reshape(-1, 3)
    acc = np.empty_like(pts)
    pot = np.empty(len(pts)) if potential else None
    lib().orc_gravity(mesh.handle, _dp(pts), len(pts), _dp(acc), _dp(pot) if potential else None, threads)
    return (acc, pot) if potential else acc


def gravity_direct(mesh: Mesh, pts):
    """Literal Tsoulis LN/AN transcription (cross-check of `gravity`)."""
    pts = _f64(pts).reshape(-1, 3)
    acc = np.empty_like(pts)
    lib().orc_gravity_direct(mesh.handle, _dp(pts), len(pts), _dp(acc), None)
    return acc


def gravity_tiers(mesh: Mesh, pts):
    """Census of the paths `gravity` takes for these points: dict(far_rows, mid_rows, term_rows, ln_terms, atan2_terms)."""
    pts = _f64(pts).reshape(-1, 3)
    out = (C.c_int64 * 5)()
    lib().orc_gravity_tiers(mesh.handle, _dp(pts), len(pts), out)
    return dict(zip(("far_rows", "mid_rows", "term_rows", "ln_terms", "atan2_terms"), [int(x) for x in out]))


MATH_FN = dict(two_atanh=0, log=1, atan2=2, sincos=3, root8=4, root4=5, atan_series8=6, div=7, sqrt=8,
               two_atanh_mm=9, two_atan_mm=10, two_atanh_far=11, two_atan_far=12, two_atanh_mid=13, two_atanh_wide=14)


def math_probe(fn: str, a, b=None):
    """orc_math.h function `fn` over arrays (tests): returns out, or (sin, cos) for sincos."""
    a = _f64(a).ravel()
    b = _f64(b).ravel() if b is not None else np.zeros_like(a)
    out, out2 = np.empty_like(a), np.empty_like(a)
    lib().orc_math_probe(MATH_FN[fn], _dp(a), _dp(b), len(a), _dp(out), _dp(out2))
    return (out, out2) if fn == "sincos" else out


def rotate_point(t, x, axis, w):
    x = _f64(x)
    axis = _f64(axis)
    out = np.empty(3)
    lib().orc_rotate_point(float(t), _dp(x), _dp(axis), float(w), _dp(out))
    return out


def compute_motion(mesh: Mesh, p: OrcParams, t, y):
    y = _f64(y)
    out = np.empty(6)
    lib().orc_compute_motion(mesh.handle, C.byref(p), float(t), _dp(y), _dp(out))
    return out


class Trajectory:
    """Result of compute_trajectory: concatenated knots + per-segment views."""

    def __init__(self, t, y, f, seg_start, intervals, collided, events, counters):
        self.t, self.y, self.f = t, y, f
        self.seg_start = seg_start
        self.intervals = intervals
        self.collision_detected = bool(collided)
        self.event_points = events
        self.counters = counters

    @property
    def n_seg(self):
        return len(self.seg_start) - 1

    def segment(self, k):
        a, b = self.seg_start[k], self.seg_start[k + 1]
        return self.t[a:b], self.y[a:b], self.f[a:b]


def compute_trajectory(mesh: Mesh, p: OrcParams, x, n_man: int, cap: int = 20000) -> Trajectory:
    x = _f64(x)
    assert len(x) == 7 + 5 * n_man
    kt = np.empty(cap)
    ky = np.empty((cap, 6))
    kf = np.empty((cap, 6))
    seg = np.zeros(n_man + 3, dtype=np.int32)
    iv = np.zeros(n_man + 3)
    n_seg = C.c_int()
    n_kn = C.c_int()
    coll = C.c_int()
    ev = np.empty((cap, 3))
    cnt = OrcCounters()
    rc = lib().orc_compute_trajectory(mesh.handle, C.byref(p), _dp(x), n_man, _dp(kt), _dp(ky), _dp(kf), cap,
                                      seg.ctypes.data_as(c_int_p), _dp(iv), C.byref(n_seg), C.byref(n_kn),
                                      C.byref(coll), _dp(ev), cap, C.byref(cnt))
    if rc != 0:
        raise RuntimeError("oracle knot capacity exceeded")
    n = n_kn.value
    ns = n_seg.value
    return Trajectory(kt[:n].copy(), ky[:n].copy(), kf[:n].copy(), seg[: ns + 1].copy(), iv[: ns + 1].copy(),
                      coll.value, ev[: min(cnt.n_event_points, cap)].copy(),
                      dict(n_rhs=cnt.n_rhs, n_accepted=cnt.n_accepted, n_rejected=cnt.n_rejected,
                           n_event_points=cnt.n_event_points, status=cnt.status))


def fixed_step(p: OrcParams, traj: Trajectory):
    N = lib().orc_num_samples(C.byref(p))
    pos = np.empty((3, N))
    vel = np.empty((3, N))
    ts = np.empty(N)
    seg = np.ascontiguousarray(traj.seg_start, dtype=np.int32)
    lib().orc_fixed_step(C.byref(p), _dp(traj.t), _dp(traj.y), _dp(traj.f), seg.ctypes.data_as(c_int_p),
                         traj.n_seg, _dp(pos), _dp(vel), _dp(ts))
    return pos, vel, ts


def dense_eval(t_knots, y_knots, f_knots, t):
    t = _f64(np.atleast_1d(t))
    out = np.empty((len(t), 6))
    lib().orc_dense_eval(_dp(_f64(t_knots)), _dp(_f64(y_knots)), _dp(_f64(f_knots)), len(t_knots), _dp(t),
                         len(t), _dp(out))
    return out


def is_outside(mesh: Mesh, pts):
    pts = _f64(pts).reshape(-1, 3)
    out = np.empty(len(pts), dtype=np.uint8)
    lib().orc_is_outside(mesh.handle, _dp(pts), len(pts), out.ctypes.data_as(C.POINTER(C.c_uint8)))
    return out.astype(bool)


def _grid_args(r, th, ph, bt):
    r, th, ph = _f64(r), _f64(th), _f64(ph)
    bt = np.ascontiguousarray(bt, dtype=np.uint8)
    assert bt.shape == (len(r), len(th), len(ph))
    return r, th, ph, bt


def get_fitness(fn: int, p: OrcParams, pos, ts, grid, return_visited=False):
    pos = _f64(pos)
    ts = _f64(ts)
    r, th, ph, bt = _grid_args(*grid)
    nv = C.c_int64(-1)
    v = lib().orc_get_fitness(int(fn), C.byref(p), _dp(pos), pos.shape[1], _dp(ts), _dp(r), len(r), _dp(th), len(th),
                              _dp(ph), len(ph), bt.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(nv))
    return (v, nv.value) if return_visited else v


def space_coverage(n_sc, axis, w, pos, ts, rmin, rmax, grid, return_visited=False, return_tensor=False):
    pos = _f64(pos)
    ts = _f64(np.atleast_1d(ts))
    axis = _f64(axis)
    r, th, ph, bt = _grid_args(*grid)
    single = 1 if pos.ndim == 1 else 0
    N = 1 if single else pos.shape[1]
    nv = C.c_int64(-1)
    out = np.empty_like(bt) if return_tensor else None
    v = lib().orc_space_coverage(int(n_sc), _dp(axis), float(w), _dp(pos), N, _dp(ts), float(rmin), float(rmax),
                                 _dp(r), len(r), _dp(th), len(th), _dp(ph), len(ph),
                                 bt.ctypes.data_as(C.POINTER(C.c_uint8)), single, C.byref(nv),
                                 out.ctypes.data_as(C.POINTER(C.c_uint8)) if return_tensor else None)
    res = (v,)
    if return_visited:
        res += (nv.value,)
    if return_tensor:
        res += (out.astype(bool),)
    return res if len(res) > 1 else v


def update_tensor(n_sc, axis, w, pos, ts, rmin, rmax, grid):
    pos = _f64(pos)
    ts = _f64(ts)
    axis = _f64(axis)
    r, th, ph, bt = _grid_args(*grid)
    bt = bt.copy()
    lib().orc_update_tensor(int(n_sc), _dp(axis), float(w), _dp(pos), pos.shape[1], _dp(ts), float(rmin), float(rmax),
                            _dp(r), len(r), _dp(th), len(th), _dp(ph), len(ph),
                            bt.ctypes.data_as(C.POINTER(C.c_uint8)))
    return bt.astype(bool)


def population_fitness(mesh: Mesh, p: OrcParams, grid, xs, fixed, n_man, threads=None):
    """udp.fitness over a population.  Returns dict(fitness, collision, counters, n_visited)."""
    xs = _f64(xs)
    pop, dim = xs.shape
    fixed = _f64(fixed)
    r, th, ph, bt = _grid_args(*grid)
    fit = np.empty(pop)
    coll = np.zeros(pop, dtype=np.int32)
    cnt = np.zeros(pop, dtype=COUNTERS_DTYPE)
    nv = np.zeros(pop, dtype=np.int64)
    if threads is None:
        threads = os.cpu_count() or 1
    lib().orc_population_fitness(mesh.handle, C.byref(p), _dp(r), len(r), _dp(th), len(th), _dp(ph), len(ph),
                                 bt.ctypes.data_as(C.POINTER(C.c_uint8)), _dp(xs), pop, dim,
                                 _dp(fixed) if len(fixed) else None, len(fixed), int(n_man),
                                 _dp(fit), coll.ctypes.data_as(C.POINTER(C.c_int32)), cnt.ctypes.data,
                                 nv.ctypes.data_as(C.POINTER(C.c_int64)), int(threads))
    return dict(fitness=fit, collision=coll, counters=cnt, n_visited=nv)
