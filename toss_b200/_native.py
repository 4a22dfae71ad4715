"""ctypes binding of libtoss_b200.so (include/toss_b200.h) -- the thin layer between the Python host and
the sm_100a kernels.  PyTorch is used for device-buffer ownership and streams only.

There is NO CPU fallback: if the shared library is missing or no B200 is visible, the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

_CSRC = Path(__file__).resolve().parent / "csrc"
LIB_PATH = Path(os.environ.get("TOSS_B200_LIB", _CSRC / "libtoss_b200.so"))   # override: developer experiments only

c_double_p = C.POINTER(C.c_double)


class TossParams(C.Structure):
    """include/toss_b200.h: toss_params"""
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


class WsLayout(C.Structure):
    """include/toss_b200.h: toss_ws_layout"""
    _fields_ = [(n, C.c_size_t) for n in ("knots_t", "knots_y", "knots_f", "seg_start", "intervals", "n_seg",
                                          "counters", "collision", "queue", "total_bytes")]


COUNTERS_DTYPE = np.dtype(
    [("n_rhs", "<i8"), ("n_accepted", "<i4"), ("n_rejected", "<i4"), ("n_event_points", "<i4"), ("status", "<i4")]
)

EXPORTS = [
    "toss_version", "toss_last_error", "toss_ctx_create", "toss_ctx_destroy", "toss_mesh_upload",
    "toss_mesh_num_faces", "toss_mesh_num_pairs", "toss_mesh_pairs", "toss_mesh_pair_records", "toss_pair_faces", "toss_grid_upload", "toss_gravity_eval", "toss_compute_motion",
    "toss_rotate_points", "toss_workspace_layout", "toss_propagate", "toss_integrate", "toss_collision_verdict", "toss_num_samples", "toss_resample", "toss_dense_eval",
    "toss_fitness_eval", "toss_update_tensor", "toss_is_outside", "toss_population_fitness",
    "toss_population_fitness_host", "toss_measure_fp64_peak", "toss_launch_count", "toss_math_probe",
    "toss_selftest_div_sqrt", "toss_set_queue_order", "toss_predict_cost", "toss_cost_order", "toss_deal_rows",
    "toss_pack_result", "toss_unpack_result", "toss_gaco_sample",
]

_lib = None


def build(force: bool = False) -> Path:
    """Compile libtoss_b200.so in-tree with nvcc for sm_100a (no GPU needed to build)."""
    import subprocess
    srcs = list(_CSRC.glob("*.cu")) + list(_CSRC.glob("*.cuh")) + list(_CSRC.glob("*.h"))
    srcs.append(_CSRC.parent.parent / "include" / "toss_b200.h")
    stale = (not LIB_PATH.exists()) or any(s.stat().st_mtime > LIB_PATH.stat().st_mtime for s in srcs)
    if force or stale:
        r = subprocess.run(["make", "-C", str(_CSRC), "-j4"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("building libtoss_b200.so failed:\n" + r.stdout + r.stderr)
    return LIB_PATH


def lib():
    """Load the shared library (raises if it has not been built: there is no fallback path)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(toss_b200 has no CPU fallback)")
    L = C.CDLL(str(LIB_PATH))
    vp, i32, i64, dbl, up = C.c_void_p, C.c_int, C.c_int64, C.c_double, C.c_size_t
    pp = C.POINTER(TossParams)
    L.toss_version.restype = C.c_char_p
    L.toss_last_error.restype = C.c_char_p
    L.toss_ctx_create.argtypes = [i32, C.POINTER(vp)]
    L.toss_ctx_destroy.argtypes = [vp]
    L.toss_mesh_upload.argtypes = [vp, c_double_p, i32, C.POINTER(C.c_int32), i32, dbl]
    L.toss_mesh_num_faces.argtypes = [vp]
    L.toss_mesh_num_pairs.argtypes = [vp]
    L.toss_mesh_pairs.argtypes = [vp, C.POINTER(C.c_int32)]
    L.toss_pair_faces.argtypes = [C.POINTER(C.c_int32), i32, i32, C.POINTER(C.c_int32)]
    L.toss_grid_upload.argtypes = [vp, c_double_p, i32, c_double_p, i32, c_double_p, i32, C.POINTER(C.c_uint8)]
    L.toss_gravity_eval.argtypes = [vp, vp, i64, vp, vp, up]
    L.toss_compute_motion.argtypes = [vp, pp, vp, vp, i64, vp, up]
    L.toss_workspace_layout.argtypes = [i32, i32, i32, C.POINTER(WsLayout)]
    L.toss_propagate.argtypes = [vp, pp, vp, i32, i32, c_double_p, i32, i32, vp, i32, up]
    L.toss_integrate.argtypes = [vp, pp, vp, i32, i32, c_double_p, i32, i32, vp, i32, up]
    L.toss_collision_verdict.argtypes = [vp, pp, vp, i32, i32, i32, up]
    L.toss_rotate_points.argtypes = [vp, c_double_p, dbl, vp, vp, i64, vp, up]
    L.toss_num_samples.restype = i64
    L.toss_num_samples.argtypes = [pp]
    L.toss_resample.argtypes = [vp, pp, vp, i32, i32, i32, vp, vp, vp, up]
    L.toss_dense_eval.argtypes = [vp, vp, vp, vp, i32, vp, i64, vp, up]
    L.toss_fitness_eval.argtypes = [vp, i32, pp, vp, i32, i64, vp, vp, vp, up]
    L.toss_update_tensor.argtypes = [vp, pp, vp, i64, vp, C.POINTER(C.c_uint8), up]
    L.toss_is_outside.argtypes = [vp, vp, i64, vp, up]
    L.toss_population_fitness.argtypes = [vp, pp, vp, i32, i32, c_double_p, i32, i32, vp, i32, vp, vp, vp, vp, up]
    L.toss_population_fitness_host.argtypes = [vp, pp, c_double_p, i32, i32, c_double_p, i32, i32, i32,
                                               c_double_p, C.POINTER(C.c_int32), vp, C.POINTER(C.c_int64)]
    L.toss_math_probe.argtypes = [vp, i32, vp, vp, i64, vp, vp, up]
    L.toss_selftest_div_sqrt.argtypes = [vp, i64, C.POINTER(C.c_int64), C.POINTER(C.c_int64)]
    L.toss_set_queue_order.argtypes = [vp, i32]
    L.toss_predict_cost.argtypes = [vp, pp, vp, i32, i32, c_double_p, i32, i32, vp, up]
    L.toss_cost_order.argtypes = [vp, vp, i32, vp, up]
    L.toss_deal_rows.argtypes = [vp, vp, i32, i32, vp, i32, i32, vp, up]
    L.toss_pack_result.argtypes = [vp, vp, vp, i32, i32, vp, up]
    L.toss_unpack_result.argtypes = [vp, vp, vp, i32, i32, i32, vp, vp, up]
    L.toss_mesh_pair_records.argtypes = [vp, c_double_p]
    L.toss_gaco_sample.argtypes = [vp, c_double_p, i32, i32, c_double_p, c_double_p, c_double_p, c_double_p,
                                   C.c_uint64, C.c_uint64, i32, i32, vp, up]
    L.toss_measure_fp64_peak.argtypes = [vp, c_double_p]
    L.toss_launch_count.restype = i64
    L.toss_launch_count.argtypes = [vp]
    _lib = L
    return L


class TossError(RuntimeError):
    pass


def _check(rc):
    if rc != 0:
        raise TossError(lib().toss_last_error().decode() or f"libtoss_b200 error {rc}")


def _dp(a):
    return a.ctypes.data_as(c_double_p)


def _f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def make_params(**kw) -> TossParams:
    """toss_params with the values of toss/resources/default_cfg.toml (+ derived fields of
    setup_parameters.py:49-56); keyword overrides by field name."""
    from math import pi
    p = TossParams()
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
    for k, v in kw.items():
        if k == "spin_axis":
            p.spin_axis[:] = tuple(float(a) for a in v)
        else:
            setattr(p, k, v)
    return p


def pair_faces(faces, n_vertices: int):
    """toss_pair_faces: the matching of a triangle list into edge-sharing pairs (host integer work, no device)."""
    f = np.ascontiguousarray(faces, dtype=np.int32)
    out = np.empty((len(f), 3), np.int32)
    n = lib().toss_pair_faces(f.ctypes.data_as(C.POINTER(C.c_int32)), len(f), int(n_vertices),
                              out.ctypes.data_as(C.POINTER(C.c_int32)))
    if n < 0:
        raise ValueError("toss_pair_faces: bad triangle list")
    return out[:n].copy()


class Context:
    """One device context (mesh + voxel grid resident in HBM).  Buffers are torch CUDA tensors."""

    def __init__(self, device: int | None = None):
        import torch
        if not torch.cuda.is_available():
            raise TossError("toss_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
        self.torch = torch
        self.device_index = torch.cuda.current_device() if device is None else int(device)
        self.device = torch.device("cuda", self.device_index)
        h = C.c_void_p()
        _check(lib().toss_ctx_create(self.device_index, C.byref(h)))
        self.h = h
        self.n_faces = 0
        self.grid_shape = None
        self._ws = None
        self._ws_key = None

    def close(self):
        if getattr(self, "h", None):
            lib().toss_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- setup ------------------------------------------------------------------------------------
    def upload_mesh(self, vertices, faces, density: float):
        v = _f64(vertices)
        f = np.ascontiguousarray(faces, dtype=np.int32)
        _check(lib().toss_mesh_upload(self.h, _dp(v), len(v), f.ctypes.data_as(C.POINTER(C.c_int32)), len(f),
                                      float(density)))
        self.n_faces = len(f)

    def mesh_pairs(self):
        """The face pairing of the gravity sum: int32 [n][3] rows (fA, qA, fB), fB = -1 for a face left single."""
        n = lib().toss_mesh_num_pairs(self.h)
        out = np.empty((n, 3), np.int32)
        _check(lib().toss_mesh_pairs(self.h, out.ctypes.data_as(C.POINTER(C.c_int32))))
        return out

    def mesh_pair_records(self):
        """The per-pair constants the device built at upload: float64 [n][43] (include/toss_b200.h)."""
        n = lib().toss_mesh_num_pairs(self.h)
        out = np.empty((n, 44))
        _check(lib().toss_mesh_pair_records(self.h, _dp(out)))
        return out[:, :43].copy()

    def upload_grid(self, r, theta, phi, bool_tensor):
        r, theta, phi = _f64(r), _f64(theta), _f64(phi)
        bt = np.ascontiguousarray(bool_tensor, dtype=np.uint8)
        if bt.shape != (len(r), len(theta), len(phi)):
            raise ValueError("bool_tensor shape does not match the grid axes")
        _check(lib().toss_grid_upload(self.h, _dp(r), len(r), _dp(theta), len(theta), _dp(phi), len(phi),
                                      bt.ctypes.data_as(C.POINTER(C.c_uint8))))
        self.grid_shape = bt.shape

    # ---- helpers ----------------------------------------------------------------------------------
    def _stream(self):
        return self.torch.cuda.current_stream(self.device).cuda_stream

    def _dev(self, a, dtype=None):
        t = self.torch
        if isinstance(a, t.Tensor):
            return a.to(device=self.device, dtype=dtype or t.float64).contiguous()
        return t.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(self.device)

    def to_device_pinned(self, a):
        """host array -> CUDA tensor through a pinned staging buffer the context keeps (grows on demand)"""
        t = self.torch
        a = np.ascontiguousarray(a, dtype=np.float64)
        if getattr(self, "_pin", None) is None or self._pin.numel() < a.size:
            self._pin = t.empty(max(a.size, 1), dtype=t.float64).pin_memory()
        stage = self._pin[:a.size].view(a.shape)
        stage.copy_(t.from_numpy(a))
        return stage.to(self.device, non_blocking=True)

    def launch_count(self) -> int:
        return int(lib().toss_launch_count(self.h))

    # ---- K1 ---------------------------------------------------------------------------------------
    def gravity(self, points, potential: bool = False):
        """package-convention acceleration (n,3) [and potential (n,)] as CUDA tensors"""
        t = self.torch
        pts = self._dev(points).reshape(-1, 3)
        acc = t.empty_like(pts)
        pot = t.empty(len(pts), dtype=t.float64, device=self.device) if potential else None
        _check(lib().toss_gravity_eval(self.h, pts.data_ptr(), len(pts), acc.data_ptr(),
                                       pot.data_ptr() if potential else None, self._stream()))
        return (acc, pot) if potential else acc

    def compute_motion(self, params: TossParams, ts, ys):
        t = self.torch
        ts = self._dev(ts).reshape(-1)
        ys = self._dev(ys).reshape(-1, 6)
        out = t.empty_like(ys)
        _check(lib().toss_compute_motion(self.h, C.byref(params), ts.data_ptr(), ys.data_ptr(), len(ys),
                                         out.data_ptr(), self._stream()))
        return out

    def rotate_points(self, axis, spin_velocity, ts, xs):
        t = self.torch
        ts = self._dev(ts).reshape(-1)
        xs = self._dev(xs).reshape(-1, 3)
        out = t.empty_like(xs)
        axis = _f64(axis)
        _check(lib().toss_rotate_points(self.h, _dp(axis), float(spin_velocity), ts.data_ptr(), xs.data_ptr(), len(xs),
                                        out.data_ptr(), self._stream()))
        return out

    def is_outside(self, points):
        t = self.torch
        pts = self._dev(points).reshape(-1, 3)
        out = t.empty(len(pts), dtype=t.uint8, device=self.device)
        _check(lib().toss_is_outside(self.h, pts.data_ptr(), len(pts), out.data_ptr(), self._stream()))
        return out

    # ---- K2 ---------------------------------------------------------------------------------------
    def workspace(self, pop: int, n_man: int, knot_cap: int):
        """The propagation workspace for this shape.  ONE buffer is cached per context and reused by every later
        propagate / integrate / population_fitness call of the same shape: a PropagationResult is a VIEW of it and is
        overwritten by the next such call -- copy what must survive (`to_host()`) before propagating again."""
        key = (pop, n_man, knot_cap)
        if self._ws_key != key:
            L = WsLayout()
            _check(lib().toss_workspace_layout(pop, n_man, knot_cap, C.byref(L)))
            self._ws = (self.torch.empty(L.total_bytes, dtype=self.torch.uint8, device=self.device), L)
            self._ws_key = key
        return self._ws

    def propagate(self, params: TossParams, xs, fixed, n_man: int, knot_cap: int = 4096):
        """compute_trajectory for a population; returns a PropagationResult over the workspace."""
        xs = self._dev(xs)
        pop, dim = xs.shape
        fixed = _f64(fixed)
        ws, L = self.workspace(pop, n_man, knot_cap)
        _check(lib().toss_propagate(self.h, C.byref(params), xs.data_ptr(), pop, dim,
                                    _dp(fixed) if len(fixed) else None, len(fixed), n_man, ws.data_ptr(), knot_cap,
                                    self._stream()))
        return PropagationResult(self, ws, L, pop, n_man, knot_cap)

    def integrate(self, params: TossParams, xs, fixed, n_man: int, knot_cap: int = 4096):
        """the stepper kernel alone (no collision verdict): what bench.py times for the roofline"""
        xs = self._dev(xs)
        pop, dim = xs.shape
        fixed = _f64(fixed)
        ws, L = self.workspace(pop, n_man, knot_cap)
        _check(lib().toss_integrate(self.h, C.byref(params), xs.data_ptr(), pop, dim,
                                    _dp(fixed) if len(fixed) else None, len(fixed), n_man, ws.data_ptr(), knot_cap,
                                    self._stream()))
        return PropagationResult(self, ws, L, pop, n_man, knot_cap)

    def collision_verdict(self, params: TossParams, prop: "PropagationResult"):
        """re-derives collision flags / event counts from the stored knots (the stepper already computed them inline)"""
        _check(lib().toss_collision_verdict(self.h, C.byref(params), prop.ws.data_ptr(), prop.pop, prop.n_man,
                                            prop.knot_cap, self._stream()))
        return prop

    def num_samples(self, params: TossParams) -> int:
        return int(lib().toss_num_samples(C.byref(params)))

    def resample(self, params: TossParams, prop: "PropagationResult", velocities: bool = True):
        t = self.torch
        N = self.num_samples(params)
        pos = t.empty((prop.pop, 3, N), dtype=t.float64, device=self.device)
        vel = t.empty((prop.pop, 3, N), dtype=t.float64, device=self.device) if velocities else None
        ts = t.empty(N, dtype=t.float64, device=self.device)
        _check(lib().toss_resample(self.h, C.byref(params), prop.ws.data_ptr(), prop.pop, prop.n_man, prop.knot_cap,
                                   pos.data_ptr(), vel.data_ptr() if velocities else None, ts.data_ptr(),
                                   self._stream()))
        return pos, vel, ts

    def resample_segments(self, params: TossParams, segments):
        """get_trajectory_fixed_step for ONE trajectory given as a list of (t (n,), y (n,6), f (n,6)) knot arrays,
        one entry per integration interval: the knots are packed into a one-trajectory propagation workspace and
        toss_resample does the rest on the device (sample -> segment bookkeeping of trajectory_tools.py:46-74
        included: csrc/fitness.cu build_seg_table).  Returns numpy positions (3,N), velocities (3,N), timesteps (N)."""
        t = self.torch
        n_seg = len(segments)
        if n_seg < 1:
            raise ValueError("resample_segments: no segments")
        counts = [len(s[0]) for s in segments]
        cap = max(4, sum(counts))
        L = WsLayout()
        _check(lib().toss_workspace_layout(1, n_seg - 1, cap, C.byref(L)))
        ws = np.zeros(L.total_bytes, dtype=np.uint8)
        kt = ws[L.knots_t:L.knots_t + 8 * cap].view(np.float64)
        ky = ws[L.knots_y:L.knots_y + 48 * cap].view(np.float64).reshape(cap, 6)
        kf = ws[L.knots_f:L.knots_f + 48 * cap].view(np.float64).reshape(cap, 6)
        seg_start = ws[L.seg_start:L.seg_start + 4 * (n_seg + 1)].view(np.int32)
        k = 0
        for i, (st, sy, sf) in enumerate(segments):
            n = counts[i]
            seg_start[i] = k
            kt[k:k + n], ky[k:k + n], kf[k:k + n] = st, sy, sf
            k += n
        seg_start[n_seg] = k
        ws[L.n_seg:L.n_seg + 4].view(np.int32)[0] = n_seg
        d_ws = t.from_numpy(ws).to(self.device)
        N = self.num_samples(params)
        out = t.empty((2, 3, max(N, 1)), dtype=t.float64, device=self.device)
        ts = t.empty(max(N, 1), dtype=t.float64, device=self.device)
        _check(lib().toss_resample(self.h, C.byref(params), d_ws.data_ptr(), 1, n_seg - 1, cap, out[0].data_ptr(),
                                   out[1].data_ptr(), ts.data_ptr(), self._stream()))
        h = out.cpu().numpy()
        return h[0, :, :N], h[1, :, :N], ts.cpu().numpy()[:N]

    def dense_eval(self, kt, ky, kf, times):
        t = self.torch
        kt, ky, kf, times = self._dev(kt).reshape(-1), self._dev(ky).reshape(-1, 6), self._dev(kf).reshape(-1, 6), \
            self._dev(times).reshape(-1)
        out = t.empty((len(times), 6), dtype=t.float64, device=self.device)
        _check(lib().toss_dense_eval(self.h, kt.data_ptr(), ky.data_ptr(), kf.data_ptr(), len(kt), times.data_ptr(),
                                     len(times), out.data_ptr(), self._stream()))
        return out

    # ---- K4 ---------------------------------------------------------------------------------------
    def fitness(self, fn: int, params: TossParams, positions, timesteps, return_visited: bool = False):
        """positions (batch,3,N) or (3,N); returns fitness tensor (batch,) [and n_visited]"""
        t = self.torch
        pos = self._dev(positions)
        if pos.dim() == 2:
            pos = pos.unsqueeze(0)
        ts = self._dev(timesteps).reshape(-1)
        B, _, N = pos.shape
        fit = t.empty(B, dtype=t.float64, device=self.device)
        nv = t.empty(B, dtype=t.int64, device=self.device)
        _check(lib().toss_fitness_eval(self.h, int(fn), C.byref(params), pos.data_ptr(), B, N, ts.data_ptr(),
                                       fit.data_ptr(), nv.data_ptr(), self._stream()))
        return (fit, nv) if return_visited else fit

    def update_tensor(self, params: TossParams, positions, timesteps):
        pos = self._dev(positions)
        ts = self._dev(timesteps).reshape(-1)
        out = np.empty(self.grid_shape, dtype=np.uint8)
        _check(lib().toss_update_tensor(self.h, C.byref(params), pos.data_ptr(), pos.shape[-1], ts.data_ptr(),
                                        out.ctypes.data_as(C.POINTER(C.c_uint8)), self._stream()))
        return out.astype(bool)

    # ---- the whole path -----------------------------------------------------------------------------
    def population_fitness(self, params: TossParams, xs, fixed, n_man: int, knot_cap: int = 4096):
        """device-resident batch_fitness: xs is a CUDA tensor (pop, dim); returns CUDA tensors."""
        t = self.torch
        xs = self._dev(xs)
        pop, dim = xs.shape
        fixed = _f64(fixed)
        ws, L = self.workspace(pop, n_man, knot_cap)
        fit = t.empty(pop, dtype=t.float64, device=self.device)
        coll = t.empty(pop, dtype=t.int32, device=self.device)
        cnt = t.empty(pop * COUNTERS_DTYPE.itemsize, dtype=t.uint8, device=self.device)
        nv = t.empty(pop, dtype=t.int64, device=self.device)
        _check(lib().toss_population_fitness(self.h, C.byref(params), xs.data_ptr(), pop, dim,
                                             _dp(fixed) if len(fixed) else None, len(fixed), n_man, ws.data_ptr(),
                                             knot_cap, fit.data_ptr(), coll.data_ptr(), cnt.data_ptr(), nv.data_ptr(),
                                             self._stream()))
        return dict(fitness=fit, collision=coll, counters=cnt, n_visited=nv)

    def population_fitness_host(self, params: TossParams, xs, fixed, n_man: int, knot_cap: int = 4096):
        """host buffers in, host buffers out (the C ABI does the copies): numpy arrays."""
        xs = _f64(xs)
        pop, dim = xs.shape
        fixed = _f64(fixed)
        fit = np.empty(pop)
        coll = np.zeros(pop, dtype=np.int32)
        cnt = np.zeros(pop, dtype=COUNTERS_DTYPE)
        nv = np.zeros(pop, dtype=np.int64)
        _check(lib().toss_population_fitness_host(self.h, C.byref(params), _dp(xs), pop, dim,
                                                  _dp(fixed) if len(fixed) else None, len(fixed), n_man, knot_cap,
                                                  _dp(fit), coll.ctypes.data_as(C.POINTER(C.c_int32)), cnt.ctypes.data,
                                                  nv.ctypes.data_as(C.POINTER(C.c_int64))))
        return dict(fitness=fit, collision=coll, counters=cnt, n_visited=nv)

    # ---- one population over several GPUs (include/toss_b200.h "one population over several GPUs") --------
    def set_queue_order(self, pre_ordered: bool):
        _check(lib().toss_set_queue_order(self.h, 1 if pre_ordered else 0))

    def predict_cost(self, params: TossParams, xs, fixed, n_man: int, out=None):
        """int32 cost key of every row of the CUDA tensor xs (pop, dim); `out` = int32 CUDA tensor of length >= pop"""
        t = self.torch
        pop, dim = xs.shape
        fixed = _f64(fixed)
        cost = t.empty(pop, dtype=t.int32, device=self.device) if out is None else out
        if pop:
            _check(lib().toss_predict_cost(self.h, C.byref(params), xs.data_ptr(), pop, dim,
                                           _dp(fixed) if len(fixed) else None, len(fixed), n_man, cost.data_ptr(),
                                           self._stream()))
        return cost

    def cost_order(self, cost):
        order = self.torch.empty(len(cost), dtype=self.torch.int32, device=self.device)
        _check(lib().toss_cost_order(self.h, cost.data_ptr(), len(cost), order.data_ptr(), self._stream()))
        return order

    def deal_rows(self, xs, order, rank: int, world: int):
        n, dim = xs.shape
        n_local = (n - rank + world - 1) // world
        out = self.torch.empty((n_local, dim), dtype=self.torch.float64, device=self.device)
        if n_local:
            _check(lib().toss_deal_rows(self.h, xs.data_ptr(), n, dim, order.data_ptr(), rank, world, out.data_ptr(),
                                        self._stream()))
        return out

    def pack_result(self, fitness, counters, n_local: int, cap: int):
        send = self.torch.empty(cap + 2, dtype=self.torch.float64, device=self.device)
        _check(lib().toss_pack_result(self.h, fitness.data_ptr() if n_local else None,
                                      counters.data_ptr() if (n_local and counters is not None) else None, n_local, cap,
                                      send.data_ptr(), self._stream()))
        return send

    def unpack_result(self, recv, order, n: int, world: int, cap: int):
        t = self.torch
        fit = t.empty(n, dtype=t.float64, device=self.device)
        flags = t.empty(2, dtype=t.float64, device=self.device)
        _check(lib().toss_unpack_result(self.h, recv.data_ptr(), order.data_ptr(), n, world, cap, fit.data_ptr(),
                                        flags.data_ptr(), self._stream()))
        return fit, flags

    def gaco_sample(self, archive_x, cum_weights, sigma, lb, ub, seed: int, generation: int, first_ant: int, n: int,
                    out=None):
        """n candidates of generation `generation` (ants first_ant ..) drawn on the device into a CUDA (n, dim) tensor"""
        t = self.torch
        archive_x = _f64(archive_x)
        ker, dim = archive_x.shape
        cum, sigma, lb, ub = _f64(cum_weights), _f64(sigma), _f64(lb), _f64(ub)
        if out is None:
            out = t.empty((n, dim), dtype=t.float64, device=self.device)
        _check(lib().toss_gaco_sample(self.h, _dp(archive_x), ker, dim, _dp(cum), _dp(sigma), _dp(lb), _dp(ub),
                                      seed & 0xFFFFFFFFFFFFFFFF, generation, first_ant, n, out.data_ptr(),
                                      self._stream()))
        return out

    def math_probe(self, fn: int, a, b=None):
        """tmath.cuh function `fn` over arrays (tests): numpy out, or (sin, cos) for fn 3"""
        t = self.torch
        a = self._dev(a).reshape(-1)
        b = self._dev(b).reshape(-1) if b is not None else t.zeros_like(a)
        out, out2 = t.empty_like(a), t.empty_like(a)
        _check(lib().toss_math_probe(self.h, int(fn), a.data_ptr(), b.data_ptr(), len(a), out.data_ptr(),
                                     out2.data_ptr(), self._stream()))
        return (out.cpu().numpy(), out2.cpu().numpy()) if fn == 3 else out.cpu().numpy()

    def selftest_div_sqrt(self, samples: int):
        bd, bs = C.c_int64(), C.c_int64()
        _check(lib().toss_selftest_div_sqrt(self.h, int(samples), C.byref(bd), C.byref(bs)))
        return bd.value, bs.value

    def measure_fp64_peak(self) -> float:
        v = C.c_double()
        _check(lib().toss_measure_fp64_peak(self.h, C.byref(v)))
        return v.value


class PropagationResult:
    """Views into the propagation workspace (device memory)."""

    def __init__(self, ctx: Context, ws, L: WsLayout, pop: int, n_man: int, knot_cap: int):
        self.ctx, self.ws, self.L, self.pop, self.n_man, self.knot_cap = ctx, ws, L, pop, n_man, knot_cap

    def _view(self, off, nbytes, dtype, shape):
        return self.ws[off:off + nbytes].view(dtype).reshape(shape)

    def to_host(self):
        """numpy copies: dict(knots_t, knots_y, knots_f, seg_start, intervals, n_seg, counters, collision)"""
        t = self.ctx.torch
        pop, cap, ns = self.pop, self.knot_cap, self.n_man + 2
        L = self.L
        out = {}
        out["knots_t"] = self._view(L.knots_t, 8 * pop * cap, t.float64, (pop, cap)).cpu().numpy()
        out["knots_y"] = self._view(L.knots_y, 48 * pop * cap, t.float64, (pop, cap, 6)).cpu().numpy()
        out["knots_f"] = self._view(L.knots_f, 48 * pop * cap, t.float64, (pop, cap, 6)).cpu().numpy()
        out["seg_start"] = self._view(L.seg_start, 4 * pop * ns, t.int32, (pop, ns)).cpu().numpy()
        out["intervals"] = self._view(L.intervals, 8 * pop * ns, t.float64, (pop, ns)).cpu().numpy()
        out["n_seg"] = self._view(L.n_seg, 4 * pop, t.int32, (pop,)).cpu().numpy()
        out["collision"] = self._view(L.collision, 4 * pop, t.int32, (pop,)).cpu().numpy()
        cnt = self.ws[L.counters:L.counters + COUNTERS_DTYPE.itemsize * pop].cpu().numpy()
        out["counters"] = cnt.view(COUNTERS_DTYPE)
        return out
