"""Host-side glue between the reference-style `args` DotMap and the device context.

`args` is a plain (picklable) config object; the device context (mesh + voxel grid in HBM, ctypes and
torch handles) lives HERE, in a per-process cache with EXPLICIT keys -- (device, mesh identity) for everything that
reads a mesh, ("bare", device) for the mesh-independent calls -- so UDP objects can be deep-copied and pickled by
pygmo (SURVEY.md 8b "Ownership") and re-attach lazily in a new process.  A context holds one voxel grid at a time;
`set_grid` (the only place that uploads one) compares the content and re-uploads when a caller with another grid
comes along, so two UDPs that share a mesh but not a grid stay correct (each pays an 1 KB upload when they alternate).
"""
from __future__ import annotations

import zlib

import numpy as np

from . import _native as N

_contexts: dict = {}


def _mesh_key(vertices, faces, density):
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    f = np.ascontiguousarray(faces, dtype=np.int32)
    return (v.shape, f.shape, zlib.crc32(v.tobytes()), zlib.crc32(f.tobytes()), float(density))


def context_for_mesh(vertices, faces, density, device=None) -> N.Context:
    import torch
    dev = torch.cuda.current_device() if device is None and torch.cuda.is_available() else (device or 0)
    key = (dev,) + _mesh_key(vertices, faces, density)
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = N.Context(dev)
        ctx.upload_mesh(vertices, faces, density)
        ctx._grid_key = None
        _contexts[key] = ctx
    return ctx


def context_for_args(args, need_grid: bool = False) -> N.Context:
    """Device context for args.mesh (+ args.problem grid when `need_grid`)."""
    ctx = context_for_mesh(args.mesh.vertices, args.mesh.faces, args.body.density)
    if need_grid:
        set_grid(ctx, args.problem.tensor_grid_r, args.problem.tensor_grid_theta, args.problem.tensor_grid_phi,
                 args.problem.bool_tensor)
    return ctx


def bare_context(device=None) -> N.Context:
    """The context of the calls that need no mesh (fitness on given positions, rotations, voxel-grid updates): one
    per device, created on first use, holding NO mesh and its own voxel grid.  It is never a mesh-bearing context
    borrowed from the cache, so such calls cannot disturb the grid a UDP evaluates against (and vice versa)."""
    import torch
    dev = torch.cuda.current_device() if device is None and torch.cuda.is_available() else (device or 0)
    key = ("bare", dev)
    ctx = _contexts.get(key)
    if ctx is None:
        ctx = N.Context(dev)
        ctx._grid_key = None
        _contexts[key] = ctx
    return ctx


def release_contexts():
    """Destroy every cached device context (tests; long-lived processes that switch meshes)."""
    for ctx in list(_contexts.values()):
        ctx.close()
    _contexts.clear()


def set_grid(ctx: N.Context, r, theta, phi, bool_tensor):
    r, theta, phi = (np.ascontiguousarray(a, dtype=np.float64) for a in (r, theta, phi))
    bt = np.ascontiguousarray(bool_tensor, dtype=np.uint8)
    key = (r.tobytes(), theta.tobytes(), phi.tobytes(), bt.tobytes())   # bool_tensor is mutated in place by callers
    if getattr(ctx, "_grid_key", None) != key:
        ctx.upload_grid(r, theta, phi, bt)
        ctx._grid_key = key


def params_from_args(args, **over) -> N.TossParams:
    """Flatten the reference's `args` DotMap into the POD `toss_params` of the C ABI."""
    from math import pi
    pr, b, it = args.problem, args.body, args.integrator
    p = N.TossParams()
    p.density = float(b.density)
    axis = np.asarray(b.spin_axis, dtype=np.float64)
    p.spin_axis[:] = tuple(float(a) for a in axis)
    p.spin_velocity = float(b.spin_velocity)
    p.activate_rotation = int(bool(pr.activate_rotation))
    p.rtol, p.atol = float(it.rtol), float(it.atol)
    p.start_time, p.final_time = float(pr.start_time), float(pr.final_time)
    p.initial_time_step = float(pr.initial_time_step)
    p.activate_event = int(bool(pr.activate_event))
    p.number_of_spacecrafts = int(pr.get("number_of_spacecrafts", 1))
    p.radius_inner = float(pr.radius_inner_bounding_sphere)
    p.radius_outer = float(pr.radius_outer_bounding_sphere)
    p.measurement_period = float(pr.measurement_period)
    p.penalty_scaling_factor = float(pr.get("penalty_scaling_factor", 1.0))
    p.target_squared_altitude = float(pr.get("target_squared_altitude", 1.0))
    p.maximal_measurement_sphere_volume = float(pr.get("maximal_measurement_sphere_volume",
                                                       (4 / 3) * pi * 35.95398913 ** 3))
    p.total_measurable_volume = float(pr.get("total_measurable_volume", 1.0))
    p.max_steps = int(pr.get("max_steps", 200000))
    p.stop_on_collision = 0
    for k, v in over.items():
        setattr(p, k, v)
    return p


def default_params(**over) -> N.TossParams:
    return N.make_params(**over)
