"""Mesh ingest and point-in-polyhedron test -- mirrors toss/mesh/mesh_utility.py.

read_pk_file :11-34 and the scaling of create_mesh :60 are reproduced (setup, host); the tetgen
tetrahedralisation (:63-64) only feeds plotting in the reference and is not part of this path, so
`create_mesh` returns None in its place.  is_outside :70-89 runs on the GPU (K5, toss_is_outside).
"""
import os
import pickle as pk
from pathlib import Path
from typing import Union

import numpy as np

MESH_SCALE = float(3126.6064453124995)   # mesh_utility.py:60
_BUNDLED = Path(__file__).resolve().parent.parent / "resources" / "meshes"
_BUNDLED_NAMES = {
    "churyumov-gerasimenko_lllp": "67p_lllp",
    "churyumov-gerasimenko_llp": "67p_llp",
    "churyumov-gerasimenko_lp": "67p_lp",
    "churyumov-gerasimenko": "67p_full",
}


def read_pk_file(filename):
    """(mesh_points, mesh_triangles) of a pickled mesh, non-dimensionalised as the reference does
    (mesh_utility.py:24-33).  If `filename` does not exist but names one of the four 67P meshes of the
    reference's 3dmeshes/ folder, the bundled re-encoded copy (resources/meshes/*.npz) is used."""
    if os.path.exists(filename):
        with open(filename, "rb") as f:
            mesh_points, mesh_triangles = pk.load(f)
        mesh_points = np.array(mesh_points)
        mesh_triangles = np.array(mesh_triangles)
    else:
        stem = Path(filename).stem
        if stem not in _BUNDLED_NAMES:
            raise FileNotFoundError(filename)
        z = np.load(_BUNDLED / f"{_BUNDLED_NAMES[stem]}.npz")
        mesh_points = np.array(z["points"], dtype=np.float64)
        mesh_triangles = np.array(z["triangles"])
    L = max(mesh_points[:, 0]) - min(mesh_points[:, 0])
    mesh_points = mesh_points / L * 2 * 0.8
    return mesh_points, mesh_triangles


def create_mesh(mesh_path: str) -> Union[None, np.ndarray, np.ndarray, float]:
    """Returns (body, vertices [m], faces, largest_body_protuberant) as the reference (:37-67); `body`
    (a tetgen object used only for plotting there) is None here."""
    mesh_points, mesh_triangles = read_pk_file(mesh_path)
    mesh_points = mesh_points * MESH_SCALE
    largest_protuberant = max(max(mesh_points[:, 0]), max(mesh_points[:, 1]), max(mesh_points[:, 2]))
    return None, mesh_points, mesh_triangles, largest_protuberant


def is_outside(points, mesh_vertices, mesh_triangles):
    """True where a point is outside the mesh (even number of +z ray hits), mesh_utility.py:70-89.
    points: (N,3) or (3,)."""
    from .._runtime import context_for_mesh
    pts = np.asarray(points, dtype=np.float64)
    single = pts.ndim == 1
    if pts.size == 0:
        return np.zeros(0, dtype=bool)
    if pts.reshape(-1, 3).shape[1] != 3:
        raise ValueError("Shape of points should be (N, 3)")
    ctx = context_for_mesh(mesh_vertices, mesh_triangles, 1.0)
    out = ctx.is_outside(pts.reshape(-1, 3)).cpu().numpy().astype(bool)
    return out[0:1] if single else out
