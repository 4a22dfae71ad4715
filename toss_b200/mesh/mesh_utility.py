"""Mesh ingest and point-in-polyhedron test -- mirrors toss/mesh/mesh_utility.py.

read_pk_file :11-34 and the scaling of create_mesh :60 are reproduced (setup, host); the tetgen
tetrahedralisation (:63-64) only feeds plotting in the reference and is not part of this path, so
`create_mesh` returns None in its place.  is_outside :70-89 runs on the GPU (K5, toss_is_outside).

Beyond the reference's pickles (README "Use Case 2": another body): `read_obj_file` (Wavefront .obj), `read_npz_file`
and `mesh_from_arrays` bring any triangulated closed surface in metres to the same (vertices, faces) pair, and
`check_mesh` says -- before anything is uploaded -- what the device path requires of it: every edge shared by exactly
two faces in opposite directions (closed, consistently oriented), outward winding, no zero-area face.  The per-pair
constants themselves are computed on the device at upload (csrc/pairs.cu).
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


def read_obj_file(filename):
    """(vertices, triangles) of a Wavefront .obj: `v x y z` and `f a b c ...` records (1-based, negative = relative,
    `a/b/c` texture / normal references ignored); polygons are fanned into triangles.  Units are the file's."""
    verts, tris = [], []
    with open(filename, "r") as f:
        for line in f:
            tok = line.split()
            if not tok or tok[0].startswith("#"):
                continue
            if tok[0] == "v":
                verts.append([float(tok[1]), float(tok[2]), float(tok[3])])
            elif tok[0] == "f":
                idx = []
                for t in tok[1:]:
                    i = int(t.split("/")[0])
                    idx.append(i - 1 if i > 0 else len(verts) + i)
                if len(idx) < 3:
                    raise ValueError(f"{filename}: face with fewer than three vertices")
                for k in range(1, len(idx) - 1):
                    tris.append([idx[0], idx[k], idx[k + 1]])
    return mesh_from_arrays(np.array(verts, dtype=np.float64).reshape(-1, 3), np.array(tris, dtype=np.int64).reshape(-1, 3))


def read_npz_file(filename):
    """(vertices, triangles) of an .npz with arrays `points` / `triangles` (or `vertices` / `faces`)."""
    z = np.load(filename)
    pts = z["points"] if "points" in z else z["vertices"]
    tri = z["triangles"] if "triangles" in z else z["faces"]
    return mesh_from_arrays(pts, tri)


def mesh_from_arrays(vertices, faces):
    """Validated (vertices float64 (V,3), faces int32 (F,3)) of a triangulated surface; raises ValueError with the
    first problem `check_mesh` finds."""
    v = np.ascontiguousarray(vertices, dtype=np.float64)
    f = np.ascontiguousarray(faces)
    if v.ndim != 2 or v.shape[1] != 3 or f.ndim != 2 or f.shape[1] != 3:
        raise ValueError("mesh_from_arrays: expected vertices (V, 3) and triangular faces (F, 3)")
    if not np.issubdtype(f.dtype, np.integer):
        raise ValueError("mesh_from_arrays: faces must hold integer vertex indices")
    f = f.astype(np.int32)
    problems = check_mesh(v, f)
    if problems:
        raise ValueError("mesh_from_arrays: " + problems[0])
    return v, f


def check_mesh(vertices, faces):
    """What toss_mesh_upload would refuse, as a list of messages (empty: the mesh is usable): index range, edges
    without an opposite partner (open or inconsistently oriented surface), inward winding, zero-area faces."""
    v = np.asarray(vertices, dtype=np.float64)
    f = np.asarray(faces, dtype=np.int64)
    out = []
    if len(v) < 4 or len(f) < 4:
        return ["a closed mesh needs at least 4 vertices and 4 faces"]
    if not np.isfinite(v).all():
        out.append("non-finite vertex coordinates")
    if f.min() < 0 or f.max() >= len(v):
        return out + ["face index out of range"]
    he = np.concatenate((f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]))              # directed half-edges
    key = he[:, 0] * len(v) + he[:, 1]
    opp = he[:, 1] * len(v) + he[:, 0]
    uniq, counts = np.unique(key, return_counts=True)
    if (counts > 1).any():
        out.append(f"{int((counts > 1).sum())} directed edge(s) used by more than one face (inconsistent orientation "
                   "or non-manifold surface)")
    missing = ~np.isin(opp, uniq)
    if missing.any():
        h = int(np.flatnonzero(missing)[0])
        out.append(f"{int(missing.sum())} face edge(s) without an opposite partner: the surface is not closed and "
                   f"consistently oriented (first: face {h % len(f)}, edge {h // len(f)})")
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    n = np.cross(b - a, c - b)
    if (np.einsum("ij,ij->i", n, n) == 0.0).any():
        out.append(f"{int((np.einsum('ij,ij->i', n, n) == 0.0).sum())} zero-area face(s)")
    if not out and np.einsum("ij,ij->", a, np.cross(b, c)) <= 0.0:
        out.append("the faces are wound clockwise seen from outside (negative volume): flip the vertex order of every face")
    return out


def read_mesh_file(filename):
    """(vertices, triangles) by extension: .pk (the reference's pickles, non-dimensionalised like read_pk_file),
    .obj, .npz (both taken as they are)."""
    ext = Path(filename).suffix.lower()
    if ext == ".obj":
        return read_obj_file(filename)
    if ext == ".npz":
        return read_npz_file(filename)
    return read_pk_file(filename)


def create_mesh(mesh_path: str) -> Union[None, np.ndarray, np.ndarray, float]:
    """Returns (body, vertices [m], faces, largest_body_protuberant) as the reference (:37-67); `body`
    (a tetgen object used only for plotting there) is None here."""
    if Path(mesh_path).suffix.lower() in (".obj", ".npz"):
        # another body (README "Use Case 2"): coordinates are taken in metres as they are -- the reference's
        # normalise-then-scale-by-3126.6 step (:24-33, :60) is specific to its 67P pickles
        mesh_points, mesh_triangles = read_mesh_file(mesh_path)
    else:
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
