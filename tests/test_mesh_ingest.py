"""Mesh ingest beyond the reference's four pickles (README "Use Case 2"): arrays, Wavefront .obj and
.npz in, validation of what the device path requires (closed, consistently oriented, outward, no zero-area face).
CPU part: the host checks.  GPU part (-m gpu): the upload refuses the same meshes loudly, the per-pair constants the
device builds equal a numpy re-computation bit for bit, a body that is not 67P evaluates like the oracle."""
import numpy as np
import pytest

from toss_b200.mesh import check_mesh, create_mesh, mesh_from_arrays, read_mesh_file, read_obj_file

CUBE_OBJ = """# a 2 km cube, quads, outward
v -1000 -1000 -1000
v  1000 -1000 -1000
v  1000  1000 -1000
v -1000  1000 -1000
v -1000 -1000  1000
v  1000 -1000  1000
v  1000  1000  1000
v -1000  1000  1000
f 1 4 3 2
f 5 6 7 8
f 1/1 2/2 6/3 5/4
f 2//1 3//1 7//1 6//1
f 3 4 8 7
f -5 -8 -4 -1
"""


def cube(tmp_path):
    p = tmp_path / "cube.obj"
    p.write_text(CUBE_OBJ)
    return p


def test_obj_reader_fans_polygons_and_resolves_index_forms(tmp_path):
    v, f = read_obj_file(cube(tmp_path))
    assert v.shape == (8, 3) and f.shape == (12, 3) and f.dtype == np.int32
    assert check_mesh(v, f) == []
    # `-5 -8 -4 -1` with 8 vertices read = 4 1 5 8 (1-based) = the -x face
    assert sorted(f[-2:].ravel().tolist()) == sorted([3, 0, 4, 3, 4, 7])
    vol = np.einsum("ij,ij->", v[f[:, 0]], np.cross(v[f[:, 1]], v[f[:, 2]])) / 6
    assert vol == pytest.approx(8e9)
    _, v2, f2, prot = create_mesh(str(cube(tmp_path)))            # .obj / .npz bodies are taken in metres as they are
    assert np.array_equal(v2, v) and np.array_equal(f2, f) and prot == 1000.0
    np.savez(tmp_path / "cube.npz", vertices=v, faces=f)
    v3, f3 = read_mesh_file(tmp_path / "cube.npz")
    assert np.array_equal(v3, v) and np.array_equal(f3, f)


def test_check_mesh_names_every_defect(tmp_path):
    v, f = read_obj_file(cube(tmp_path))
    assert "not closed" in check_mesh(v, f[:-1])[0]                                  # a hole
    g = f.copy()
    g[3] = g[3][::-1]                                                                # one face flipped
    msgs = " | ".join(check_mesh(v, g))
    assert "more than one face" in msgs and "without an opposite partner" in msgs
    assert "clockwise" in check_mesh(v, f[:, ::-1])[0]                               # inside out
    g = f.copy()
    g[0] = [0, 0, 1]
    assert any("zero-area" in m for m in check_mesh(v, g))
    g = f.copy()
    g[0, 0] = 99
    assert check_mesh(v, g) == ["face index out of range"]
    with pytest.raises(ValueError, match="not closed"):
        mesh_from_arrays(v, f[:-1])
    with pytest.raises(ValueError, match="integer"):
        mesh_from_arrays(v, f.astype(np.float64))
    with pytest.raises(ValueError, match=r"\(V, 3\)"):
        mesh_from_arrays(v[:, :2], f)


def test_the_four_reference_meshes_pass():
    from bench import MESH_PATHS
    for name, path in MESH_PATHS.items():
        _, v, f, _ = create_mesh(path)
        assert check_mesh(v, f) == [], name


# ------------------------------------------------------------------------------------------------------ GPU
def _numpy_pair_records(v, f, pairs):
    """the constants of csrc/pairs.cu in numpy: the same IEEE operations in the same order"""
    def consts(w0, w1, w2):
        w = (w0, w1, w2)
        G = [w[(q + 1) % 3] - w[q] for q in range(3)]
        N = np.array([G[0][1] * G[1][2] - G[0][2] * G[1][1], G[0][2] * G[1][0] - G[0][0] * G[1][2],
                      G[0][0] * G[1][1] - G[0][1] * G[1][0]])
        nn = np.sqrt(N[0] * N[0] + N[1] * N[1] + N[2] * N[2])
        Nu = N / nn
        n, ln = [], []
        for q in range(3):
            g = G[q]
            ln.append(np.sqrt(g[0] * g[0] + g[1] * g[1] + g[2] * g[2]))
            nq = np.array([g[1] * Nu[2] - g[2] * Nu[1], g[2] * Nu[0] - g[0] * Nu[2], g[0] * Nu[1] - g[1] * Nu[0]])
            n.append(nq / np.sqrt(nq[0] * nq[0] + nq[1] * nq[1] + nq[2] * nq[2]))
        return Nu, n, ln, nn
    out = np.zeros((len(pairs), 43))
    for i, (fa, qa, fb) in enumerate(pairs):
        i0, i1, i2 = f[fa][qa], f[fa][(qa + 1) % 3], f[fa][(qa + 2) % 3]
        i3 = i2
        if fb >= 0:
            qb = next(q for q in range(3) if f[fb][q] == i1 and f[fb][(q + 1) % 3] == i0)
            i3 = f[fb][(qb + 2) % 3]
        r = out[i]
        r[0:3], r[3:6], r[6:9], r[9:12] = v[i0], v[i1], v[i2], v[i3]
        N, n, ln, twoA = consts(v[i0], v[i1], v[i2])
        r[12:15] = N
        r[18:27] = np.concatenate(n)
        r[36:39] = ln
        r[41] = twoA
        if fb >= 0:
            N, n, ln, twoA = consts(v[i1], v[i0], v[i3])
            r[15:18] = N
            r[27:36] = np.concatenate(n)
            r[39], r[40], r[42] = ln[1], ln[2], twoA
    return out


@pytest.mark.gpu
@pytest.mark.parametrize("mesh_name", ["lllp", "llp", "lp"])
def test_device_built_pair_constants_equal_the_numpy_recomputation_bit_for_bit(mesh_name):
    from tests.test_gpu_parity import ctx_for
    c, m = ctx_for(mesh_name)
    pairs = c.mesh_pairs()
    assert np.array_equal(pairs, m.pairs())                       # the oracle's own matching
    got = c.mesh_pair_records()
    want = _numpy_pair_records(m.vertices, m.faces, pairs)
    assert got.shape == want.shape and np.array_equal(got, want)


@pytest.mark.gpu
def test_another_body_from_obj_evaluates_like_the_oracle_and_defects_are_refused(tmp_path):
    from oracle import oracle as O
    from toss_b200 import _native as N
    v, f = read_obj_file(cube(tmp_path))
    c = N.Context(0)
    c.upload_mesh(v, f, 2000.0)
    m = O.Mesh(v, f, 2000.0)
    rng = np.random.default_rng(2)
    pts = rng.normal(size=(300, 3))
    pts *= (rng.uniform(1100.0, 9000.0, 300) / np.abs(pts).max(axis=1))[:, None]     # outside the cube, some 100 m above a face
    got = c.gravity(pts).cpu().numpy()
    want = O.gravity(m, pts)
    assert np.max(np.linalg.norm(got - want, axis=1) / np.linalg.norm(want, axis=1)) < 1e-12
    far = np.array([[0.0, 3.0e5, 4.0e5]])                          # point-mass limit: |g| -> G rho V / r^2
    g = c.gravity(far).cpu().numpy()[0]
    assert np.linalg.norm(g) == pytest.approx(6.67430e-11 * 2000.0 * 8e9 / 2.5e11, rel=1e-4)
    # (+z rays through an edge of the triangulation are excluded: there the reference's own answer is ambiguous)
    inside = c.is_outside(np.array([[123.0, -456.0, 0.0], [123.0, -456.0, 1500.0], [999.0, 998.0, 997.0]])).cpu().numpy()
    assert inside.astype(bool).tolist() == [False, True, False]
    for bad, word in ((f[:-1], "not a closed"), (f[:, ::-1].copy(), "clockwise")):
        with pytest.raises(RuntimeError, match=word):
            c.upload_mesh(v, np.ascontiguousarray(bad, dtype=np.int32), 2000.0)
    g2 = f.copy()
    g2[3] = g2[3][::-1]
    with pytest.raises(RuntimeError, match="not a closed"):
        c.upload_mesh(v, g2, 2000.0)
    c.upload_mesh(v, f, 2000.0)                                    # the context is still usable afterwards
    assert np.array_equal(c.gravity(pts).cpu().numpy(), got)
