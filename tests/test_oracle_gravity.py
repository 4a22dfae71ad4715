"""Gravity alone has no golden in the reference.  The oracle's gravity is pinned
instead by (a) a 50-digit mpmath evaluation of the literal Tsoulis formula, (b) the literal
double-precision transcription (orc_gravity_direct), (c) the far-field point-mass limit, and (d)
through the golden trajectories in test_oracle_golden.py.  CPU only."""
import numpy as np
import pytest

from oracle import oracle as O

G_CONST = 6.67430e-11


def _points(n, seed, rlo=4000.0, rhi=12500.0):
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1)[:, None]
    return d * rng.uniform(rlo, rhi, size=(n, 1))


def _mp_gravity(verts, faces, P, Grho):
    """Tsoulis (2012) eqs. in 50 digits, literal LN / AN / sing-A form."""
    import mpmath as mp
    mp.mp.dps = 50
    V = [[mp.mpf(float(x)) for x in v] for v in verts]
    P = [mp.mpf(float(x)) for x in P]
    sub = lambda a, b: [a[i] - b[i] for i in range(3)]
    dot = lambda a, b: a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
    cross = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
    norm = lambda a: mp.sqrt(dot(a, a))
    acc = [mp.mpf(0)] * 3
    for f in faces:
        v = [V[int(i)] for i in f]
        Gs = [sub(v[(q + 1) % 3], v[q]) for q in range(3)]
        N = cross(Gs[0], Gs[1])
        nn = norm(N)
        N = [x / nn for x in N]
        r = [sub(v[i], P) for i in range(3)]
        l = [norm(x) for x in r]
        hp_s = dot(N, r[0])
        hp = abs(hp_s)
        S, npos = mp.mpf(0), 0
        for q in range(3):
            gl = norm(Gs[q])
            g = [x / gl for x in Gs[q]]
            n = cross(Gs[q], N)
            nl = norm(n)
            n = [x / nl for x in n]
            hq_s = dot(n, r[q])
            hq, sg = abs(hq_s), mp.sign(hq_s)
            s1, s2 = dot(g, r[q]), dot(g, r[(q + 1) % 3])
            l1, l2 = l[q], l[(q + 1) % 3]
            LN = mp.log((s2 + l2) / (s1 + l1))
            AN = mp.atan(hp * s2 / (hq * l2)) - mp.atan(hp * s1 / (hq * l1)) if hp != 0 and hq != 0 else 0
            S += hq_s * LN + hp * sg * AN
            npos += sg > 0
        if npos == 3:
            S -= 2 * mp.pi * hp
        acc = [acc[i] + N[i] * S for i in range(3)]
    return np.array([float(Grho * a) for a in acc])


@pytest.mark.parametrize("name,npts,tol", [("lllp", 6, 2e-14), ("llp", 4, 2e-14)])
def test_gravity_vs_mpmath(name, npts, tol):
    import mpmath as mp
    v, f = O.load_mesh(name)
    m = O.Mesh(v, f)
    pts = _points(npts, 11)
    pts[0] = pts[0] / np.linalg.norm(pts[0]) * 900.0       # a point INSIDE the body
    a = O.gravity(m, pts)
    d = O.gravity_direct(m, pts)
    for i in range(npts):
        truth = _mp_gravity(v, f, pts[i], mp.mpf(G_CONST) * 533)
        nt = np.linalg.norm(truth)
        assert np.linalg.norm(a[i] - truth) / nt < tol
        assert np.linalg.norm(d[i] - truth) / nt < 5e-12


def test_gravity_near_the_surface_vs_mpmath():
    """Points 3 m .. 600 m above face centroids: some edge / solid-angle terms of the nearby pairs are outside the
    range of the short polynomials and fall back, each on its own, to the ln-based atanh / full-range atan2
    (pair_term); far terms of the same pair stay on the shared division.  Every mixture has to stay at rounding
    level against the 50-digit evaluation."""
    import mpmath as mp
    v, f = O.load_mesh("llp")
    m = O.Mesh(v, f)
    rng = np.random.default_rng(3)
    pts = []
    for h in (3.0, 30.0, 120.0, 300.0, 600.0):
        i = int(rng.integers(len(f)))
        a, b, c = v[f[i, 0]], v[f[i, 1]], v[f[i, 2]]
        n = np.cross(b - a, c - a)
        pts.append((a + b + c) / 3 + h * n / np.linalg.norm(n) + rng.normal(size=3))
    pts = np.array(pts)
    acc = O.gravity(m, pts)
    for i in range(len(pts)):
        truth = _mp_gravity(v, f, pts[i], mp.mpf(G_CONST) * 533)
        assert np.linalg.norm(acc[i] - truth) / np.linalg.norm(truth) < 5e-14


@pytest.mark.parametrize("name,n,noise", [("lllp", 4000, 1e-11), ("llp", 4000, 5e-11), ("lp", 2000, 1e-10),
                                          ("full", 300, 5e-10)])
def test_gravity_closed_forms_equal_literal_tsoulis(name, n, noise):
    """Same numbers from the cancellation-free closed forms and the literal transcription; the
    difference is the literal form's rounding noise (grows with mesh resolution), which is why
    1e-12 parity is anchored on the closed forms."""
    v, f = O.load_mesh(name)
    m = O.Mesh(v, f)
    pts = _points(n, 5)
    a = O.gravity(m, pts, threads=8)
    d = O.gravity_direct(m, pts)
    rel = np.linalg.norm(a - d, axis=1) / np.linalg.norm(a, axis=1)
    assert rel.max() < noise


@pytest.mark.parametrize("name", ["lllp", "llp", "lp", "full"])
def test_face_pairing(name):
    """Every face in exactly one record, partners really share the stated edge with opposite directions, almost
    no face is left single; the library's own (host-only) matching gives the same list."""
    from toss_b200 import _native as N
    v, f = O.load_mesh(name)
    m = O.Mesh(v, f, 533.0)
    pr = m.pairs()
    single = pr[:, 2] < 0
    assert np.array_equal(np.sort(np.concatenate([pr[:, 0], pr[~single, 2]])), np.arange(len(f)))
    assert np.all(np.diff(pr[:, 0]) > 0) and np.all(pr[~single, 0] < pr[~single, 2])
    for fa, qa, fb in pr[~single]:
        a, b = int(f[fa][qa]), int(f[fa][(qa + 1) % 3])
        assert any(int(f[fb][q]) == b and int(f[fb][(q + 1) % 3]) == a for q in range(3))
    assert single.sum() <= 0.02 * len(f) + 2
    assert np.array_equal(N.pair_faces(f, len(v)), pr)


def test_pairing_of_an_open_or_inconsistent_surface_degrades_to_singles():
    """edges with one face, or with two faces running the same way, simply do not pair"""
    from toss_b200 import _native as N
    f = np.array([[0, 1, 2], [0, 1, 3], [4, 5, 6]], np.int32)       # 0-1 is traversed the same way twice
    pr = N.pair_faces(f, 7)
    assert np.array_equal(pr, [[0, 0, -1], [1, 0, -1], [2, 0, -1]])
    f = np.array([[0, 1, 2], [1, 0, 3]], np.int32)
    assert np.array_equal(N.pair_faces(f, 4), [[0, 0, 1]])


def test_gravity_far_field_point_mass():
    v, f = O.load_mesh("llp")
    m = O.Mesh(v, f)
    # volume by the divergence theorem
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    vol = np.sum(np.einsum("ij,ij->i", a, np.cross(b, c))) / 6.0
    com = np.sum((a + b + c) / 4.0 * np.einsum("ij,ij->i", a, np.cross(b, c))[:, None] / 6.0, axis=0) / vol
    GM = G_CONST * 533 * vol
    P = np.array([[3.0e6, -2.0e6, 1.5e6]])
    acc_pkg, pot = O.gravity(m, P, potential=True)
    r = P[0] - com
    expect = -GM * r / np.linalg.norm(r) ** 3            # attraction
    assert np.allclose(-acc_pkg[0], expect, rtol=1e-5)   # TOSS uses the negative of the package value
    assert np.isclose(pot[0], GM / np.linalg.norm(r), rtol=1e-5)


def test_compute_motion_quirks():
    """equations_of_motion.py:84-95: gravity at the rotated position, NOT rotated back; sign flip :59."""
    v, f = O.load_mesh("llp")
    m = O.Mesh(v, f)
    p = O.default_params()
    y = np.array([-135.13402075, -4089.53592604, 6050.17636635, 0.1, -0.2, 0.3])
    t = 12345.0
    out = O.compute_motion(m, p, t, y)
    rot = O.rotate_point(t, y[:3], np.array(p.spin_axis[:]), p.spin_velocity)
    assert np.array_equal(out[:3], y[3:])
    assert np.allclose(out[3:], -O.gravity(m, rot)[0], rtol=0, atol=0)


@pytest.mark.parametrize("name", ["lp", "full"])
def test_gravity_vs_committed_mpmath_truth(name):
    """The meshes the benchmark runs on, at the north-star tolerance: 44 field points each (shell, 1-3.5 km above the
    surface, 3..600 m above faces and vertices, inside the body) against tests/golden/gravity_truth_<mesh>.npz --
    50-digit mpmath, literal Tsoulis formula, written by tools/gen_gravity_truth.py with nothing shared with the
    oracle.  Every tier and every fall-back of the production arithmetic must have been taken."""
    from pathlib import Path
    d = np.load(Path(__file__).parent / "golden" / f"gravity_truth_{name}.npz")
    v, f = O.load_mesh(name)
    m = O.Mesh(v, f)
    acc = O.gravity(m, d["points"])
    rel = np.linalg.norm(acc - d["acc"], axis=1) / np.linalg.norm(d["acc"], axis=1)
    assert rel.max() < 5e-14, rel.max()
    census = O.gravity_tiers(m, d["points"])
    assert min(census.values()) > 0, census           # far, middle, per-term rows; ln-based atanh; full-range atan2
    # the literal double-precision transcription cannot serve as the 1e-12 anchor on these meshes
    lit = O.gravity_direct(m, d["points"])
    rel_lit = np.linalg.norm(lit - d["acc"], axis=1) / np.linalg.norm(d["acc"], axis=1)
    assert rel_lit.max() < 2e-11 and rel_lit.max() > 20 * rel.max()
