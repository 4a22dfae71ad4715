"""Generates tests/golden/gravity_truth_<mesh>.npz: the polyhedral gravity of the 67P meshes at ~44 field points each,
evaluated with 50-digit mpmath from the LITERAL Tsoulis (2012) line-integral formula (LN / AN /
sing-A terms, no closed forms, no face pairs, no polynomial tiers -- nothing shared with oracle/ or toss_b200/csrc).
The points are chosen so that every path of the production arithmetic is exercised: a shell of far points, points
1-3 km above the surface (middle tier), points 3..600 m above a face (per-term rows, ln-based edge terms, full-range
atan2), points inside the body.  Independent truth for `tests/test_oracle_gravity.py` (oracle, CPU) and
`tests/test_gpu_gravity_truth.py` (K1a / K1b directly, GPU).

    python tools/gen_gravity_truth.py lp full        # ~1 min / ~10 min on 8 cores
"""
import sys
from multiprocessing import Pool
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
G_CONST = "6.67430e-11"
DENSITY = 533


def load_mesh(name):
    import toss_b200 as T
    from bench import MESH_PATHS
    _, v, f, _ = T.create_mesh(MESH_PATHS[name])      # vertices after the reference's scaling (mesh_utility.py:37-67)
    return np.asarray(v, dtype=np.float64), np.asarray(f, dtype=np.int64)


def field_points(v, f, seed):
    rng = np.random.default_rng(seed)
    pts, kind = [], []

    def unit(n):
        d = rng.normal(size=(n, 3))
        return d / np.linalg.norm(d, axis=1)[:, None]
    for r in np.linspace(4000.0, 12500.0, 14):                      # shell (configs[1] radii)
        pts.append(unit(1)[0] * r), kind.append("shell")
    a, b, c = v[f[:, 0]], v[f[:, 1]], v[f[:, 2]]
    cen = (a + b + c) / 3
    nrm = np.cross(b - a, c - a)
    nrm /= np.linalg.norm(nrm, axis=1)[:, None]
    for h in (3.0, 7.0, 15.0, 30.0, 60.0, 120.0, 200.0, 300.0, 450.0, 600.0):    # hugging the surface
        i = int(rng.integers(len(f)))
        pts.append(cen[i] + h * nrm[i] + rng.normal(size=3)), kind.append("near")
    for h in (800.0, 1000.0, 1300.0, 1600.0, 2000.0, 2500.0, 3000.0, 3500.0):     # middle distances
        i = int(rng.integers(len(f)))
        pts.append(cen[i] + h * nrm[i]), kind.append("mid")
    for s in (0.0, 0.1, 0.2, 0.3, 0.4, 0.5):                                        # inside the body
        i = int(rng.integers(len(f)))
        pts.append(cen[i] * s + rng.normal(size=3) * 20.0), kind.append("inside")
    for h in (40.0, 90.0, 150.0, 250.0, 400.0, 700.0):                              # above a VERTEX (many faces close)
        i = int(rng.integers(len(v)))
        d = v[i] / np.linalg.norm(v[i])
        pts.append(v[i] + h * d), kind.append("vertex")
    return np.array(pts), np.array(kind)


_MESH = {}


def _init(name):
    import mpmath as mp
    mp.mp.dps = 50
    v, f = load_mesh(name)
    V = [[mp.mpf(float(x)) for x in p] for p in v]
    sub = lambda a, b: [a[i] - b[i] for i in range(3)]
    dot = lambda a, b: a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
    cross = lambda a, b: [a[1] * b[2] - a[2] * b[1], a[2] * b[0] - a[0] * b[2], a[0] * b[1] - a[1] * b[0]]
    faces = []
    for tri in f:
        w = [V[int(i)] for i in tri]
        Gs = [sub(w[(q + 1) % 3], w[q]) for q in range(3)]
        N = cross(Gs[0], Gs[1])
        nn = mp.sqrt(dot(N, N))
        N = [x / nn for x in N]
        g, n = [], []
        for q in range(3):
            gl = mp.sqrt(dot(Gs[q], Gs[q]))
            g.append([x / gl for x in Gs[q]])
            nq = cross(Gs[q], N)
            nl = mp.sqrt(dot(nq, nq))
            n.append([x / nl for x in nq])
        faces.append((w, N, g, n))
    _MESH["faces"] = faces


def _truth(P):
    """Tsoulis (2012) eqs. in 50 digits: g = G rho sum_p N_p [ sum_q sigma_pq h_pq LN_pq + h_p (sum_q sigma_pq AN_pq + singA_p) ]"""
    import mpmath as mp
    P = [mp.mpf(float(x)) for x in P]
    dot = lambda a, b: a[0] * b[0] + a[1] * b[1] + a[2] * b[2]
    acc = [mp.mpf(0)] * 3
    for w, N, g, n in _MESH["faces"]:
        r = [[w[i][k] - P[k] for k in range(3)] for i in range(3)]
        l = [mp.sqrt(dot(x, x)) for x in r]
        hp_s = dot(N, r[0])
        hp = abs(hp_s)
        S, npos = mp.mpf(0), 0
        for q in range(3):
            hq_s = dot(n[q], r[q])
            hq, sg = abs(hq_s), mp.sign(hq_s)
            s1, s2 = dot(g[q], r[q]), dot(g[q], r[(q + 1) % 3])
            l1, l2 = l[q], l[(q + 1) % 3]
            LN = mp.log((s2 + l2) / (s1 + l1))
            AN = mp.atan(hp * s2 / (hq * l2)) - mp.atan(hp * s1 / (hq * l1)) if hp != 0 and hq != 0 else 0
            S += hq_s * LN + hp * sg * AN
            npos += sg > 0
        if npos == 3:
            S -= 2 * mp.pi * hp
        acc = [acc[i] + N[i] * S for i in range(3)]
    Grho = mp.mpf(G_CONST) * DENSITY
    return [mp.nstr(Grho * a, 40) for a in acc]


def main():
    for name in sys.argv[1:] or ["lp", "full"]:
        v, f = load_mesh(name)
        pts, kind = field_points(v, f, 20261020 + len(f))
        with Pool(8, initializer=_init, initargs=(name,)) as pool:
            rows = pool.map(_truth, list(pts), chunksize=1)
        acc = np.array([[float(x) for x in r] for r in rows])
        out = ROOT / "tests" / "golden" / f"gravity_truth_{name}.npz"
        np.savez_compressed(out, points=pts, kind=kind, acc=acc, acc_40_digits=np.array(rows),
                            note="package sign convention (TOSS uses the negative); G = 6.67430e-11, rho = 533; "
                                 "mpmath 50 digits, literal Tsoulis formula; tools/gen_gravity_truth.py")
        print(name, len(pts), "points ->", out, flush=True)


if __name__ == "__main__":
    main()
