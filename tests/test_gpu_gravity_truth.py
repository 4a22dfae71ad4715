"""Gravity kernels against INDEPENDENT truth (run with -m gpu): the thread-per-point kernel (K1a), the warp-per-point
kernel (K1b) and compute_motion, called through the C ABI, compared directly with the committed 50-digit mpmath
evaluation of the literal Tsoulis formula (tests/golden/gravity_truth_<mesh>.npz, tools/gen_gravity_truth.py) on the
meshes the benchmark runs on -- no oracle in between.  Bar (north_star): 1e-12 relative; observed ~1e-14."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLDEN = Path(__file__).parent / "golden"


def _ctx(name):
    from toss_b200 import _native as N
    import toss_b200 as T
    from bench import MESH_PATHS
    c = N.Context(0)
    _, v, f, _ = T.create_mesh(MESH_PATHS[name])
    c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
    return c, N


def _rel(a, b):
    return np.linalg.norm(a - b, axis=1) / np.linalg.norm(b, axis=1)


@pytest.mark.parametrize("name", ["lp", "full"])
def test_gravity_kernels_vs_mpmath_truth(name):
    d = np.load(GOLDEN / f"gravity_truth_{name}.npz")
    pts, truth = d["points"], d["acc"]
    c, N = _ctx(name)
    # K1b: fewer than 4096 points -> warp per point (the stepper's summation order)
    warp = c.gravity(pts).cpu().numpy()
    assert _rel(warp, truth).max() < 1e-12
    assert _rel(warp, truth).max() < 5e-14            # what the arithmetic actually delivers
    # K1a: >= 4096 points -> thread per point, TMA ring (the truth points tiled; every copy must agree)
    reps = -(-4096 // len(pts)) + 3
    tiled = np.tile(pts, (reps, 1))
    thread = c.gravity(tiled).cpu().numpy()
    assert _rel(thread, np.tile(truth, (reps, 1))).max() < 1e-12
    assert _rel(thread, np.tile(truth, (reps, 1))).max() < 1e-13
    # compute_motion without rotation: [v, -g]
    p = N.make_params(activate_rotation=0)
    y = np.concatenate([pts, np.zeros_like(pts)], axis=1)
    out = c.compute_motion(p, np.zeros(len(pts)), y).cpu().numpy()
    assert _rel(-out[:, 3:], truth).max() < 5e-14
