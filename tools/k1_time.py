"""developer probe: time of the thread-per-point gravity kernel (K1a) for 1 M points on a mesh, and the tier mix it sees.
usage: k1_time.py [mesh ...]"""
import sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch
from toss_b200 import _native as N
import toss_b200 as T, numpy as np
from bench import gravity_points, MESH_PATHS
for mesh in (sys.argv[1:] or ["lp", "llp", "full"]):
    c = N.Context(0); _, v, f, _ = T.create_mesh(MESH_PATHS[mesh]); c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
    pts = torch.from_numpy(gravity_points(1_000_000, 0)).cuda()
    for _ in range(3):
        c.gravity(pts)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        c.gravity(pts)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    print(mesh, "faces", len(f), "ms per 1M points %.3f" % ms, "evals/s %.3e" % (1e9 / ms), "alg TFLOP/s %.2f" % (1e6 * len(f) * 127 / ms / 1e9), flush=True)
    c.close()
