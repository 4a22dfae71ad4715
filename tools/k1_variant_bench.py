"""developer experiment: time the thread-per-point gravity kernel (1 M points, lp) for every variant in build/variants"""
import os, subprocess, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, json; sys.path.insert(0, %r)
import numpy as np, torch
from toss_b200 import _native as N
import toss_b200 as T
from bench import gravity_points, MESH_PATHS
c = N.Context(0); _, v, f, _ = T.create_mesh(MESH_PATHS[sys.argv[1]]); c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
pts = torch.from_numpy(gravity_points(1_000_000, 0)).cuda()
for _ in range(3): c.gravity(pts)
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c.gravity(pts); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps(dict(ms=round(best, 3))))
''' % str(ROOT)
mesh = sys.argv[1] if len(sys.argv) > 1 else "lp"
for n in sys.argv[2].split(","):
    env = dict(os.environ, TOSS_B200_LIB=str(ROOT / "build" / "variants" / f"libtoss_{n}.so"))
    r = subprocess.run([sys.executable, "-c", CHILD, mesh], env=env, capture_output=True, text=True, timeout=300)
    print(f"{n:10s}", r.stdout.strip() or r.stderr.strip()[-300:], flush=True)
