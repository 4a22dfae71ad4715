"""developer A/B: warps per trajectory (1 / 2 / 4) at populations of a few trajectories per warp"""
import os, subprocess, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, json; sys.path.insert(0, %r)
import numpy as np, torch
from toss_b200 import _native as N
import toss_b200 as T
from bench import init_like_population, converged_like_population, FIXED_POS, MESH_PATHS
mesh, pop, kind = sys.argv[1], int(sys.argv[2]), sys.argv[3]
c = N.Context(0); _, v, f, _ = T.create_mesh(MESH_PATHS[mesh]); c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
p = N.make_params(stop_on_collision=1)
xs = torch.from_numpy((init_like_population if kind == "init" else converged_like_population)(pop, 2, 1005)).cuda()
c.integrate(p, xs, FIXED_POS, 2, 1024); torch.cuda.synchronize()
best = 1e9
for _ in range(2):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c.integrate(p, xs, FIXED_POS, 2, 1024); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(json.dumps(dict(ms=round(best, 2))))
''' % str(ROOT)
cases = [tuple(a.split(":")) for a in sys.argv[1:]] or [("lp", "4096"), ("full", "4096"), ("lp", "8192"), ("full", "1024")]
for mesh, pop in cases:
    for kind in ("init", "conv"):
        out = []
        for team in ("1", "2", "4"):
            env = dict(os.environ, TOSS_B200_TEAM=team)
            r = subprocess.run([sys.executable, "-c", CHILD, mesh, pop, kind], env=env, capture_output=True, text=True, timeout=900)
            out.append(f"team{team} " + (r.stdout.strip() or r.stderr.strip()[-200:]))
        print(f"{mesh} {pop} {kind}: " + "  ".join(out), flush=True)
