"""developer experiment: time the population path for every stepper variant under build/variants (one subprocess each)"""
import os, subprocess, sys, json
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
CHILD = r'''
import sys, json; sys.path.insert(0, %r)
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, converged_like_population, FIXED_POS
mesh, pop, kind = sys.argv[1], int(sys.argv[2]), sys.argv[3]
c = N.Context(0); v, f = O.load_mesh(mesh); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(stop_on_collision=1); xs = torch.from_numpy((init_like_population if kind == 'init' else converged_like_population)(pop, 2, 1005)).cuda()
c.integrate(p, xs, FIXED_POS, 2, 2048); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c.integrate(p, xs, FIXED_POS, 2, 2048); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
h = c.integrate(p, xs, FIXED_POS, 2, 2048).to_host()
print(json.dumps(dict(ms=best, n_rhs=int(h["counters"]["n_rhs"].sum()), bad=int((h["counters"]["status"]!=0).sum()))))
''' % str(ROOT)
mesh = sys.argv[1] if len(sys.argv) > 1 else "lp"
pop = sys.argv[2] if len(sys.argv) > 2 else "4096"
names = sys.argv[3].split(",") if len(sys.argv) > 3 else sorted(p.stem[8:] for p in (ROOT / "build" / "variants").glob("libtoss_*.so"))
kinds = sys.argv[4].split(",") if len(sys.argv) > 4 else ["init"]
for kind in kinds:
    for n in names:
        env = dict(os.environ, TOSS_B200_LIB=str(ROOT / "build" / "variants" / f"libtoss_{n}.so"))
        env.setdefault("TOSS_B200_LPT", "0")     # kernel A/B: no queue reordering unless asked for
        r = subprocess.run([sys.executable, "-c", CHILD, mesh, pop, kind], env=env, capture_output=True, text=True, timeout=600)
        print(f"{kind:5s} {n:12s}", r.stdout.strip() or r.stderr.strip()[-400:], flush=True)
