import os, sys; from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ["TOSS_B200_LIB"] = str(ROOT / "build" / "variants" / "libtoss_prof.so")
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
c = N.Context(0); v, f = O.load_mesh("lp"); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(); pop = 16384
rng = np.random.default_rng(0)
for name, rad in (("circular 10 km", 10000.0), ("circular 30 km", 30000.0), ("circular 5 km", 5000.0)):
    d = rng.normal(size=(pop, 3)); d /= np.linalg.norm(d, axis=1)[:, None]
    t = np.cross(d, rng.normal(size=(pop, 3))); t /= np.linalg.norm(t, axis=1)[:, None]
    vc = np.sqrt(1728.0 / rad)
    cols = [d * rad, np.full((pop, 1), vc), t]
    for _ in range(2):
        cols += [rng.uniform(1, 604799, (pop, 1)), np.full((pop, 1), vc), t + 0.02 * rng.normal(size=(pop, 3))]
    xs = torch.from_numpy(np.hstack(cols)).cuda()
    c.integrate(p, xs, np.zeros(0), 2, 2048); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pr = c.integrate(p, xs, np.zeros(0), 2, 2048); e1.record(); torch.cuda.synchronize()
    q = pr.ws[pr.L.queue + 16:pr.L.queue + 64].view(torch.int64).cpu().numpy()
    wait, comp, total, n, wait0, idle = q
    h = pr.to_host(); nr = h["counters"]["n_rhs"]
    print(f"{name}: {e0.elapsed_time(e1):.0f} ms evals/s {nr.sum() / e0.elapsed_time(e1) * 1e3:.3e} mean_rhs {nr.mean():.0f} max {nr.max()} bad {(h['counters']['status'] != 0).sum()} "
          f"wait {wait / total:.3f} compute {comp / total:.3f} other {(total - wait - comp) / total:.3f}", flush=True)
