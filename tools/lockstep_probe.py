"""developer probe: how much of the stepper's time is lost to the sixteen warps of a CTA moving in lock step?  A population of
32768 COPIES of one chromosome has no imbalance inside a tile ring (every warp does the same work); its gravity
evaluations/s against the mixed population's is an upper bound of what a decoupled ring could recover.
usage: lockstep_probe.py [n_chromosomes]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
import toss_b200 as T
from bench import init_like_population, FIXED_POS, MESH_PATHS

c = N.Context(0)
_, v, f, _ = T.create_mesh(MESH_PATHS["lp"])
c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
a = T.setup_parameters()
c.upload_grid(a.problem.tensor_grid_r, a.problem.tensor_grid_theta, a.problem.tensor_grid_phi, np.asarray(a.problem.bool_tensor))
p = N.make_params(stop_on_collision=1)
pop = 32768
xs_h = init_like_population(pop, 2, 1005)


def run(x):
    xs = torch.from_numpy(np.ascontiguousarray(x)).cuda()
    c.integrate(p, xs, FIXED_POS, 2, 1024); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); r = c.integrate(p, xs, FIXED_POS, 2, 1024); e1.record(); torch.cuda.synchronize()
    n = int(r.to_host()["counters"]["n_rhs"].sum())
    return e0.elapsed_time(e1), n


ms, n = run(xs_h)
print("mixed population: %.1f ms, %.2f M evaluations/s" % (ms, n / ms / 1e3), flush=True)
h = c.integrate(p, torch.from_numpy(xs_h[:4096]).cuda(), FIXED_POS, 2, 1024).to_host()
nr = h["counters"]["n_rhs"]
k = int(sys.argv[1]) if len(sys.argv) > 1 else 6
# chromosomes spread over the length distribution (survivors only)
cand = np.argsort(nr)
cand = cand[nr[cand] > 1500]
pick = cand[np.linspace(0, len(cand) - 1, k).astype(int)]
rates = []
for i in pick:
    ms, n = run(np.repeat(xs_h[i:i + 1], pop, axis=0))
    rates.append(n / ms / 1e3)
    print("32768 copies of chromosome %d (%d evaluations each): %.1f ms, %.2f M evaluations/s" % (i, nr[i], ms, rates[-1]), flush=True)
print("mean over the copies: %.2f M evaluations/s" % np.mean(rates))
