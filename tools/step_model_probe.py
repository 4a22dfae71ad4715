import sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, FIXED_POS
c = N.Context(0); v, f = O.load_mesh("lp"); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(); pop = 1024
xs = init_like_population(pop, 2, 1005)
h = c.integrate(p, torch.from_numpy(xs).cuda(), FIXED_POS, 2, 2048).to_host()
R, DT, V, T = [], [], [], []
for i in range(pop):
    ns = h["n_seg"][i]; seg = h["seg_start"][i]
    for s in range(ns):
        a, b = seg[s], seg[s + 1]
        t = h["knots_t"][i, a:b]; y = h["knots_y"][i, a:b]
        if b - a > 8:   # skip the ramp-up
            R.append(np.linalg.norm(y[6:-1, :3], axis=1)); DT.append(np.diff(t)[6:]); V.append(np.linalg.norm(y[6:-1, 3:], axis=1)); T.append(t[6:-1])
np.savez_compressed("gpurun_out/step_model.npz", r=np.concatenate(R), dt=np.concatenate(DT), v=np.concatenate(V), t=np.concatenate(T),
                    n_rhs=h["counters"]["n_rhs"], xs=xs)
print("saved", sum(len(a) for a in R))
