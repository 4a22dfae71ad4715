import sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, FIXED_POS
mesh = sys.argv[1] if len(sys.argv) > 1 else "lp"
c = N.Context(0); v, f = O.load_mesh(mesh); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(stop_on_collision=int(sys.argv[3]) if len(sys.argv) > 3 else 0)
for pop in [int(a) for a in sys.argv[2].split(",")] if len(sys.argv) > 2 else (4096, 32768):
    xs = torch.from_numpy(init_like_population(pop, 2, 1005)).cuda()
    c.integrate(p, xs, FIXED_POS, 2, 2048); torch.cuda.synchronize()
    best = 1e9
    for _ in range(2):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); pr = c.integrate(p, xs, FIXED_POS, 2, 2048); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    h = pr.to_host()
    n_rhs = h["counters"]["n_rhs"].astype(float)
    q = pr.ws[pr.L.queue:pr.L.queue + 4 * (4 + 2 * pop)].view(torch.int32).cpu().numpy()
    order, cost = q[4:4 + pop], q[4 + pop:4 + 2 * pop].astype(float)
    ideal = n_rhs.sum() / (2 * 148 * 8)
    print(f"{mesh} pop {pop}: {best:.1f} ms  traj/s {pop / best * 1e3:.0f}  evals/s {n_rhs.sum() / best * 1e3:.3e}  "
          f"corr(pred,actual) {np.corrcoef(cost, n_rhs)[0, 1]:.3f}  max/ideal {n_rhs.max() / ideal:.2f} "
          f"order ok {sorted(order.tolist()) == list(range(pop))} bad {(h['counters']['status'] != 0).sum()}", flush=True)
