import os, sys; from pathlib import Path
os.environ["TOSS_B200_LPT"] = "1"; os.environ["TOSS_B200_TEAM"] = "1"
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, FIXED_POS
c = N.Context(0); v, f = O.load_mesh("lp"); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(stop_on_collision=int(sys.argv[1])); pop = int(sys.argv[2])
xs = torch.from_numpy(init_like_population(pop, 2, 1005)).cuda()
pr = c.integrate(p, xs, FIXED_POS, 2, 2048); torch.cuda.synchronize()
h = pr.to_host(); n_rhs = h["counters"]["n_rhs"].astype(float)
q = pr.ws[pr.L.queue:pr.L.queue + 4 * (4 + 2 * pop)].view(torch.int32).cpu().numpy()
cost = q[4 + pop:4 + 2 * pop].astype(float)
print("pred: mean %.0f min %.0f max %.0f ; actual mean %.0f min %.0f max %.0f" % (cost.mean(), cost.min(), cost.max(), n_rhs.mean(), n_rhs.min(), n_rhs.max()))
print("corr all", np.corrcoef(cost, n_rhs)[0, 1], "corr non-collided", np.corrcoef(cost[h["collision"] == 0], n_rhs[h["collision"] == 0])[0, 1])
print("spearman", np.corrcoef(np.argsort(np.argsort(cost)), np.argsort(np.argsort(n_rhs)))[0, 1])
for i in range(12): print(int(cost[i]), int(n_rhs[i]), int(h["collision"][i]))
