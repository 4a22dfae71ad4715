"""developer probe: (1) the length distribution of the bench population (n_rhs, proxy cost -> gpurun_out/dist_<mesh>.npz),
(2) throughput of ONE SM as a function of the trajectories it holds and of the team size, on trajectories of (nearly)
equal length -- the calibration of tools/schedule_sim.py.  usage: calib_probe.py [mesh]"""
import os, sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
import toss_b200 as T
from bench import init_like_population, FIXED_POS, MESH_PATHS

mesh = sys.argv[1] if len(sys.argv) > 1 else "lp"
pop = 32768
c = N.Context(0)
_, v, f, _ = T.create_mesh(MESH_PATHS[mesh])
c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
a = T.setup_parameters()
c.upload_grid(a.problem.tensor_grid_r, a.problem.tensor_grid_theta, a.problem.tensor_grid_phi, np.asarray(a.problem.bool_tensor))
p = N.make_params(stop_on_collision=1)
xs_h = init_like_population(pop, 2, 1005)
xs = torch.from_numpy(xs_h).cuda()


def t(fn, reps=2):
    fn(); torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); r = fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best, r


full_ms, res = t(lambda: c.population_fitness(p, xs, FIXED_POS, 2, 1024))
cnt = res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)
n_rhs = cnt["n_rhs"].astype(np.int64)
coll = res["collision"].cpu().numpy()
cost = c.predict_cost(p, xs, FIXED_POS, 2).cpu().numpy()
Path("gpurun_out").mkdir(exist_ok=True)
np.savez_compressed(f"gpurun_out/dist_{mesh}.npz", n_rhs=n_rhs, cost=cost, collision=coll,
                    n_acc=cnt["n_accepted"], n_rej=cnt["n_rejected"], n_ev=cnt["n_event_points"])
out = {"mesh": mesh, "full_ms": full_ms, "n_rhs_sum": int(n_rhs.sum()), "n_rhs_max": int(n_rhs.max()),
       "n_rhs_pct": [int(x) for x in np.percentile(n_rhs, [0, 1, 10, 25, 50, 75, 90, 95, 99, 99.9, 100])]}
print(out, flush=True)

# trajectories of nearly equal length around the median, no collision
med = np.median(n_rhs[coll == 0])
idx = np.argsort(np.abs(n_rhs - med) + 1e9 * (coll != 0), kind="stable")[:148 * 16]
rng = np.random.default_rng(0)
rng.shuffle(idx)
out["calib_len_minmax"] = [int(n_rhs[idx].min()), int(n_rhs[idx].max())]
calib = []
plans = {"w16": ((1, (1, 2, 3, 4, 6, 8, 10, 12, 14, 16)), (2, (1, 2, 3, 4, 6, 8)), (4, (1, 2, 3, 4))),
         "w8": ((1, (1, 2, 3, 4, 5, 6, 7, 8)), (2, (1, 2, 4)), (4, (1, 2)))}
for team, ks in plans[sys.argv[2] if len(sys.argv) > 2 else "w16"]:
    os.environ["TOSS_B200_TEAM"] = str(team)
    for k in ks:
        sel = idx[:148 * k]
        x = xs[torch.from_numpy(sel).cuda()].contiguous()
        ms, r = t(lambda: c.integrate(p, x, FIXED_POS, 2, 1024))
        ev = int(n_rhs[sel].sum())
        row = {"team": team, "traj_per_sm": k, "ms": ms, "evals": ev, "evals_per_s": ev / ms * 1e3,
               "us_per_eval_per_traj": ms * 1e3 / float(n_rhs[sel].max())}
        calib.append(row)
        print(row, flush=True)
del os.environ["TOSS_B200_TEAM"]
out["calib"] = calib
Path(f"gpurun_out/calib_{mesh}_{sys.argv[2] if len(sys.argv) > 2 else 'w16'}.json").write_text(json.dumps(out, indent=1))
