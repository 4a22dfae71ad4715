"""Developer timing probe (not the contract bench): K1 on 1 M points and the population path."""
import sys, time, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import torch
from toss_b200 import _native as N
from oracle import oracle as O
from tests.test_gpu_parity import random_population, bound_population, FIXED_POS, points

def ev_time(fn, reps=3):
    best = 1e99
    for _ in range(reps):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best

out = {}
c = N.Context(0)
out["fp64_peak_tflops"] = c.measure_fp64_peak()
print("fp64 peak", out["fp64_peak_tflops"], flush=True)
for mesh in sys.argv[1].split(",") if len(sys.argv) > 1 else ["llp", "lp"]:
    v, f = O.load_mesh(mesh)
    c.upload_mesh(v, f, 533.0)
    c.upload_grid(*O.default_grid())
    F = len(f)
    pts = torch.from_numpy(points(1_000_000, 0)).cuda()
    c.gravity(pts)
    t = ev_time(lambda: c.gravity(pts))
    out[f"k1_{mesh}"] = dict(s=t, evals_per_s=1e6 / t, face_pairs_per_s=1e6 * F / t,
                              alg_tflops=1e6 * F * 127 / t / 1e12)
    print(mesh, "K1", out[f"k1_{mesh}"], flush=True)
    for kind, gen in (("init", random_population), ("bound", bound_population)):
        pop = int(sys.argv[2]) if len(sys.argv) > 2 else (4096 if F < 2000 else 512)
        p = N.make_params()
        xs = torch.from_numpy(gen(pop, 2, 1005)).cuda()
        c.population_fitness(p, xs, FIXED_POS, 2)
        torch.cuda.synchronize()
        t = ev_time(lambda: c.population_fitness(p, xs, FIXED_POS, 2), reps=2)
        res = c.population_fitness(p, xs, FIXED_POS, 2)
        cnt = res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)
        nrhs = int(cnt["n_rhs"].sum())
        out[f"pop_{mesh}_{kind}"] = dict(pop=pop, s=t, traj_per_s=pop / t, grav_evals_per_s=nrhs / t,
                                         alg_tflops=nrhs * F * 127 / t / 1e12, mean_rhs=nrhs / pop,
                                         max_rhs=int(cnt["n_rhs"].max()), status_bad=int((cnt["status"] != 0).sum()),
                                         collided=int(res["collision"].sum().item()),
                                         max_knots=int((cnt["n_accepted"]).max()))
        print(mesh, kind, out[f"pop_{mesh}_{kind}"], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path("gpurun_out/quick_bench.json").write_text(json.dumps(out, indent=1))
