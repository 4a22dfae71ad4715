"""developer probe: strong-scaling efficiency of ONE GPU's share, measured on one GPU.  The 32768 population is dealt
for world = 2 / 4 / 8 exactly as parallel.py does; shard 0 is evaluated alone (pre-ordered queue) and compared with
1/world of the whole population's time.  usage: shard_probe.py [mesh] [pop]"""
import sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
import toss_b200 as T
from bench import init_like_population, FIXED_POS, MESH_PATHS

mesh = sys.argv[1] if len(sys.argv) > 1 else "lp"
pop = int(sys.argv[2]) if len(sys.argv) > 2 else 32768
c = N.Context(0)
_, v, f, _ = T.create_mesh(MESH_PATHS[mesh])
c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
a = T.setup_parameters()
c.upload_grid(a.problem.tensor_grid_r, a.problem.tensor_grid_theta, a.problem.tensor_grid_phi, np.asarray(a.problem.bool_tensor))
p = N.make_params(stop_on_collision=1)
xs = torch.from_numpy(init_like_population(pop, 2, 1005)).cuda()

def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r

full_ms, res = t(lambda: c.population_fitness(p, xs, FIXED_POS, 2, 1024))
cnt = res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)
out = {"mesh": mesh, "pop": pop, "full_ms": full_ms, "n_rhs": int(cnt["n_rhs"].sum())}
print(out, flush=True)
cost_ms, cost = t(lambda: c.predict_cost(p, xs, FIXED_POS, 2))
order = c.cost_order(cost)
out["predict_cost_full_ms"] = cost_ms
for world in (2, 4, 8):
    xl = c.deal_rows(xs, order, 0, world)
    c.set_queue_order(True)
    ms, r = t(lambda: c.population_fitness(p, xl, FIXED_POS, 2, 1024))
    c.set_queue_order(False)
    lc = r["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)
    share = lc["n_rhs"].sum() / cnt["n_rhs"].sum()
    out[f"world{world}"] = {"shard_ms": ms, "ideal_ms": full_ms / world, "efficiency": full_ms / world / ms,
                            "work_share_x_world": float(share * world), "predict_ms_per_rank": cost_ms / world}
    print(world, out[f"world{world}"], flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
Path(f"gpurun_out/shard_probe_{mesh}.json").write_text(json.dumps(out, indent=1))
