"""developer probe: like shard_probe.py, for the schedule options of a shard (static-round layout, capped CTAs) -- the
options are read from the environment at every launch, so one process tries them all.
usage: shard_probe2.py [world ...]"""
import os, sys, json
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
import toss_b200 as T
from bench import init_like_population, FIXED_POS, MESH_PATHS

worlds = [int(a) for a in sys.argv[1:]] or [8, 4]
c = N.Context(0)
_, v, f, _ = T.create_mesh(MESH_PATHS["lp"])
c.upload_mesh(v, np.ascontiguousarray(f, dtype=np.int32), 533.0)
a = T.setup_parameters()
c.upload_grid(a.problem.tensor_grid_r, a.problem.tensor_grid_theta, a.problem.tensor_grid_phi, np.asarray(a.problem.bool_tensor))
p = N.make_params(stop_on_collision=1)
pop = 32768
xs = torch.from_numpy(init_like_population(pop, 2, 1005)).cuda()


def t(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): r = fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps, r


full_ms, res = t(lambda: c.population_fitness(p, xs, FIXED_POS, 2, 1024), 2)
cnt = res["counters"].cpu().numpy().view(N.COUNTERS_DTYPE)
cost_ms, cost = t(lambda: c.predict_cost(p, xs, FIXED_POS, 2))
ch = cost.cpu().numpy()
surv = cnt["n_rhs"] > 1400
print(json.dumps({"full_ms": full_ms, "predict_ms": cost_ms, "corr_all": float(np.corrcoef(ch, cnt["n_rhs"])[0, 1]),
                  "corr_survivors": float(np.corrcoef(ch[surv], cnt["n_rhs"][surv])[0, 1])}), flush=True)
Path("gpurun_out").mkdir(exist_ok=True)
np.savez("gpurun_out/dist_lp_r2g.npz", n_rhs=cnt["n_rhs"], cost=ch)
order = c.cost_order(cost)
variants = [("default", {}), ("old", {"TOSS_B200_CAPS": "0", "TOSS_B200_INTERLEAVE": "0"}),
            ("interleave", {"TOSS_B200_CAPS": "0", "TOSS_B200_INTERLEAVE": "1"}),
            ("caps4x8", {"TOSS_B200_CAPS": "4,8"}), ("caps8x8", {"TOSS_B200_CAPS": "8,8"}),
            ("caps16x8", {"TOSS_B200_CAPS": "16,8"}), ("caps8x12", {"TOSS_B200_CAPS": "8,12"}),
            ("caps16x12", {"TOSS_B200_CAPS": "16,12"}), ("caps8x4", {"TOSS_B200_CAPS": "8,4"})]
for world in worlds:
    xl = c.deal_rows(xs, order, 0, world)
    c.set_queue_order(True)
    for name, env in variants:
        for k in ("TOSS_B200_CAPS", "TOSS_B200_INTERLEAVE"):
            os.environ.pop(k, None)
        os.environ.update(env)
        ms, r = t(lambda: c.population_fitness(p, xl, FIXED_POS, 2, 1024))
        print(world, name, round(ms, 2), "eff", round(full_ms / world / ms, 4), flush=True)
    c.set_queue_order(False)
