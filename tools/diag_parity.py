import sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from tests.test_gpu_parity import *
from toss_b200._native import COUNTERS_DTYPE
for mesh_name, pop, n_man, kind in [("llp", 96, 2, "init"), ("llp", 64, 2, "bound"), ("lllp", 16, 1, "init"), ("lp", 24, 2, "init")]:
    c, m = ctx_for(mesh_name)
    ft = 604800.0 if mesh_name != "lp" else 86400.0
    p, po = both_params(final_time=ft)
    xs = (random_population if kind == "init" else bound_population)(pop, n_man, 1000 + pop, ft)
    res = c.population_fitness(p, xs, FIXED_POS, n_man)
    fit = res["fitness"].cpu().numpy(); coll = res["collision"].cpu().numpy(); nvis = res["n_visited"].cpu().numpy()
    cnt = res["counters"].cpu().numpy().view(COUNTERS_DTYPE)
    ref = O.population_fitness(m, po, O.default_grid(), xs, FIXED_POS, n_man)
    ok = coll == 0
    rel = np.abs(fit - ref["fitness"]) / np.abs(ref["fitness"])
    print(mesh_name, kind, "coll equal", np.array_equal(coll, ref["collision"]), "ncoll", coll.sum())
    print("  dacc", (cnt["n_accepted"] - ref["counters"]["n_accepted"]).tolist())
    print("  drej", (cnt["n_rejected"] - ref["counters"]["n_rejected"]).tolist())
    print("  dev ", (cnt["n_event_points"] - ref["counters"]["n_event_points"]).tolist())
    print("  rel fitness sorted", np.sort(rel[ok])[::-1][:12])
    print("  nvis diff", (nvis - ref["n_visited"])[ok].tolist())
    h = c.propagate(p, xs, FIXED_POS, n_man).to_host()
    d = []
    for i in range(pop):
        tr = O.compute_trajectory(m, po, np.concatenate((FIXED_POS, xs[i])), n_man)
        n = h["seg_start"][i, h["n_seg"][i]]
        d.append(np.max(np.abs(h["knots_y"][i, n - 1] - tr.y[-1]) / (np.abs(tr.y[-1]) + 1e-3)))
    print("  final-state rel diff max", max(d))
