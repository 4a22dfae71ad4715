"""tiny run of every kernel for compute-sanitizer (memcheck / racecheck): small meshes, short windows"""
import os, sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
from toss_b200 import _native as N
from oracle import oracle as O
from tests.test_gpu_parity import random_population, FIXED_POS, points
for mesh, pop, ft in (("lllp", 9, 20000.0), ("lp", 10, 3000.0)):
    c = N.Context(0); v, f = O.load_mesh(mesh); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
    p = N.make_params(final_time=ft, stop_on_collision=1)
    xs = random_population(pop, 2, 3, ft)
    r = c.population_fitness_host(p, xs, FIXED_POS, 2, knot_cap=256)
    ref = O.population_fitness(O.Mesh(v, f), O.default_params(final_time=ft, stop_on_collision=1), O.default_grid(), xs, FIXED_POS, 2)
    assert np.array_equal(r["fitness"], ref["fitness"]), mesh
    prop = c.propagate(p, xs, FIXED_POS, 2, 256)
    c.resample(p, prop); c.collision_verdict(p, prop)
    pts = points(4200, 1)
    c.gravity(pts, potential=True); c.gravity(pts[:7]); c.is_outside(pts[:50])
    c.compute_motion(p, np.zeros(5), np.hstack([pts[:5], pts[:5]]))
    print(mesh, "ok", flush=True)
    c.close()
