"""developer probe: HBM roofline of the materialising resampler (K3) and the fitness-from-positions kernel (K4)"""
import sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, FIXED_POS
c = N.Context(0); v, f = O.load_mesh("llp"); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
p = N.make_params(); pop = 4096
xs = torch.from_numpy(init_like_population(pop, 2, 1005)).cuda()
prop = c.propagate(p, xs, FIXED_POS, 2, 1024)
def t(fn, reps=5):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3
Nsamp = c.num_samples(p)
s3 = t(lambda: c.resample(p, prop))
pos, vel, ts = c.resample(p, prop)
s4 = t(lambda: c.fitness(9, p, pos, ts))
b3 = pop * 6 * Nsamp * 8; b4 = pop * 3 * Nsamp * 8
print(f"K3 resample: {s3*1e3:.2f} ms, writes {b3/1e9:.2f} GB -> {b3/s3/1e9:.0f} GB/s (peak 6539)")
print(f"K4 fitness(9) from positions: {s4*1e3:.2f} ms, reads {b4/1e9:.2f} GB -> {b4/s4/1e9:.0f} GB/s")
