import os, sys; from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from bench import init_like_population, FIXED_POS
c = N.Context(0); v, f = O.load_mesh("lp"); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
pop = int(sys.argv[2]); p = N.make_params(stop_on_collision=int(sys.argv[1]))
xs = torch.from_numpy(init_like_population(pop, 2, 1005)).cuda()
c.integrate(p, xs, FIXED_POS, 2, 1024); torch.cuda.synchronize()
best = 1e9
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); c.integrate(p, xs, FIXED_POS, 2, 1024); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print(f"LPT={os.environ.get('TOSS_B200_LPT','0')} TEAM={os.environ.get('TOSS_B200_TEAM','auto')} stop={sys.argv[1]} pop={pop}: {best:.1f} ms", flush=True)
