"""-DK2_PROFILE build: where an owner warp's clocks go per RHS evaluation, by section of the stepper loop
(small populations: the per-evaluation latency is what bounds BASELINE configs[0] and configs[2])."""
import os, sys; from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
os.environ["TOSS_B200_LIB"] = str(ROOT / "build" / "variants" / "libtoss_prof.so")
os.environ["TOSS_B200_LPT"] = "0"          # the queue's order area doubles as the profile mailbox
sys.path.insert(0, str(ROOT))
import numpy as np, torch
from toss_b200 import _native as N
from oracle import oracle as O
from tests.test_gpu_parity import random_population, FIXED_POS
NAMES = ["advance/fetch", "stage point", "shuffle+rotation", "publish+ray", "gravity sweep", "reduce+k store"]
for mesh, pop in (("lllp", 64), ("llp", 256), ("lp", 512), ("lp", 4096)):
    c = N.Context(0); v, f = O.load_mesh(mesh); c.upload_mesh(v, f, 533.0); c.upload_grid(*O.default_grid())
    p = N.make_params(); xs = torch.from_numpy(random_population(pop, 2, 1005)).cuda()
    c.integrate(p, xs, FIXED_POS, 2, 1024); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); pr = c.integrate(p, xs, FIXED_POS, 2, 1024); e1.record(); torch.cuda.synchronize()
    q = pr.ws[pr.L.queue + 16:pr.L.queue + 16 + 8 * 30].view(torch.int64).cpu().numpy()
    sec = q[24:30].astype(float); nr = pr.to_host()["counters"]["n_rhs"]
    print(f"{mesh} pop {pop}: {e0.elapsed_time(e1):.2f} ms  max_rhs {nr.max()}  {e0.elapsed_time(e1) * 1e3 / nr.max():.2f} us per RHS of the longest; "
          f"cycles per RHS evaluation (owner warps): " + "  ".join(f"{n} {s / nr.sum():.0f}" for n, s in zip(NAMES, sec)) + f"  total {sec.sum() / nr.sum():.0f}", flush=True)
    c.close()
